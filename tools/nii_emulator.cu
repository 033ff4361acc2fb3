// tools/nii_emulator.cu — TEST INFRASTRUCTURE.  Replays, for one frame on the CPU, what an (alpha lane, beta lane)
// pair of the non-parity "nii" decoder (modulations_b200/csrc/decode_nii.cu) does — prep fused into the single
// "in" pass, bit-reversed labels for the backward lane, checkpoints, meet in the middle, recompute windows with the
// re-associated a-posteriori maxima, float32 epilogue, boundary metrics carried to the next iteration — with the
// kernel's own arithmetic header (nii_core.cuh compiled for the host), and compares `iters` chained SISOs with the
// naive model oracle/nii_model.c bit for bit.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../modulations_b200/csrc/nii_core.cuh"

extern "C" {
void orc_build_trellis(int32_t *next_state, int32_t *out_W, int32_t *out_Y, int32_t *prev_state,
                       int32_t *prev_input, int32_t *G);
void nii_siso(const float *Lc_A, const float *Lc_B, const float *Lc_W, const float *Lc_Y, const float *La_A,
              const float *La_B, const int32_t *next_st, const int32_t *out_W, const int32_t *out_Y,
              const int32_t *prev_st, const int32_t *prev_inp, int N, float sf, float *a0, float *b0, float *Le_A,
              float *Le_B, float *scratch);
}

using namespace b200dvb;
using tpf::bwd_step;
using tpf::pass_step;
using tpf::rho4;
constexpr int W = 4;

static double urand() { return (double)rand() / RAND_MAX; }

static void window(int len, int w0, float (&X)[16], float (&Z)[16], const std::vector<float> &rec,
                   const std::vector<float> &Y, float sf, float *LeA, float *LeB)
{
    float store[W][16];
    for (int jj = 0; jj < len; ++jj) {
        const int k = w0 + len - 1 - jj;
        for (int s = 0; s < 16; ++s) store[len - 1 - jj][s] = Z[s];     // beta[k+1]
        float g[8]; memcpy(g, &rec[k * 8], sizeof g);
        bwd_step(Z, g);
    }
    for (int jj = 0; jj < len; ++jj) {
        const int k = w0 + jj;
        float g[8]; memcpy(g, &rec[k * 8], sizeof g);
        float uv[4];
        nii::app_maxima(X, store[jj], g, uv);
        pass_step(X, g, false);
        nii::make_extrinsic(uv, Y[2 * k], Y[2 * k + 1], sf, LeA[k], LeB[k]);
    }
}

static int run(int N, unsigned seed, double scale, int iters)
{
    srand(seed);
    std::vector<float> LcA(N), LcB(N), LcW(N), LcY(N), LaA(N), LaB(N), refA(N), refB(N), LeA(N), LeB(N);
    for (int k = 0; k < N; ++k) {
        LcA[k] = (float)((urand() - 0.5) * scale); LcB[k] = (float)((urand() - 0.5) * scale);
        LcW[k] = (k % 3 == 0) ? 0.f : (float)((urand() - 0.5) * scale);
        LcY[k] = (float)((urand() - 0.5) * scale);
        LaA[k] = 0.f; LaB[k] = 0.f;
    }
    int32_t ns[64], oW[64], oY[64], ps[64], pi[64], G[16];
    orc_build_trellis(ns, oW, oY, ps, pi, G);
    std::vector<float> scratch((size_t)N * 64 + 2 * (N + 1) * 16);
    float ra0[16] = {0}, rb0[16] = {0};                 // model's boundary metrics
    float ea0[16] = {0}, eb0[16] = {0};                 // emulated lanes': alpha[0] (natural), beta[N] (natural)
    int bad = 0;
    for (int it = 0; it < iters; ++it) {
        const float sf = it + 1 < iters ? 0.7f : 1.0f;
        nii_siso(LcA.data(), LcB.data(), LcW.data(), LcY.data(), LaA.data(), LaB.data(), ns, oW, oY, ps, pi, N, sf,
                 ra0, rb0, refA.data(), refB.data(), scratch.data());
        // ---- emulated warp schedule ----
        const int M = N / 2, nfull = M / W, r = M % W, nslots = nfull + (r ? 1 : 0);
        std::vector<float> rec((size_t)N * 8), Y((size_t)N * 2);
        auto prep = [&](int k) {
            Y[2 * k] = tpf::f_add(LcA[k], LaA[k]); Y[2 * k + 1] = tpf::f_add(LcB[k], LaB[k]);
            float g[8]; nii::make_record(Y[2 * k], Y[2 * k + 1], LcW[k], LcY[k], g);
            memcpy(&rec[k * 8], g, sizeof g);
        };
        float va[16], vb[16];
        for (int s = 0; s < 16; ++s) { va[s] = ea0[s]; vb[s] = eb0[rho4(s)]; }      // beta lane works in rho4 labels
        std::vector<float> ckA((size_t)nslots * 16), ckB((size_t)nslots * 16);
        for (int j = 0; j < M; ++j) {                   // the single "in" pass, checkpoints on the way
            int slot = -1;
            if ((M - j) % W == 0) slot = (M - j) / W - 1;
            else if (j == 0) slot = nfull;
            if (slot >= 0)
                for (int s = 0; s < 16; ++s) { ckA[slot * 16 + s] = va[s]; ckB[slot * 16 + rho4(s)] = vb[s]; }
            prep(j); prep(N - 1 - j);
            float g[8];
            memcpy(g, &rec[j * 8], sizeof g); pass_step(va, g, false);
            memcpy(g, &rec[(N - 1 - j) * 8], sizeof g); pass_step(vb, g, true);
        }
        float Ra[16], Rb[16];
        for (int s = 0; s < 16; ++s) { Ra[rho4(s)] = vb[s]; Rb[s] = va[s]; }
        for (int i = 0; i < nslots; ++i) {
            const bool ragged = (i == nfull);
            const int len = ragged ? r : W;
            float X[16], Z[16];
            const int wa = ragged ? 0 : M - (i + 1) * W;
            for (int s = 0; s < 16; ++s) { Z[s] = Ra[s]; X[s] = ckA[i * 16 + s]; }
            window(len, wa, X, Z, rec, Y, sf, LeA.data(), LeB.data());
            for (int s = 0; s < 16; ++s) Ra[s] = Z[s];
            const int wb = M + i * W;
            for (int s = 0; s < 16; ++s) { X[s] = Rb[s]; Z[s] = ckB[i * 16 + s]; }
            window(len, wb, X, Z, rec, Y, sf, LeA.data(), LeB.data());
            for (int s = 0; s < 16; ++s) Rb[s] = X[s];
        }
        // boundary metrics for the next iteration: the alpha lane ends with beta[0], the beta lane with alpha[N]
        for (int s = 0; s < 16; ++s) { eb0[s] = Ra[s]; ea0[s] = Rb[s]; }
        for (int k = 0; k < N; ++k)
            if (memcmp(&LeA[k], &refA[k], 4) || memcmp(&LeB[k], &refB[k], 4)) {
                if (bad < 5) printf("  N=%d it=%d k=%d: got (%.9g, %.9g) want (%.9g, %.9g)\n", N, it, k, LeA[k], LeB[k], refA[k], refB[k]);
                ++bad;
            }
        if (memcmp(ea0, ra0, sizeof ea0) || memcmp(eb0, rb0, sizeof eb0)) { printf("  N=%d it=%d: boundary metrics differ\n", N, it); ++bad; }
        // feed the extrinsics back as the next a-priori values (a self-concatenated loop: enough to make every
        // iteration see different inputs and the carried boundary metrics)
        for (int k = 0; k < N; ++k) { LaA[k] = refA[(k * 7 + 3) % N]; LaB[k] = refB[(k * 5 + 1) % N]; }
    }
    printf("N=%d seed=%u scale=%g: %s (%d mismatches)\n", N, seed, scale, bad ? "FAIL" : "ok", bad);
    return bad;
}

int main()
{
    int bad = 0;
    const int Ns[] = {48, 64, 212, 220, 16, 20, 12};
    for (int N : Ns)
        for (unsigned seed = 1; seed <= 3; ++seed) bad += run(N, seed, seed == 3 ? 90.0 : 8.0, 4);
    return bad ? 1 : 0;
}
