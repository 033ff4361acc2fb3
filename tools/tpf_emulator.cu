// tools/tpf_emulator.cu — TEST INFRASTRUCTURE.  Replays the schedule of the thread-per-frame
// decoder (modulations_b200/csrc/decode_tpf.cu) for one frame on the CPU, using the very
// same arithmetic core (tpf_core.cuh), and checks one SISO against oracle/turbo_oracle.c
// bit for bit.  Build + run:  make -C tools tpf_emulator   (or see tools/Makefile)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../modulations_b200/csrc/tpf_core.cuh"

extern "C" {
void orc_build_trellis(int32_t *next_state, int32_t *out_W, int32_t *out_Y, int32_t *prev_state,
                       int32_t *prev_input, int32_t *G);
void orc_bcjr_max_log_map(const float *Lc_A, const float *Lc_B, const float *Lc_W, const float *Lc_Y,
                          const double *La_A, const double *La_B, const int32_t *next_st,
                          const int32_t *out_W, const int32_t *out_Y, const int32_t *prev_st,
                          const int32_t *prev_inp, int N, double scaling_factor, double *Le_A,
                          double *Le_B, float *scratch);
}

using namespace b200dvb::tpf;
constexpr int W = 4;

static double urand() { return (double)rand() / RAND_MAX; }

template <int LEN>
static void window(int N, int w0, float (&X)[16], float (&Z)[16], const std::vector<float> &rec,
                   const std::vector<double> &Y, double sf, double *LeA, double *LeB)
{
    float store[LEN][16];
    for (int jj = 0; jj < LEN; ++jj) {
        const int k = w0 + LEN - 1 - jj;
        for (int s = 0; s < 16; ++s) store[LEN - 1 - jj][s] = Z[s];     // beta[k+1]
        float g[8]; memcpy(g, &rec[k * 8], sizeof g);
        bwd_step(Z, g);
    }
    for (int jj = 0; jj < LEN; ++jj) {
        const int k = w0 + jj;
        float g[8]; memcpy(g, &rec[k * 8], sizeof g);
        float uv[4];
        ext_step(X, store[jj], g, uv);
        make_extrinsic(uv, Y[2 * k], Y[2 * k + 1], sf, LeA[k], LeB[k]);
    }
}

static int run(int N, unsigned seed, double scale)
{
    srand(seed);
    std::vector<float> LcA(N), LcB(N), LcW(N), LcY(N);
    std::vector<double> LaA(N), LaB(N), refA(N), refB(N), LeA(N), LeB(N);
    for (int k = 0; k < N; ++k) {
        LcA[k] = (float)((urand() - 0.5) * scale); LcB[k] = (float)((urand() - 0.5) * scale);
        LcW[k] = (k % 3 == 0) ? 0.f : (float)((urand() - 0.5) * scale);
        LcY[k] = (float)((urand() - 0.5) * scale);
        LaA[k] = (urand() - 0.5) * scale * 1.37; LaB[k] = (urand() - 0.5) * scale * 0.73;
    }
    int32_t ns[64], oW[64], oY[64], ps[64], pi[64], G[16];
    orc_build_trellis(ns, oW, oY, ps, pi, G);
    std::vector<float> scratch((size_t)N * 64 + 2 * (N + 1) * 16);
    const double sf = 0.7;
    orc_bcjr_max_log_map(LcA.data(), LcB.data(), LcW.data(), LcY.data(), LaA.data(), LaB.data(), ns, oW, oY,
                         ps, pi, N, sf, refA.data(), refB.data(), scratch.data());

    // ---- emulated warp schedule: one alpha lane + one beta lane of the same frame ----
    const int M = N / 2, nfull = M / W, r = M % W, nslots = nfull + (r ? 1 : 0);
    std::vector<float> rec((size_t)N * 8);
    std::vector<double> Y((size_t)N * 2);
    auto prep = [&](int k) {
        Y[2 * k] = d_add((double)LcA[k], LaA[k]); Y[2 * k + 1] = d_add((double)LcB[k], LaB[k]);
        float g[8]; make_record(Y[2 * k], Y[2 * k + 1], LcW[k], LcY[k], g);
        memcpy(&rec[k * 8], g, sizeof g);
    };
    float va[16] = {0}, vb[16] = {0};
    for (int j = 0; j < N; ++j) {                       // pass 1
        const int ka = j, kb = N - 1 - j;
        if (j < M) { prep(ka); prep(kb); }
        float g[8];
        memcpy(g, &rec[ka * 8], sizeof g); pass_step(va, g, false);
        memcpy(g, &rec[kb * 8], sizeof g); pass_step(vb, g, true);
    }
    std::vector<float> ckA((size_t)nslots * 16), ckB((size_t)nslots * 16);
    for (int j = 0; j < M; ++j) {                       // pass 2 up to the crossing point
        int slot = -1;
        if ((M - j) % W == 0) slot = (M - j) / W - 1;
        else if (j == 0) slot = nfull;
        if (slot >= 0)
            for (int s = 0; s < 16; ++s) { ckA[slot * 16 + s] = va[s]; ckB[slot * 16 + rho4(s)] = vb[s]; }
        float g[8];
        memcpy(g, &rec[j * 8], sizeof g); pass_step(va, g, false);
        memcpy(g, &rec[(N - 1 - j) * 8], sizeof g); pass_step(vb, g, true);
    }
    // hand-over at the crossing point: the alpha lane continues beta downwards, the beta lane alpha upwards
    float Ra[16], Rb[16];                               // running vector of the alpha lane (beta[M]) / beta lane (alpha[M])
    for (int s = 0; s < 16; ++s) { Ra[rho4(s)] = vb[s]; Rb[s] = va[s]; }
    for (int i = 0; i < nslots; ++i) {
        const bool ragged = (i == nfull);
        const int len = ragged ? r : W;
        float X[16], Z[16];
        // alpha lane: window [wa, wa+len) of [0, M): beta running, alpha from its checkpoint
        const int wa = ragged ? 0 : M - (i + 1) * W;
        for (int s = 0; s < 16; ++s) { Z[s] = Ra[s]; X[s] = ckA[i * 16 + s]; }
        switch (len) {
            case 8: window<8>(N, wa, X, Z, rec, Y, sf, LeA.data(), LeB.data()); break;
            case 6: window<6>(N, wa, X, Z, rec, Y, sf, LeA.data(), LeB.data()); break;
            case 4: window<4>(N, wa, X, Z, rec, Y, sf, LeA.data(), LeB.data()); break;
            case 2: window<2>(N, wa, X, Z, rec, Y, sf, LeA.data(), LeB.data()); break;
        }
        for (int s = 0; s < 16; ++s) Ra[s] = Z[s];
        // beta lane: window [wb, wb+len) of [M, N): alpha running, beta from its checkpoint
        const int wb = M + i * W;
        for (int s = 0; s < 16; ++s) { X[s] = Rb[s]; Z[s] = ckB[i * 16 + s]; }
        switch (len) {
            case 8: window<8>(N, wb, X, Z, rec, Y, sf, LeA.data(), LeB.data()); break;
            case 6: window<6>(N, wb, X, Z, rec, Y, sf, LeA.data(), LeB.data()); break;
            case 4: window<4>(N, wb, X, Z, rec, Y, sf, LeA.data(), LeB.data()); break;
            case 2: window<2>(N, wb, X, Z, rec, Y, sf, LeA.data(), LeB.data()); break;
        }
        for (int s = 0; s < 16; ++s) Rb[s] = X[s];
    }
    int bad = 0;
    for (int k = 0; k < N; ++k)
        if (memcmp(&LeA[k], &refA[k], 8) || memcmp(&LeB[k], &refB[k], 8)) {
            if (bad < 5) printf("  N=%d k=%d: got (%.17g, %.17g) want (%.17g, %.17g)\n", N, k, LeA[k], LeB[k], refA[k], refB[k]);
            ++bad;
        }
    printf("N=%d seed=%u scale=%g: %s (%d mismatches)\n", N, seed, scale, bad ? "FAIL" : "ok", bad);
    return bad;
}

int main()
{
    int bad = 0;
    const int Ns[] = {48, 64, 212, 220, 228, 16, 20, 12};
    for (int N : Ns)
        for (unsigned seed = 1; seed <= 3; ++seed) bad += run(N, seed, seed == 3 ? 90.0 : 8.0);
    return bad ? 1 : 0;
}
