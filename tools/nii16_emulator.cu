// tools/nii16_emulator.cu — TEST INFRASTRUCTURE.  The lane schedule of the non-parity "nii16" decoder mode (two frames
// per 32-bit register, DPX add-compare-select) replayed on the CPU with the kernel's own arithmetic header
// (nii16_core.cuh; its packed operations have host definitions) and compared with the naive integer model
// oracle/nii16_model.c, frame by frame, four chained SISOs with the boundary metrics carried over.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../modulations_b200/csrc/nii16_core.cuh"

extern "C" {
void orc_build_trellis(int32_t *next_state, int32_t *out_W, int32_t *out_Y, int32_t *prev_state,
                       int32_t *prev_input, int32_t *G);
int nii16_siso_ext(const int *Lc_A, const int *Lc_B, const int *Lc_W, const int *Lc_Y, const int *La_A, const int *La_B,
                   const int32_t *next_st, const int32_t *out_W, const int32_t *out_Y, const int32_t *prev_st,
                   const int32_t *prev_inp, int N, int sf_q, int *a0, int *b0, int *Le_A, int *Le_B, int *scratch);
}

using namespace b200dvb;
using namespace b200dvb::nii16;
using tpf::rho4;
constexpr int W = 4;

static void window(int len, int w0, p16 (&X)[16], p16 (&Z)[16], const std::vector<p16> &rec, const std::vector<p16> &Y,
                   int sf, p16 *LeA, p16 *LeB)
{
    p16 store[W][16];
    for (int jj = 0; jj < len; ++jj) {
        const int k = w0 + len - 1 - jj;
        for (int s = 0; s < 16; ++s) store[len - 1 - jj][s] = Z[s];
        p16 g[8]; memcpy(g, &rec[k * 8], sizeof g);
        bwd_step(Z, g);
    }
    for (int jj = 0; jj < len; ++jj) {
        const int k = w0 + jj;
        p16 g[8]; memcpy(g, &rec[k * 8], sizeof g);
        p16 uv[4];
        app_maxima(X, store[jj], g, uv);
        pass_step(X, g, false);
        make_extrinsic(uv, Y[2 * k], Y[2 * k + 1], sf, LeA[k], LeB[k]);
    }
}

static int run(int N, unsigned seed, int amp, int iters)
{
    srand(seed);
    // two independent frames f = 0, 1 packed into the low / high halves
    std::vector<int> LcA[2], LcB[2], LcW[2], LcY[2], LaA[2], LaB[2], refA[2], refB[2];
    for (int f = 0; f < 2; ++f) {
        LcA[f].resize(N); LcB[f].resize(N); LcW[f].resize(N); LcY[f].resize(N);
        LaA[f].assign(N, 0); LaB[f].assign(N, 0); refA[f].resize(N); refB[f].resize(N);
        for (int k = 0; k < N; ++k) {
            auto r = [&]() { int v = rand() % (2 * amp + 1) - amp; return amp == kChanMax && (rand() & 3) == 0 ? (v < 0 ? -kChanMax : kChanMax) : v; };
            LcA[f][k] = r(); LcB[f][k] = r(); LcW[f][k] = (k % 3 == 0) ? 0 : r(); LcY[f][k] = r();
        }
    }
    int32_t ns[64], oW[64], oY[64], ps[64], pi[64], G[16];
    orc_build_trellis(ns, oW, oY, ps, pi, G);
    std::vector<int> scratch((size_t)N * 64 + 2 * (N + 1) * 16);
    int ra0[2][16] = {{0}}, rb0[2][16] = {{0}};
    p16 ea0[16] = {0}, eb0[16] = {0};
    std::vector<p16> LeA(N), LeB(N);
    int bad = 0;
    for (int it = 0; it < iters; ++it) {
        const int sf = it + 1 < iters ? kSfInner : kSfLast;
        for (int f = 0; f < 2; ++f)
            if (nii16_siso_ext(LcA[f].data(), LcB[f].data(), LcW[f].data(), LcY[f].data(), LaA[f].data(), LaB[f].data(), ns, oW, oY,
                               ps, pi, N, sf, ra0[f], rb0[f], refA[f].data(), refB[f].data(), scratch.data())) { printf("  model overflow\n"); ++bad; }
        const int M = N / 2, nfull = M / W, r = M % W, nslots = nfull + (r ? 1 : 0);
        std::vector<p16> rec((size_t)N * 8), Y((size_t)N * 2);
        auto prep = [&](int k) {
            Y[2 * k] = add2(pack16(LcA[0][k], LcA[1][k]), pack16(LaA[0][k], LaA[1][k]));
            Y[2 * k + 1] = add2(pack16(LcB[0][k], LcB[1][k]), pack16(LaB[0][k], LaB[1][k]));
            p16 g[8]; make_record(Y[2 * k], Y[2 * k + 1], pack16(LcW[0][k], LcW[1][k]), pack16(LcY[0][k], LcY[1][k]), g);
            memcpy(&rec[k * 8], g, sizeof g);
        };
        p16 va[16], vb[16];
        for (int s = 0; s < 16; ++s) { va[s] = ea0[s]; vb[s] = eb0[rho4(s)]; }
        std::vector<p16> ckA((size_t)nslots * 16), ckB((size_t)nslots * 16);
        for (int j = 0; j < M; ++j) {
            int slot = -1;
            if ((M - j) % W == 0) slot = (M - j) / W - 1;
            else if (j == 0) slot = nfull;
            if (slot >= 0)
                for (int s = 0; s < 16; ++s) { ckA[slot * 16 + s] = va[s]; ckB[slot * 16 + rho4(s)] = vb[s]; }
            prep(j); prep(N - 1 - j);
            p16 g[8];
            memcpy(g, &rec[j * 8], sizeof g); pass_step(va, g, false);
            memcpy(g, &rec[(N - 1 - j) * 8], sizeof g); pass_step(vb, g, true);
        }
        p16 Ra[16], Rb[16];
        for (int s = 0; s < 16; ++s) { Ra[rho4(s)] = vb[s]; Rb[s] = va[s]; }
        for (int i = 0; i < nslots; ++i) {
            const bool ragged = (i == nfull);
            const int len = ragged ? r : W;
            p16 X[16], Z[16];
            const int wa = ragged ? 0 : M - (i + 1) * W;
            for (int s = 0; s < 16; ++s) { Z[s] = Ra[s]; X[s] = ckA[i * 16 + s]; }
            window(len, wa, X, Z, rec, Y, sf, LeA.data(), LeB.data());
            for (int s = 0; s < 16; ++s) Ra[s] = Z[s];
            const int wb = M + i * W;
            for (int s = 0; s < 16; ++s) { X[s] = Rb[s]; Z[s] = ckB[i * 16 + s]; }
            window(len, wb, X, Z, rec, Y, sf, LeA.data(), LeB.data());
            for (int s = 0; s < 16; ++s) Rb[s] = X[s];
        }
        for (int s = 0; s < 16; ++s) { eb0[s] = Ra[s]; ea0[s] = Rb[s]; }
        for (int k = 0; k < N; ++k) {
            const int g[4] = {lo16(LeA[k]), hi16(LeA[k]), lo16(LeB[k]), hi16(LeB[k])};
            const int w[4] = {refA[0][k], refA[1][k], refB[0][k], refB[1][k]};
            if (memcmp(g, w, sizeof g)) {
                if (bad < 5) printf("  N=%d it=%d k=%d: got (%d,%d | %d,%d) want (%d,%d | %d,%d)\n", N, it, k, g[0], g[1], g[2], g[3], w[0], w[1], w[2], w[3]);
                ++bad;
            }
        }
        for (int s = 0; s < 16; ++s)
            if (lo16(ea0[s]) != ra0[0][s] || hi16(ea0[s]) != ra0[1][s] || lo16(eb0[s]) != rb0[0][s] || hi16(eb0[s]) != rb0[1][s]) {
                printf("  N=%d it=%d: boundary metrics differ at state %d\n", N, it, s); ++bad; break;
            }
        for (int f = 0; f < 2; ++f)
            for (int k = 0; k < N; ++k) { LaA[f][k] = refA[f][(k * 7 + 3) % N]; LaB[f][k] = refB[f][(k * 5 + 1) % N]; }
    }
    printf("N=%d seed=%u amp=%d: %s (%d mismatches)\n", N, seed, amp, bad ? "FAIL" : "ok", bad);
    return bad;
}

int main()
{
    int bad = 0;
    const int Ns[] = {48, 64, 212, 220, 16, 20, 12};
    for (int N : Ns)
        for (unsigned seed = 1; seed <= 3; ++seed) bad += run(N, seed, seed == 3 ? kChanMax : 24, 4);
    return bad ? 1 : 0;
}
