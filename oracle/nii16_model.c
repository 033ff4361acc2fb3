/*
 * oracle/nii16_model.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C integer model of the NON-PARITY decoder mode "nii16" of modulations_b200 (the single-pass
 * next-iteration-initialisation decoder in 16-bit fixed point, two frames per register on the GPU;
 * modulations_b200/csrc/nii16_core.cuh states the format).  It is NOT a restatement of the reference (that is
 * oracle/turbo_oracle.c).  It exists so that the CUDA kernel of that mode can be checked bit for bit against a
 * naive implementation of the SAME definition: full gamma[N][16][4] / alpha / beta arrays in int32, 4-way maxima
 * over the reference's prev/next tables, no merged branches, no packing.  Every value that the kernel keeps in 16
 * bits is range-checked here: a result outside [-32768, 32767] makes nii16_decode return -2, so an unsafe format
 * cannot pass the parity tests silently.
 *
 * Parity status: "parity unpinned" BY DESIGN (no reference output exists for this mode); pinned by kernel == model
 * and by BER/FER against the parity mode inside confidence intervals (tests/test_gpu_nii.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NS 16
#define CHAN_MAX 127
#define EXT_MAX 255

static int ovf;
static inline int chk(int v) { if (v < -32768 || v > 32767) ovf = 1; return v; }

static void nii16_siso(const int *Lc_A, const int *Lc_B, const int *Lc_W, const int *Lc_Y, const int *La_A,
                       const int *La_B, const int32_t *next_st, const int32_t *out_W, const int32_t *out_Y,
                       const int32_t *prev_st, const int32_t *prev_inp, int N, int sf_q, int *a0, int *b0,
                       int *Le_A, int *Le_B, int *scratch /* N*64 + 2*(N+1)*16 ints */)
{
    const int NEG = -(1 << 30);
    int *gamma = scratch, *alpha = gamma + (size_t)N * 64, *beta = alpha + (size_t)(N + 1) * 16;
    for (int k = 0; k < N; ++k) {
        const int in_A = chk(Lc_A[k] + La_A[k]), in_B = chk(Lc_B[k] + La_B[k]);
        for (int s = 0; s < NS; ++s)
            for (int inp = 0; inp < 4; ++inp) {
                const int bit_A = (inp >> 1) & 1, bit_B = inp & 1;
                const int bit_W = out_W[s * 4 + inp], bit_Y = out_Y[s * 4 + inp];
                gamma[(size_t)k * 64 + s * 4 + inp] = chk((bit_A ? -in_A : in_A) + (bit_B ? -in_B : in_B) +
                                                          (bit_W ? -Lc_W[k] : Lc_W[k]) + (bit_Y ? -Lc_Y[k] : Lc_Y[k]));
            }
    }
    memcpy(alpha, a0, sizeof(int) * NS);
    for (int k = 0; k < N; ++k) {
        for (int ns = 0; ns < NS; ++ns) {
            int mx = NEG;
            for (int idx = 0; idx < 4; ++idx) {
                const int ps = prev_st[ns * 4 + idx], inp = prev_inp[ns * 4 + idx];
                const int t = chk(alpha[k * 16 + ps] + gamma[(size_t)k * 64 + ps * 4 + inp]);
                if (t > mx) mx = t;
            }
            alpha[(k + 1) * 16 + ns] = mx;
        }
        const int norm = alpha[(k + 1) * 16];
        for (int s = 0; s < NS; ++s) alpha[(k + 1) * 16 + s] = chk(alpha[(k + 1) * 16 + s] - norm);
    }
    memcpy(beta + (size_t)N * 16, b0, sizeof(int) * NS);
    for (int k = N - 1; k >= 0; --k) {
        for (int s = 0; s < NS; ++s) {
            int mx = NEG;
            for (int inp = 0; inp < 4; ++inp) {
                const int t = chk(beta[(k + 1) * 16 + next_st[s * 4 + inp]] + gamma[(size_t)k * 64 + s * 4 + inp]);
                if (t > mx) mx = t;
            }
            beta[k * 16 + s] = mx;
        }
        const int norm = beta[k * 16];
        for (int s = 0; s < NS; ++s) beta[k * 16 + s] = chk(beta[k * 16 + s] - norm);
    }
    memcpy(a0, alpha + (size_t)N * 16, sizeof(int) * NS);
    memcpy(b0, beta, sizeof(int) * NS);
    for (int k = 0; k < N; ++k) {
        int app[4] = {NEG, NEG, NEG, NEG};
        for (int s = 0; s < NS; ++s)
            for (int inp = 0; inp < 4; ++inp) {
                const int m = chk(chk(alpha[k * 16 + s] + beta[(k + 1) * 16 + next_st[s * 4 + inp]]) +
                                  gamma[(size_t)k * 64 + s * 4 + inp]);
                if (m > app[inp]) app[inp] = m;
            }
        const int pA0 = app[0] > app[1] ? app[0] : app[1], pA1 = app[2] > app[3] ? app[2] : app[3];
        const int pB0 = app[0] > app[2] ? app[0] : app[2], pB1 = app[1] > app[3] ? app[1] : app[3];
        const int YA = Lc_A[k] + La_A[k], YB = Lc_B[k] + La_B[k];
        int a = (((pA0 - pA1) - 2 * YA) * sf_q + 64) >> 7;      /* arithmetic shift (gcc: sign-propagating) */
        int b = (((pB0 - pB1) - 2 * YB) * sf_q + 64) >> 7;
        if (a > EXT_MAX) a = EXT_MAX;
        if (a < -EXT_MAX) a = -EXT_MAX;
        if (b > EXT_MAX) b = EXT_MAX;
        if (b < -EXT_MAX) b = -EXT_MAX;
        Le_A[k] = a; Le_B[k] = b;
    }
}

static int quant(float llr)
{
    int q = (int)rintf(llr * 4.0f);             /* round to nearest even, like cvt.rni on the device */
    return q > CHAN_MAX ? CHAN_MAX : (q < -CHAN_MAX ? -CHAN_MAX : q);
}

/* Full decode of one frame.  -1: llr too short; -2: a 16-bit quantity overflowed. */
int nii16_decode(int N, int iterations, const int32_t *next_st, const int32_t *out_W, const int32_t *out_Y,
                 const int32_t *prev_st, const int32_t *prev_inp, const int32_t *perm, const int32_t *inv_perm,
                 const uint8_t *punct, int period, int sf_inner_q, int sf_last_q, const float *llr, int n_llr,
                 int32_t *decoded)
{
    const size_t nf = (size_t)N;
    int *f = (int *)calloc(nf * 16 + nf * 64 + 2 * (nf + 1) * 16, sizeof(int));
    int *Lc_A = f, *Lc_B = f + nf, *Lc_W1 = f + 2 * nf, *Lc_Y1 = f + 3 * nf, *Lc_W2 = f + 4 * nf;
    int *Lc_Y2 = f + 5 * nf, *Lc_Ai = f + 6 * nf, *Lc_Bi = f + 7 * nf;
    int *La_A = f + 8 * nf, *La_B = f + 9 * nf, *Le1_A = f + 10 * nf, *Le1_B = f + 11 * nf;
    int *La2_A = f + 12 * nf, *La2_B = f + 13 * nf, *Le2_A = f + 14 * nf, *Le2_B = f + 15 * nf;
    int *scratch = f + 16 * nf;
    int st[4][NS];
    memset(st, 0, sizeof st);
    int idx = 0, rc = 0;
    ovf = 0;
    for (int i = 0; i < N; ++i) {
        const int p = i % period;
        if (idx + 2 > n_llr) { rc = -1; goto done; }
        Lc_A[i] = quant(llr[idx++]); Lc_B[i] = quant(llr[idx++]);
        if (punct[0 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_W1[i] = quant(llr[idx++]); }
        if (punct[1 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_Y1[i] = quant(llr[idx++]); }
        if (punct[2 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_W2[i] = quant(llr[idx++]); }
        if (punct[3 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_Y2[i] = quant(llr[idx++]); }
    }
    for (int i = 0; i < N; ++i) { Lc_Ai[i] = Lc_A[perm[i]]; Lc_Bi[i] = Lc_B[perm[i]]; }
    for (int it = 0; it < iterations; ++it) {
        const int sf = (it < iterations - 1) ? sf_inner_q : sf_last_q;
        nii16_siso(Lc_A, Lc_B, Lc_W1, Lc_Y1, La_A, La_B, next_st, out_W, out_Y, prev_st, prev_inp, N, sf,
                   st[0], st[1], Le1_A, Le1_B, scratch);
        for (int i = 0; i < N; ++i) { La2_A[i] = Le1_A[perm[i]]; La2_B[i] = Le1_B[perm[i]]; }
        nii16_siso(Lc_Ai, Lc_Bi, Lc_W2, Lc_Y2, La2_A, La2_B, next_st, out_W, out_Y, prev_st, prev_inp, N, sf,
                   st[2], st[3], Le2_A, Le2_B, scratch);
        for (int i = 0; i < N; ++i) { La_A[i] = Le2_A[inv_perm[i]]; La_B[i] = Le2_B[inv_perm[i]]; }
    }
    for (int i = 0; i < N; ++i) {
        const int LA = chk(chk(Lc_A[i] + La_A[i]) + Le1_A[i]), LB = chk(chk(Lc_B[i] + La_B[i]) + Le1_B[i]);
        decoded[2 * i] = LA < 0 ? 1 : 0;
        decoded[2 * i + 1] = LB < 0 ? 1 : 0;
    }
    if (ovf) rc = -2;
done:
    free(f);
    return rc;
}

int nii16_decode_batch(int B, int N, int iterations, const int32_t *next_st, const int32_t *out_W,
                       const int32_t *out_Y, const int32_t *prev_st, const int32_t *prev_inp, const int32_t *perm,
                       const int32_t *inv_perm, const uint8_t *punct, int period, int sf_inner_q, int sf_last_q,
                       const float *llr, int n_llr, int32_t *decoded)
{
    int worst = 0;
    for (int b = 0; b < B; ++b) {
        const int rc = nii16_decode(N, iterations, next_st, out_W, out_Y, prev_st, prev_inp, perm, inv_perm, punct,
                                    period, sf_inner_q, sf_last_q, llr + (size_t)b * n_llr, n_llr,
                                    decoded + (size_t)b * 2 * N);
        if (rc < worst) worst = rc;
    }
    return worst;
}

/* One SISO on already-quantised inputs (for the CPU emulator test of nii16_core.cuh). */
int nii16_siso_ext(const int *Lc_A, const int *Lc_B, const int *Lc_W, const int *Lc_Y, const int *La_A, const int *La_B,
                   const int32_t *next_st, const int32_t *out_W, const int32_t *out_Y, const int32_t *prev_st,
                   const int32_t *prev_inp, int N, int sf_q, int *a0, int *b0, int *Le_A, int *Le_B, int *scratch)
{
    ovf = 0;
    nii16_siso(Lc_A, Lc_B, Lc_W, Lc_Y, La_A, La_B, next_st, out_W, out_Y, prev_st, prev_inp, N, sf_q, a0, b0, Le_A, Le_B, scratch);
    return ovf ? -2 : 0;
}
