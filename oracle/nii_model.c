/*
 * oracle/nii_model.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C model of the NON-PARITY decoder mode "nii" of modulations_b200 (single pass per SISO with
 * next-iteration initialisation of the circular boundary, float32 throughout, re-associated a-posteriori
 * maxima; modulations_b200/csrc/nii_core.cuh states the definition).  It is NOT a restatement of the reference:
 * the reference's decoder is oracle/turbo_oracle.c.  This file exists so that the CUDA kernel of that mode
 * (csrc/decode_nii.cu) can be checked bit for bit against an independent, naive implementation of the SAME
 * definition — full gamma[N][16][4], alpha[N+1][16], beta[N+1][16] arrays, 4-way maxima over the reference's
 * prev/next tables, no merged branches, no checkpoints — in tests/ only.
 *
 * Parity status: "parity unpinned" BY DESIGN — there is no reference output for this mode.  What pins it is
 * (a) kernel == this model, bit for bit, and (b) BER/FER of this mode against the parity mode inside binomial
 * confidence intervals (tests/test_gpu_nii.py, tools/ber_compare.py).
 *
 * Structure follows dvb_rcs2_turbo.py:116-281 / :464-537 with the three stated differences.
 * Build with -ffp-contract=off (oracle/Makefile).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NS 16

/* One SISO.  a0 / b0 [16]: in = alpha[0] / beta[N] to start from, out = alpha[N] / beta[0] reached. */
void nii_siso(const float *Lc_A, const float *Lc_B, const float *Lc_W, const float *Lc_Y,
              const float *La_A, const float *La_B, const int32_t *next_st, const int32_t *out_W,
              const int32_t *out_Y, const int32_t *prev_st, const int32_t *prev_inp, int N, float sf,
              float *a0, float *b0, float *Le_A, float *Le_B, float *scratch /* N*64 + 2*(N+1)*16 floats */)
{
    const float NEG = -1e9f;
    float *gamma = scratch, *alpha = gamma + (size_t)N * 64, *beta = alpha + (size_t)(N + 1) * 16;
    for (int k = 0; k < N; ++k) {
        const float in_A = Lc_A[k] + La_A[k], in_B = Lc_B[k] + La_B[k];
        for (int s = 0; s < NS; ++s)
            for (int inp = 0; inp < 4; ++inp) {
                const int bit_A = (inp >> 1) & 1, bit_B = inp & 1;
                const int bit_W = out_W[s * 4 + inp], bit_Y = out_Y[s * 4 + inp];
                float m = 0.0f;
                m += in_A * (bit_A == 0 ? 0.5f : -0.5f);
                m += in_B * (bit_B == 0 ? 0.5f : -0.5f);
                m += Lc_W[k] * (bit_W == 0 ? 0.5f : -0.5f);
                m += Lc_Y[k] * (bit_Y == 0 ? 0.5f : -0.5f);
                gamma[(size_t)k * 64 + s * 4 + inp] = m;
            }
    }
    memcpy(alpha, a0, sizeof(float) * NS);
    for (int k = 0; k < N; ++k) {
        for (int ns = 0; ns < NS; ++ns) {
            float mx = NEG;
            for (int idx = 0; idx < 4; ++idx) {
                const int ps = prev_st[ns * 4 + idx], inp = prev_inp[ns * 4 + idx];
                const float t = alpha[k * 16 + ps] + gamma[(size_t)k * 64 + ps * 4 + inp];
                if (t > mx) mx = t;
            }
            alpha[(k + 1) * 16 + ns] = mx;
        }
        const float norm = alpha[(k + 1) * 16];
        for (int s = 0; s < NS; ++s) alpha[(k + 1) * 16 + s] -= norm;
    }
    memcpy(beta + (size_t)N * 16, b0, sizeof(float) * NS);
    for (int k = N - 1; k >= 0; --k) {
        for (int s = 0; s < NS; ++s) {
            float mx = NEG;
            for (int inp = 0; inp < 4; ++inp) {
                const float t = beta[(k + 1) * 16 + next_st[s * 4 + inp]] + gamma[(size_t)k * 64 + s * 4 + inp];
                if (t > mx) mx = t;
            }
            beta[k * 16 + s] = mx;
        }
        const float norm = beta[k * 16];
        for (int s = 0; s < NS; ++s) beta[k * 16 + s] -= norm;
    }
    memcpy(a0, alpha + (size_t)N * 16, sizeof(float) * NS);
    memcpy(b0, beta, sizeof(float) * NS);
    for (int k = 0; k < N; ++k) {
        float app[4] = {NEG, NEG, NEG, NEG};
        for (int s = 0; s < NS; ++s)
            for (int inp = 0; inp < 4; ++inp) {
                float metric = alpha[k * 16 + s] + beta[(k + 1) * 16 + next_st[s * 4 + inp]];   /* re-associated */
                metric = metric + gamma[(size_t)k * 64 + s * 4 + inp];
                if (metric > app[inp]) app[inp] = metric;
            }
        const float pA0 = app[0] > app[1] ? app[0] : app[1], pA1 = app[2] > app[3] ? app[2] : app[3];
        const float pB0 = app[0] > app[2] ? app[0] : app[2], pB1 = app[1] > app[3] ? app[1] : app[3];
        float a = ((pA0 - pA1) - (Lc_A[k] + La_A[k])) * sf;
        float b = ((pB0 - pB1) - (Lc_B[k] + La_B[k])) * sf;
        if (a > 300.0f) a = 300.0f;
        if (a < -300.0f) a = -300.0f;
        if (b > 300.0f) b = 300.0f;
        if (b < -300.0f) b = -300.0f;
        Le_A[k] = a; Le_B[k] = b;
    }
}

/* Full decode of one frame.  Returns -1 when llr is shorter than the depuncturer consumes. */
int nii_decode(int N, int iterations, const int32_t *next_st, const int32_t *out_W, const int32_t *out_Y,
               const int32_t *prev_st, const int32_t *prev_inp, const int32_t *perm, const int32_t *inv_perm,
               const uint8_t *punct, int period, float sf_inner, float sf_last, const float *llr, int n_llr,
               int32_t *decoded)
{
    const size_t nf = (size_t)N;
    float *f = (float *)calloc(nf * 16 + nf * 64 + 2 * (nf + 1) * 16, sizeof(float));
    float *Lc_A = f, *Lc_B = f + nf, *Lc_W1 = f + 2 * nf, *Lc_Y1 = f + 3 * nf, *Lc_W2 = f + 4 * nf;
    float *Lc_Y2 = f + 5 * nf, *Lc_Ai = f + 6 * nf, *Lc_Bi = f + 7 * nf;
    float *La_A = f + 8 * nf, *La_B = f + 9 * nf, *Le1_A = f + 10 * nf, *Le1_B = f + 11 * nf;
    float *La2_A = f + 12 * nf, *La2_B = f + 13 * nf, *Le2_A = f + 14 * nf, *Le2_B = f + 15 * nf;
    float *scratch = f + 16 * nf;
    float st[4][NS];                    /* a0, b0 of SISO 1; a0, b0 of SISO 2: zeros in the first iteration */
    memset(st, 0, sizeof st);
    int idx = 0, rc = 0;
    for (int i = 0; i < N; ++i) {
        const int p = i % period;
        if (idx + 2 > n_llr) { rc = -1; goto done; }
        Lc_A[i] = llr[idx++]; Lc_B[i] = llr[idx++];
        if (punct[0 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_W1[i] = llr[idx++]; }
        if (punct[1 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_Y1[i] = llr[idx++]; }
        if (punct[2 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_W2[i] = llr[idx++]; }
        if (punct[3 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_Y2[i] = llr[idx++]; }
    }
    for (int i = 0; i < N; ++i) { Lc_Ai[i] = Lc_A[perm[i]]; Lc_Bi[i] = Lc_B[perm[i]]; }
    for (int it = 0; it < iterations; ++it) {
        const float sf = (it < iterations - 1) ? sf_inner : sf_last;
        nii_siso(Lc_A, Lc_B, Lc_W1, Lc_Y1, La_A, La_B, next_st, out_W, out_Y, prev_st, prev_inp, N, sf,
                 st[0], st[1], Le1_A, Le1_B, scratch);
        for (int i = 0; i < N; ++i) { La2_A[i] = Le1_A[perm[i]]; La2_B[i] = Le1_B[perm[i]]; }
        nii_siso(Lc_Ai, Lc_Bi, Lc_W2, Lc_Y2, La2_A, La2_B, next_st, out_W, out_Y, prev_st, prev_inp, N, sf,
                 st[2], st[3], Le2_A, Le2_B, scratch);
        for (int i = 0; i < N; ++i) { La_A[i] = Le2_A[inv_perm[i]]; La_B[i] = Le2_B[inv_perm[i]]; }
    }
    for (int i = 0; i < N; ++i) {
        const float LA = (Lc_A[i] + La_A[i]) + Le1_A[i], LB = (Lc_B[i] + La_B[i]) + Le1_B[i];
        decoded[2 * i] = LA < 0 ? 1 : 0;
        decoded[2 * i + 1] = LB < 0 ? 1 : 0;
    }
done:
    free(f);
    return rc;
}

int nii_decode_batch(int B, int N, int iterations, const int32_t *next_st, const int32_t *out_W,
                     const int32_t *out_Y, const int32_t *prev_st, const int32_t *prev_inp, const int32_t *perm,
                     const int32_t *inv_perm, const uint8_t *punct, int period, float sf_inner, float sf_last,
                     const float *llr, int n_llr, int32_t *decoded)
{
    int bad = 0;
    for (int b = 0; b < B; ++b)
        bad |= nii_decode(N, iterations, next_st, out_W, out_Y, prev_st, prev_inp, perm, inv_perm, punct, period,
                          sf_inner, sf_last, llr + (size_t)b * n_llr, n_llr, decoded + (size_t)b * 2 * N) != 0;
    return bad ? -1 : 0;
}
