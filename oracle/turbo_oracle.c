/*
 * oracle/turbo_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the reference's 16-state duo-binary circular
 * turbo codec (reference: dvb_rcs2_turbo.py).  It exists only to check the
 * CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
 * --impl reference leg).  Nothing under modulations_b200/ may import, link or
 * call it.
 *
 * Parity status: PINNED.  The reference's own tests assert nothing, so the pin
 * is "outputs of the reference itself, run in the authoring container":
 * oracle/make_golden.py imports the unmodified reference from /root/reference,
 * runs it on seeded inputs and commits the vectors under tests/golden/;
 * tests/test_oracle_golden.py checks this file against them bit for bit.
 *
 * Arithmetic follows numba's typing of the reference exactly
 * (dvb_rcs2_turbo.py:116-281):
 *   - branch metric: float64 left-to-right sum, rounded once to float32
 *   - recursions / normalisation / APP metric / L_post: float32
 *   - extrinsic: float64
 * Build with -ffp-contract=off and without -ffast-math (see oracle/Makefile).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NS 16

/* ---- trellis: dvb_rcs2_turbo.py:327-396 (_init_trellis) ----------------- */
void orc_build_trellis(int32_t *next_state, int32_t *out_W, int32_t *out_Y,
                       int32_t *prev_state, int32_t *prev_input, int32_t *G)
{
    memset(G, 0, 16 * sizeof(int32_t));
    for (int s = 0; s < NS; ++s) {
        int s0 = s & 1, s1 = (s >> 1) & 1, s2 = (s >> 2) & 1, s3 = (s >> 3) & 1;
        for (int inp = 0; inp < 4; ++inp) {
            int A = (inp >> 1) & 1, B = inp & 1;
            int dk = A ^ B ^ s2 ^ s3;          /* :351 */
            int w = dk ^ s0 ^ s1 ^ s3;         /* :355 */
            int y = dk ^ s1 ^ s2 ^ s3;         /* :359 */
            int ns = (s2 << 3) | (s1 << 2) | (s0 << 1) | dk;   /* :366 */
            next_state[s * 4 + inp] = ns;
            out_W[s * 4 + inp] = w;
            out_Y[s * 4 + inp] = y;
        }
    }
    G[0 * 4 + 2] = 1; G[0 * 4 + 3] = 1;        /* :380-383 */
    G[1 * 4 + 0] = 1; G[2 * 4 + 1] = 1; G[3 * 4 + 2] = 1;
    int counts[NS] = {0};
    for (int i = 0; i < NS * 4; ++i) { prev_state[i] = -1; prev_input[i] = -1; }
    for (int s = 0; s < NS; ++s)               /* :389-396 */
        for (int inp = 0; inp < 4; ++inp) {
            int ns = next_state[s * 4 + inp];
            int idx = counts[ns];
            if (idx < 4) {
                prev_state[ns * 4 + idx] = s;
                prev_input[ns * 4 + idx] = inp;
                counts[ns]++;
            }
        }
}

/* ---- interleaver: dvb_rcs2_turbo.py:311-324 (perm only; inv_perm is
 *      np.argsort(perm) on the host, see oracle.py) ----------------------- */
void orc_interleaver(int N, int P, int Q0, int Q1, int Q2, int Q3, int32_t *perm)
{
    for (int i = 0; i < N; ++i) {
        int r = i % 4, d = 0;
        if (r == 1) d = Q0; else if (r == 2) d = Q1; else if (r == 3) d = Q2;
        perm[i] = (int32_t)(((long long)P * (i + d + Q3 * (i / 4))) % N);
    }
}

/* ---- GF(2) helpers: dvb_rcs2_turbo.py:37-114 ---------------------------- */
void orc_mat_mul_gf2(const int32_t *A, const int32_t *B, int32_t *C)
{
    int32_t T[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            int v = 0;
            for (int k = 0; k < 4; ++k) v ^= (A[i * 4 + k] & B[k * 4 + j]);
            T[i * 4 + j] = v;
        }
    memcpy(C, T, sizeof T);
}

void orc_mat_pow_gf2(const int32_t *A, long long power, int32_t *R)
{
    int32_t res[16] = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1}, base[16];
    memcpy(base, A, sizeof base);
    while (power > 0) {
        if (power % 2 == 1) orc_mat_mul_gf2(res, base, res);
        orc_mat_mul_gf2(base, base, base);
        power /= 2;
    }
    memcpy(R, res, sizeof res);
}

int orc_solve_circular_state_gf2(const int32_t *G_pow_N, int Z_N)
{
    int32_t M[4][5];
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) M[i][j] = ((i == j) + G_pow_N[i * 4 + j]) % 2;
        M[i][4] = (Z_N >> i) & 1;
    }
    for (int i = 0; i < 4; ++i) {              /* :85-99 */
        if (M[i][i] == 0)
            for (int k = i + 1; k < 4; ++k)
                if (M[k][i] == 1) {
                    for (int j = 0; j < 5; ++j) { int t = M[i][j]; M[i][j] = M[k][j]; M[k][j] = t; }
                    break;
                }
        if (M[i][i] == 1)
            for (int k = i + 1; k < 4; ++k)
                if (M[k][i] == 1)
                    for (int j = 0; j < 5; ++j) M[k][j] ^= M[i][j];
    }
    int x[4] = {0, 0, 0, 0};
    for (int i = 3; i >= 0; --i) {             /* :103-107 */
        int sum = M[i][4];
        for (int j = i + 1; j < 4; ++j) sum ^= (M[i][j] & x[j]);
        x[i] = sum;
    }
    int state = 0;
    for (int i = 0; i < 4; ++i) if (x[i]) state |= (1 << i);
    return state;
}

/* ---- encoder: dvb_rcs2_turbo.py:404-462 --------------------------------- */
static void encode_component(int N, const int32_t *next_state, const int32_t *out_W,
                             const int32_t *out_Y, const int32_t *G,
                             const int32_t *A, const int32_t *B, int32_t *W, int32_t *Y,
                             int *start_state_out)
{
    int state = 0;
    for (int i = 0; i < N; ++i) state = next_state[state * 4 + ((A[i] << 1) | B[i])];
    int32_t GN[16];
    orc_mat_pow_gf2(G, N, GN);
    int start = orc_solve_circular_state_gf2(GN, state);
    if (start_state_out) *start_state_out = start;
    state = start;
    for (int i = 0; i < N; ++i) {
        int inp = (A[i] << 1) | B[i];
        W[i] = out_W[state * 4 + inp];
        Y[i] = out_Y[state * 4 + inp];
        state = next_state[state * 4 + inp];
    }
}

/* punct: uint8 [4][period] rows W1,Y1,W2,Y2.  coded must hold the emitted
 * length (2N + transmitted parities); returns that length.  circ[2] (nullable)
 * receives the two circular start states. */
int orc_encode(int N, const int32_t *next_state, const int32_t *out_W, const int32_t *out_Y,
               const int32_t *G, const int32_t *perm, const uint8_t *punct, int period,
               const int32_t *bits, int32_t *coded, int32_t *circ)
{
    int32_t *buf = (int32_t *)calloc((size_t)N * 8, sizeof(int32_t));
    int32_t *A = buf, *B = buf + N, *Ai = buf + 2 * N, *Bi = buf + 3 * N;
    int32_t *W1 = buf + 4 * N, *Y1 = buf + 5 * N, *W2 = buf + 6 * N, *Y2 = buf + 7 * N;
    for (int i = 0; i < N; ++i) { A[i] = bits[2 * i]; B[i] = bits[2 * i + 1]; }
    int c1, c2;
    encode_component(N, next_state, out_W, out_Y, G, A, B, W1, Y1, &c1);
    for (int i = 0; i < N; ++i) { Ai[i] = A[perm[i]]; Bi[i] = B[perm[i]]; }
    encode_component(N, next_state, out_W, out_Y, G, Ai, Bi, W2, Y2, &c2);
    if (circ) { circ[0] = c1; circ[1] = c2; }
    int n = 0;
    for (int i = 0; i < N; ++i) {
        int p = i % period;
        coded[n++] = A[i]; coded[n++] = B[i];
        if (punct[0 * period + p]) coded[n++] = W1[i];
        if (punct[1 * period + p]) coded[n++] = Y1[i];
        if (punct[2 * period + p]) coded[n++] = W2[i];
        if (punct[3 * period + p]) coded[n++] = Y2[i];
    }
    free(buf);
    return n;
}

/* ---- SISO: dvb_rcs2_turbo.py:116-281 (bcjr_max_log_map) ------------------ */
void orc_bcjr_max_log_map(const float *Lc_A, const float *Lc_B, const float *Lc_W,
                          const float *Lc_Y, const double *La_A, const double *La_B,
                          const int32_t *next_st, const int32_t *out_W, const int32_t *out_Y,
                          const int32_t *prev_st, const int32_t *prev_inp,
                          int N, double scaling_factor, double *Le_A, double *Le_B,
                          float *scratch /* (N*64 + 2*(N+1)*16) floats */)
{
    const double NEG_INF_VAL = -1e9;
    float *gamma = scratch;                    /* [N][16][4]  :129 */
    float *alpha = gamma + (size_t)N * 64;     /* [N+1][16]   :163 */
    float *beta = alpha + (size_t)(N + 1) * 16;/* [N+1][16]   :200 */

    for (int k = 0; k < N; ++k) {              /* :131-160 */
        double in_A = (double)Lc_A[k] + La_A[k];
        double in_B = (double)Lc_B[k] + La_B[k];
        float par_W = Lc_W[k], par_Y = Lc_Y[k];
        for (int s = 0; s < NS; ++s)
            for (int inp = 0; inp < 4; ++inp) {
                int bit_A = (inp >> 1) & 1, bit_B = inp & 1;
                int bit_W = out_W[s * 4 + inp], bit_Y = out_Y[s * 4 + inp];
                double m = 0.0;
                m += in_A * (bit_A == 0 ? 0.5 : -0.5);
                m += in_B * (bit_B == 0 ? 0.5 : -0.5);
                m += (double)par_W * (bit_W == 0 ? 0.5 : -0.5);
                m += (double)par_Y * (bit_Y == 0 ? 0.5 : -0.5);
                gamma[(size_t)k * 64 + s * 4 + inp] = (float)m;
            }
    }

    memset(alpha, 0, sizeof(float) * (N + 1) * 16);
    for (int pass = 0; pass < 2; ++pass) {     /* :167-197 */
        for (int k = 0; k < N; ++k) {
            for (int ns = 0; ns < NS; ++ns) {
                double max_val = NEG_INF_VAL;
                for (int idx = 0; idx < 4; ++idx) {
                    int ps = prev_st[ns * 4 + idx], inp = prev_inp[ns * 4 + idx];
                    float tmp = alpha[k * 16 + ps] + gamma[(size_t)k * 64 + ps * 4 + inp];
                    if ((double)tmp > max_val) max_val = (double)tmp;
                }
                alpha[(k + 1) * 16 + ns] = (float)max_val;
            }
            float norm = alpha[(k + 1) * 16];
            for (int s = 0; s < NS; ++s) alpha[(k + 1) * 16 + s] -= norm;
        }
        if (pass == 0)
            for (int s = 0; s < NS; ++s) alpha[s] = alpha[N * 16 + s];   /* :182-183 */
    }

    memset(beta, 0, sizeof(float) * (N + 1) * 16);
    for (int pass = 0; pass < 2; ++pass) {     /* :203-230 */
        for (int k = N - 1; k >= 0; --k) {
            for (int s = 0; s < NS; ++s) {
                double max_val = NEG_INF_VAL;
                for (int inp = 0; inp < 4; ++inp) {
                    int ns = next_st[s * 4 + inp];
                    float tmp = beta[(k + 1) * 16 + ns] + gamma[(size_t)k * 64 + s * 4 + inp];
                    if ((double)tmp > max_val) max_val = (double)tmp;
                }
                beta[k * 16 + s] = (float)max_val;
            }
            float norm = beta[k * 16];
            for (int s = 0; s < NS; ++s) beta[k * 16 + s] -= norm;
        }
        if (pass == 0)
            for (int s = 0; s < NS; ++s) beta[N * 16 + s] = beta[s];     /* :216-217 */
    }

    for (int k = 0; k < N; ++k) {              /* :239-279 */
        float app[4] = {(float)NEG_INF_VAL, (float)NEG_INF_VAL, (float)NEG_INF_VAL, (float)NEG_INF_VAL};
        for (int s = 0; s < NS; ++s)
            for (int inp = 0; inp < 4; ++inp) {
                int ns = next_st[s * 4 + inp];
                float metric = alpha[k * 16 + s] + gamma[(size_t)k * 64 + s * 4 + inp];
                metric = metric + beta[(k + 1) * 16 + ns];
                if (metric > app[inp]) app[inp] = metric;
            }
        float pA0 = app[0] > app[1] ? app[0] : app[1];
        float pA1 = app[2] > app[3] ? app[2] : app[3];
        float pB0 = app[0] > app[2] ? app[0] : app[2];
        float pB1 = app[1] > app[3] ? app[1] : app[3];
        float LpA = pA0 - pA1, LpB = pB0 - pB1;
        double a = (double)LpA - ((double)Lc_A[k] + La_A[k]);
        double b = (double)LpB - ((double)Lc_B[k] + La_B[k]);
        a *= scaling_factor; b *= scaling_factor;
        if (a > 300.0) a = 300.0;
        if (a < -300.0) a = -300.0;
        if (b > 300.0) b = 300.0;
        if (b < -300.0) b = -300.0;
        Le_A[k] = a; Le_B[k] = b;
    }
}

/* ---- full decoder: dvb_rcs2_turbo.py:464-537 (decode) --------------------
 * llr: the reference casts its argument to float32 first (:466); the caller
 * passes float32.  n_coded_in = number of LLRs available (the reference would
 * raise IndexError if too few; we return -1).  trace (nullable) receives
 * Le1_A,Le1_B,Le2_A,Le2_B of every iteration: [iterations][4][N] doubles.
 * Lfinal (nullable): [2][N] doubles. */
int orc_decode(int N, int iterations, const int32_t *next_st, const int32_t *out_W,
               const int32_t *out_Y, const int32_t *prev_st, const int32_t *prev_inp,
               const int32_t *perm, const int32_t *inv_perm, const uint8_t *punct, int period,
               const float *llr, int n_llr, int32_t *decoded, double *trace, double *Lfinal)
{
    size_t nf = (size_t)N;
    float *f = (float *)calloc(nf * 8 + nf * 64 + 2 * (nf + 1) * 16, sizeof(float));
    double *d = (double *)calloc(nf * 8, sizeof(double));
    float *Lc_A = f, *Lc_B = f + nf, *Lc_W1 = f + 2 * nf, *Lc_Y1 = f + 3 * nf;
    float *Lc_W2 = f + 4 * nf, *Lc_Y2 = f + 5 * nf, *Lc_Ai = f + 6 * nf, *Lc_Bi = f + 7 * nf;
    float *scratch = f + 8 * nf;
    double *La_A = d, *La_B = d + nf, *Le1_A = d + 2 * nf, *Le1_B = d + 3 * nf;
    double *La2_A = d + 4 * nf, *La2_B = d + 5 * nf, *Le2_A = d + 6 * nf, *Le2_B = d + 7 * nf;
    int idx = 0, rc = 0;
    for (int i = 0; i < N; ++i) {              /* :476-487 */
        int p = i % period;
        if (idx + 2 > n_llr) { rc = -1; goto done; }
        Lc_A[i] = llr[idx++]; Lc_B[i] = llr[idx++];
        if (punct[0 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_W1[i] = llr[idx++]; }
        if (punct[1 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_Y1[i] = llr[idx++]; }
        if (punct[2 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_W2[i] = llr[idx++]; }
        if (punct[3 * period + p]) { if (idx >= n_llr) { rc = -1; goto done; } Lc_Y2[i] = llr[idx++]; }
    }
    for (int i = 0; i < N; ++i) { Lc_Ai[i] = Lc_A[perm[i]]; Lc_Bi[i] = Lc_B[perm[i]]; }
    for (int it = 0; it < iterations; ++it) {  /* :493-524 */
        double sf = (it < iterations - 1) ? 0.7 : 1.0;
        orc_bcjr_max_log_map(Lc_A, Lc_B, Lc_W1, Lc_Y1, La_A, La_B, next_st, out_W, out_Y,
                             prev_st, prev_inp, N, sf, Le1_A, Le1_B, scratch);
        for (int i = 0; i < N; ++i) { La2_A[i] = Le1_A[perm[i]]; La2_B[i] = Le1_B[perm[i]]; }
        orc_bcjr_max_log_map(Lc_Ai, Lc_Bi, Lc_W2, Lc_Y2, La2_A, La2_B, next_st, out_W, out_Y,
                             prev_st, prev_inp, N, sf, Le2_A, Le2_B, scratch);
        for (int i = 0; i < N; ++i) { La_A[i] = Le2_A[inv_perm[i]]; La_B[i] = Le2_B[inv_perm[i]]; }
        if (trace) {
            double *t = trace + (size_t)it * 4 * nf;
            memcpy(t, Le1_A, nf * 8); memcpy(t + nf, Le1_B, nf * 8);
            memcpy(t + 2 * nf, Le2_A, nf * 8); memcpy(t + 3 * nf, Le2_B, nf * 8);
        }
    }
    for (int i = 0; i < N; ++i) {              /* :529-535 */
        double LA = ((double)Lc_A[i] + La_A[i]) + Le1_A[i];
        double LB = ((double)Lc_B[i] + La_B[i]) + Le1_B[i];
        decoded[2 * i] = LA < 0 ? 1 : 0;
        decoded[2 * i + 1] = LB < 0 ? 1 : 0;
        if (Lfinal) { Lfinal[i] = LA; Lfinal[nf + i] = LB; }
    }
done:
    free(f); free(d);
    return rc;
}

/* Batched wrappers: serial loops over frames.  oracle.py runs them from a
 * thread pool over frame chunks (ctypes releases the GIL), which is how the CPU
 * baseline uses every host core.  Used only for the CPU baseline timing and
 * for bulk parity checks. */
int orc_decode_batch(int B, int N, int iterations, const int32_t *next_st, const int32_t *out_W,
                     const int32_t *out_Y, const int32_t *prev_st, const int32_t *prev_inp,
                     const int32_t *perm, const int32_t *inv_perm, const uint8_t *punct,
                     int period, const float *llr, int n_coded, int32_t *decoded)
{
    int bad = 0;
    for (int b = 0; b < B; ++b)
        bad |= orc_decode(N, iterations, next_st, out_W, out_Y, prev_st, prev_inp, perm, inv_perm,
                          punct, period, llr + (size_t)b * n_coded, n_coded,
                          decoded + (size_t)b * 2 * N, NULL, NULL) != 0;
    return bad ? -1 : 0;
}

int orc_encode_batch(int B, int N, const int32_t *next_state, const int32_t *out_W,
                     const int32_t *out_Y, const int32_t *G, const int32_t *perm,
                     const uint8_t *punct, int period, const int32_t *bits, int32_t *coded,
                     int n_emit)
{
    int bad = 0;
    for (int b = 0; b < B; ++b)
        bad |= orc_encode(N, next_state, out_W, out_Y, G, perm, punct, period,
                          bits + (size_t)b * 2 * N, coded + (size_t)b * n_emit, NULL) != n_emit;
    return bad ? -1 : 0;
}
