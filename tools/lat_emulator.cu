// tools/lat_emulator.cu — TEST INFRASTRUCTURE.  Replays on the CPU what the low-latency decoder
// (modulations_b200/csrc/decode_lat.cu) does for one SISO of one frame, lane by lane, with the kernel's own arithmetic
// core (tpf_core.cuh compiled for the host), and checks the extrinsics against oracle/turbo_oracle.c bit for bit:
//   * one state metric per lane: lane s of the alpha run combines v[t], v[8+t] (s = 2t + b), lane s of the beta run
//     z[2t], z[2t+1] (s = 8 hi + t), both normalised by a redundantly computed n[0] — the per-lane operand and record
//     selection of LaneRec::init / LaneRec::step;
//   * lap 1 in four speculative segments (zeros W steps before the segment), verified and, where the guess had not
//     re-joined, recomputed by one carrier; lap 2 from lap 1's end state until it re-joins (segmented_recursion).
// The claim under test: the stored metrics equal the reference's second lap for EVERY warm-up length W, including W
// that make nearly every guess wrong.  Also prints how many steps the carrier had to recompute.
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
constexpr int kSegs = 4;

static double urand() { return (double)rand() / RAND_MAX; }

// the 16 lanes of one recursion (decode_lat.cu LaneRec<BETA>)
struct Lanes {
    bool beta;
    int N;
    const float *rec;
    int la[16], lb[16], l0b, c2[16];
    bool swp[16];
    void init(bool beta_, int N_, const float *rec_)
    {
        beta = beta_; N = N_; rec = rec_; l0b = beta ? 1 : 8;
        for (int s = 0; s < 16; ++s) {
            const int t = beta ? (s & 7) : (s >> 1);
            la[s] = beta ? 2 * t : t; lb[s] = beta ? 2 * t + 1 : 8 + t;
            swp[s] = (((t >> 2) ^ (beta ? (s >> 3) : s)) & 1) != 0;
            c2[s] = 2 * cls(t);
        }
    }
    int kk(int i) const { return beta ? N - 1 - i : i; }
    void step(float (&v)[16], int i) const
    {
        const float *r = rec + 8 * kk(i);
        float n[16];
        for (int s = 0; s < 16; ++s) {
            const float a = v[la[s]], bq = v[lb[s]], v0 = v[0], v8 = v[l0b];           // the four shuffles
            const float pcx = r[c2[s]], pcy = r[c2[s] + 1], p0x = r[0], p0y = r[1];
            const float X = swp[s] ? pcy : pcx, Yv = swp[s] ? pcx : pcy;
            n[s] = f_sub(f_max(f_add(a, X), f_add(bq, Yv)), f_max(f_add(v0, p0x), f_add(v8, p0y)));
        }
        memcpy(v, n, sizeof n);
    }
};

static int seg_len(int N, int W) { return (N + (kSegs - 1) * W + kSegs - 1) / kSegs; }
static int seg_start(int j, int N, int W)
{
    if (j <= 0) return 0;
    if (j >= kSegs) return N;
    const int L = seg_len(N, W);
    if (L <= W) return N;
    const int p = L + (j - 1) * (L - W);
    return p < N ? p : N;
}

// cells[16 * kk(i) + s] = state before step i; returns the number of steps the carrier recomputed in lap 1
static int segmented(const Lanes &R, int W, std::vector<float> &cells, int &lap2_steps)
{
    const int N = R.N;
    float es[kSegs][16];
    for (int j = 0; j < kSegs; ++j) {                                   // phase A
        const int p0 = seg_start(j, N, W), p1 = seg_start(j + 1, N, W);
        if (p0 >= p1) continue;
        const int i0 = j == 0 ? 0 : (p0 > W ? p0 - W : 0);
        float v[16] = {0};
        for (int i = i0; i < p1; ++i) {
            if (i >= p0) memcpy(&cells[16 * R.kk(i)], v, sizeof v);
            R.step(v, i);
        }
        memcpy(es[j], v, sizeof v);
    }
    auto carry = [&](float (&v)[16], int i0, int i1, int &steps) -> bool {   // phase B building block
        for (int i = i0; i < i1; ++i) {
            if (!memcmp(&cells[16 * R.kk(i)], v, sizeof v)) return true;
            memcpy(&cells[16 * R.kk(i)], v, sizeof v);
            R.step(v, i);
            ++steps;
        }
        return false;
    };
    float v[16];
    memcpy(v, es[0], sizeof v);
    int redo = 0;
    for (int seg = 1; seg < kSegs; ++seg) {
        const int q0 = seg_start(seg, N, W), q1 = seg_start(seg + 1, N, W);
        if (q0 >= q1) continue;
        if (carry(v, q0, q1, redo)) memcpy(v, es[seg], sizeof v);
    }
    lap2_steps = 0;
    carry(v, 0, N, lap2_steps);                                        // lap 2
    return redo;
}

static int run(int N, unsigned seed, double scale, int W)
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
    // P0: records
    std::vector<float> rec((size_t)N * 8);
    std::vector<double> Y((size_t)N * 2);
    for (int k = 0; k < N; ++k) {
        Y[2 * k] = d_add((double)LcA[k], LaA[k]); Y[2 * k + 1] = d_add((double)LcB[k], LaB[k]);
        float g[8]; make_record(Y[2 * k], Y[2 * k + 1], LcW[k], LcY[k], g);
        memcpy(&rec[k * 8], g, sizeof g);
    }
    // P1: alpha into Al[16 k], beta into Be[16 (k + 1)]
    std::vector<float> Al((size_t)N * 16), Be((size_t)(N + 1) * 16);
    Lanes A, B;
    A.init(false, N, rec.data()); B.init(true, N, rec.data());
    std::vector<float> cb((size_t)N * 16);
    int l2a = 0, l2b = 0;
    const int ra = segmented(A, W, Al, l2a), rb = segmented(B, W, cb, l2b);
    memcpy(&Be[16], cb.data(), sizeof(float) * N * 16);
    // P2
    for (int k = 0; k < N; ++k) {
        float x[16], zs[16], g[8], uv[4];
        memcpy(x, &Al[16 * k], sizeof x); memcpy(zs, &Be[16 * (k + 1)], sizeof zs); memcpy(g, &rec[8 * k], sizeof g);
        ext_step(x, zs, g, uv);
        make_extrinsic(uv, Y[2 * k], Y[2 * k + 1], sf, LeA[k], LeB[k]);
    }
    int bad = 0;
    for (int k = 0; k < N; ++k)
        if (memcmp(&LeA[k], &refA[k], 8) || memcmp(&LeB[k], &refB[k], 8)) {
            if (bad < 3) printf("  N=%d k=%d: got (%.17g, %.17g) want (%.17g, %.17g)\n", N, k, LeA[k], LeB[k], refA[k], refB[k]);
            ++bad;
        }
    printf("N=%d seed=%u scale=%g W=%d: %s (%d mismatches; carrier recomputed %d + %d lap-1 steps, lap 2 ran %d + %d of %d)\n",
           N, seed, scale, W, bad ? "FAIL" : " ok ", bad, ra, rb, l2a, l2b, N);
    return bad;
}

int main()
{
    int bad = 0;
    const int Ns[] = {48, 64, 68, 212, 220, 424, 752, 16, 12};
    const int Ws[] = {1, 7, 33, 53, 64, 96, 300, 2000};
    for (int N : Ns)
        for (int W : Ws)
            for (unsigned seed = 1; seed <= 2; ++seed) bad += run(N, seed + 10 * W, seed == 2 ? 90.0 : 8.0, W);
    return bad ? 1 : 0;
}
