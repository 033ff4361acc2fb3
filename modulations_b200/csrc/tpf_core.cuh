// tpf_core.cuh — arithmetic core of the thread-per-frame ("TPF") max-log-MAP decoder.
//
// One thread holds all 16 state metrics of one frame and one direction in registers:
// no shuffles, no shared-memory exchange, every step is straight-line FADD/FMNMX code.
// These functions are __host__ __device__ so that tools/tpf_emulator.cu can replay the
// kernel's schedule on the CPU against the oracle before any GPU time is spent.
//
// Reference arithmetic (dvb_rcs2_turbo.py:116-281): branch metrics are float64 sums
// rounded once to float32 (:131-160); recursions, normalisation by state 0 and the APP
// metrics are float32 (:162-248); the extrinsic is float64 (:250-279).
//
// Trellis facts used (dvb_rcs2_turbo.py:327-396; state s = (s3 s2 s1 s0) = (x, t2 t1 t0)):
//   * ns = 2*(s & 7) + dk: the predecessors of (t, d) are (0, t) and (1, t) — a butterfly.
//   * inputs {00, 11} from s share the next state (dk = s2 ^ s3) and the parity class
//     c(s) = (s0^s1^s2, s1); inputs {01, 10} go to the other next state with class ~c.
//     Rounding is monotone, so max(fl(a+g1), fl(a+g2)) == fl(a + max(g1, g2)): per step only
//     the 8 merged metrics GP[c] = max(P[c], -P[~c]), GM[c] = max(M[~c], -M[c]) are needed
//     (P / M = rounded sums that start from a+b / a-b); the smaller member of a pair is
//     -GP[~c] / -GM[~c], and which member is input 00 (01) is the sign of a+b (a-b).
//   * the backward recursion under bit-reversed state labels has the same register wiring
//     as the forward one; only the GP/GM roles of the classes 1 and 2 swap.  A warp can
//     therefore run alpha lanes and beta lanes through ONE instruction stream.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TPF_HD __host__ __device__ __forceinline__
#else
#define TPF_HD inline
#endif

namespace b200dvb {
namespace tpf {

TPF_HD float f_add(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
TPF_HD float f_sub(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
TPF_HD float f_max(float a, float b)
{
#ifdef __CUDA_ARCH__
    return fmaxf(a, b);
#else
    return a > b ? a : b;
#endif
}
TPF_HD double d_add(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
TPF_HD double d_sub(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
TPF_HD double d_mul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
TPF_HD float d_to_f(double a)
{
#ifdef __CUDA_ARCH__
    return __double2float_rn(a);
#else
    return (float)a;
#endif
}

// parity class of the butterfly t = s & 7: c = 2*(s0^s1^s2) + s1
TPF_HD constexpr int cls(int t) { return 2 * ((t ^ (t >> 1) ^ (t >> 2)) & 1) + ((t >> 1) & 1); }
// 4-bit reversal: natural state index <-> label used by the beta lanes during the passes
TPF_HD constexpr int rho4(int r) { return ((r & 1) << 3) | ((r & 2) << 1) | ((r & 4) >> 1) | ((r & 8) >> 3); }

// Branch-metric record of one trellis step: g[2c] = GP[c], g[2c+1] = GM[c]
// (dvb_rcs2_turbo.py:131-160: float64 left-to-right sums, rounded once to float32).
TPF_HD void make_record(double YA, double YB, float pW, float pY, float (&g)[8])
{
    const double a = d_mul(YA, 0.5), b = d_mul(YB, 0.5);
    const double w = d_mul((double)pW, 0.5), y = d_mul((double)pY, 0.5);
    const double s = d_add(a, b), d = d_sub(a, b);
    float P[4], Mv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double sw = (c & 2) ? -w : w, sy = (c & 1) ? -y : y;
        P[c] = d_to_f(d_add(d_add(s, sw), sy));
        Mv[c] = d_to_f(d_add(d_add(d, sw), sy));
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        g[2 * c] = f_max(P[c], -P[3 - c]);
        g[2 * c + 1] = f_max(Mv[3 - c], -Mv[c]);
    }
}

// One step of the two "in" passes (dvb_rcs2_turbo.py:167-179 forward, :203-213 backward)
// in the shared wiring out[2t+o] = max_i(in[8i+t] + role(i^o)):
//   alpha lanes (isb = false): v = alpha[k] in natural labels  -> alpha[k+1]
//   beta lanes  (isb = true):  v = beta[k+1] in rho4 labels     -> beta[k]
// followed by the normalisation by state 0 (:178-179, :212-213; rho4(0) == 0).
TPF_HD void pass_step(float (&v)[16], const float (&g)[8], bool isb)
{
    float G[8];
    G[0] = g[0]; G[1] = g[1]; G[6] = g[6]; G[7] = g[7];
    G[2] = isb ? g[3] : g[2]; G[3] = isb ? g[2] : g[3];
    G[4] = isb ? g[5] : g[4]; G[5] = isb ? g[4] : g[5];
    float n[16];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int c = cls(t);
        const float A = (t & 4) ? G[2 * c + 1] : G[2 * c];
        const float B = (t & 4) ? G[2 * c] : G[2 * c + 1];
        n[2 * t] = f_max(f_add(v[t], A), f_add(v[8 + t], B));
        n[2 * t + 1] = f_max(f_add(v[t], B), f_add(v[8 + t], A));
    }
    const float z = n[0];
    v[0] = 0.f;                                       // fl(z - z): metrics are finite (the reference clips its inputs)
#pragma unroll
    for (int s = 1; s < 16; ++s) v[s] = f_sub(n[s], z);
}

// Backward step in natural labels: z = beta[k+1] -> beta[k] (dvb_rcs2_turbo.py:203-213).
TPF_HD void bwd_step(float (&z)[16], const float (&g)[8])
{
    float n[16];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int c = cls(t);
        const float A = (t & 4) ? g[2 * c + 1] : g[2 * c];
        const float B = (t & 4) ? g[2 * c] : g[2 * c + 1];
        n[t] = f_max(f_add(z[2 * t], A), f_add(z[2 * t + 1], B));
        n[8 + t] = f_max(f_add(z[2 * t], B), f_add(z[2 * t + 1], A));
    }
    const float q = n[0];
    z[0] = 0.f;
#pragma unroll
    for (int s = 1; s < 16; ++s) z[s] = f_sub(n[s], q);
}

// Extrinsic maxima for step k (dvb_rcs2_turbo.py:239-248) fused with the forward step:
// x = alpha[k] -> alpha[k+1]; zs = beta[k+1]; both in natural labels.
// uv = (U0, U3, V1, V2): max over states of the larger / smaller member of the {00,11}
// pair and of the {01,10} pair, each summed as fl(fl(alpha + gamma) + beta) (:252-253).
TPF_HD void ext_step(float (&x)[16], const float (&zs)[16], const float (&g)[8], float (&uv)[4])
{
    float n[16];
    float U0 = 0.f, U3 = 0.f, V1 = 0.f, V2 = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int c = cls(t);
        const float gp = g[2 * c], gm = g[2 * c + 1];
        const float hp = g[2 * (3 - c)], hm = g[2 * (3 - c) + 1];
        const int t2 = (t >> 2) & 1;
        // state (0, t): P-branch -> 2t + t2, M-branch -> 2t + 1 - t2;  state (1, t): the other way
        const int nP0 = 2 * t + t2, nM0 = 2 * t + 1 - t2;
        const float p0 = f_add(x[t], gp), m0 = f_add(x[t], gm);
        const float p1 = f_add(x[8 + t], gp), m1 = f_add(x[8 + t], gm);
        const float ub0 = f_add(p0, zs[nP0]), ub1 = f_add(p1, zs[nM0]);
        const float vb0 = f_add(m0, zs[nM0]), vb1 = f_add(m1, zs[nP0]);
        const float us0 = f_add(f_sub(x[t], hp), zs[nP0]), us1 = f_add(f_sub(x[8 + t], hp), zs[nM0]);
        const float vs0 = f_add(f_sub(x[t], hm), zs[nM0]), vs1 = f_add(f_sub(x[8 + t], hm), zs[nP0]);
        if (t == 0) {
            U0 = f_max(ub0, ub1); U3 = f_max(us0, us1); V1 = f_max(vb0, vb1); V2 = f_max(vs0, vs1);
        } else {
            U0 = f_max(U0, f_max(ub0, ub1)); U3 = f_max(U3, f_max(us0, us1));
            V1 = f_max(V1, f_max(vb0, vb1)); V2 = f_max(V2, f_max(vs0, vs1));
        }
        n[nP0] = f_max(p0, m1);
        n[nM0] = f_max(m0, p1);
    }
    uv[0] = U0; uv[1] = U3; uv[2] = V1; uv[3] = V2;
    const float q = n[0];
    x[0] = 0.f;
#pragma unroll
    for (int s = 1; s < 16; ++s) x[s] = f_sub(n[s], q);
}

// Extrinsic epilogue for one step (dvb_rcs2_turbo.py:250-279).
TPF_HD void make_extrinsic(const float (&uv)[4], double YA, double YB, double sf, double &ea, double &eb)
{
    // sign of fl(YA/2 + YB/2) == sign of fl(YA + YB): halving commutes with rounding
    const bool sP = d_add(YA, YB) < 0.0, sM = d_sub(YA, YB) < 0.0;
    const float app0 = sP ? uv[1] : uv[0], app3 = sP ? uv[0] : uv[1];
    const float app1 = sM ? uv[3] : uv[2], app2 = sM ? uv[2] : uv[3];
    const float LA = f_sub(f_max(app0, app1), f_max(app2, app3));
    const float LB = f_sub(f_max(app0, app2), f_max(app1, app3));
    ea = d_mul(d_sub((double)LA, YA), sf);
    eb = d_mul(d_sub((double)LB, YB), sf);
    ea = ea > 300.0 ? 300.0 : ea; ea = ea < -300.0 ? -300.0 : ea;
    eb = eb > 300.0 ? 300.0 : eb; eb = eb < -300.0 ? -300.0 : eb;
}

}  // namespace tpf
}  // namespace b200dvb
