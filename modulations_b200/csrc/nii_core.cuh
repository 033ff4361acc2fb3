// nii_core.cuh — arithmetic of the NON-PARITY decoder mode "nii" (BASELINE north_star: "windows ...
// use next-iteration circular-state initialisation").  NOT the reference's arithmetic: a labelled variant whose
// results are judged on BER/FER, and whose kernel is checked bit for bit against its own plain-C model
// (oracle/nii_model.c).  Differences to the parity mode (tpf_core.cuh, dvb_rcs2_turbo.py:116-281):
//   * single pass per SISO: the circular boundary is not resolved by a first full sweep (:167-183, :203-217);
//     alpha[0] / beta[N] start from the metrics the SAME constituent decoder ended with one iteration earlier
//     (zeros in the first iteration);
//   * float32 throughout: a-priori / extrinsic values, branch-metric sums and the extrinsic epilogue (the
//     reference mixes float64 sums rounded to float32 with a float64 epilogue);
//   * the a-posteriori maxima are re-associated: max_s((alpha_s + beta_ns) + gamma) instead of
//     max_s((alpha_s + gamma) + beta_ns) (:246-247) — the branch metric is added AFTER the per-class maximum, which
//     rounding monotonicity makes exact with respect to the re-associated definition.
// Shared with tpf_core.cuh: the merged-branch record, the butterfly wiring, the bit-reversed beta labels, the
// normalisation by state 0 every step, the extrinsic scaling and the +-300 clip.
#pragma once
#include "tpf_core.cuh"

namespace b200dvb {
namespace nii {

using tpf::cls;
using tpf::f_add;
using tpf::f_max;
using tpf::f_sub;

TPF_HD float f_mul(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}

// Branch-metric record of one step in float32: g[2c] = GP[c], g[2c+1] = GM[c] (same roles as tpf::make_record).
// gamma(s, u) = ((+-YA/2 +- YB/2) +- W/2) +- Y/2, summed left to right in float32.
TPF_HD void make_record(float YA, float YB, float pW, float pY, float (&g)[8])
{
    const float a = f_mul(YA, 0.5f), b = f_mul(YB, 0.5f), w = f_mul(pW, 0.5f), y = f_mul(pY, 0.5f);
    const float s = f_add(a, b), d = f_sub(a, b);
    float P[4], Mv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float sw = (c & 2) ? -w : w, sy = (c & 1) ? -y : y;
        P[c] = f_add(f_add(s, sw), sy);
        Mv[c] = f_add(f_add(d, sw), sy);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        g[2 * c] = f_max(P[c], -P[3 - c]);
        g[2 * c + 1] = f_max(Mv[3 - c], -Mv[c]);
    }
}

// A-posteriori maxima of step k from x = alpha[k], zs = beta[k+1] (natural labels) and the step's record:
// uv = (U0, U3, V1, V2) = max over states of the larger / smaller member of the {00,11} pair and of the {01,10}
// pair.  48 FADD + 36 FMNMX against 128 + 60 for the reference's association.
TPF_HD void app_maxima(const float (&x)[16], const float (&zs)[16], const float (&g)[8], float (&uv)[4])
{
    float TP[4], TM[4];
    bool seen[4] = {false, false, false, false};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int c = cls(t);
        const int t2 = (t >> 2) & 1;
        const int nP0 = 2 * t + t2, nM0 = 2 * t + 1 - t2;       // state (0,t): P-branch -> nP0; state (1,t): P-branch -> nM0
        const float p = f_max(f_add(x[t], zs[nP0]), f_add(x[8 + t], zs[nM0]));
        const float m = f_max(f_add(x[t], zs[nM0]), f_add(x[8 + t], zs[nP0]));
        if (!seen[c]) { TP[c] = p; TM[c] = m; seen[c] = true; }
        else          { TP[c] = f_max(TP[c], p); TM[c] = f_max(TM[c], m); }
    }
    float U0 = f_add(TP[0], g[0]), U3 = f_sub(TP[0], g[6]), V1 = f_add(TM[0], g[1]), V2 = f_sub(TM[0], g[7]);
#pragma unroll
    for (int c = 1; c < 4; ++c) {
        U0 = f_max(U0, f_add(TP[c], g[2 * c]));
        U3 = f_max(U3, f_sub(TP[c], g[2 * (3 - c)]));
        V1 = f_max(V1, f_add(TM[c], g[2 * c + 1]));
        V2 = f_max(V2, f_sub(TM[c], g[2 * (3 - c) + 1]));
    }
    uv[0] = U0; uv[1] = U3; uv[2] = V1; uv[3] = V2;
}

// Extrinsic epilogue of one step, float32: Le = clip((L_post - (Lc + La)) * sf, +-300).
TPF_HD void make_extrinsic(const float (&uv)[4], float YA, float YB, float sf, float &ea, float &eb)
{
    const bool sP = f_add(YA, YB) < 0.f, sM = f_sub(YA, YB) < 0.f;
    const float app0 = sP ? uv[1] : uv[0], app3 = sP ? uv[0] : uv[1];
    const float app1 = sM ? uv[3] : uv[2], app2 = sM ? uv[2] : uv[3];
    const float LA = f_sub(f_max(app0, app1), f_max(app2, app3));
    const float LB = f_sub(f_max(app0, app2), f_max(app1, app3));
    ea = f_mul(f_sub(LA, YA), sf);
    eb = f_mul(f_sub(LB, YB), sf);
    ea = ea > 300.f ? 300.f : ea; ea = ea < -300.f ? -300.f : ea;
    eb = eb > 300.f ? 300.f : eb; eb = eb < -300.f ? -300.f : eb;
}

}  // namespace nii
}  // namespace b200dvb
