// nii16_core.cuh — arithmetic of the NON-PARITY decoder mode "nii16" ("fp16x2/DPX (VIADDMNMX)
// fixed-point ACS"): the single-pass next-iteration-initialisation decoder of nii_core.cuh in 16-bit fixed point,
// TWO frames per 32-bit register (s16x2), add-compare-select as ONE DPX instruction per branch pair
// (`__viaddmax_s16x2(a, b, c) = max(a + b, c)`, SASS VIADDMNMX.S16x2).  Not the reference's arithmetic: a labelled
// variant judged on BER/FER, whose kernel is checked bit for bit against its own integer model
// (oracle/nii16_model.c).  Integer arithmetic is associative, so the model can be naive and the kernel merged.
//
// Fixed-point format (all quantities in units of 1/4 LLR):
//   channel LLR   q = clamp(rint(4 llr), +-127)                      (+-31.75; the reference clips at +-50 upstream)
//   extrinsic     clamp(.., +-255)                                   (+-63.75)
//   Y = Lc + La   |Y| <= 382
//   branch metric Gamma = +-YA +- YB +- W +- Yp  (NOT halved: every path metric is 2x its float counterpart)
//                 |Gamma| <= 1018; state metrics normalised by state 0 every step lie within 4 steps x 2 |Gamma| =
//                 +-8144, alpha + beta + Gamma within +-17306 < 2^15: no 16-bit overflow anywhere.
//   epilogue      Le = clamp((((LA2 - 2 YA) * sf_q) + 64) >> 7, +-255) with LA2 the doubled a-posteriori LLR and
//                 sf_q = 45 (0.703) / 64 (1.0) the extrinsic scaling in Q6; >> is an arithmetic shift.
// Register convention: the kernel moves values through `float` registers and float2 / float4 containers (bit casts are
// free); every function here takes and returns the raw 32 bits.  Low half = frame f, high half = frame f + 16 of a
// 32-frame tile.
#pragma once
#include <stdint.h>

#include "tpf_core.cuh"

namespace b200dvb {
namespace nii16 {

using tpf::cls;

typedef uint32_t p16;           // two int16 lanes

constexpr int kChanMax = 127, kExtMax = 255, kSfShift = 7, kSfInner = 45, kSfLast = 64;

TPF_HD int lo16(p16 a) { return (int)(int16_t)(a & 0xffffu); }
TPF_HD int hi16(p16 a) { return (int)(int16_t)(a >> 16); }
TPF_HD p16 pack16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }

TPF_HD p16 add2(p16 a, p16 b)
{
#ifdef __CUDA_ARCH__
    p16 d;
    asm("add.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
#else
    return pack16(lo16(a) + lo16(b), hi16(a) + hi16(b));
#endif
}
TPF_HD p16 max2(p16 a, p16 b)
{
#ifdef __CUDA_ARCH__
    p16 d;
    asm("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
#else
    return pack16(lo16(a) > lo16(b) ? lo16(a) : lo16(b), hi16(a) > hi16(b) ? hi16(a) : hi16(b));
#endif
}
// max(a + b, c): the add-compare-select of one branch pair, one DPX instruction
TPF_HD p16 addmax2(p16 a, p16 b, p16 c)
{
#ifdef __CUDA_ARCH__
    return __viaddmax_s16x2(a, b, c);
#else
    return max2(add2(a, b), c);
#endif
}
TPF_HD p16 neg2(p16 a) { return add2(~a, 0x00010001u); }

// ---- float-register views (the kernel's registers are `float`) ------------------------------------------------------
#ifdef __CUDA_ARCH__
TPF_HD p16 bits(float x) { return __float_as_uint(x); }
TPF_HD float fl(p16 x) { return __uint_as_float(x); }
#else
inline p16 bits(float x) { p16 u; __builtin_memcpy(&u, &x, 4); return u; }
inline float fl(p16 x) { float f; __builtin_memcpy(&f, &x, 4); return f; }
#endif

// channel LLR -> fixed point (round to nearest even like rintf / cvt.rni, then clamp)
TPF_HD int quant(float llr)
{
#ifdef __CUDA_ARCH__
    int q = __float2int_rn(llr * 4.0f);
#else
    int q = (int)__builtin_rintf(llr * 4.0f);
#endif
    return q > kChanMax ? kChanMax : (q < -kChanMax ? -kChanMax : q);
}

// Merged branch-metric record (same roles as tpf::make_record / nii::make_record): g[2c] = GP[c], g[2c+1] = GM[c].
TPF_HD void make_record(p16 YA, p16 YB, p16 pW, p16 pY, p16 (&g)[8])
{
    const p16 s = add2(YA, YB), d = add2(YA, neg2(YB)), ns = neg2(s), nd = neg2(d);
    const p16 nw = neg2(pW), ny = neg2(pY);
    p16 a[4];                                   // a[c] = (+-W) + (+-Yp): c&2 -> -W, c&1 -> -Yp; a[3-c] = -a[c]
    a[0] = add2(pW, pY); a[1] = add2(pW, ny); a[2] = add2(nw, pY); a[3] = add2(nw, ny);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        g[2 * c] = max2(add2(s, a[c]), add2(ns, a[c]));             // max(P[c], -P[3-c])
        g[2 * c + 1] = max2(add2(d, a[3 - c]), add2(nd, a[3 - c])); // max(M[3-c], -M[c])
    }
}

// One step of the "in" pass in the shared wiring (tpf::pass_step), normalised by state 0.
TPF_HD void pass_step(p16 (&v)[16], const p16 (&g)[8], bool isb)
{
    p16 G[8];
    G[0] = g[0]; G[1] = g[1]; G[6] = g[6]; G[7] = g[7];
    G[2] = isb ? g[3] : g[2]; G[3] = isb ? g[2] : g[3];
    G[4] = isb ? g[5] : g[4]; G[5] = isb ? g[4] : g[5];
    p16 n[16];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int c = cls(t);
        const p16 A = (t & 4) ? G[2 * c + 1] : G[2 * c];
        const p16 B = (t & 4) ? G[2 * c] : G[2 * c + 1];
        n[2 * t] = addmax2(v[8 + t], B, add2(v[t], A));
        n[2 * t + 1] = addmax2(v[8 + t], A, add2(v[t], B));
    }
    const p16 nz = neg2(n[0]);
    v[0] = 0u;
#pragma unroll
    for (int s = 1; s < 16; ++s) v[s] = add2(n[s], nz);
}

// Backward step in natural labels (tpf::bwd_step).
TPF_HD void bwd_step(p16 (&z)[16], const p16 (&g)[8])
{
    p16 n[16];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int c = cls(t);
        const p16 A = (t & 4) ? g[2 * c + 1] : g[2 * c];
        const p16 B = (t & 4) ? g[2 * c] : g[2 * c + 1];
        n[t] = addmax2(z[2 * t + 1], B, add2(z[2 * t], A));
        n[8 + t] = addmax2(z[2 * t + 1], A, add2(z[2 * t], B));
    }
    const p16 nq = neg2(n[0]);
    z[0] = 0u;
#pragma unroll
    for (int s = 1; s < 16; ++s) z[s] = add2(n[s], nq);
}

// A-posteriori maxima (nii::app_maxima): uv = (U0, U3, V1, V2).
TPF_HD void app_maxima(const p16 (&x)[16], const p16 (&zs)[16], const p16 (&g)[8], p16 (&uv)[4])
{
    p16 TP[4], TM[4];
    bool seen[4] = {false, false, false, false};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int c = cls(t);
        const int t2 = (t >> 2) & 1;
        const int nP0 = 2 * t + t2, nM0 = 2 * t + 1 - t2;
        const p16 p = addmax2(x[8 + t], zs[nM0], add2(x[t], zs[nP0]));
        const p16 m = addmax2(x[8 + t], zs[nP0], add2(x[t], zs[nM0]));
        if (!seen[c]) { TP[c] = p; TM[c] = m; seen[c] = true; }
        else          { TP[c] = max2(TP[c], p); TM[c] = max2(TM[c], m); }
    }
    p16 U0 = add2(TP[0], g[0]), U3 = add2(TP[0], neg2(g[6])), V1 = add2(TM[0], g[1]), V2 = add2(TM[0], neg2(g[7]));
#pragma unroll
    for (int c = 1; c < 4; ++c) {
        U0 = addmax2(TP[c], g[2 * c], U0);
        U3 = addmax2(TP[c], neg2(g[2 * (3 - c)]), U3);
        V1 = addmax2(TM[c], g[2 * c + 1], V1);
        V2 = addmax2(TM[c], neg2(g[2 * (3 - c) + 1]), V2);
    }
    uv[0] = U0; uv[1] = U3; uv[2] = V1; uv[3] = V2;
}

// scalar epilogue of one frame: doubled a-posteriori maxima -> scaled, clamped extrinsic pair
TPF_HD void extrinsic1(int U0, int U3, int V1, int V2, int YA, int YB, int sf_q, int &ea, int &eb)
{
    const bool sP = YA + YB < 0, sM = YA - YB < 0;
    const int app0 = sP ? U3 : U0, app3 = sP ? U0 : U3, app1 = sM ? V2 : V1, app2 = sM ? V1 : V2;
    const int LA2 = (app0 > app1 ? app0 : app1) - (app2 > app3 ? app2 : app3);
    const int LB2 = (app0 > app2 ? app0 : app2) - (app1 > app3 ? app1 : app3);
    ea = ((LA2 - 2 * YA) * sf_q + 64) >> kSfShift;
    eb = ((LB2 - 2 * YB) * sf_q + 64) >> kSfShift;
    ea = ea > kExtMax ? kExtMax : (ea < -kExtMax ? -kExtMax : ea);
    eb = eb > kExtMax ? kExtMax : (eb < -kExtMax ? -kExtMax : eb);
}
TPF_HD void make_extrinsic(const p16 (&uv)[4], p16 YA, p16 YB, int sf_q, p16 &ea, p16 &eb)
{
    int a0, b0, a1, b1;
    extrinsic1(lo16(uv[0]), lo16(uv[1]), lo16(uv[2]), lo16(uv[3]), lo16(YA), lo16(YB), sf_q, a0, b0);
    extrinsic1(hi16(uv[0]), hi16(uv[1]), hi16(uv[2]), hi16(uv[3]), hi16(YA), hi16(YB), sf_q, a1, b1);
    ea = pack16(a0, a1);
    eb = pack16(b0, b1);
}

}  // namespace nii16
}  // namespace b200dvb
