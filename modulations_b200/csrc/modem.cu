// modem.cu — table-driven mapper, hard slicer and max-log soft demapper.
//
// Reference: SDRModem.modulate / demodulate (sdr_modem.py:101-266), Modulator
// (modulators.py:119-200) and compute_llr (test_sdr_with_coding.py:200-225):
//     llr[i*bps+b] = (min_{c: bit_b(c)=0} |s_i - c|^2 - min_{c: bit_b(c)=1} |s_i - c|^2) / max(nv, 0.005)
//     clipped to +-30, bit b = 0 is the MSB of the label, positive LLR <=> bit 1.
//
// Two demapper kernels, both HBM-streaming (8 B in, 4*bps B out per symbol):
//   * demap_generic<BPS>: any constellation.  Table in shared memory, one thread
//     per symbol, all M distances, per-bit minima with compile-time label bits.
//   * demap_pwl<HALF>: square constellations whose label splits into an I half and
//     a Q half (all of the reference's QAM tables and QPSK).  |s-c|^2 = dI^2 + dQ^2,
//     the other axis' term is common to both minima and cancels, so each bit's
//     max-log LLR depends on one coordinate only and is piecewise linear in it:
//     between consecutive mid-points of the axis levels the nearest level of each
//     subset is fixed and (x-Xa)^2 - (x-Xb)^2 = 2(Xb-Xa)x + (Xa^2-Xb^2).  The host
//     tabulates (slope, intercept) per bit and segment from the table it is given —
//     2L-2 segments per axis — so 256QAM costs ~10 flops per LLR instead of 256
//     distances: the kernel stays on the HBM roofline.
#include "common.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace b200dvb {

namespace {

constexpr int kThreads = 256;

template <typename OutT> struct Cplx;
template <> struct Cplx<float> { using type = float2; };
template <> struct Cplx<double> { using type = double2; };

template <typename OutT>
__global__ void map_kernel(size_t n, int bps, const uint8_t *__restrict__ bits,
                           const double *__restrict__ table, typename Cplx<OutT>::type *__restrict__ iq)
{
    __shared__ double2 tab[256];
    const int M = 1 << bps;
    for (int i = threadIdx.x; i < M; i += blockDim.x) tab[i] = make_double2(table[2 * i], table[2 * i + 1]);
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int lab = 0;
        for (int b = 0; b < bps; ++b) lab = (lab << 1) | (bits[i * bps + b] & 1);
        const double2 p = tab[lab];
        typename Cplx<OutT>::type o;
        o.x = (OutT)p.x; o.y = (OutT)p.y;       // complex128 -> complex64 rounds exactly like np.array(.., complex64)
        iq[i] = o;
    }
}

template <typename InT>
__global__ void hard_kernel(size_t n, int bps, const typename Cplx<InT>::type *__restrict__ iq,
                            const double *__restrict__ table, uint8_t *__restrict__ bits)
{
    __shared__ double2 tab[256];
    const int M = 1 << bps;
    for (int i = threadIdx.x; i < M; i += blockDim.x) tab[i] = make_double2(table[2 * i], table[2 * i + 1]);
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double x = (double)iq[i].x, y = (double)iq[i].y;
        double best = 1e300; int arg = 0;
        for (int c = 0; c < M; ++c) {
            const double dr = x - tab[c].x, di = y - tab[c].y;
            const double d = dr * dr + di * di;
            if (d < best) { best = d; arg = c; }
        }
        for (int b = 0; b < bps; ++b) bits[i * bps + b] = (uint8_t)((arg >> (bps - 1 - b)) & 1);
    }
}

template <int BPS>
__device__ __forceinline__ void store_llr(float *__restrict__ out, const float (&v)[BPS])
{
    if (BPS == 4) {
        *reinterpret_cast<float4 *>(out) = make_float4(v[0], v[1], v[2], v[3]);
    } else if (BPS == 8) {
        *reinterpret_cast<float4 *>(out) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(out + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else if (BPS == 2 || BPS == 6) {
#pragma unroll
        for (int b = 0; b < BPS; b += 2) *reinterpret_cast<float2 *>(out + b) = make_float2(v[b], v[b + 1]);
    } else {
#pragma unroll
        for (int b = 0; b < BPS; ++b) out[b] = v[b];
    }
}

template <int BPS>
__global__ void __launch_bounds__(kThreads)
demap_generic(size_t n, const float2 *__restrict__ iq, const float *__restrict__ table,
              float inv_nv, float scale, float *__restrict__ llr)
{
    constexpr int M = 1 << BPS;
    __shared__ float2 tab[M];
    for (int i = threadIdx.x; i < M; i += blockDim.x) tab[i] = make_float2(table[2 * i], table[2 * i + 1]);
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float2 s = __ldcs(iq + i);
        float d0[BPS], d1[BPS];
#pragma unroll
        for (int b = 0; b < BPS; ++b) { d0[b] = 3.0e38f; d1[b] = 3.0e38f; }
#pragma unroll
        for (int c = 0; c < M; ++c) {
            const float dr = s.x - tab[c].x, di = s.y - tab[c].y;
            const float d = fmaf(di, di, dr * dr);
#pragma unroll
            for (int b = 0; b < BPS; ++b) {
                if ((c >> (BPS - 1 - b)) & 1) d1[b] = fminf(d1[b], d);
                else                          d0[b] = fminf(d0[b], d);
            }
        }
        float v[BPS];
#pragma unroll
        for (int b = 0; b < BPS; ++b)
            v[b] = fminf(fmaxf((d0[b] - d1[b]) * inv_nv, -30.f), 30.f) * scale;
        store_llr<BPS>(llr + i * BPS, v);
    }
}

// Per-axis piecewise-linear demapper.  coef: float2[2][HALF][nseg].
template <int HALF>
__global__ void __launch_bounds__(kThreads)
demap_pwl(size_t n, const float2 *__restrict__ iq, const float2 *__restrict__ coef, int nseg,
          float x0a, float invda, float x0b, float invdb, int first_is_q, float inv_nv, float scale,
          float *__restrict__ llr)
{
    extern __shared__ float2 sc[];
    for (int i = threadIdx.x; i < 2 * HALF * nseg; i += blockDim.x) sc[i] = coef[i];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float2 s = __ldcs(iq + i);
        const float xa = first_is_q ? s.y : s.x, xb = first_is_q ? s.x : s.y;
        const int ua = min(max(__float2int_rd((xa - x0a) * invda), 0), nseg - 1);
        const int ub = min(max(__float2int_rd((xb - x0b) * invdb), 0), nseg - 1);
        float v[2 * HALF];
#pragma unroll
        for (int b = 0; b < HALF; ++b) {
            const float2 ca = sc[b * nseg + ua], cb = sc[(HALF + b) * nseg + ub];
            v[b] = fminf(fmaxf(fmaf(ca.x, xa, ca.y) * inv_nv, -30.f), 30.f) * scale;
            v[HALF + b] = fminf(fmaxf(fmaf(cb.x, xb, cb.y) * inv_nv, -30.f), 30.f) * scale;
        }
        store_llr<2 * HALF>(llr + i * 2 * HALF, v);
    }
}

int grid_for(size_t n, int threads, int per_sm)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t need = (n + threads - 1) / threads;
    size_t cap = (size_t)sms * per_sm;
    return (int)std::max<size_t>(1, std::min(need, cap));
}

}  // namespace

// Host: tabulate (slope, intercept) per axis, bit and segment (see header comment).
int modem_build_pwl(Modem &m)
{
    const int L = m.nlev, half = m.half;
    m.pwl = 0;
    if (!m.separable || L < 2) return B200DVB_OK;
    const int nseg = 2 * L - 2;
    std::vector<float> coef((size_t)2 * half * nseg * 2);
    for (int ax = 0; ax < 2; ++ax) {
        // ax 0 = axis selected by the first half of the label, ax 1 = second half
        std::vector<double> lev(L);
        for (int j = 0; j < L; ++j) {
            const int lab = (ax == 0) ? j * L : j;
            const bool is_q = (ax == 0) ? (m.separable == 2) : (m.separable == 1);
            lev[j] = m.h_table[2 * lab + (is_q ? 1 : 0)];
        }
        std::vector<double> sorted = lev;
        std::sort(sorted.begin(), sorted.end());
        const double span = sorted[L - 1] - sorted[0];
        if (!(span > 0)) return B200DVB_OK;
        const double step = span / (L - 1);
        for (int j = 0; j < L; ++j)
            if (fabs(sorted[j] - (sorted[0] + step * j)) > 1e-6 * span) return B200DVB_OK;   // not uniform: generic path
        const double delta = step * 0.5, x0 = sorted[0];
        m.pwl_x0[ax] = (float)x0;
        m.pwl_invd[ax] = (float)(1.0 / delta);
        for (int b = 0; b < half; ++b)
            for (int u = 0; u < nseg; ++u) {
                const double xc = (u == 0) ? x0 : (u == nseg - 1 ? sorted[L - 1] : x0 + delta * (u + 0.5));
                int a0 = -1, a1 = -1;
                for (int j = 0; j < L; ++j) {
                    const int bit = (j >> (half - 1 - b)) & 1;
                    int &best = bit ? a1 : a0;
                    if (best < 0 || fabs(xc - lev[j]) < fabs(xc - lev[best])) best = j;
                }
                if (a0 < 0 || a1 < 0) return B200DVB_OK;
                const double Xa = lev[a0], Xb = lev[a1];
                const size_t o = (((size_t)ax * half + b) * nseg + u) * 2;
                coef[o] = (float)(2.0 * (Xb - Xa));
                coef[o + 1] = (float)(Xa * Xa - Xb * Xb);
            }
    }
    B2_CUDA(cudaMalloc(&m.d_pwl, coef.size() * sizeof(float)));
    B2_CUDA(cudaMemcpy(m.d_pwl, coef.data(), coef.size() * sizeof(float), cudaMemcpyHostToDevice));
    m.nseg = nseg;
    m.pwl = 1;
    return B200DVB_OK;
}

int launch_map(const Modem &m, size_t n, const uint8_t *bits, void *iq, int out_f64, cudaStream_t s)
{
    if (n == 0) return B200DVB_OK;
    const int grid = grid_for(n, kThreads, 8);
    if (out_f64) map_kernel<double><<<grid, kThreads, 0, s>>>(n, m.bps, bits, m.d_table64, (double2 *)iq);
    else         map_kernel<float><<<grid, kThreads, 0, s>>>(n, m.bps, bits, m.d_table64, (float2 *)iq);
    B2_CUDA(cudaGetLastError());
    return B200DVB_OK;
}

int launch_hard(const Modem &m, size_t n, const void *iq, int in_f64, uint8_t *bits, cudaStream_t s)
{
    if (n == 0) return B200DVB_OK;
    const int grid = grid_for(n, kThreads, 8);
    if (in_f64) hard_kernel<double><<<grid, kThreads, 0, s>>>(n, m.bps, (const double2 *)iq, m.d_table64, bits);
    else        hard_kernel<float><<<grid, kThreads, 0, s>>>(n, m.bps, (const float2 *)iq, m.d_table64, bits);
    B2_CUDA(cudaGetLastError());
    return B200DVB_OK;
}

int launch_demap(const Modem &m, size_t n, const void *iq_, float noise_var, float scale, float *llr,
                 cudaStream_t s)
{
    if (n == 0) return B200DVB_OK;
    const float2 *iq = (const float2 *)iq_;
    const float nv = noise_var > 0.005f ? noise_var : 0.005f;      // test_sdr_with_coding.py:202
    const float inv_nv = 1.0f / nv;
    const int grid = grid_for(n, kThreads, 8);
    if (m.pwl) {
        const float2 *coef = (const float2 *)m.d_pwl;
        const size_t sm = (size_t)2 * m.half * m.nseg * sizeof(float2);
        const int fq = (m.separable == 2);
#define PWL(H) demap_pwl<H><<<grid, kThreads, sm, s>>>(n, iq, coef, m.nseg, m.pwl_x0[0], m.pwl_invd[0], \
                                                       m.pwl_x0[1], m.pwl_invd[1], fq, inv_nv, scale, llr)
        switch (m.half) {
        case 1: PWL(1); break;
        case 2: PWL(2); break;
        case 3: PWL(3); break;
        case 4: PWL(4); break;
        default: return B200DVB_ENOSPEC;
        }
#undef PWL
    } else {
        switch (m.bps) {
        case 1: demap_generic<1><<<grid, kThreads, 0, s>>>(n, iq, m.d_table32, inv_nv, scale, llr); break;
        case 2: demap_generic<2><<<grid, kThreads, 0, s>>>(n, iq, m.d_table32, inv_nv, scale, llr); break;
        case 3: demap_generic<3><<<grid, kThreads, 0, s>>>(n, iq, m.d_table32, inv_nv, scale, llr); break;
        case 4: demap_generic<4><<<grid, kThreads, 0, s>>>(n, iq, m.d_table32, inv_nv, scale, llr); break;
        case 6: demap_generic<6><<<grid, kThreads, 0, s>>>(n, iq, m.d_table32, inv_nv, scale, llr); break;
        case 8: demap_generic<8><<<grid, kThreads, 0, s>>>(n, iq, m.d_table32, inv_nv, scale, llr); break;
        default: return B200DVB_ENOSPEC;
        }
    }
    B2_CUDA(cudaGetLastError());
    return B200DVB_OK;
}

}  // namespace b200dvb
