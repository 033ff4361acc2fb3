// microbench.cu — measured issue rates of the instructions the decoder is made of
// (lane-ops per clock per SM).  bench.py quotes the ALU roofline against the
// nominal 64 ACS/clk/SM of SURVEY §8(d) AND against these measured pipes.
//
// results[0] FADD   results[1] FMNMX   results[2] ACS mix (2 FADD + 1 FMNMX)
// results[3] SHFL   results[4] DADD    results[5] F2F (f64<->f32 conversions)
// results[6] add.f32x2 on packed registers (counted as 2 lane-ops)   results[7] SM clock in MHz
#include "common.cuh"

namespace b200dvb {

namespace {

constexpr int kIter = 4096;
constexpr int kChains = 8;

template <int OP>
__global__ void __launch_bounds__(1024) probe(float seed, long long *cycles, float *sink)
{
    float x[kChains];
    double d[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { x[i] = seed + i + threadIdx.x; d[i] = x[i]; }
    const float c = seed * 0.5f, c2 = seed * 0.25f;
    unsigned long long pk[kChains], qk;
#pragma unroll
    for (int i = 0; i < kChains; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pk[i]) : "f"(x[i]), "f"(x[i] + 1.f));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(qk) : "f"(c));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIter; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c));
            if (OP == 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c));
            if (OP == 2) {
                float a, b;
                asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(a) : "f"(x[i]), "f"(c));
                asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(b) : "f"(x[(i + 1) % kChains]), "f"(c2));
                asm volatile("max.f32 %0, %1, %2;" : "=f"(x[i]) : "f"(a), "f"(b));
            }
            if (OP == 3) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+f"(x[i]));
            if (OP == 4) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"((double)c));
            if (OP == 5) {
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(x[i]) : "d"(d[i]));
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(x[i]));
            }
            if (OP == 6) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pk[i]) : "l"(qk));
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(pk[i]));
        acc += x[i] + (float)d[i] + lo + hi;
    }
    if (acc == 123.456f) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
int run_one(int sms, long long *d_cycles, float *d_sink, double ops_per_iter, double *out)
{
    probe<OP><<<sms, 1024>>>(1.0f, d_cycles, d_sink);   // warm-up
    probe<OP><<<sms, 1024>>>(1.0f, d_cycles, d_sink);
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaDeviceSynchronize());
    long long *h = (long long *)malloc(sizeof(long long) * sms);
    B2_CUDA(cudaMemcpy(h, d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    free(h);
    *out = 1024.0 * kIter * kChains * ops_per_iter / avg;
    return B200DVB_OK;
}

}  // namespace

int run_microbench(double *r)
{
    int dev = 0, sms = 0, khz = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B2_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    long long *d_cycles = nullptr;
    float *d_sink = nullptr;
    B2_CUDA(cudaMalloc(&d_cycles, sizeof(long long) * sms));
    B2_CUDA(cudaMalloc(&d_sink, sizeof(float)));
    int rc = B200DVB_OK;
    if (rc == 0) rc = run_one<0>(sms, d_cycles, d_sink, 1, r + 0);
    if (rc == 0) rc = run_one<1>(sms, d_cycles, d_sink, 1, r + 1);
    if (rc == 0) rc = run_one<2>(sms, d_cycles, d_sink, 1, r + 2);   // one ACS per iteration
    if (rc == 0) rc = run_one<3>(sms, d_cycles, d_sink, 1, r + 3);
    if (rc == 0) rc = run_one<4>(sms, d_cycles, d_sink, 1, r + 4);
    if (rc == 0) rc = run_one<5>(sms, d_cycles, d_sink, 2, r + 5);   // two conversions per iteration
    if (rc == 0) rc = run_one<6>(sms, d_cycles, d_sink, 2, r + 6);
    r[7] = khz / 1000.0;
    cudaFree(d_cycles);
    cudaFree(d_sink);
    return rc;
}

}  // namespace b200dvb

// ---------------------------------------------------------------------------
// TMEM as a lane-private scratchpad shared by the warps of one lane quadrant:
// self-test of the tcgen05 alloc / st / ld / dealloc protocol the decoder uses for
// its recursion checkpoints (warps w and w+4 address the same 32 TMEM lanes).
// ---------------------------------------------------------------------------
namespace b200dvb {
namespace {

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}

__global__ void __launch_bounds__(256) tmem_selftest_kernel(int ncols, int *errors)
{
    __shared__ unsigned s_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_base)), "r"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned base = s_base;
    const unsigned quad = (unsigned)(warp & 3) * 32u;
    if (warp >= 4) {                                   // writers
        for (int c = 0; c < ncols; c += 4) {
            const unsigned v0 = (quad + lane) * 1000u + c, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3;
            const unsigned taddr = base + (quad << 16) + c;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                         ::"r"(taddr), "r"(v0), "r"(v1), "r"(v2), "r"(v3));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    int bad = 0;
    if (warp < 4) {                                    // readers in the same quadrant
        for (int c = 0; c < ncols; c += 4) {
            unsigned r0, r1, r2, r3;
            const unsigned taddr = base + (quad << 16) + c;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            const unsigned e = (quad + lane) * 1000u + c;
            bad += (r0 != e) + (r1 != e + 1) + (r2 != e + 2) + (r3 != e + 3);
        }
    }
    if (bad) atomicAdd(errors, bad);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols));
}

}  // namespace

int run_tmem_selftest(int *result_h)
{
    int *d = nullptr;
    B2_CUDA(cudaMalloc(&d, sizeof(int)));
    B2_CUDA(cudaMemset(d, 0, sizeof(int)));
    int sms = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    tmem_selftest_kernel<<<sms, 256>>>(512, d);
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaDeviceSynchronize());
    B2_CUDA(cudaMemcpy(result_h, d, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return B200DVB_OK;
}

}  // namespace b200dvb

// ---------------------------------------------------------------------------
// Second probe set: packed 16-bit arithmetic for the non-parity decoder modes: is an
// add-compare-select in f16x2 / s16x2 (two frames per lane) or as a DPX VIADDMNMX cheaper per ACS than
// the FADD + FMNMX pair of the parity decoder?  results16[] in lane-INSTRUCTIONS per clock per SM (a
// packed instruction counts once):
//  0 add.f16x2   1 max.f16x2   2 step mix f16x2 (2 add + 1 max + 1 sub)   3 step mix f32 (same shape)
//  4 add.s32     5 max.s32     6 viaddmax.s32 (DPX)   7 add.s16x2   8 max.s16x2   9 viaddmax.s16x2 (DPX)
// 10 step mix s16x2 (add + viaddmax + sub)   11 step mix s32 (add + viaddmax + sub)   12..15 reserved
// ---------------------------------------------------------------------------
#include <cuda_fp16.h>
namespace b200dvb {
namespace {

template <int OP>
__global__ void __launch_bounds__(1024) probe2(unsigned seed, long long *cycles, unsigned *sink)
{
    unsigned x[kChains];
    float f[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { x[i] = seed * 0x3c003c00u + i + threadIdx.x; f[i] = (float)(seed + i); }
    const unsigned c = seed * 0x34003400u, c2 = seed * 0x30003000u;
    const float fc = seed * 0.5f, fc2 = seed * 0.25f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIter; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            const int j = (i + 1) % kChains;
            if (OP == 0) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
            if (OP == 1) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
            if (OP == 2) {
                unsigned a, b, m;
                asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(a) : "r"(x[i]), "r"(c));
                asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(b) : "r"(x[j]), "r"(c2));
                asm volatile("max.f16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(b));
                asm volatile("sub.rn.f16x2 %0, %1, %2;" : "=r"(x[i]) : "r"(m), "r"(c2));
            }
            if (OP == 3) {
                float a, b, m;
                asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(a) : "f"(f[i]), "f"(fc));
                asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(b) : "f"(f[j]), "f"(fc2));
                asm volatile("max.f32 %0, %1, %2;" : "=f"(m) : "f"(a), "f"(b));
                asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(f[i]) : "f"(m), "f"(fc2));
            }
            if (OP == 4) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
            if (OP == 5) asm volatile("max.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
            if (OP == 6) x[i] = (unsigned)__viaddmax_s32((int)x[i], (int)c, (int)c2);
            if (OP == 7) asm volatile("add.s16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
            if (OP == 8) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
            if (OP == 9) x[i] = __viaddmax_s16x2(x[i], c, c2);
            if (OP == 10) {
                unsigned a, m;
                asm volatile("add.s16x2 %0, %1, %2;" : "=r"(a) : "r"(x[i]), "r"(c));
                m = __viaddmax_s16x2(x[j], c2, a);
                asm volatile("add.s16x2 %0, %1, %2;" : "=r"(x[i]) : "r"(m), "r"(c));   // normalisation: add of the packed negated metric
            }
            if (OP == 11) {
                int a = (int)x[i] + (int)c;
                int m = __viaddmax_s32((int)x[j], (int)c2, a);
                x[i] = (unsigned)(m - (int)c2);
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc += x[i] + __float_as_uint(f[i]);
    if (acc == 0x12345u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
int run_two(int sms, long long *d_cycles, unsigned *d_sink, double ops_per_iter, double *out)
{
    probe2<OP><<<sms, 1024>>>(1u, d_cycles, d_sink);
    probe2<OP><<<sms, 1024>>>(1u, d_cycles, d_sink);
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaDeviceSynchronize());
    long long *h = (long long *)malloc(sizeof(long long) * sms);
    B2_CUDA(cudaMemcpy(h, d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    free(h);
    *out = 1024.0 * kIter * kChains * ops_per_iter / avg;
    return B200DVB_OK;
}

}  // namespace

int run_microbench2(double *r)
{
    int dev = 0, sms = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long *d_cycles = nullptr;
    unsigned *d_sink = nullptr;
    B2_CUDA(cudaMalloc(&d_cycles, sizeof(long long) * sms));
    B2_CUDA(cudaMalloc(&d_sink, sizeof(unsigned)));
    for (int i = 0; i < 16; ++i) r[i] = 0.0;
    int rc = B200DVB_OK;
    if (rc == 0) rc = run_two<0>(sms, d_cycles, d_sink, 1, r + 0);
    if (rc == 0) rc = run_two<1>(sms, d_cycles, d_sink, 1, r + 1);
    if (rc == 0) rc = run_two<2>(sms, d_cycles, d_sink, 4, r + 2);
    if (rc == 0) rc = run_two<3>(sms, d_cycles, d_sink, 4, r + 3);
    if (rc == 0) rc = run_two<4>(sms, d_cycles, d_sink, 1, r + 4);
    if (rc == 0) rc = run_two<5>(sms, d_cycles, d_sink, 1, r + 5);
    if (rc == 0) rc = run_two<6>(sms, d_cycles, d_sink, 1, r + 6);
    if (rc == 0) rc = run_two<7>(sms, d_cycles, d_sink, 1, r + 7);
    if (rc == 0) rc = run_two<8>(sms, d_cycles, d_sink, 1, r + 8);
    if (rc == 0) rc = run_two<9>(sms, d_cycles, d_sink, 1, r + 9);
    if (rc == 0) rc = run_two<10>(sms, d_cycles, d_sink, 3, r + 10);
    if (rc == 0) rc = run_two<11>(sms, d_cycles, d_sink, 3, r + 11);
    cudaFree(d_cycles);
    cudaFree(d_sink);
    return rc;
}

}  // namespace b200dvb
