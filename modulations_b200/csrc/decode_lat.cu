// decode_lat.cu — the LOW-LATENCY decode path: one CTA per frame, bit-exact with the reference
// (dvb_rcs2_turbo.py:116-281 bcjr_max_log_map, :464-537 decode), for the reference's own call shape — one frame
// per decode() (turbo_test_suite.py:138-164, test.py:56-92).
//
// Why a third kernel.  A frame is 16 SISOs of strictly sequential recursions: the throughput kernels amortise that
// over 64 frames per SM, and a batch of ONE frame then costs what a full wave costs (thread-per-frame kernel: 0.99 ms
// at N=212; quad kernel: 0.65 ms; profiles/r02_measure_pack1.txt).  Here everything that is NOT sequential is spread
// over the 256 threads of a CTA, and the two recursions run as two lone warps, one state metric per lane (4 shuffles +
// 7 FP32 operations per step: 77 cycles, against 120 for one thread holding all 16 states, whose 63 operations are
// issue-bound), and the second lap of a recursion stops as soon as it has provably re-joined the first:
//   P0 (all threads, one trellis step each): a-priori gather, Y = Lc + La in float64, the merged branch-metric record;
//   P1 (two warps): warp A runs alpha around the circular trellis, storing alpha[k]; the second lap (:182-183: it
//       starts from the first lap's end state) overwrites them and STOPS where its 16 metrics equal, bit for bit,
//       what the first lap stored at that position: from there on it would reproduce the first lap (deterministic
//       recursion, same inputs), whose values are already in place.  Exact; on noisy N=212 codewords the laps
//       re-join after 40 steps on average (p99: 108, measured with the oracle), so a SISO is ~N + 50 sequential
//       steps instead of 2 N.  Warp B does the same for beta (natural state labels), concurrently;
//   P2 (all threads, one step each): the 64 a-posteriori sums of a step and the float64 extrinsic epilogue.
// The whole frame lives in shared memory (240 bytes per couple: 51 KB at N=212, 204 KB at N=848), so the same kernel
// serves every N of the reference's table.  It is a latency path, not a throughput path: api.cu dispatches batches
// of up to a few hundred frames here, larger ones to the wave kernels.
#include "common.cuh"
#include "tpf_core.cuh"

namespace b200dvb {

namespace {

using namespace tpf;

constexpr int kLatThreads = 256;

struct LatArgs {
    int N, B, iterations, n_llr, num_sms;
    double sf_inner, sf_last;
    const int16_t *tab;             // [7][N]: perm, inv_perm, offA, offW1, offY1, offW2, offY2
    const float *llr;
    long long llr_stride;
    int32_t *bits;
    uint32_t *packed;
    const uint8_t *ref_bits;
    unsigned long long *counters;
};

__device__ __forceinline__ void ld16(const float *p, float (&v)[16])
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = reinterpret_cast<const float4 *>(p)[q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}
__device__ __forceinline__ void st16(float *p, const float (&v)[16])
{
#pragma unroll
    for (int q = 0; q < 4; ++q) reinterpret_cast<float4 *>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ void ld8(const float *p, float (&g)[8])
{
    const float4 a = reinterpret_cast<const float4 *>(p)[0], b = reinterpret_cast<const float4 *>(p)[1];
    g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
}


// phase timers (SM cycles of thread 0, summed over CTAs): 0 tables + de-puncture, 1 P0, 2 P1, 3 P2, 4 hard decision, 5 total
__device__ unsigned long long g_lat_cycles[8];

template <bool TIMED>
__global__ void __launch_bounds__(kLatThreads)
lat_kernel(const LatArgs A)
{
    long long ph[6] = {0, 0, 0, 0, 0, 0};
    const long long t_begin = TIMED ? clock64() : 0;
    extern __shared__ __align__(16) unsigned char sm[];
    const int N = A.N, tid = threadIdx.x;
    // shared-memory frame: [perm | inv] int16, then per couple: L1, L2 (float4), Le1, Le2, Y (double2), record (8 floats),
    // alpha (16 floats), beta (16 floats; index k holds beta[k], k = 1 .. N)
    int16_t *perm = reinterpret_cast<int16_t *>(sm);
    int16_t *inv = perm + N;
    unsigned char *p = sm + (((size_t)4 * N + 15) / 16) * 16;
    float4 *L1 = reinterpret_cast<float4 *>(p); p += (size_t)N * 16;
    float4 *L2 = reinterpret_cast<float4 *>(p); p += (size_t)N * 16;
    double2 *Le1 = reinterpret_cast<double2 *>(p); p += (size_t)N * 16;
    double2 *Le2 = reinterpret_cast<double2 *>(p); p += (size_t)N * 16;
    double2 *Y = reinterpret_cast<double2 *>(p); p += (size_t)N * 16;
    float *rec = reinterpret_cast<float *>(p); p += (size_t)N * 32;
    float *Al = reinterpret_cast<float *>(p); p += (size_t)N * 64;
    float *Be = reinterpret_cast<float *>(p); p += (size_t)(N + 1) * 64;
    unsigned *words = reinterpret_cast<unsigned *>(p);              // packed hard decisions, ceil(2N/32) words
    __shared__ int s_err[2];
    // the two lone recursion threads sit in different warps (= different schedulers); a second CTA on the same SM
    // uses the other two schedulers
    const int wa = ((blockIdx.x / A.num_sms) & 1) ? 64 : 0, wb = wa + 32;

    for (int frame = blockIdx.x; frame < A.B; frame += gridDim.x) {
        __syncthreads();
        long long tA = TIMED ? clock64() : 0;
        const float *row = A.llr + (size_t)frame * A.llr_stride;
        for (int i = tid; i < 2 * N; i += kLatThreads) perm[i] = A.tab[i];
        if (tid < 2) s_err[tid] = 0;
        for (int i = tid; i < (2 * N + 31) / 32; i += kLatThreads) words[i] = 0u;
        __syncthreads();
        // ---- de-puncture (:466-487) and gather the interleaved systematic pair (:507-512) ----
        const int16_t *g_off = A.tab + 2 * N;
        for (int k = tid; k < N; k += kLatThreads) {
            const int oa = g_off[k], op = g_off[perm[k]];
            const int o0 = g_off[N + k], o1 = g_off[2 * N + k], o2 = g_off[3 * N + k], o3 = g_off[4 * N + k];
            L1[k] = make_float4(__ldg(row + oa), __ldg(row + oa + 1), o0 >= 0 ? __ldg(row + o0) : 0.f, o1 >= 0 ? __ldg(row + o1) : 0.f);
            L2[k] = make_float4(__ldg(row + op), __ldg(row + op + 1), o2 >= 0 ? __ldg(row + o2) : 0.f, o3 >= 0 ? __ldg(row + o3) : 0.f);
        }
        __syncthreads();
        if (TIMED) { const long long t = clock64(); ph[0] += t - tA; tA = t; }
        for (int h = 0; h < 2 * A.iterations; ++h) {
            const bool second = (h & 1) != 0, first = h == 0;
            const double sf = (h >> 1) < A.iterations - 1 ? A.sf_inner : A.sf_last;
            // ---- P0: a-priori gather, Y = Lc + La, branch-metric record (:131-160) ----
            for (int k = tid; k < N; k += kLatThreads) {
                const float4 x = second ? L2[k] : L1[k];
                double2 la = make_double2(0.0, 0.0);
                if (!first) la = second ? Le1[perm[k]] : Le2[inv[k]];
                const double YA = d_add((double)x.x, la.x), YB = d_add((double)x.y, la.y);
                float g[8];
                make_record(YA, YB, x.z, x.w, g);
                Y[k] = make_double2(YA, YB);
                reinterpret_cast<float4 *>(rec + 8 * k)[0] = make_float4(g[0], g[1], g[2], g[3]);
                reinterpret_cast<float4 *>(rec + 8 * k)[1] = make_float4(g[4], g[5], g[6], g[7]);
            }
            __syncthreads();
            if (TIMED) { const long long t = clock64(); ph[1] += t - tA; tA = t; }
            // ---- P1: the two recursions, twice around the circular trellis (:162-230) ----
            // One state per lane (16 lanes of a warp, the upper half-warp mirrors the lower): the new metric of a state
            // needs two old ones and the normaliser n[0] two more, so a step is 4 shuffles -> 2 x (add, add, max) ->
            // subtract: a chain of dependent latencies instead of the ~120 issue cycles one thread needs for all 16
            // states.  Same operations on the same operands as tpf::pass_step / tpf::bwd_step, so the bits agree.
            // The second lap stops where it has re-joined the first: the recursion is deterministic, so once the 16
            // metrics of lap 2 equal, bit for bit, what lap 1 left at the same trellis position, every later position
            // of lap 2 would reproduce lap 1's values, which are already in place.  Measured with the oracle on
            // N=212 noisy codewords: re-joined after 40 steps on average (p99 108); a lap that never re-joins runs
            // to the end as before.  Exact, not an approximation: the exit is taken on verified equality only.
            // Measured (profiles/r02_latency.txt): 77 cycles per step, of which 59 are the bare dependent chain (shuffle
            // ~36 + add + max + subtract; timing-only builds without the bookkeeping).  Exchanging the metrics through
            // the shared-memory cell they are stored in anyway was 10 % slower than the shuffles, and hand-ordered
            // schedules (shuffles first; bookkeeping in the shadow of the arithmetic; loads made dependent on an add)
            // all lost 3-14 % to the compiler's: a load issued after the shuffles holds the adds back.
            if ((tid >> 5) == (wa >> 5)) {
                const int s = tid & 15, t = s >> 1, c2 = 2 * cls(t);
                const bool swp = (((t >> 2) ^ s) & 1) != 0;         // X (for v[t]) is the odd entry of the class pair
                const float2 *rp = reinterpret_cast<const float2 *>(rec) + (c2 >> 1);
                const float2 *r0 = reinterpret_cast<const float2 *>(rec);
                float v = 0.f;
                float2 pc = rp[0], p0 = r0[0];
                for (int lap = 0; lap < 2; ++lap) {
                    bool joined = false;
                    for (int k = 0; k < N; ++k) {
                        if (joined) break;                          // (tested one step late: the vote stays off the chain)
                        const int kn = k + 1 < N ? k + 1 : 0;
                        const float2 pn = rp[4 * kn], p0n = r0[4 * kn];
                        if (lap) joined = __all_sync(0xffffffffu, __float_as_uint(Al[16 * k + s]) == __float_as_uint(v));
                        if ((tid & 16) == 0) Al[16 * k + s] = v;
                        const float a = __shfl_sync(0xffffffffu, v, t, 16), bq = __shfl_sync(0xffffffffu, v, 8 + t, 16);
                        const float v0 = __shfl_sync(0xffffffffu, v, 0, 16), v8 = __shfl_sync(0xffffffffu, v, 8, 16);
                        const float X = swp ? pc.y : pc.x, Yv = swp ? pc.x : pc.y;
                        const float n = f_max(f_add(a, X), f_add(bq, Yv));
                        const float n0 = f_max(f_add(v0, p0.x), f_add(v8, p0.y));
                        v = f_sub(n, n0);
                        pc = pn; p0 = p0n;
                    }
                }
            } else if ((tid >> 5) == (wb >> 5)) {
                const int s = tid & 15, t = s & 7, c2 = 2 * cls(t);
                const bool swp = (((t >> 2) ^ (s >> 3)) & 1) != 0;
                const float2 *rp = reinterpret_cast<const float2 *>(rec) + (c2 >> 1);
                const float2 *r0 = reinterpret_cast<const float2 *>(rec);
                float z = 0.f;
                float2 pc = rp[4 * (N - 1)], p0 = r0[4 * (N - 1)];
                for (int lap = 0; lap < 2; ++lap) {
                    bool joined = false;
                    for (int k = N - 1; k >= 0; --k) {
                        if (joined) break;
                        const int kn = k > 0 ? k - 1 : N - 1;
                        const float2 pn = rp[4 * kn], p0n = r0[4 * kn];
                        if (lap) joined = __all_sync(0xffffffffu, __float_as_uint(Be[16 * (k + 1) + s]) == __float_as_uint(z));
                        if ((tid & 16) == 0) Be[16 * (k + 1) + s] = z;
                        const float a = __shfl_sync(0xffffffffu, z, 2 * t, 16), bq = __shfl_sync(0xffffffffu, z, 2 * t + 1, 16);
                        const float z0 = __shfl_sync(0xffffffffu, z, 0, 16), z1 = __shfl_sync(0xffffffffu, z, 1, 16);
                        const float X = swp ? pc.y : pc.x, Yv = swp ? pc.x : pc.y;
                        const float n = f_max(f_add(a, X), f_add(bq, Yv));
                        const float n0 = f_max(f_add(z0, p0.x), f_add(z1, p0.y));
                        z = f_sub(n, n0);
                        pc = pn; p0 = p0n;
                    }
                }
            }
            __syncthreads();
            if (TIMED) { const long long t = clock64(); ph[2] += t - tA; tA = t; }
            // ---- P2: a-posteriori maxima and the float64 extrinsic (:232-281) ----
            double2 *LeOut = second ? Le2 : Le1;
            for (int k = tid; k < N; k += kLatThreads) {
                float x[16], zs[16], g[8], uv[4];
                ld16(Al + 16 * k, x);
                ld16(Be + 16 * (k + 1), zs);
                ld8(rec + 8 * k, g);
                ext_step(x, zs, g, uv);
                const double2 y = Y[k];
                double ea, eb;
                make_extrinsic(uv, y.x, y.y, sf, ea, eb);
                LeOut[k] = make_double2(ea, eb);
            }
            __syncthreads();
            if (TIMED) { const long long t = clock64(); ph[3] += t - tA; tA = t; }
        }
        // ---- hard decision (:526-537) + optional error counting ----
        int my_err = 0;
        for (int k = tid; k < N; k += kLatThreads) {
            const float4 ab = L1[k];
            const double2 la = Le2[inv[k]], e1 = Le1[k];
            const double LA = d_add(d_add((double)ab.x, la.x), e1.x);
            const double LB = d_add(d_add((double)ab.y, la.y), e1.y);
            const int bA = LA < 0.0, bB = LB < 0.0;
            if (A.bits) *reinterpret_cast<int2 *>(A.bits + (size_t)frame * 2 * N + 2 * k) = make_int2(bA, bB);
            atomicOr(&words[k >> 4], (unsigned)(bA | (bB << 1)) << (2 * (k & 15)));
            if (A.ref_bits) {
                const uint8_t *r = A.ref_bits + (size_t)frame * 2 * N + 2 * k;
                my_err += (bA != r[0]) + (bB != r[1]);
            }
        }
        if (my_err) atomicAdd(&s_err[0], my_err);
        __syncthreads();
        const int wpf = (2 * N + 31) / 32;
        if (A.packed)
            for (int i = tid; i < wpf; i += kLatThreads) A.packed[(size_t)frame * wpf + i] = words[i];
        if (tid == 0 && A.counters) {
            if (A.ref_bits) {
                if (s_err[0]) { atomicAdd(A.counters + 0, (unsigned long long)s_err[0]); atomicAdd(A.counters + 1, 1ull); }
            }
            atomicAdd(A.counters + 2, 1ull);
            atomicAdd(A.counters + 3, 2ull * N);
        }
        if (TIMED) { const long long t = clock64(); ph[4] += t - tA; tA = t; }
    }
    if (TIMED && tid == 0) {
        ph[5] = clock64() - t_begin;
        for (int i = 0; i < 6; ++i) atomicAdd(&g_lat_cycles[i], (unsigned long long)ph[i]);
    }
}

}  // namespace

int lat_read_phase_cycles(double *out_h, int reset)
{
    unsigned long long h[8];
    B2_CUDA(cudaMemcpyFromSymbol(h, g_lat_cycles, sizeof h));
    for (int i = 0; i < 8; ++i) out_h[i] = (double)h[i];
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        B2_CUDA(cudaMemcpyToSymbol(g_lat_cycles, z, sizeof z));
    }
    return B200DVB_OK;
}

size_t lat_smem_bytes(int N)
{
    return (((size_t)4 * N + 15) / 16) * 16 + (size_t)N * (16 + 16 + 16 + 16 + 16 + 32 + 64) + (size_t)(N + 1) * 64 +
           (size_t)((2 * N + 31) / 32) * 4 + 64;
}

int lat_configure(Codec &c)
{
    c.lat_enabled = 0;
    int dev = 0;
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaGetDeviceProperties(&prop, dev));
    const size_t need = lat_smem_bytes(c.N);
    cudaFuncAttributes fa;
    B2_CUDA(cudaFuncGetAttributes(&fa, lat_kernel<false>));
    const size_t cap = (size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes;   // the static words count against the limit
    if (need > cap) return B200DVB_OK;
    B2_CUDA(cudaFuncSetAttribute(lat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
    B2_CUDA(cudaFuncSetAttribute(lat_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
    int occ = 0;
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lat_kernel<false>, kLatThreads, need));
    if (occ < 1) return B200DVB_OK;
    c.lat_enabled = 1;
    c.lat_frames_per_wave = (occ < 2 ? occ : 2) * prop.multiProcessorCount;   // at most two CTAs per SM: one lone thread per scheduler
    return B200DVB_OK;
}

int lat_launch_decode(const Codec &c, int B, const float *llr, long long llr_stride, int32_t *bits, uint32_t *packed,
                      const uint8_t *ref_bits, unsigned long long *counters, cudaStream_t s)
{
    if (B == 0) return B200DVB_OK;
    LatArgs A{};
    A.N = c.N; A.B = B; A.iterations = c.iterations; A.n_llr = c.n_llr; A.num_sms = c.num_sms;
    A.sf_inner = c.sf_inner; A.sf_last = c.sf_last; A.tab = c.d_tab;
    A.llr = llr; A.llr_stride = llr_stride; A.bits = bits; A.packed = packed; A.ref_bits = ref_bits; A.counters = counters;
    const int grid = B < c.lat_frames_per_wave ? B : c.lat_frames_per_wave;
    if (c.opt_phase_timers) lat_kernel<true><<<grid, kLatThreads, lat_smem_bytes(c.N), s>>>(A);
    else                    lat_kernel<false><<<grid, kLatThreads, lat_smem_bytes(c.N), s>>>(A);
    B2_CUDA(cudaGetLastError());
    return B200DVB_OK;
}

}  // namespace b200dvb
