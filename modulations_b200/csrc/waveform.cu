// waveform.cu — the stage on either side of the mapper / demapper:
// root-raised-cosine pulse shaping of symbols (modulators.py:85-100: upfirdn(h, syms, up=sps))
// and the receive matched filter + symbol-rate decimation (modulators.py:102-117:
// convolve(samples, h, 'full')[2*delay::sps]).  Both are HBM-streaming FIRs over complex64:
//
//   pulse shaping   8 B of symbol in, 8*sps B of samples out per symbol: write bound.  Polyphase:
//                   out[q*sps + p] = sum_j s[q-j] * h[p + j*sps]; a thread owns one output sample,
//                   the <= ceil(ntaps/sps) symbols it needs are shared by the sps threads around it
//                   (L1), the taps sit in shared memory.
//   matched filter  8*sps B of samples in, 8 B of symbol out per symbol: read bound.  A block owns
//                   256 consecutive outputs; the samples they span are staged in shared memory
//                   PHASE-MAJOR ([i mod sps][i / sps]): for a given tap every thread of a warp then
//                   reads consecutive 8-byte words (no bank conflicts — sample-major staging would be
//                   a 2*sps-word stride), and every input sample is read from HBM exactly once.
//
// Arithmetic: float32 FMAs, taps rounded once to float32; the reference convolves in float64, so
// parity is a stated tolerance (tests/test_gpu_waveform.py), not bit-exactness.
#include "common.cuh"

namespace b200dvb {

namespace {

constexpr int kMfTile = 256;        // outputs per block = threads per block

__global__ void __launch_bounds__(256)
pulse_shape_kernel(size_t n_sym, const float2 *__restrict__ sym, const float *__restrict__ taps, int ntaps,
                   int sps, size_t n_out, float2 *__restrict__ out)
{
    extern __shared__ float sh[];
    for (int i = threadIdx.x; i < ntaps; i += blockDim.x) sh[i] = taps[i];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (size_t)gridDim.x * blockDim.x) {
        const size_t q = i / sps;
        const int p = (int)(i - q * sps);
        float ar = 0.f, ai = 0.f;
        for (int j = 0, t = p; t < ntaps; ++j, t += sps) {
            if (q >= (size_t)j && q - j < n_sym) {
                const float2 s = __ldg(sym + (q - j));
                ar = fmaf(s.x, sh[t], ar);
                ai = fmaf(s.y, sh[t], ai);
            }
        }
        out[i] = make_float2(ar, ai);
    }
}

// Compile-time sps: a thread owns ALL sps output samples of one symbol period q — the Q = ceil(ntaps/sps) symbols
// they depend on are loaded once (coalesced: consecutive threads, consecutive symbols) and the sps*Q complex MACs run
// out of registers with broadcast tap reads; the thread's sps samples are 8*sps contiguous bytes, stored 16 B at a time.
template <int SPS>
__global__ void __launch_bounds__(256)
pulse_shape_sps_kernel(size_t n_sym, const float2 *__restrict__ sym, const float *__restrict__ taps, int ntaps,
                       size_t n_out, float2 *__restrict__ out)
{
    constexpr int kMaxQ = 16;
    extern __shared__ float sh[];                                   // taps, zero-padded to Q * SPS
    const int Q = (ntaps + SPS - 1) / SPS;
    for (int i = threadIdx.x; i < Q * SPS; i += blockDim.x) sh[i] = i < ntaps ? taps[i] : 0.f;
    __syncthreads();
    const size_t nq = (n_out + SPS - 1) / SPS;                      // symbol periods that hold output samples
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (size_t)gridDim.x * blockDim.x) {
        float2 s[kMaxQ];
#pragma unroll
        for (int j = 0; j < kMaxQ; ++j)
            s[j] = (j < Q && q >= (size_t)j && q - j < n_sym) ? __ldg(sym + (q - j)) : make_float2(0.f, 0.f);
        float2 acc[SPS];
#pragma unroll
        for (int p = 0; p < SPS; ++p) acc[p] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kMaxQ; ++j) {
            if (j < Q) {
#pragma unroll
                for (int p = 0; p < SPS; ++p) {
                    const float hh = sh[j * SPS + p];
                    acc[p].x = fmaf(s[j].x, hh, acc[p].x);
                    acc[p].y = fmaf(s[j].y, hh, acc[p].y);
                }
            }
        }
        float2 *o = out + q * SPS;
        if ((q + 1) * SPS <= n_out && (SPS % 2) == 0) {             // whole period in range: 16-byte stores (q*SPS*8 B is 16-aligned)
#pragma unroll
            for (int p = 0; p < SPS; p += 2)
                *reinterpret_cast<float4 *>(o + p) = make_float4(acc[p].x, acc[p].y, acc[p + 1].x, acc[p + 1].y);
        } else {
#pragma unroll
            for (int p = 0; p < SPS; ++p)
                if (q * SPS + p < n_out) o[p] = acc[p];
        }
    }
}

// out[m] = sum_t h[t] * x[start + m*sps - t],  x[i] = 0 outside [0, n)
__global__ void __launch_bounds__(kMfTile)
matched_filter_kernel(size_t n, const float2 *__restrict__ x, const float *__restrict__ taps, int ntaps,
                      int sps, long long start, size_t n_out, float2 *__restrict__ out, int pitch)
{
    extern __shared__ float sh[];
    float *h = sh;                                                  // [ntaps]
    float2 *xs = reinterpret_cast<float2 *>(sh + ((ntaps + 1) & ~1)); // [sps][pitch], phase-major
    for (int i = threadIdx.x; i < ntaps; i += blockDim.x) h[i] = taps[i];
    for (size_t m0 = (size_t)blockIdx.x * kMfTile; m0 < n_out; m0 += (size_t)gridDim.x * kMfTile) {
        // samples spanned by outputs m0 .. m0 + 255: [lo, hi]; the staging origin is lo rounded DOWN to a multiple
        // of sps so that (i - base) mod sps is the phase of sample i for every block
        const long long hi = start + (long long)(m0 + kMfTile - 1) * sps;
        const long long lo = start + (long long)m0 * sps - (ntaps - 1);
        const long long base = (lo >= 0 ? lo / sps : -((-lo + sps - 1) / sps)) * sps;
        const int span = (int)(hi - base + 1);
        __syncthreads();                                            // previous tile fully consumed
        for (int r = threadIdx.x; r < span; r += blockDim.x) {
            const long long i = base + r;
            const float2 v = (i >= 0 && (size_t)i < n) ? __ldg(x + i) : make_float2(0.f, 0.f);
            xs[(r % sps) * pitch + r / sps] = v;
        }
        __syncthreads();
        const size_t m = m0 + threadIdx.x;
        const int r0 = (int)(start + (long long)m0 * sps - base);   // staged index of this tile's first output sample
        float ar = 0.f, ai = 0.f;
        // staged index of tap t for this thread: r0 - t (+ threadIdx.x * sps): phase and position walk down by one
        // sample per tap — warp-uniform bookkeeping, no division in the loop
        int ph = r0 % sps, off = ph * pitch + r0 / sps + threadIdx.x;
#pragma unroll 7
        for (int t = 0; t < ntaps; ++t) {
            const float2 v = xs[off];
            ar = fmaf(v.x, h[t], ar);
            ai = fmaf(v.y, h[t], ai);
            if (ph == 0) { ph = sps - 1; off += (sps - 1) * pitch - 1; }    // previous sample: last phase, one position back
            else         { --ph; off -= pitch; }
        }
        if (m < n_out) out[m] = make_float2(ar, ai);
    }
}

}  // namespace

int launch_pulse_shape(size_t n_sym, const void *sym, const float *taps, int ntaps, int sps, void *out, cudaStream_t s)
{
    if (n_sym == 0) return B200DVB_OK;
    const size_t n_out = (n_sym - 1) * (size_t)sps + ntaps;
    int dev = 0, sms = 148;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int Q = (ntaps + sps - 1) / sps;
    const bool al16 = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if ((sps == 8 || sps == 4 || sps == 2) && Q <= 16 && al16) {
        size_t blocks = ((n_out + sps - 1) / sps + 255) / 256;
        if (blocks > (size_t)sms * 16) blocks = (size_t)sms * 16;
        const size_t smem = (size_t)Q * sps * sizeof(float);
        const float2 *sy = reinterpret_cast<const float2 *>(sym);
        float2 *o = reinterpret_cast<float2 *>(out);
        if (sps == 8)      pulse_shape_sps_kernel<8><<<(unsigned)blocks, 256, smem, s>>>(n_sym, sy, taps, ntaps, n_out, o);
        else if (sps == 4) pulse_shape_sps_kernel<4><<<(unsigned)blocks, 256, smem, s>>>(n_sym, sy, taps, ntaps, n_out, o);
        else               pulse_shape_sps_kernel<2><<<(unsigned)blocks, 256, smem, s>>>(n_sym, sy, taps, ntaps, n_out, o);
        B2_CUDA(cudaGetLastError());
        return B200DVB_OK;
    }
    size_t blocks = (n_out + 255) / 256;
    if (blocks > (size_t)sms * 16) blocks = (size_t)sms * 16;
    pulse_shape_kernel<<<(unsigned)blocks, 256, ntaps * sizeof(float), s>>>(
        n_sym, reinterpret_cast<const float2 *>(sym), taps, ntaps, sps, n_out, reinterpret_cast<float2 *>(out));
    B2_CUDA(cudaGetLastError());
    return B200DVB_OK;
}

int launch_matched_filter(size_t n, const void *x, const float *taps, int ntaps, int sps, long long start,
                          size_t n_out, void *out, cudaStream_t s)
{
    if (n_out == 0) return B200DVB_OK;
    // staged positions per phase: 256 outputs + the taps' reach + the rounding of the origin; odd pitch (in 8-byte
    // words) keeps the sps rows of the staging pass in different banks
    const int pitch = (kMfTile + (ntaps + sps - 1) / sps + 2) | 1;
    const size_t smem = (size_t)((ntaps + 1) & ~1) * sizeof(float) + (size_t)sps * pitch * sizeof(float2);
    if (smem > 200 * 1024) return B200DVB_EINVAL;
    int dev = 0, sms = 148;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (smem > 48 * 1024)
        B2_CUDA(cudaFuncSetAttribute(matched_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    size_t blocks = (n_out + kMfTile - 1) / kMfTile;
    if (blocks > (size_t)sms * 8) blocks = (size_t)sms * 8;
    matched_filter_kernel<<<(unsigned)blocks, kMfTile, smem, s>>>(
        n, reinterpret_cast<const float2 *>(x), taps, ntaps, sps, start, n_out, reinterpret_cast<float2 *>(out), pitch);
    B2_CUDA(cudaGetLastError());
    return B200DVB_OK;
}

}  // namespace b200dvb
