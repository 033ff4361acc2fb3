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
        for (int t = 0; t < ntaps; ++t) {
            const int r = r0 - t;                                   // >= 0 by construction of base
            const float2 v = xs[(r % sps) * pitch + r / sps + threadIdx.x];
            ar = fmaf(v.x, h[t], ar);
            ai = fmaf(v.y, h[t], ai);
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
