// common.cuh — shared declarations for libb200dvb.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/b200dvb.h"

namespace b200dvb {

// ---- error plumbing --------------------------------------------------------
void set_cuda_error(cudaError_t e, const char *where);
#define B2_CUDA(call)                                                     \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) {                                         \
            ::b200dvb::set_cuda_error(e__, #call);                        \
            return B200DVB_ECUDA;                                         \
        }                                                                 \
    } while (0)

constexpr int kMaxN = 2048;

// ---- decoder geometry (decode_quad.cu) --------------------------------------
constexpr int kFramesPerCta = 8;    // one 4-lane "quad" per frame and direction
constexpr int kCtaThreads = 64;     // per group: one forward (alpha) warp + one backward (beta) warp
constexpr int kMaxGroups = 7;       // groups (of 8 frames) per CTA
constexpr int kMaxCtaThreads = 448; // 14 warps: 2 per group + helper warps for the data-parallel phases
constexpr int kWin = 8;             // checkpoint spacing / recompute window (steps)

struct QuadGeom {
    int N;            // couples per frame
    unsigned magic;   // floor(2^32 / N) + 1: flat position -> (frame, step) without a divide
    int M;            // crossing point: alpha warp owns [M,N), beta warp owns [0,M)
    int nckA;         // alpha checkpoints (alpha[0], alpha[8], ... alpha[M-8])
    int nckB;         // beta checkpoints, one per alpha-warp window end
    int rec_stride;   // floats between two frames' branch-metric records (8N + 8)
    int ck_stride;    // floats between two frames' checkpoint areas
    int groups;       // (alpha warp, beta warp) pairs per CTA
    int threads;      // CTA size: 64 * groups recursion threads + helper warps
    int use_tmem;     // recursion checkpoints live in tensor memory instead of shared memory
    int tmem_cols;    // TMEM columns allocated per CTA (power of two >= 32)
    int y_slots;      // > 0: Y = Lc + La is parked in TMEM, this many positions per thread
    int frames;       // frames per CTA (8 * groups, or 4/2/1 when even one group does not fit)
    int ctas_per_sm;
    int grec;         // branch-metric records live in the (L1/L2-cached) global workspace, not in shared memory
    size_t smem_bytes;
};

// ---- thread-per-frame decoder geometry (decode_tpf.cu) ----------------------
constexpr int kTpfFrames = 16;      // frames per warp: lanes 0-15 run alpha, lanes 16-31 beta of the same frames
constexpr int kTpfWarps = 4;        // one warp per SM sub-partition, each with its own TMEM lane quadrant
constexpr int kTpfWin = 4;          // checkpoint spacing / recompute window (steps)

struct TpfGeom {
    int enabled;      // this codec is decoded by the thread-per-frame kernel
    int N, M;         // couples per frame; crossing point M = N / 2
    int T;            // branch-metric records per frame end kept in tensor memory ([0,T) and [N-T,N))
    int mid;          // records per frame kept in shared memory ([T, N-T))
    int nfull, rag;   // full recompute windows per half and length of the ragged last one (M % kTpfWin)
    int nslots;       // checkpoints per lane
    int tmem_cols;    // TMEM columns allocated per CTA (power of two >= 8 T, >= 32)
    size_t smem_bytes;
    size_t off_l1, off_l2, off_le, off_lef, off_y, off_ck, off_init;   // byte offsets in a warp's workspace (off_init: nii mode only)
    size_t ws_per_warp;
};

struct Codec {
    int N = 0, period = 0, iterations = 0, n_llr = 0;
    double sf_inner = 0.7, sf_last = 1.0;
    QuadGeom geom{};
    QuadGeom geom_g{};            // long frames: the same kernel with its records in global memory (geom_g.grec = 1) or frames = 0
    TpfGeom tpf{};
    TpfGeom nii{};                // geometry of the non-parity "nii" mode (decode_nii.cu)
    int num_sms = 0;
    int vec_ab = 0, vec_wy = 0;   // (A,B) / (W,Y) LLR pairs are adjacent and even-aligned in the stream
    // development / test switches (b200dvb_codec_set_option); all 0 in production
    int opt_kernel = 0;           // 0: automatic choice per batch, 1: quad kernel, 2: thread-per-frame kernel, 3: low-latency kernel
    int opt_no_row_staging = 0;   // 1: thread-per-frame transposition without the cp.async row staging
    int opt_phase_timers = 0;     // 1: run the kernel instances that keep per-phase cycle counters
    int lat_pad = 0;              // low-latency kernel: conflict-free (padded) shared-memory strides fit for this N
    int lat_enabled = 0, lat_frames_per_wave = 0;   // low-latency kernel (decode_lat.cu): usable for this N; frames it takes at once
    int opt_mode = 0;             // decoder arithmetic: 0 = parity (the reference's), 1 = non-parity "nii" (B200DVB_MODE_NII)
    // device tables
    int16_t *d_tab = nullptr;     // [7][N] int16: perm, inv_perm, offA, offW1, offY1, offW2, offY2
    // host tables
    int32_t next_state[64], out_W[64], out_Y[64];
    int32_t circ_lut[16];
    int16_t *h_tab = nullptr;
};

struct Modem {
    int mod_id = 0, bps = 0, M = 0;
    int separable = 0;            // 1: first half of the label selects I, 2: selects Q, 0: not separable
    int half = 0, nlev = 0;       // bits / levels per axis when separable
    int pwl = 0;                  // piecewise-linear per-axis demapper usable (uniform PAM axes)
    int nseg = 0;                 // segments per axis (2*nlev - 2)
    float pwl_x0[2] = {0, 0}, pwl_invd[2] = {0, 0};   // per axis (0: first half of label, 1: second)
    double *d_table64 = nullptr;  // double2[M]
    float *d_table32 = nullptr;   // float2[M]
    float *d_pwl = nullptr;       // float2[2][half][nseg] (slope, intercept)
    double h_table[512];
};

// launchers (each returns a B200DVB_* code)
int launch_decode(const Codec &c, int B, const float *llr, long long llr_stride, int32_t *bits,
                  uint32_t *packed, const uint8_t *ref_bits, unsigned long long *counters,
                  void *ws, size_t ws_bytes, cudaStream_t s);
int launch_siso(const Codec &c, int B, const float *Lc_A, const float *Lc_B, const float *Lc_W,
                const float *Lc_Y, const double *La_A, const double *La_B, double sf,
                double *Le_A, double *Le_B, void *ws, size_t ws_bytes, cudaStream_t s);
size_t decode_workspace_bytes(const Codec &c, int B);
size_t siso_workspace_bytes(const Codec &c, int B);
int quad_configure(Codec &c);
int tpf_configure(Codec &c);
int tpf_launch_decode(const Codec &c, int B, const float *llr, long long llr_stride, int32_t *bits,
                      uint32_t *packed, const uint8_t *ref_bits, unsigned long long *counters,
                      void *ws, size_t ws_bytes, cudaStream_t s);
size_t tpf_workspace_bytes(const Codec &c, int B);
int tpf_read_phase_cycles(double *out_h, int reset);
int lat_configure(Codec &c);
int lat_launch_decode(const Codec &c, int B, const float *llr, long long llr_stride, int32_t *bits, uint32_t *packed,
                      const uint8_t *ref_bits, unsigned long long *counters, cudaStream_t s);
int nii_configure(Codec &c);
int nii_launch_decode(const Codec &c, int B, const float *llr, long long llr_stride, int32_t *bits,
                      uint32_t *packed, const uint8_t *ref_bits, unsigned long long *counters,
                      void *ws, size_t ws_bytes, cudaStream_t s);
size_t nii_workspace_bytes(const Codec &c, int B);
int nii_read_phase_cycles(double *out_h, int reset);
int lat_read_phase_cycles(double *out_h, int reset);
int read_phase_cycles(double *out_h, int reset);

int launch_encode(const Codec &c, int B, const uint8_t *info, uint8_t *coded, uint8_t *circ,
                  cudaStream_t s);
int launch_mc_bpsk(const Codec &c, int B, float noise_var, unsigned long long seed,
                   unsigned long long frame_offset, uint8_t *info, uint8_t *coded, float *llr,
                   cudaStream_t s);

int launch_awgn_complex(size_t n, float sigma, unsigned long long seed, unsigned long long offset,
                        void *iq, cudaStream_t s);

int launch_map(const Modem &m, size_t n, const uint8_t *bits, void *iq, int out_f64, cudaStream_t s);
int launch_demap(const Modem &m, size_t n, const void *iq, float noise_var, float scale,
                 float *llr, cudaStream_t s);
int launch_demap_bf16(const Modem &m, size_t n, const void *iq, float noise_var, float scale,
                      float *llr, cudaStream_t s);
int launch_hard(const Modem &m, size_t n, const void *iq, int in_f64, uint8_t *bits, cudaStream_t s);

constexpr int kMaxFirTaps = 448;    // FIR taps travel in the kernel parameter space (constant bank): 8 B per tap of 4 KB
int launch_pulse_shape(size_t n_sym, const void *sym, const double *taps_h, int ntaps, int sps, void *out, cudaStream_t s);
int launch_matched_filter(size_t n, const void *x, const double *taps_h, int ntaps, int sps, long long start,
                          size_t n_out, void *out, cudaStream_t s);

void set_mf_variant(int v);
void set_lat_warm(int v);
void set_map_variant(int v);
int modem_build_pwl(Modem &m);
int run_microbench(double *results_h);
int run_microbench2(double *results16_h);
int run_tmem_selftest(int *result_h);

}  // namespace b200dvb
