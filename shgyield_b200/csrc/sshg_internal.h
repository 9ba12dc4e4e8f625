// Internal declarations shared by the translation units of libsshg.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/sshg.h"

// Tracing: every compute entry point of the C ABI is an NVTX range (header-only NVTX 3; a no-op unless a
// profiler is attached), e.g. `ncu --nvtx --nvtx-include "sshg_yield/" ...` profiles only the yield kernels.
#include <nvtx3/nvToolsExt.h>
struct sshg_trace_range {
    explicit sshg_trace_range(const char* name) { nvtxRangePushA(name); }
    ~sshg_trace_range() { nvtxRangePop(); }
    sshg_trace_range(const sshg_trace_range&) = delete;
    sshg_trace_range& operator=(const sshg_trace_range&) = delete;
};
#define SSHG_TRACE(name) sshg_trace_range sshg_trace_range_(name)

#define SSHG_NSCRATCH 32
#define SSHG_TAP_HOST 72
#define SCR_TAPS 20      // Gaussian taps of the last sigma (cached)
#define SCR_SPLFAC 17    // spline elimination coefficients of the last axis (cached)
#define SCR_XPHI 18      // phi-contracted chi2 of the sparse-phi kernel (slot 19: thickness-axis info)
#define SCR_BROAD 16     // broadened rows handed from K1 to the spline kernels
#define SCR_ND 21        // ping-pong buffer of the N-D Gaussian filter
#define SCR_PTS_IN 22    // small point-list calls: theta | phi | thick | eidx, one upload
#define SCR_PTS_Y 23     // ... un-broadened yields [4][n]
#define SCR_PTS_OUT 24   // ... broadened yields [4][n]
#define SCR_SPLTAB 25    // spline: cyclic-reduction coefficient tables of the cached axis
#define SCR_BCAST 26     // sshg_yield_broadcast: result layout buffers (slots 26, 27) of calls up to 256 MB

// Test / tuning switches of a context (sshg_ctx_set_option, include/sshg.h); -1 = the library decides.
struct sshg_options {
    int k2_sparse;      // 0 / 1: force the dense / sparse-phi yield kernel
    int k2_nt;          // 64 / 256: CTA width of the dense yield kernel
    int k2_recurrence;  // 0: no phase recurrence on uniform thickness axes
    int k3;             // 0: never fuse the output broadening (K2 -> K1 by slabs)
    int k3_nt;          // 64 / 128: CTA width of the fused kernel
    int k3_fixup;       // 0: leave the harmonic sums next to azimuthal nodes as they are (tests of the error bound)
    int nseg;           // phi segments per column (1 .. 16)
    int k1_generic;     // 1: never use the fixed-radius FIR kernels
    int graphs;         // 0: small calls are launched kernel by kernel instead of replaying a CUDA graph
    int quads;          // 0: no four-row items (phi, 180 - phi, phi + 180, 360 - phi) in the azimuthal sweeps
    int k1_bulk;        // 0: the fixed-radius FIR kernels stage their tiles with per-thread loads / stores (no bulk copies)
    int pair_copy;      // 0: host-output sweeps copy every row over PCIe (no host-side duplication of the equal sS rows)
};

// A recorded small call: copy in -> yield_points_kernel -> FIR passes -> copy out, keyed on everything the recorded
// kernel arguments depend on.
#define SSHG_NGRAPH 4
struct sshg_small_key {
    int64_t n, nE;
    int32_t ndim;
    int64_t shape[8];
    double sigma_out;
    sshg_physics phys;
    const void *energy, *eps1w, *eps2w, *chi2;      // resident spectra
    uint64_t epoch;
};
struct sshg_small_graph {
    sshg_small_key key;
    cudaGraphExec_t exec;       // null: slot free
    uint64_t used;              // LRU stamp
    int launches;               // kernels per replay
};

struct sshg_ctx {
    int device;
    sshg_options opt;
    cudaStream_t own_stream;
    cudaStream_t stream;        // own_stream or an adopted caller stream
    cudaStream_t copy_stream;   // D2H of finished slabs while the next slab is computed
    cudaEvent_t ev_done[2];     // slab k computed
    cudaEvent_t ev_copied[2];   // slab k copied to the host
    cudaEvent_t ev_switch;      // orders the scratch buffers across a change of stream
    int num_sms;
    int64_t launches;
    double tap_sigma;           // sigma whose taps are in scratch[SCR_TAPS] (tap_ptr guards reallocation)
    void* tap_ptr;
    double tap_host[SSHG_TAP_HOST];   // host copy of the first taps (kernel arguments of the fixed-radius K1)
    uint64_t spl_hash;          // axis whose spline factorisation is in scratch[SCR_SPLFAC] (0: none)
    int64_t spl_n;
    void* spl_ptr;
    // growable device scratch buffers (inputs copied from the host, staged outputs)
    void* scratch[SSHG_NSCRATCH];
    size_t scratch_bytes[SSHG_NSCRATCH];
    uint64_t epoch;             // bumped whenever a scratch or staging buffer moves: recorded graphs of older epochs are stale
    // small point-list calls (sshg_yield_points_broadened): page-locked staging + recorded CUDA graphs
    void* pin_in;
    void* pin_out;
    size_t pin_in_bytes, pin_out_bytes;
    void* stage[2];             // page-locked staging buffers of the threaded device-to-host copy (sshg_copy_out)
    size_t stage_bytes;
    sshg_small_graph graphs[SSHG_NGRAPH];
    sshg_small_key seen[SSHG_NGRAPH];   // combinations that ran once: a graph is recorded when one of them comes again
    int seen_valid[SSHG_NGRAPH], seen_next;
    uint64_t graph_clock;
    int64_t graph_replays;
};

void sshg_set_error(const char* fmt, ...);
int sshg_cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int sshg_scratch(sshg_ctx* ctx, int slot, size_t bytes, void** out);   // grows if needed

#define SSHG_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) return sshg_cuda_fail(e_, #call, __FILE__, __LINE__);      \
    } while (0)

#define SSHG_REQUIRE(cond, ...)                                                           \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            sshg_set_error(__VA_ARGS__);                                                  \
            return SSHG_EINVAL;                                                           \
        }                                                                                 \
    } while (0)

// Device-level stages (all pointers on the device, queued on ctx->stream)
int sshg_gauss_device(sshg_ctx* ctx, const double* din, double* dout, int64_t rows, int64_t n, int64_t row_stride,
                      int64_t elem_stride, double sigma);
int sshg_gauss_device_ex(sshg_ctx* ctx, const double* din, double* dout, int64_t rows, int64_t n, int64_t row_stride,
                         int64_t elem_stride, int64_t inner_rows, int64_t outer_stride, double sigma);
int sshg_gauss_nd_device(sshg_ctx* ctx, const double* din, double* dout, int lead, int ndim, const int64_t* shape,
                         double sigma);
int sshg_spline_device(sshg_ctx* ctx, const double* dx, const double* dy, int64_t rows, int64_t n, const double* dq,
                       int64_t nq, double* dout, int* dflag, uint64_t axis_hash);

// scipy _gaussian_kernel1d on the host: w[k] = exp(-k^2 / (2 sigma^2)) / sum over -r..r, k = 0..r (w[0] centre)
void sshg_gaussian_taps(double sigma, long long r, double* w);

// Device -> host copy of a large result, blocking, ordered after everything queued on ctx->stream.  Page-locked
// destinations take one cudaMemcpyAsync; pageable ones (a fresh NumPy array) go through two page-locked staging buffers and
// are written by several host threads, so that the page faults of first touch and the copy out of the staging buffer run in
// parallel (the driver's own pageable path is one thread: 3.8 GB/s into fresh memory on the test boxes, this 25-35).
int sshg_copy_out(sshg_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

// Host -> device copy ordered with ctx->stream; large pageable sources go through the page-locked staging buffers, filled
// by several host threads.  The source may be changed as soon as the call returns.
int sshg_copy_in(sshg_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);

// `ncols` blocks of `bytes` bytes, block c from base + c pitch + src_off to base + c pitch + dst_off, by several host threads.
void sshg_parallel_dup(char* base, size_t pitch, size_t src_off, size_t dst_off, size_t bytes, int64_t ncols);

// Copies a host array to a scratch slot (or passes a device pointer through).
int sshg_stage_in(sshg_ctx* ctx, int slot, const void* src, size_t bytes, int on_device,
                  const void** dev);
