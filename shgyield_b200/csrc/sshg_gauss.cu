// K1: Gaussian broadening of real signals (sm_100a, FP64) -- shared-memory tiled 1-D FIR.
//
// Replaces broad / broadC of the reference (shgyield/shg.py:26-35), i.e. one axis of
// scipy.ndimage.gaussian_filter: taps w_k = exp(-k^2 / (2 sigma^2)) / sum, radius
// int(4 sigma + 0.5), boundary 'reflect' (d c b a | a b c d | d c b a, repeated when the radius
// exceeds the signal), sigma <= 1e-15 -> copy.
//
// Contiguous signals (the spectra rows and the energy axis of a yield cube): a CTA stages a tile + halo in shared
// memory, each thread then produces a run of consecutive outputs from a register window -- per staged sample one
// conflict-free LDS feeds up to 13 DFMAs.
//   gauss_fixed_bulk_kernel<RAD>   every radius up to 64: taps in the constant bank, straight-line window, tile + halo
//                                  and results moved by the bulk-copy engine (cp.async.bulk, one thread issues);
//                                  FP64-pipe bound from sigma = 5, HBM bound below
//   gauss_fixed_kernel<RAD>        the same with per-thread staging (rows that are not 16-byte aligned / of odd length)
//   gauss_stream_kernel            any radius: taps as a rotating register window fed from shared memory
// Non-contiguous axes of an N-D array (the reference's broad() filters EVERY axis of a broadcast result):
//   gauss_column_kernel            [positions along the axis + halo] x [32 neighbouring signals] per CTA, the same
//                                  register-window FIR down the columns
//   gauss_tiled_kernel / gauss_direct_kernel   narrow trailing axes, very short axes, huge radii (SciPy's accumulation
//                                  order: centre tap, then pairs from the farthest inwards)
//
// HBM traffic: 8 B read + 8 B written per sample (halo re-reads hit L2).
#include "sshg_internal.h"
#include "sshg_fir.cuh"

#include <math.h>
#include <stdlib.h>
#include <type_traits>

namespace {

__device__ __forceinline__ long long reflect(long long i, long long n) {
    long long m = i % (2 * n);
    if (m < 0) m += 2 * n;
    return m >= n ? 2 * n - 1 - m : m;
}

// ---- contiguous rows: register-window FIR (window arithmetic in sshg_fir.cuh) ------------------
constexpr int ST = 128;             // threads per CTA
constexpr int SR = sshg::SR;        // consecutive outputs per thread
constexpr int STILE = ST * SR;      // outputs per CTA

// smem: xs[STILE + 2r + SR] samples (zero padded), then wext[2r + 1 + 3 SR] taps:
// wext[k] = w_full[k - (SR-1)] for 0 <= k - (SR-1) <= 2r, else 0.
// Taps that do not exist are skipped rather than multiplied by zero, so that an Inf/NaN sample
// never reaches outputs beyond the true radius (the reference's NaN pattern).
__global__ void __launch_bounds__(ST) gauss_stream_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                          const double* __restrict__ taps, long long n,
                                                          long long row_stride, int r, int tiles_per_row) {
    extern __shared__ double sm[];
    const int span = STILE + 2 * r + SR;           // staged samples (incl. zero padding)
    const int used = STILE + 2 * r;                // samples that feed an output of this tile
    const int ntap = 2 * r + 1 + 3 * SR;
    double* xs = sm;
    double* wext = sm + span;
    const int tid = threadIdx.x;
    const long long row = blockIdx.x / tiles_per_row;
    const long long i0 = (long long)(blockIdx.x % tiles_per_row) * STILE;
    const double* src = in + row * row_stride;
    double* dst = out + row * row_stride;
    const bool interior = (i0 - r >= 0) && (i0 + STILE + r <= n);
    constexpr int NB = 10;                         // loads in flight per thread (one batch covers radius <= 59)
    for (int k0 = 0; k0 < span; k0 += ST * NB) {
        double v[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int k = k0 + b * ST + tid;
            v[b] = 0.0;
            if (k < used) {
                const long long i = i0 - r + k;
                if (interior || (i >= 0 && i < n)) v[b] = src[i];
                else if (i < n + r) v[b] = src[reflect(i, n)];
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int k = k0 + b * ST + tid;
            if (k < span) xs[k] = v[b];
        }
    }
    for (int k = tid; k < ntap; k += ST) {
        const int j = k - (SR - 1);                // index into the full symmetric kernel
        wext[k] = (j >= 0 && j <= 2 * r) ? taps[j <= r ? r - j : j - r] : 0.0;
    }
    __syncthreads();

    const double* xw = xs + tid * SR;              // window of this thread: xw[t], t = 0 .. 2r + SR - 1
    double acc[SR], wr[SR];
#pragma unroll
    for (int q = 0; q < SR; ++q) {
        acc[q] = 0.0;
        wr[q] = wext[q];
    }
    sshg::fir_window(xw, wext, r, acc, wr);
    __syncthreads();                               // every window read is done: reuse xs for the results
#pragma unroll
    for (int q = 0; q < SR; ++q) xs[tid * SR + q] = acc[q];
    __syncthreads();
    if (i0 + STILE <= n) {
#pragma unroll
        for (int k = 0; k < SR; ++k) dst[i0 + k * ST + tid] = xs[k * ST + tid];
    } else {
        for (int k = tid; k < STILE; k += ST)
            if (i0 + k < n) dst[i0 + k] = xs[k];
    }
}

// ---- contiguous rows, common radii: taps as kernel parameters ----------------------------------
// For the radii that sigma = 1, 1.5, 2, ... produce (even radii 4 .. 24, 28, 32, 40; the reference's
// default sigma_out = 5 gives 20) the radius is a template parameter and the RAD + 1 distinct taps
// travel as a by-value kernel argument, i.e. in the constant bank: every tap is an immediate
// constant-bank operand of its DFMA, no tap is loaded or held in a register, and the whole window
// (2 RAD + FR steps x up to FR outputs) is straight-line code.  That removes a quarter of the
// shared-memory wavefronts of the generic kernel and frees the registers for FR = 13 outputs per
// thread, which moves the kernel from shared-memory bound to FP64 bound.
constexpr int FR = 13;              // outputs per thread (odd: conflict-free stride)
constexpr int FT_MAX = 256;         // threads per CTA: 64 .. 256, chosen per launch (smallest ragged remainder)

template <int RAD>
struct FixedTaps {
    double w[RAD + 1];              // w[0] centre
};

template <int RAD>
__global__ void __launch_bounds__(FT_MAX) gauss_fixed_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                             const FixedTaps<RAD> taps, long long n,
                                                             long long row_stride, int tiles_per_row) {
    extern __shared__ double sm[];
    const int FTN = blockDim.x;
    const int tile_n = FTN * FR;
    const int used = tile_n + 2 * RAD;             // samples that feed an output of this tile
    double* xs = sm;
    const int tid = threadIdx.x;
    const long long row = blockIdx.x / tiles_per_row;
    const long long i0 = (long long)(blockIdx.x % tiles_per_row) * tile_n;
    const double* src = in + row * row_stride;
    double* dst = out + row * row_stride;
    const bool interior = (i0 - RAD >= 0) && (i0 + tile_n + RAD <= n);
    constexpr int NB = 8;                          // loads in flight per thread
    for (int k0 = 0; k0 < used; k0 += FTN * NB) {
        double v[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int k = k0 + b * FTN + tid;
            v[b] = 0.0;
            if (k < used) {
                const long long i = i0 - RAD + k;
                if (interior || (i >= 0 && i < n)) v[b] = src[i];
                else if (i < n + RAD) v[b] = src[reflect(i, n)];
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int k = k0 + b * FTN + tid;
            if (k < used) xs[k] = v[b];
        }
    }
    __syncthreads();

    const double* xw = xs + tid * FR;              // xw[t] = x[o0 - RAD + t], o0 = first output of this thread
    double acc[FR];
#pragma unroll
    for (int q = 0; q < FR; ++q) acc[q] = 0.0;
#pragma unroll
    for (int t = 0; t < 2 * RAD + FR; ++t) {
        const double v = xw[t];
#pragma unroll
        for (int q = 0; q < FR; ++q) {
            const int j = t - q;                   // tap index 0 .. 2 RAD of the full kernel; absent taps are skipped
            if (j >= 0 && j <= 2 * RAD) acc[q] = fma(taps.w[j <= RAD ? RAD - j : j - RAD], v, acc[q]);
        }
    }
    __syncthreads();                               // every window read is done: reuse xs for the results
#pragma unroll
    for (int q = 0; q < FR; ++q) xs[tid * FR + q] = acc[q];
    __syncthreads();
    if (i0 + tile_n <= n) {
#pragma unroll
        for (int k = 0; k < FR; ++k) dst[i0 + k * FTN + tid] = xs[k * FTN + tid];
    } else {
        for (int k = tid; k < tile_n; k += FTN)
            if (i0 + k < n) dst[i0 + k] = xs[k];
    }
}

// ---- the same with the tile moved by the bulk-copy engine (TMA, cp.async.bulk) ---------------------
// ncu on the kernel above (profiles/r01_k1d_fixed_radius_ncu_full.txt): 1225 warp instructions per thread of which 533
// are the DFMAs -- the staging loop (64-bit index arithmetic, bounds and reflect tests, LDG + STS per sample) and the
// store loop (LDS + STG per sample) take 40 % of the issue slots of a kernel whose FP64 pipe already needs 50 % of
// them.  Here ONE thread issues one cp.async.bulk global -> shared for the whole tile + halo (completion on an
// mbarrier) and one cp.async.bulk shared -> global for the results; the reflected samples at the two ends of a row are
// filled from the staged ones.  Needs 16-byte aligned rows of even length (decided on the host); other launches take the
// kernel above (the radii of sigma = 1, 1.5, 2, ...) or the generic kernel.  Instantiated for EVERY radius up to 64: an odd
// radius stages one more sample on each side so that the copy stays 16-byte aligned.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RAD>
__global__ void __launch_bounds__(FT_MAX) gauss_fixed_bulk_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                                  const FixedTaps<RAD> taps, long long n,
                                                                  long long row_stride, int tiles_per_row) {
    extern __shared__ __align__(128) double smb[];
    __shared__ __align__(8) unsigned long long bar;
    const int FTN = blockDim.x;
    const int tile_n = FTN * FR;
    constexpr int PAD = RAD & 1;                   // odd radii: one more sample on each side keeps the copy 16-byte aligned
    double* xs = smb;                              // xs[j] = x[lo + j], j < tile_n + 2 (RAD + PAD)
    const int tid = threadIdx.x;
    const long long row = blockIdx.x / tiles_per_row;
    const long long i0 = (long long)(blockIdx.x % tiles_per_row) * tile_n;
    const double* src = in + row * row_stride;
    double* dst = out + row * row_stride;
    const long long lo = i0 - RAD - PAD;
    const long long g0 = lo < 0 ? 0 : lo;                                   // samples [g0, g1) exist
    const long long g1 = i0 + tile_n + RAD + PAD < n ? i0 + tile_n + RAD + PAD : n;
    const unsigned bar_a = smem_addr(&bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned bytes = (unsigned)(g1 - g0) * 8u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(xs + (g0 - lo))),
                     "l"(src + g0), "r"(bytes), "r"(bar_a)
                     : "memory");
    }
    __syncthreads();                               // the barrier is initialised before anybody polls it
    {
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar_a)
                : "memory");
        }
    }
    // 'reflect' boundary (d c b a | a b c d | d c b a) from the staged samples; the host guarantees n >= 2 RAD
    const bool left = i0 == 0, right = i0 + tile_n + RAD > n;
    if (left || right) {
        for (int k = tid; k < RAD; k += FTN) {
            if (left) xs[RAD + PAD - 1 - k] = xs[RAD + PAD + k];            // x[-1 - k] = x[k]
            const long long i = n + k;                                      // x[n + k] = x[n - 1 - k]
            if (right && i < i0 + tile_n + RAD) xs[i - lo] = xs[2 * n - 1 - i - lo];
        }
        __syncthreads();
    }

    const double* xw = xs + PAD + tid * FR;        // xw[t] = x[o0 - RAD + t], o0 = first output of this thread
    double acc[FR];
#pragma unroll
    for (int q = 0; q < FR; ++q) acc[q] = 0.0;
#pragma unroll
    for (int t = 0; t < 2 * RAD + FR; ++t) {
        const double v = xw[t];
#pragma unroll
        for (int q = 0; q < FR; ++q) {
            const int j = t - q;                   // tap index 0 .. 2 RAD of the full kernel; absent taps are skipped
            if (j >= 0 && j <= 2 * RAD) acc[q] = fma(taps.w[j <= RAD ? RAD - j : j - RAD], v, acc[q]);
        }
    }
    __syncthreads();                               // every window read is done: reuse xs for the results
#pragma unroll
    for (int q = 0; q < FR; ++q) xs[tid * FR + q] = acc[q];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // results visible to the copy engine
    __syncthreads();
    if (tid == 0) {
        const long long valid = n - i0 < tile_n ? n - i0 : tile_n;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + i0), "r"(smem_addr(xs)),
                     "r"((unsigned)valid * 8u)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // shared memory is read before the CTA retires
    }
}

// Returns 1 when this launch cannot take the fixed-radius kernels (the caller falls through to the generic kernel).
// PLAIN: the per-thread-staging kernel exists for this radius as well (the radii of sigma = 1, 1.5, 2, ...); the other
// radii up to 64 have the bulk-copy kernel only.
template <int RAD, bool PLAIN>
int launch_fixed(sshg_ctx* ctx, const double* din, double* dout, int64_t rows, int64_t n, int64_t row_stride,
                 const double* w_host) {
    FixedTaps<RAD> taps;
    for (int k = 0; k <= RAD; ++k) taps.w[k] = w_host[k];
    // CTA width: the multiple of 32 in [64, 256] that leaves the smallest ragged remainder in the last tile
    int ft = FT_MAX;
    long long best = -1;
    for (int t = 64; t <= FT_MAX; t += 32) {
        const long long tile = (long long)t * FR;
        const long long waste = (n + tile - 1) / tile * tile - n;
        if (best < 0 || waste <= best) {
            best = waste;
            ft = t;
        }
    }
    const long long tile = (long long)ft * FR;
    const long long tiles = (n + tile - 1) / tile;
    SSHG_REQUIRE(rows * tiles < (1LL << 31), "sshg_gauss1d: too many tiles");
    // bulk-copy staging: 16-byte aligned rows of even length (tile widths are even), simple reflection
    const bool bulk = ctx->opt.k1_bulk != 0 && ((uintptr_t)din | (uintptr_t)dout) % 16 == 0 && n % 2 == 0 &&
                      (rows == 1 || row_stride % 2 == 0) && n >= 2 * RAD;
    if (bulk) {
        const size_t smem = (size_t)(tile + 2 * (RAD + (RAD & 1))) * 8;
        SSHG_CUDA(cudaFuncSetAttribute(gauss_fixed_bulk_kernel<RAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        gauss_fixed_bulk_kernel<RAD><<<(unsigned)(rows * tiles), ft, smem, ctx->stream>>>(din, dout, taps, n, row_stride, (int)tiles);
    } else if constexpr (PLAIN) {
        const size_t smem = (size_t)(tile + 2 * RAD) * 8;
        SSHG_CUDA(cudaFuncSetAttribute(gauss_fixed_kernel<RAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        gauss_fixed_kernel<RAD><<<(unsigned)(rows * tiles), ft, smem, ctx->stream>>>(din, dout, taps, n, row_stride, (int)tiles);
    } else {
        return 1;
    }
    SSHG_CUDA(cudaGetLastError());
    ctx->launches++;
    return SSHG_OK;
}

// ---- strided rows: SciPy-order tiled kernel -------------------------------------------------------
constexpr int GT = 256;          // threads per CTA
constexpr int GPT = 4;           // outputs per thread
constexpr int GTILE = GT * GPT;  // outputs per CTA

// smem: [GTILE + 2r] samples, then [r + 1] taps (w[0] = centre)
// Row `row` starts at (row / inner_rows) * outer_stride + (row % inner_rows) * row_stride: filtering a middle axis of an
// N-D array (rows = outer x inner signals) is one launch; plain row lists pass inner_rows = rows.
__device__ __forceinline__ long long row_start(long long row, long long inner_rows, long long outer_stride, long long row_stride) {
    return (row / inner_rows) * outer_stride + (row % inner_rows) * row_stride;
}

__global__ void __launch_bounds__(GT) gauss_tiled_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                         const double* __restrict__ taps, long long rows,
                                                         long long n, long long row_stride, long long elem_stride,
                                                         long long inner_rows, long long outer_stride,
                                                         int r, int tiles_per_row) {
    extern __shared__ double sm[];
    double* xs = sm;
    double* ws = sm + GTILE + 2 * r;
    const long long row = blockIdx.x / tiles_per_row;
    const long long i0 = (long long)(blockIdx.x % tiles_per_row) * GTILE;
    const long long start = row_start(row, inner_rows, outer_stride, row_stride);
    const double* src = in + start;
    double* dst = out + start;
    const int span = GTILE + 2 * r;
    for (int k = threadIdx.x; k < span; k += GT) {
        long long i = i0 - r + k;
        double v = 0.0;                        // samples beyond n - 1 + r feed no valid output
        if (i >= 0 && i < n) v = src[i * elem_stride];
        else if (i < n + r) v = src[reflect(i, n) * elem_stride];
        xs[k] = v;
    }
    for (int k = threadIdx.x; k <= r; k += GT) ws[k] = taps[k];
    __syncthreads();
    double acc[GPT];
#pragma unroll
    for (int q = 0; q < GPT; ++q) acc[q] = xs[r + threadIdx.x + q * GT] * ws[0];
    for (int j = r; j >= 1; --j) {
        const double w = ws[j];
#pragma unroll
        for (int q = 0; q < GPT; ++q) {
            const int c = r + threadIdx.x + q * GT;
            acc[q] = fma(xs[c - j] + xs[c + j], w, acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < GPT; ++q) {
        const long long i = i0 + threadIdx.x + q * GT;
        if (i < n) dst[i * elem_stride] = acc[q];
    }
}

// ---- a non-contiguous axis of an N-D array: column kernel ------------------------------------------------------------
// broad() of an N-D result filters EVERY axis (scipy.ndimage.gaussian_filter; the reference's shgyield() does that to
// broadcast results, shg.py:433-436).  Along an axis that is not the last one, the elements of a signal are `inner`
// doubles apart, but `inner` neighbouring signals are contiguous: a CTA stages a tile [positions along the axis + halo] x
// [32 neighbouring signals] (coalesced 256-byte rows), lane = signal, warp = a group of SR consecutive outputs along the
// axis, and every thread runs the register-window FIR of the contiguous kernels down its column of the tile
// (conflict-free: the lanes of a warp read neighbouring words).  The strided kernel above stages 1024 + 2 r samples per
// CTA with `inner`-strided loads (one 32-byte sector per sample) and fills its CTAs badly on short axes (90 thetas).
constexpr int CW = 32;              // signals per CTA (one per lane)
constexpr int CWARPS = 8;           // warps per CTA
constexpr int CG = CWARPS;          // output groups per CTA (one per warp): CG * SR = 72 outputs along the axis.  (Two
                                    // groups per warp halve the share of the halo rows but cost registers and resident
                                    // CTAs and measured 30 % slower.)

// BULK: the rows of a tile (256 contiguous bytes each, `inner` doubles apart) are fetched by the bulk-copy engine, one
// cp.async.bulk per row issued by one thread each, all in flight at once, completion on an mbarrier -- instead of a
// load / store pair per element through the registers of a warp that handles its rows four at a time.  Needs 16-byte
// aligned rows (even `inner`, aligned base); tiles at the ragged end of the signals take the per-thread loads.
template <bool BULK>
__global__ void __launch_bounds__(CW * CWARPS) gauss_column_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                               const double* __restrict__ taps, long long n, long long inner,
                                                               int r, int tiles_a, int tiles_c) {
    extern __shared__ __align__(128) double sm[];
    __shared__ __align__(8) unsigned long long bar;
    constexpr int TA = CG * SR;                    // outputs along the axis per CTA
    const int span_max = TA + 2 * r + SR;          // row count the shared-memory layout is planned for
    const int ntap = 2 * r + 1 + 3 * SR;
    double* xs = sm;                               // [span][CW]
    double* wext = sm + (size_t)span_max * CW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned blk = blockIdx.x;               // 32-bit index arithmetic: a 64-bit division costs ~100 instructions
    const unsigned tc = blk % (unsigned)tiles_c, rest = blk / (unsigned)tiles_c;
    const unsigned ta = rest % (unsigned)tiles_a, o = rest / (unsigned)tiles_a;
    const long long a0 = (long long)ta * TA, col = (long long)tc * CW + lane;
    // output groups of this tile that exist (a short axis, or the last tile of a long one): the others stage and compute nothing
    const int groups = (int)((((n - a0 < TA) ? n - a0 : TA) + SR - 1) / SR);
    const int used = groups * SR + 2 * r;          // staged positions that feed an output of this tile
    const int span = used + SR;                    // incl. zero padding (the window of the last output group over-reads)
    const bool colok = col < inner;
    const double* src = in + (long long)o * n * inner;
    double* dst = out + (long long)o * n * inner;
    const bool full_tile = (long long)tc * CW + CW <= inner;       // CTA-uniform
    if (BULK && full_tile) {
        // rows that exist: j < used and a0 - r + j < n + r
        const long long lim = n + r - (a0 - r);
        const int rows_in = (int)(lim < used ? lim : used);
        const unsigned bar_a = smem_addr(&bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((unsigned)rows_in * CW * 8u) : "memory");
        }
        __syncthreads();
        for (int j = threadIdx.x; j < span; j += CW * CWARPS) {
            if (j < rows_in) {
                const long long a = a0 - r + j;
                const long long ra = (a >= 0 && a < n) ? a : reflect(a, n);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_addr(xs + j * CW)),
                             "l"(src + ra * inner + (long long)tc * CW), "r"((unsigned)(CW * 8)), "r"(bar_a)
                             : "memory");
            } else {                               // positions beyond n - 1 + r feed no valid output
#pragma unroll 4
                for (int c = 0; c < CW; ++c) xs[j * CW + c] = 0.0;
            }
        }
    } else {
        for (int j = warp; j < span; j += CWARPS) {
            const long long a = a0 - r + j;
            double v = 0.0;                        // positions beyond n - 1 + r feed no valid output
            if (j < used && colok && a < n + r) v = src[((a >= 0 && a < n) ? a : reflect(a, n)) * inner + col];
            xs[j * CW + lane] = v;
        }
    }
    for (int k = threadIdx.x; k < ntap; k += CW * CWARPS) {
        const int j = k - (SR - 1);                // index into the full symmetric kernel
        wext[k] = (j >= 0 && j <= 2 * r) ? taps[j <= r ? r - j : j - r] : 0.0;
    }
    __syncthreads();
    if (BULK && full_tile) {
        const unsigned bar_a = smem_addr(&bar);
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar_a)
                : "memory");
        }
    }
    if (warp >= groups) return;
    double acc[SR], wr[SR];
#pragma unroll
    for (int q = 0; q < SR; ++q) {
        acc[q] = 0.0;
        wr[q] = wext[q];
    }
    sshg::fir_window<CW>(xs + (warp * SR) * CW + lane, wext, r, acc, wr);
    if (colok) {
#pragma unroll
        for (int q = 0; q < SR; ++q) {
            const long long a = a0 + warp * SR + q;
            if (a < n) dst[a * inner + col] = acc[q];
        }
    }
}

// Fallback for radii whose halo does not fit in shared memory.
__global__ void gauss_direct_kernel(const double* __restrict__ in, double* __restrict__ out,
                                    const double* __restrict__ taps, long long rows, long long n,
                                    long long row_stride, long long elem_stride, long long inner_rows,
                                    long long outer_stride, long long r) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= rows * n) return;
    const long long row = g / n, i = g % n;
    const long long start = row_start(row, inner_rows, outer_stride, row_stride);
    const double* src = in + start;
    double acc = src[i * elem_stride] * taps[0];
    for (long long j = r; j >= 1; --j)
        acc = fma(src[reflect(i - j, n) * elem_stride] + src[reflect(i + j, n) * elem_stride], taps[j], acc);
    out[start + i * elem_stride] = acc;
}

__global__ void copy_strided_kernel(const double* __restrict__ in, double* __restrict__ out, long long rows,
                                    long long n, long long row_stride, long long elem_stride, long long inner_rows,
                                    long long outer_stride) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= rows * n) return;
    const long long off = row_start(g / n, inner_rows, outer_stride, row_stride) + (g % n) * elem_stride;
    out[off] = in[off];
}

constexpr size_t SMEM_CAP = 200 * 1024;

}  // namespace

// scipy _gaussian_kernel1d: exp(-0.5 / sigma^2 * x^2), normalised by the sum over -r..r
void sshg_gaussian_taps(double sigma, long long r, double* w) {
    const double f = -0.5 / (sigma * sigma);
    double sum = 0.0;
    for (long long k = r; k >= 1; --k) {
        w[k] = exp(f * (double)(k * k));
        sum += 2.0 * w[k];
    }
    w[0] = 1.0;
    sum += 1.0;
    for (long long k = 0; k <= r; ++k) w[k] /= sum;
}

// Device-level filter: din / dout are device pointers; queued on ctx->stream.
int sshg_gauss_device(sshg_ctx* ctx, const double* din, double* dout, int64_t rows, int64_t n, int64_t row_stride,
                      int64_t elem_stride, double sigma) {
    return sshg_gauss_device_ex(ctx, din, dout, rows, n, row_stride, elem_stride, rows, 0, sigma);
}

// The same with the row addressing of an N-D axis: row k starts at (k / inner_rows) * outer_stride +
// (k % inner_rows) * row_stride.
int sshg_gauss_device_ex(sshg_ctx* ctx, const double* din, double* dout, int64_t rows, int64_t n, int64_t row_stride,
                         int64_t elem_stride, int64_t inner_rows, int64_t outer_stride, double sigma) {
    const long long total = (long long)rows * n;
    const bool plain = inner_rows >= rows;                  // rows are row_stride apart: the contiguous kernels apply
    if (!(sigma > 1e-15)) {
        copy_strided_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(din, dout, rows, n, row_stride,
                                                                                   elem_stride, inner_rows, outer_stride);
        SSHG_CUDA(cudaGetLastError());
        ctx->launches++;
        return SSHG_OK;
    }
    const long long r = (long long)(4.0 * sigma + 0.5);
    SSHG_REQUIRE(r < (1LL << 28), "sshg_gauss1d: sigma %g is too large", sigma);
    void* dw;
    if (int rc = sshg_scratch(ctx, SCR_TAPS, (size_t)(r + 1) * 8, &dw)) return rc;
    if (!(ctx->tap_sigma == sigma && ctx->tap_ptr == dw)) {
        double* w = (double*)malloc((size_t)(r + 1) * 8);
        if (!w) {
            sshg_set_error("out of host memory");
            return SSHG_ENOMEM;
        }
        sshg_gaussian_taps(sigma, r, w);
        for (long long k = 0; k <= r && k < SSHG_TAP_HOST; ++k) ctx->tap_host[k] = w[k];
        // pageable source: the runtime copies it to a staging buffer before returning
        cudaError_t e = cudaMemcpyAsync(dw, w, (size_t)(r + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
        free(w);
        if (e != cudaSuccess) return sshg_cuda_fail(e, "tap upload", __FILE__, __LINE__);
        ctx->tap_sigma = sigma;
        ctx->tap_ptr = dw;
    }
    if (elem_stride == 1 && plain && ctx->opt.k1_generic != 1) {
        int rc = 1;                                // 1: not taken
        switch (r) {                               // radii with a compile-time specialisation
#define SSHG_FIXED(R_) case R_: rc = launch_fixed<R_, true>(ctx, din, dout, rows, n, row_stride, ctx->tap_host); break;
#define SSHG_BULK(R_) case R_: rc = launch_fixed<R_, false>(ctx, din, dout, rows, n, row_stride, ctx->tap_host); break;
            SSHG_FIXED(4) SSHG_FIXED(6) SSHG_FIXED(8) SSHG_FIXED(10) SSHG_FIXED(12) SSHG_FIXED(14) SSHG_FIXED(16)
            SSHG_FIXED(18) SSHG_FIXED(20) SSHG_FIXED(22) SSHG_FIXED(24) SSHG_FIXED(26) SSHG_FIXED(28) SSHG_FIXED(30)
            SSHG_FIXED(32) SSHG_FIXED(36) SSHG_FIXED(40) SSHG_FIXED(48) SSHG_FIXED(56) SSHG_FIXED(64)
            // every other radius up to 64 (sigma <= 16): bulk-copy staging only
            SSHG_BULK(1) SSHG_BULK(2) SSHG_BULK(3) SSHG_BULK(5) SSHG_BULK(7) SSHG_BULK(9) SSHG_BULK(11) SSHG_BULK(13)
            SSHG_BULK(15) SSHG_BULK(17) SSHG_BULK(19) SSHG_BULK(21) SSHG_BULK(23) SSHG_BULK(25) SSHG_BULK(27) SSHG_BULK(29)
            SSHG_BULK(31) SSHG_BULK(33) SSHG_BULK(34) SSHG_BULK(35) SSHG_BULK(37) SSHG_BULK(38) SSHG_BULK(39) SSHG_BULK(41)
            SSHG_BULK(42) SSHG_BULK(43) SSHG_BULK(44) SSHG_BULK(45) SSHG_BULK(46) SSHG_BULK(47) SSHG_BULK(49) SSHG_BULK(50)
            SSHG_BULK(51) SSHG_BULK(52) SSHG_BULK(53) SSHG_BULK(54) SSHG_BULK(55) SSHG_BULK(57) SSHG_BULK(58) SSHG_BULK(59)
            SSHG_BULK(60) SSHG_BULK(61) SSHG_BULK(62) SSHG_BULK(63)
#undef SSHG_FIXED
#undef SSHG_BULK
            default: break;
        }
        if (rc != 1) return rc;
    }
    // a non-contiguous axis of a C-contiguous N-D array (neighbouring signals contiguous): column kernel
    const size_t smem_col = ((size_t)(CG * SR + 2 * r + SR) * CW + (size_t)(2 * r + 1 + 3 * SR)) * 8;
    // (rows k = o * inner + c at o * n * inner + c, elements `inner` apart; a single block of interleaved rows -- the
    // FIRST axis of an array -- arrives as rows == inner_rows)
    const long long inner_c = inner_rows < rows ? inner_rows : rows;
    if (row_stride == 1 && elem_stride == inner_c && elem_stride > 1 && (rows == inner_c || outer_stride == n * inner_c) &&
        inner_c >= 8 && rows % inner_c == 0 && smem_col <= SMEM_CAP) {
        const long long outer = rows / inner_c;
        const long long inner_rows = inner_c;
        const long long tiles_a = (n + CG * SR - 1) / (CG * SR), tiles_c = (inner_rows + CW - 1) / CW;
        if (outer * tiles_a * tiles_c < (1LL << 31)) {
            const bool bulk = ctx->opt.k1_bulk != 0 && inner_rows % 2 == 0 && (uintptr_t)din % 16 == 0;
            if (bulk) {
                SSHG_CUDA(cudaFuncSetAttribute(gauss_column_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP));
                gauss_column_kernel<true><<<(unsigned)(outer * tiles_a * tiles_c), CW * CWARPS, smem_col, ctx->stream>>>(
                    din, dout, (const double*)dw, n, inner_rows, (int)r, (int)tiles_a, (int)tiles_c);
            } else {
                SSHG_CUDA(cudaFuncSetAttribute(gauss_column_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP));
                gauss_column_kernel<false><<<(unsigned)(outer * tiles_a * tiles_c), CW * CWARPS, smem_col, ctx->stream>>>(
                    din, dout, (const double*)dw, n, inner_rows, (int)r, (int)tiles_a, (int)tiles_c);
            }
            SSHG_CUDA(cudaGetLastError());
            ctx->launches++;
            return SSHG_OK;
        }
    }
    const size_t smem_stream = (size_t)(STILE + 2 * r + SR + 2 * r + 1 + 3 * SR) * 8;
    const size_t smem_tiled = (size_t)(GTILE + 2 * r + r + 1) * 8;
    if (elem_stride == 1 && plain && smem_stream <= SMEM_CAP) {
        const int tiles = (int)((n + STILE - 1) / STILE);
        SSHG_REQUIRE(rows * tiles < (1LL << 31), "sshg_gauss1d: too many tiles");
        SSHG_CUDA(cudaFuncSetAttribute(gauss_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP));
        gauss_stream_kernel<<<(unsigned)(rows * tiles), ST, smem_stream, ctx->stream>>>(din, dout, (const double*)dw, n,
                                                                                      row_stride, (int)r, tiles);
    } else if (smem_tiled <= SMEM_CAP && n >= 64) {
        // (short strided axes -- the four-entry axis of a [4, P, E] stack, an axis of extent 1 of a broadcast result -- would
        // make one CTA stage 1024 + 2 r samples for a handful of outputs per row: they take the direct kernel below)
        const int tiles = (int)((n + GTILE - 1) / GTILE);
        SSHG_REQUIRE(rows * tiles < (1LL << 31), "sshg_gauss1d: too many tiles");
        SSHG_CUDA(cudaFuncSetAttribute(gauss_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP));
        gauss_tiled_kernel<<<(unsigned)(rows * tiles), GT, smem_tiled, ctx->stream>>>(din, dout, (const double*)dw, rows, n,
                                                                                    row_stride, elem_stride, inner_rows,
                                                                                    outer_stride, (int)r, tiles);
    } else {
        gauss_direct_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(din, dout, (const double*)dw, rows, n,
                                                                                   row_stride, elem_stride, inner_rows,
                                                                                   outer_stride, r);
    }
    SSHG_CUDA(cudaGetLastError());
    ctx->launches++;
    return SSHG_OK;
}

extern "C" int sshg_gauss1d(sshg_ctx* ctx, const double* in, double* out, int64_t rows, int64_t n,
                            int64_t row_stride, int64_t elem_stride, double sigma, int where) {
    SSHG_TRACE("sshg_gauss1d");
    SSHG_REQUIRE(ctx && in && out, "sshg_gauss1d: null argument");
    SSHG_REQUIRE(in != out, "sshg_gauss1d: in-place filtering is not supported");
    SSHG_REQUIRE(rows > 0 && n > 0, "sshg_gauss1d: rows and n must be positive");
    SSHG_REQUIRE(elem_stride >= 1 && row_stride >= 1, "sshg_gauss1d: strides must be positive");
    SSHG_REQUIRE(sigma >= 0.0 && isfinite(sigma), "sshg_gauss1d: sigma must be finite and >= 0 (got %g)", sigma);
    SSHG_CUDA(cudaSetDevice(ctx->device));
    const int in_dev = (where & SSHG_IN_DEVICE) != 0, out_dev = (where & SSHG_OUT_DEVICE) != 0;
    // extent of the block of memory the rows live in (rows interleave when a non-last axis is filtered)
    const int64_t extent = (rows - 1) * row_stride + (n - 1) * elem_stride + 1;
    const size_t bytes = (size_t)extent * 8;
    const void* din = in;
    if (int rc = sshg_stage_in(ctx, 10, in, bytes, in_dev, &din)) return rc;
    void* dout = out;
    if (!out_dev) {
        if (int rc = sshg_scratch(ctx, 11, bytes, &dout)) return rc;
    }
    if (int rc = sshg_gauss_device(ctx, (const double*)din, (double*)dout, rows, n, row_stride, elem_stride, sigma)) return rc;
    if (!out_dev) return sshg_copy_out(ctx, out, dout, bytes);      // large pageable results: threaded staged copy
    return SSHG_OK;
}

// Every axis of a C-contiguous N-D array, as scipy.ndimage.gaussian_filter(data, sigma) does it (broad(), shg.py:26-28,
// on an N-D result): one pass per axis, ping-pong between `out` and one scratch buffer, the array crosses PCIe once
// each way.  `lead` leading axes are NOT filtered (the four polarisations of a yield stack).  All on the device.
int sshg_gauss_nd_device(sshg_ctx* ctx, const double* din, double* dout, int lead, int ndim, const int64_t* shape,
                         double sigma) {
    int64_t total = 1;
    for (int a = 0; a < ndim; ++a) total *= shape[a];
    const int passes = ndim - lead;
    if (passes <= 0 || !(sigma > 1e-15) || total == 0) {
        if (total) SSHG_CUDA(cudaMemcpyAsync(dout, din, (size_t)total * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        return SSHG_OK;
    }
    void* tmp = nullptr;
    if (passes > 1) {
        if (int rc = sshg_scratch(ctx, SCR_ND, (size_t)total * 8, &tmp)) return rc;
    }
    // the last pass must land in dout: passes alternate dout / tmp backwards from there
    const double* src = din;
    for (int k = 0; k < passes; ++k) {
        const int a = lead + k;
        double* dst = ((passes - 1 - k) % 2 == 0) ? dout : (double*)tmp;
        int64_t outer = 1, inner = 1;
        for (int j = 0; j < a; ++j) outer *= shape[j];
        for (int j = a + 1; j < ndim; ++j) inner *= shape[j];
        const int64_t n = shape[a];
        int rc;
        if (inner == 1) rc = sshg_gauss_device(ctx, src, dst, outer, n, n, 1, sigma);
        else rc = sshg_gauss_device_ex(ctx, src, dst, outer * inner, n, 1, inner, inner, n * inner, sigma);
        if (rc) return rc;
        src = dst;
    }
    return SSHG_OK;
}

extern "C" int sshg_gauss_nd(sshg_ctx* ctx, const double* in, double* out, int32_t ndim, const int64_t* shape, double sigma,
                             int where) {
    SSHG_TRACE("sshg_gauss_nd");
    SSHG_REQUIRE(ctx && in && out && (shape || ndim == 0), "sshg_gauss_nd: null argument");
    SSHG_REQUIRE(in != out, "sshg_gauss_nd: in-place filtering is not supported");
    SSHG_REQUIRE(ndim >= 0 && ndim <= 16, "sshg_gauss_nd: ndim must be in 0 .. 16 (got %d)", (int)ndim);
    SSHG_REQUIRE(sigma >= 0.0 && isfinite(sigma), "sshg_gauss_nd: sigma must be finite and >= 0 (got %g)", sigma);
    int64_t total = 1;
    for (int a = 0; a < ndim; ++a) {
        SSHG_REQUIRE(shape[a] > 0, "sshg_gauss_nd: extents must be positive");
        total *= shape[a];
    }
    SSHG_CUDA(cudaSetDevice(ctx->device));
    const int in_dev = (where & SSHG_IN_DEVICE) != 0, out_dev = (where & SSHG_OUT_DEVICE) != 0;
    const size_t bytes = (size_t)total * 8;
    const void* din = in;
    if (int rc = sshg_stage_in(ctx, 10, in, bytes, in_dev, &din)) return rc;
    void* dout = out;
    if (!out_dev) {
        if (int rc = sshg_scratch(ctx, 11, bytes, &dout)) return rc;
    }
    if (int rc = sshg_gauss_nd_device(ctx, (const double*)din, (double*)dout, 0, ndim, shape, sigma)) return rc;
    if (!out_dev) return sshg_copy_out(ctx, out, dout, bytes);      // large pageable results: threaded staged copy
    return SSHG_OK;
}
