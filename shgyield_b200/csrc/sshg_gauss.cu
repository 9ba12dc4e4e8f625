// K1: Gaussian broadening of real signals (sm_100a, FP64) -- a shared-memory tiled 1-D FIR.
//
// Replaces broad / broadC of the reference (shgyield/shg.py:26-35), i.e. one axis of
// scipy.ndimage.gaussian_filter: taps w_k = exp(-k^2 / (2 sigma^2)) / sum, radius
// int(4 sigma + 0.5), boundary 'reflect' (d c b a | a b c d | d c b a, repeated when the radius
// exceeds the signal), sigma <= 1e-15 -> copy.  The accumulation order is SciPy's symmetric
// correlate1d loop: centre tap first, then tap pairs from the farthest inwards.
//
// HBM traffic: 8 B read + 8 B written per sample (halo re-reads hit L2).
#include "sshg_internal.h"

#include <math.h>
#include <stdlib.h>

namespace {

constexpr int GT = 256;          // threads per CTA
constexpr int GPT = 4;           // outputs per thread
constexpr int GTILE = GT * GPT;  // outputs per CTA

__device__ __forceinline__ long long reflect(long long i, long long n) {
    long long m = i % (2 * n);
    if (m < 0) m += 2 * n;
    return m >= n ? 2 * n - 1 - m : m;
}

// smem: [GTILE + 2r] samples, then [r + 1] taps (w[0] = centre)
__global__ void __launch_bounds__(GT) gauss_tiled_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                         const double* __restrict__ taps, long long rows,
                                                         long long n, long long row_stride, long long elem_stride,
                                                         int r, int tiles_per_row) {
    extern __shared__ double sm[];
    double* xs = sm;
    double* ws = sm + GTILE + 2 * r;
    const long long row = blockIdx.x / tiles_per_row;
    const long long i0 = (long long)(blockIdx.x % tiles_per_row) * GTILE;
    const double* src = in + row * row_stride;
    double* dst = out + row * row_stride;
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

// Fallback for radii whose halo does not fit in shared memory.
__global__ void gauss_direct_kernel(const double* __restrict__ in, double* __restrict__ out,
                                    const double* __restrict__ taps, long long rows, long long n,
                                    long long row_stride, long long elem_stride, long long r) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= rows * n) return;
    const long long row = g / n, i = g % n;
    const double* src = in + row * row_stride;
    double acc = src[i * elem_stride] * taps[0];
    for (long long j = r; j >= 1; --j)
        acc = fma(src[reflect(i - j, n) * elem_stride] + src[reflect(i + j, n) * elem_stride], taps[j], acc);
    out[row * row_stride + i * elem_stride] = acc;
}

__global__ void copy_strided_kernel(const double* __restrict__ in, double* __restrict__ out, long long rows,
                                    long long n, long long row_stride, long long elem_stride) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= rows * n) return;
    const long long off = (g / n) * row_stride + (g % n) * elem_stride;
    out[off] = in[off];
}

}  // namespace

extern "C" int sshg_gauss1d(sshg_ctx* ctx, const double* in, double* out, int64_t rows, int64_t n,
                            int64_t row_stride, int64_t elem_stride, double sigma, int where) {
    SSHG_REQUIRE(ctx && in && out, "sshg_gauss1d: null argument");
    SSHG_REQUIRE(in != out, "sshg_gauss1d: in-place filtering is not supported");
    SSHG_REQUIRE(rows > 0 && n > 0, "sshg_gauss1d: rows and n must be positive");
    SSHG_REQUIRE(elem_stride >= 1 && row_stride >= 1, "sshg_gauss1d: strides must be positive");
    SSHG_REQUIRE(sigma >= 0.0 && isfinite(sigma), "sshg_gauss1d: sigma must be finite and >= 0 (got %g)", sigma);
    SSHG_CUDA(cudaSetDevice(ctx->device));
    const int in_dev = (where & SSHG_IN_DEVICE) != 0, out_dev = (where & SSHG_OUT_DEVICE) != 0;

    // extent of the (possibly strided) block of memory the rows live in
    int64_t extent = 0;
    {
        // rows may interleave (row_stride < n * elem_stride when filtering a non-last axis)
        const int64_t a = (rows - 1) * row_stride + (n - 1) * elem_stride + 1;
        extent = a;
    }
    const size_t bytes = (size_t)extent * 8;
    const void* din = in;
    if (int rc = sshg_stage_in(ctx, 10, in, bytes, in_dev, &din)) return rc;
    void* dout = out;
    if (!out_dev) {
        if (int rc = sshg_scratch(ctx, 11, bytes, &dout)) return rc;
    }

    const long long total = (long long)rows * n;
    if (!(sigma > 1e-15)) {
        copy_strided_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
            (const double*)din, (double*)dout, rows, n, row_stride, elem_stride);
    } else {
        const long long r = (long long)(4.0 * sigma + 0.5);
        SSHG_REQUIRE(r < (1LL << 28), "sshg_gauss1d: sigma %g is too large", sigma);
        double* w = (double*)malloc((size_t)(r + 1) * 8);
        if (!w) {
            sshg_set_error("out of host memory");
            return SSHG_ENOMEM;
        }
        // scipy _gaussian_kernel1d: exp(-0.5 / sigma^2 * x^2), normalised by the sum over -r..r
        const double f = -0.5 / (sigma * sigma);
        double sum = 0.0;
        for (long long k = r; k >= 1; --k) {
            w[k] = exp(f * (double)(k * k));
            sum += 2.0 * w[k];
        }
        w[0] = 1.0;
        sum += 1.0;
        for (long long k = 0; k <= r; ++k) w[k] /= sum;
        void* dw;
        int rc = sshg_scratch(ctx, 12, (size_t)(r + 1) * 8, &dw);
        if (rc == SSHG_OK) {
            cudaError_t e = cudaMemcpyAsync(dw, w, (size_t)(r + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // w is freed below
            if (e != cudaSuccess) rc = sshg_cuda_fail(e, "tap upload", __FILE__, __LINE__);
        }
        free(w);
        if (rc) return rc;
        const size_t smem = (size_t)(GTILE + 2 * r + r + 1) * 8;
        if (smem <= 200 * 1024) {
            const int tiles = (int)((n + GTILE - 1) / GTILE);
            SSHG_REQUIRE(rows * tiles < (1LL << 31), "sshg_gauss1d: too many tiles");
            SSHG_CUDA(cudaFuncSetAttribute(gauss_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            gauss_tiled_kernel<<<(unsigned)(rows * tiles), GT, smem, ctx->stream>>>(
                (const double*)din, (double*)dout, (const double*)dw, rows, n, row_stride, elem_stride, (int)r, tiles);
        } else {
            gauss_direct_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
                (const double*)din, (double*)dout, (const double*)dw, rows, n, row_stride, elem_stride, r);
        }
    }
    SSHG_CUDA(cudaGetLastError());
    ctx->launches++;
    if (!out_dev) {
        SSHG_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return SSHG_OK;
}
