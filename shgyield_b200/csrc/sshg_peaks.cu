// Roofline denominators measured in the same run as the bench:
// a pure streaming-store kernel (the yield cube is write-only traffic, while
// MEASURED_PEAKS.json's figure is a read+write copy) and a register-resident DFMA loop.
#include "sshg_internal.h"

namespace {

// Each warp writes 4 consecutive 512-byte runs per iteration (the access pattern of the yield kernel:
// st.global.cs.v2.f64, a warp covers 512 contiguous bytes per store).
__global__ void __launch_bounds__(256) store_kernel(double2* __restrict__ p, long long n2, double v) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (long long i = warp * 128 + lane; i < n2; i += stride) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long j = i + 32 * k;
            if (j < n2) __stcs(p + j, make_double2(v, v + (double)k));
        }
    }
}

// Device-to-device copy by the bulk-copy engine: one elected thread per CTA moves a 16 KB tile global -> shared
// (cp.async.bulk, completion on an mbarrier) and shared -> global (bulk group): what K1's staging does, without the FIR.
constexpr int COPY_TILE = 16 * 1024;
__global__ void __launch_bounds__(32) copy_bulk_kernel(const char* __restrict__ src, char* __restrict__ dst, long long bytes) {
    __shared__ __align__(128) char buf[COPY_TILE];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x != 0) return;
    const long long off = (long long)blockIdx.x * COPY_TILE;
    const unsigned n = (unsigned)(bytes - off < COPY_TILE ? bytes - off : COPY_TILE);
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar), buf_a = (unsigned)__cvta_generic_to_shared(buf);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(n) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf_a),
                 "l"(src + off), "r"(n), "r"(bar_a)
                 : "memory");
    unsigned done = 0;
    while (!done)
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar_a)
            : "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(buf_a), "r"(n) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

constexpr int DFMA_ILP = 8;
constexpr int DFMA_ITERS = 4096;

__global__ void __launch_bounds__(256) dfma_kernel(double* __restrict__ sink, double a, double b) {
    double x[DFMA_ILP];
#pragma unroll
    for (int k = 0; k < DFMA_ILP; ++k) x[k] = a + (double)(threadIdx.x + k);
#pragma unroll 4
    for (int it = 0; it < DFMA_ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < DFMA_ILP; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < DFMA_ILP; ++k) s += x[k];
    if (s == 12345.678) sink[0] = s;           // never true; keeps the loop alive
}

}  // namespace

extern "C" int sshg_peak_store(sshg_ctx* ctx, int64_t bytes, int iters, double* gbs) {
    SSHG_REQUIRE(ctx && gbs && bytes >= 1024 && iters > 0, "sshg_peak_store: bad argument");
    SSHG_CUDA(cudaSetDevice(ctx->device));
    void* buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        sshg_set_error("sshg_peak_store: device allocation of %lld bytes failed", (long long)bytes);
        return SSHG_ENOMEM;
    }
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const long long n2 = bytes / 16;
    const int grid = ctx->num_sms * 16;
    float best = 1e30f;
    for (int it = 0; it < iters + 2; ++it) {
        cudaEventRecord(t0, ctx->stream);
        store_kernel<<<grid, 256, 0, ctx->stream>>>((double2*)buf, n2, 1.0 + it);
        cudaEventRecord(t1, ctx->stream);
        cudaEventSynchronize(t1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        if (it >= 2 && ms < best) best = ms;
        ctx->launches++;
    }
    e = cudaGetLastError();
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(buf);
    if (e != cudaSuccess) return sshg_cuda_fail(e, "store_kernel", __FILE__, __LINE__);
    *gbs = (double)(n2 * 16) / (best * 1e-3) / 1e9;
    return SSHG_OK;
}

extern "C" int sshg_peak_copy(sshg_ctx* ctx, int64_t bytes, int iters, double* gbs) {
    SSHG_REQUIRE(ctx && gbs && bytes >= 1024 && bytes % 16 == 0 && iters > 0, "sshg_peak_copy: bad argument");
    SSHG_CUDA(cudaSetDevice(ctx->device));
    void *src = nullptr, *dst = nullptr;
    if (cudaMalloc(&src, (size_t)bytes) != cudaSuccess || cudaMalloc(&dst, (size_t)bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        if (src) cudaFree(src);
        sshg_set_error("sshg_peak_copy: device allocation of 2 x %lld bytes failed", (long long)bytes);
        return SSHG_ENOMEM;
    }
    cudaMemsetAsync(src, 1, (size_t)bytes, ctx->stream);
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const long long tiles = (bytes + COPY_TILE - 1) / COPY_TILE;
    float best = 1e30f;
    for (int it = 0; it < iters + 2; ++it) {
        cudaEventRecord(t0, ctx->stream);
        copy_bulk_kernel<<<(unsigned)tiles, 32, 0, ctx->stream>>>((const char*)src, (char*)dst, bytes);
        cudaEventRecord(t1, ctx->stream);
        cudaEventSynchronize(t1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        if (it >= 2 && ms < best) best = ms;
        ctx->launches++;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(src);
    cudaFree(dst);
    if (e != cudaSuccess) return sshg_cuda_fail(e, "copy_bulk_kernel", __FILE__, __LINE__);
    *gbs = 2.0 * (double)bytes / (best * 1e-3) / 1e9;          // read + written
    return SSHG_OK;
}

extern "C" int sshg_peak_dfma(sshg_ctx* ctx, int iters, double* tflops) {
    SSHG_REQUIRE(ctx && tflops && iters > 0, "sshg_peak_dfma: bad argument");
    SSHG_CUDA(cudaSetDevice(ctx->device));
    void* sink;
    if (int rc = sshg_scratch(ctx, 15, 16, &sink)) return rc;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int grid = ctx->num_sms * 16, block = 256;
    float best = 1e30f;
    for (int it = 0; it < iters + 2; ++it) {
        cudaEventRecord(t0, ctx->stream);
        dfma_kernel<<<grid, block, 0, ctx->stream>>>((double*)sink, 1.0000001, 1e-9);
        cudaEventRecord(t1, ctx->stream);
        cudaEventSynchronize(t1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        if (it >= 2 && ms < best) best = ms;
        ctx->launches++;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    if (e != cudaSuccess) return sshg_cuda_fail(e, "dfma_kernel", __FILE__, __LINE__);
    const double flops = 2.0 * DFMA_ILP * DFMA_ITERS * (double)grid * block;
    *tflops = flops / (best * 1e-3) / 1e12;
    return SSHG_OK;
}
