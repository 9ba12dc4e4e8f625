// Context management, error reporting and host/device staging for libsshg.so.
#include "sshg_internal.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

static thread_local char g_err[512] = "";

void sshg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sshg_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    const char* base = strrchr(file, '/');
    sshg_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), base ? base + 1 : file, line, what);
    return e == cudaErrorMemoryAllocation ? SSHG_ENOMEM : SSHG_ECUDA;
}

int sshg_scratch(sshg_ctx* ctx, int slot, size_t bytes, void** out) {
    if (slot < 0 || slot >= SSHG_NSCRATCH) {
        sshg_set_error("internal: scratch slot %d", slot);
        return SSHG_EINVAL;
    }
    if (bytes == 0) bytes = 16;
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) {
            // the old buffer may still be in use by work queued on the streams
            SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
            SSHG_CUDA(cudaStreamSynchronize(ctx->copy_stream));
            SSHG_CUDA(cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_bytes[slot] = 0;
        }
        size_t cap = bytes + bytes / 4;
        cap = (cap + 255) & ~(size_t)255;
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], cap);
        if (e != cudaSuccess) {
            cap = (bytes + 255) & ~(size_t)255;
            e = cudaMalloc(&ctx->scratch[slot], cap);
        }
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            sshg_set_error("device allocation of %zu bytes failed: %s", cap, cudaGetErrorString(e));
            return SSHG_ENOMEM;
        }
        ctx->scratch_bytes[slot] = cap;
        ctx->epoch++;
    }
    *out = ctx->scratch[slot];
    return SSHG_OK;
}

int sshg_stage_in(sshg_ctx* ctx, int slot, const void* src, size_t bytes, int on_device, const void** dev) {
    if (on_device) {
        *dev = src;
        return SSHG_OK;
    }
    void* d;
    if (int rc = sshg_scratch(ctx, slot, bytes, &d)) return rc;
    if (int rc = sshg_copy_in(ctx, d, src, bytes)) return rc;
    *dev = d;
    return SSHG_OK;
}

// ---- threaded device-to-host copy into pageable memory -------------------------------------------------------------
namespace {
constexpr size_t STAGE_BYTES = (size_t)64 << 20;       // per staging buffer
constexpr size_t STAGED_MIN = (size_t)32 << 20;        // smaller copies: one cudaMemcpy

int copy_threads() {
    unsigned n = std::thread::hardware_concurrency();   // the CPUs this process may run on (glibc: sched_getaffinity)
    if (n == 0) n = 4;
    return (int)(n > 16 ? 16 : n);
}
}  // namespace

int sshg_copy_out(sshg_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    if (bytes == 0) return SSHG_OK;
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, dst_host) == cudaSuccess && at.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();                           // un-registered host memory may report an error on old drivers
    if (pinned || bytes < STAGED_MIN) {
        SSHG_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
        return SSHG_OK;
    }
    if (!ctx->stage[0]) {
        for (int k = 0; k < 2; ++k) {
            cudaError_t e = cudaMallocHost(&ctx->stage[k], STAGE_BYTES);
            if (e != cudaSuccess) {                     // no page-locked memory left: the driver's own pageable path
                (void)cudaGetLastError();
                if (k == 1) cudaFreeHost(ctx->stage[0]);
                ctx->stage[0] = ctx->stage[1] = nullptr;
                SSHG_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
                SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
                return SSHG_OK;
            }
        }
        ctx->stage_bytes = STAGE_BYTES;
    }
    // the copies run on the copy stream, after everything queued on the compute stream so far
    SSHG_CUDA(cudaEventRecord(ctx->ev_switch, ctx->stream));
    SSHG_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_switch, 0));
    const size_t n_chunks = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
    auto chunk_len = [&](size_t i) { return i + 1 < n_chunks ? STAGE_BYTES : bytes - i * STAGE_BYTES; };
    const int T = copy_threads();
    // workers: chunk i may be copied once ready > i; a worker bumps done[i % 2] when its part of chunk i is out
    std::atomic<long long> ready(0);
    std::atomic<int> done[2];
    done[0] = done[1] = 0;
    std::atomic<bool> abort_flag(false);
    std::vector<std::thread> pool;
    pool.reserve(T);
    bool spawned = true;
    for (int t = 0; t < T && spawned; ++t) {
      try {
        pool.emplace_back([&, t]() {
            for (size_t i = 0; i < n_chunks; ++i) {
                while (ready.load(std::memory_order_acquire) <= (long long)i) {
                    if (abort_flag.load(std::memory_order_relaxed)) return;
                    std::this_thread::yield();
                }
                const size_t len = chunk_len(i);
                const size_t per = ((len + T - 1) / T + 4095) & ~(size_t)4095;       // page-sized shares
                const size_t lo = (size_t)t * per, hi = lo + per < len ? lo + per : len;
                if (lo < hi)
                    memcpy((char*)dst_host + i * STAGE_BYTES + lo, (const char*)ctx->stage[i % 2] + lo, hi - lo);
                done[i % 2].fetch_add(1, std::memory_order_release);
            }
        });
      } catch (...) {                                   // no more threads (resource limits): nothing crosses the C ABI
        spawned = false;
      }
    }
    if (!spawned) {                                     // the workers that did start leave; the driver's own path copies
        abort_flag.store(true);
        for (auto& th : pool) th.join();
        SSHG_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
        return SSHG_OK;
    }
    cudaError_t err = cudaSuccess;
    auto issue = [&](size_t i) {
        if (err != cudaSuccess) return;
        err = cudaMemcpyAsync(ctx->stage[i % 2], (const char*)src_dev + i * STAGE_BYTES, chunk_len(i), cudaMemcpyDeviceToHost,
                              ctx->copy_stream);
        if (err == cudaSuccess) err = cudaEventRecord(ctx->ev_copied[i % 2], ctx->copy_stream);
    };
    issue(0);
    for (size_t i = 0; i < n_chunks && err == cudaSuccess; ++i) {
        if (i + 1 < n_chunks) {
            if (i >= 1) {                               // buffer (i + 1) % 2 held chunk i - 1: every worker is out of it
                while (done[(i + 1) % 2].load(std::memory_order_acquire) < T) std::this_thread::yield();
                done[(i + 1) % 2].store(0, std::memory_order_relaxed);
            }
            issue(i + 1);
        }
        if (err == cudaSuccess) err = cudaEventSynchronize(ctx->ev_copied[i % 2]);
        if (err == cudaSuccess) ready.store((long long)i + 1, std::memory_order_release);
    }
    if (err != cudaSuccess) abort_flag.store(true);
    for (auto& th : pool) th.join();
    if (err != cudaSuccess) return sshg_cuda_fail(err, "staged device-to-host copy", __FILE__, __LINE__);
    return SSHG_OK;
}

// Host -> device copy of a large pageable array (sshg_stage_in): host threads fill the page-locked staging buffers, the
// copy engine drains them -- the mirror image of sshg_copy_out.  Queued work on ctx->stream that reads `dst_dev` must be
// issued after this returns (the last chunk is synchronised).
int sshg_copy_in(sshg_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, src_host) == cudaSuccess && at.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    bool staged = !pinned && bytes >= STAGED_MIN;
    if (staged && !ctx->stage[0]) {
        for (int k = 0; k < 2 && staged; ++k) {
            if (cudaMallocHost(&ctx->stage[k], STAGE_BYTES) != cudaSuccess) {
                (void)cudaGetLastError();
                if (k == 1) cudaFreeHost(ctx->stage[0]);
                ctx->stage[0] = ctx->stage[1] = nullptr;
                staged = false;
            }
        }
        if (staged) ctx->stage_bytes = STAGE_BYTES;
    }
    if (!staged) {
        // pageable source: the runtime stages the copy before returning, so the caller's buffer is free to change as soon
        // as the API call returns
        SSHG_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return SSHG_OK;
    }
    // the device buffer may still be read by work queued on the compute stream: the copies wait for it; this call returns
    // when the last copy has landed
    SSHG_CUDA(cudaEventRecord(ctx->ev_switch, ctx->stream));
    SSHG_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_switch, 0));
    const size_t n_chunks = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
    int T = copy_threads();
    for (size_t i = 0; i < n_chunks; ++i) {
        const size_t len = i + 1 < n_chunks ? STAGE_BYTES : bytes - i * STAGE_BYTES;
        if (i >= 2) SSHG_CUDA(cudaEventSynchronize(ctx->ev_copied[i % 2]));       // the engine has drained this buffer
        // fill stage[i % 2] with T threads (blocking), then hand it to the copy engine; the other buffer is on the wire
        {
            const size_t per = ((len + T - 1) / T + 4095) & ~(size_t)4095;
            char* dstb = (char*)ctx->stage[i % 2];
            const char* srcb = (const char*)src_host + i * STAGE_BYTES;
            auto work = [&](int t) {
                const size_t lo = (size_t)t * per, hi = lo + per < len ? lo + per : len;
                if (lo < hi) memcpy(dstb + lo, srcb + lo, hi - lo);
            };
            std::vector<std::thread> pool;
            pool.reserve(T - 1);
            int started = 1;
            for (int t = 1; t < T; ++t) {
                try {
                    pool.emplace_back(work, t);
                    started = t + 1;
                } catch (...) {
                    break;
                }
            }
            work(0);
            for (int t = started; t < T; ++t) work(t);
            for (auto& th : pool) th.join();
        }
        SSHG_CUDA(cudaMemcpyAsync((char*)dst_dev + i * STAGE_BYTES, ctx->stage[i % 2], len, cudaMemcpyHostToDevice, ctx->copy_stream));
        SSHG_CUDA(cudaEventRecord(ctx->ev_copied[i % 2], ctx->copy_stream));
    }
    SSHG_CUDA(cudaStreamSynchronize(ctx->copy_stream));                           // staging buffers free, data on the device
    return SSHG_OK;
}

void sshg_parallel_dup(char* base, size_t pitch, size_t src_off, size_t dst_off, size_t bytes, int64_t ncols) {
    const size_t total = bytes * (size_t)ncols;
    if (total == 0) return;
    int T = copy_threads();
    if (T > 8) T = 8;                                   // the PCIe copy of the next slab runs meanwhile
    if (total < ((size_t)4 << 20)) T = 1;
    const size_t share = ((total + T - 1) / T + 4095) & ~(size_t)4095;
    auto work = [&](int t) {
        size_t lo = (size_t)t * share, hi = lo + share < total ? lo + share : total;
        while (lo < hi) {                               // virtual byte range [lo, hi) of the concatenated blocks
            const size_t c = lo / bytes, off = lo - c * bytes;
            const size_t n = (bytes - off < hi - lo) ? bytes - off : hi - lo;
            memcpy(base + c * pitch + dst_off + off, base + c * pitch + src_off + off, n);
            lo += n;
        }
    };
    if (T == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(T - 1);
    int started = 1;                                    // shares [0, started) have a thread (this one takes share 0)
    for (int t = 1; t < T; ++t) {
        try {
            pool.emplace_back(work, t);
            started = t + 1;
        } catch (...) {                                 // no more threads: this thread does the remaining shares
            break;
        }
    }
    work(0);
    for (int t = started; t < T; ++t) work(t);
    for (auto& th : pool) th.join();
}

extern "C" {

int sshg_abi_version(void) { return SSHG_ABI_VERSION; }

const char* sshg_last_error(void) { return g_err; }

int sshg_device_count(int* count) {
    SSHG_REQUIRE(count, "sshg_device_count: null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *count = 0;
        return sshg_cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    }
    *count = n;
    return SSHG_OK;
}

int sshg_ctx_create(int device, sshg_ctx** out) {
    SSHG_REQUIRE(out, "sshg_ctx_create: null argument");
    *out = nullptr;
    int n = 0;
    if (int rc = sshg_device_count(&n)) return rc;
    if (n == 0) {
        sshg_set_error("no CUDA device: libsshg has no CPU fallback");
        return SSHG_ECUDA;
    }
    SSHG_REQUIRE(device >= 0 && device < n, "sshg_ctx_create: device %d of %d", device, n);
    SSHG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SSHG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        sshg_set_error("device %d is sm_%d%d; libsshg is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return SSHG_ECUDA;
    }
    sshg_ctx* ctx = (sshg_ctx*)calloc(1, sizeof(sshg_ctx));
    if (!ctx) {
        sshg_set_error("out of host memory");
        return SSHG_ENOMEM;
    }
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->opt = sshg_options{-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_copied[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_switch, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        sshg_ctx_destroy(ctx);
        return sshg_cuda_fail(e, "stream/event creation", __FILE__, __LINE__);
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return SSHG_OK;
}

void sshg_ctx_destroy(sshg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < SSHG_NSCRATCH; ++i)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    for (int k = 0; k < 2; ++k) {
        if (ctx->ev_done[k]) cudaEventDestroy(ctx->ev_done[k]);
        if (ctx->ev_copied[k]) cudaEventDestroy(ctx->ev_copied[k]);
    }
    for (int k = 0; k < SSHG_NGRAPH; ++k)
        if (ctx->graphs[k].exec) cudaGraphExecDestroy(ctx->graphs[k].exec);
    for (int k = 0; k < 2; ++k)
        if (ctx->stage[k]) cudaFreeHost(ctx->stage[k]);
    if (ctx->pin_in) cudaFreeHost(ctx->pin_in);
    if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
    if (ctx->ev_switch) cudaEventDestroy(ctx->ev_switch);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    free(ctx);
}

int sshg_ctx_set_stream(sshg_ctx* ctx, void* cuda_stream) {
    SSHG_REQUIRE(ctx, "sshg_ctx_set_stream: null context");
    cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    if (next == ctx->stream) return SSHG_OK;
    SSHG_CUDA(cudaSetDevice(ctx->device));
    // The scratch buffers are ordered on the old stream: the new stream waits (on the device, no host
    // synchronisation) for everything queued there.  Callers that adopt a temporary stream switch back to the
    // context's own stream (NULL) before that stream is destroyed -- the Python binding does so in a finally clause.
    SSHG_CUDA(cudaEventRecord(ctx->ev_switch, ctx->stream));
    SSHG_CUDA(cudaStreamWaitEvent(next, ctx->ev_switch, 0));
    ctx->stream = next;
    return SSHG_OK;
}

namespace {
typedef int sshg_options::*OptionField;
struct OptionKey {
    const char* key;
    OptionField field;
    int lo, hi;             // allowed values besides -1
    int only_a, only_b;     // when >= 0: the only two allowed values
};
const OptionKey kOptions[] = {
    {"k2_sparse", &sshg_options::k2_sparse, 0, 1, -1, -1},
    {"k2_nt", &sshg_options::k2_nt, 0, 0, 64, 256},
    {"k2_recurrence", &sshg_options::k2_recurrence, 0, 1, -1, -1},
    {"k3", &sshg_options::k3, 0, 1, -1, -1},
    {"k3_nt", &sshg_options::k3_nt, 0, 0, 64, 128},
    {"k3_fixup", &sshg_options::k3_fixup, 0, 1, -1, -1},
    {"nseg", &sshg_options::nseg, 1, 16, -1, -1},
    {"k1_generic", &sshg_options::k1_generic, 0, 1, -1, -1},
    {"graphs", &sshg_options::graphs, 0, 1, -1, -1},
    {"quads", &sshg_options::quads, 0, 1, -1, -1},
    {"k1_bulk", &sshg_options::k1_bulk, 0, 1, -1, -1},
    {"pair_copy", &sshg_options::pair_copy, 0, 1, -1, -1},
};
const OptionKey* find_option(const char* key) {
    for (const OptionKey& o : kOptions)
        if (strcmp(o.key, key) == 0) return &o;
    return nullptr;
}
}  // namespace

int sshg_ctx_set_option(sshg_ctx* ctx, const char* key, int64_t value) {
    SSHG_REQUIRE(ctx && key, "sshg_ctx_set_option: null argument");
    const OptionKey* o = find_option(key);
    SSHG_REQUIRE(o, "sshg_ctx_set_option: unknown option '%s'", key);
    const bool ok = value == -1 || (o->only_a >= 0 ? (value == o->only_a || value == o->only_b)
                                                   : (value >= o->lo && value <= o->hi));
    SSHG_REQUIRE(ok, "sshg_ctx_set_option: value %lld is not allowed for '%s'", (long long)value, key);
    ctx->opt.*(o->field) = (int)value;
    return SSHG_OK;
}

int sshg_ctx_get_option(sshg_ctx* ctx, const char* key, int64_t* value) {
    SSHG_REQUIRE(ctx && key && value, "sshg_ctx_get_option: null argument");
    const OptionKey* o = find_option(key);
    SSHG_REQUIRE(o, "sshg_ctx_get_option: unknown option '%s'", key);
    *value = ctx->opt.*(o->field);
    return SSHG_OK;
}

void* sshg_ctx_stream(sshg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int sshg_ctx_sync(sshg_ctx* ctx) {
    SSHG_REQUIRE(ctx, "sshg_ctx_sync: null context");
    SSHG_CUDA(cudaSetDevice(ctx->device));
    SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
    SSHG_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return SSHG_OK;
}

int64_t sshg_ctx_launches(sshg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int sshg_host_alloc(void** ptr, int64_t bytes) {
    SSHG_REQUIRE(ptr && bytes > 0, "sshg_host_alloc: bad argument");
    cudaError_t e = cudaMallocHost(ptr, (size_t)bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *ptr = nullptr;
        sshg_set_error("pinned host allocation of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
        return SSHG_ENOMEM;
    }
    return SSHG_OK;
}

int sshg_host_free(void* ptr) {
    if (!ptr) return SSHG_OK;
    SSHG_CUDA(cudaFreeHost(ptr));
    return SSHG_OK;
}

int sshg_download(sshg_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes) {
    SSHG_REQUIRE(ctx && dst_host && src_dev && bytes > 0, "sshg_download: bad argument");
    SSHG_CUDA(cudaSetDevice(ctx->device));
    return sshg_copy_out(ctx, dst_host, src_dev, (size_t)bytes);
}

int sshg_upload(sshg_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes) {
    SSHG_REQUIRE(ctx && dst_dev && src_host && bytes > 0, "sshg_upload: bad argument");
    SSHG_CUDA(cudaSetDevice(ctx->device));
    SSHG_CUDA(cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
    return SSHG_OK;
}

// xxHash64 construction (Y. Collet's public specification): four 64-bit lanes over 32-byte stripes, merged and
// avalanched.  Host code; used as a content key, not for security.
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

uint64_t sshg_hash64(const void* data, int64_t bytes, uint64_t seed) {
    const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
                   P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
    if (!data || bytes <= 0) return seed ^ P5;
    const unsigned char* p = (const unsigned char*)data;
    const unsigned char* end = p + bytes;
    auto rd64 = [](const unsigned char* q) { uint64_t v; memcpy(&v, q, 8); return v; };
    auto round = [&](uint64_t acc, uint64_t in) { return rotl64(acc + in * P2, 31) * P1; };
    uint64_t h;
    if (bytes >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        const unsigned char* limit = end - 32;
        do {
            v1 = round(v1, rd64(p));
            v2 = round(v2, rd64(p + 8));
            v3 = round(v3, rd64(p + 16));
            v4 = round(v4, rd64(p + 24));
            p += 32;
        } while (p <= limit);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        auto merge = [&](uint64_t acc, uint64_t v) { return (acc ^ round(0, v)) * P1 + P4; };
        h = merge(h, v1); h = merge(h, v2); h = merge(h, v3); h = merge(h, v4);
    } else {
        h = seed + P5;
    }
    h += (uint64_t)bytes;
    while (p + 8 <= end) {
        h = rotl64(h ^ round(0, rd64(p)), 27) * P1 + P4;
        p += 8;
    }
    while (p < end) {
        h = rotl64(h ^ (*p * P5), 11) * P1;
        ++p;
    }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

// ---- peer memory for the store-through gather (include/sshg.h) ---------------------------------------
int sshg_device_alloc(sshg_ctx* ctx, void** ptr, int64_t bytes) {
    SSHG_REQUIRE(ctx && ptr && bytes > 0, "sshg_device_alloc: bad argument");
    *ptr = nullptr;
    SSHG_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        sshg_set_error("device allocation of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
        return SSHG_ENOMEM;
    }
    return SSHG_OK;
}

int sshg_device_free(sshg_ctx* ctx, void* ptr) {
    SSHG_REQUIRE(ctx, "sshg_device_free: null context");
    if (!ptr) return SSHG_OK;
    SSHG_CUDA(cudaSetDevice(ctx->device));
    SSHG_CUDA(cudaFree(ptr));
    return SSHG_OK;
}

static_assert(sizeof(cudaIpcMemHandle_t) == SSHG_IPC_HANDLE_BYTES, "IPC handle size");

int sshg_ipc_export(sshg_ctx* ctx, const void* dev_ptr, unsigned char handle[SSHG_IPC_HANDLE_BYTES]) {
    SSHG_REQUIRE(ctx && dev_ptr && handle, "sshg_ipc_export: null argument");
    SSHG_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    SSHG_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle, &h, sizeof(h));
    return SSHG_OK;
}

int sshg_ipc_open(sshg_ctx* ctx, const unsigned char handle[SSHG_IPC_HANDLE_BYTES], void** dev_ptr) {
    SSHG_REQUIRE(ctx && handle && dev_ptr, "sshg_ipc_open: null argument");
    *dev_ptr = nullptr;
    SSHG_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    SSHG_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SSHG_OK;
}

int sshg_ipc_close(sshg_ctx* ctx, void* dev_ptr) {
    SSHG_REQUIRE(ctx, "sshg_ipc_close: null context");
    if (!dev_ptr) return SSHG_OK;
    SSHG_CUDA(cudaSetDevice(ctx->device));
    SSHG_CUDA(cudaStreamSynchronize(ctx->stream));     // stores of queued kernels have left this GPU
    SSHG_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return SSHG_OK;
}

}  // extern "C"
