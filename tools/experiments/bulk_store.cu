// Experiment: K2's address pattern (every CTA writes 1 KB runs of rows that are 160 KB apart) with the row segments
// staged in shared memory and written by cp.async.bulk (shared -> global, one elected thread), against the same
// pattern with per-thread 128-bit streaming stores.  No arithmetic: this measures the ceiling of the pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_store bulk_store.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// direct: per-thread st.global.cs.v2.f64 (what yield_grid_kernel does)
template <int NT>
__global__ void direct(double* out, long long nE, long long nP, int n_tiles, int H) {
    const long long blk = blockIdx.x;
    const int tile = blk % n_tiles;
    const long long col = blk / n_tiles;
    const long long e0 = ((long long)tile * NT + threadIdx.x) * 2;
    if (e0 >= nE) return;
    const long long rows = nP * nE;
    double* base = out + col * 4 * rows + e0;
    for (int pol = 0; pol < 4; ++pol) {
        double* lo = base + pol * rows;
        double* hi = lo + (long long)H * nE;
        for (int i = 0; i < H; ++i) {
            __stcs((double2*)lo, make_double2(1.0 + i, 2.0));
            __stcs((double2*)hi, make_double2(3.0, 4.0 + i));
            lo += nE; hi += nE;
        }
    }
}

// bulk: B items (2 B rows of NT x 16 bytes) per batch, two batches in flight
template <int NT, int B>
__global__ void bulk(double* out, long long nE, long long nP, int n_tiles, int H) {
    __shared__ __align__(128) double buf[2][2 * B][NT * 2];
    const long long blk = blockIdx.x;
    const int tile = blk % n_tiles;
    const long long col = blk / n_tiles;
    const long long tile_e0 = (long long)tile * NT * 2;
    long long valid = nE - tile_e0;                              // energies of this tile that exist
    if (valid > NT * 2) valid = NT * 2;
    const unsigned row_bytes = (unsigned)valid * 8;              // multiple of 16: nE even
    const long long rows = nP * nE;
    double* base = out + col * 4 * rows + tile_e0;
    int slot = 0;
    for (int pol = 0; pol < 4; ++pol) {
        double* lo = base + pol * rows;
        double* hi = lo + (long long)H * nE;
        for (int i0 = 0; i0 < H; i0 += B) {
            const int nb = H - i0 < B ? H - i0 : B;
            if (threadIdx.x == 0) bulk_wait_read<1>();           // the batch issued two iterations ago has left shared memory
            __syncthreads();
            for (int i = 0; i < nb; ++i) {
                *reinterpret_cast<double2*>(&buf[slot][2 * i][threadIdx.x * 2]) = make_double2(1.0 + i0 + i, 2.0);
                *reinterpret_cast<double2*>(&buf[slot][2 * i + 1][threadIdx.x * 2]) = make_double2(3.0, 4.0 + i0 + i);
            }
            fence_async();
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int i = 0; i < nb; ++i) {
                    bulk_store(lo + (long long)(i0 + i) * nE, &buf[slot][2 * i][0], row_bytes);
                    bulk_store(hi + (long long)(i0 + i) * nE, &buf[slot][2 * i + 1][0], row_bytes);
                }
                bulk_commit();
            }
            slot ^= 1;
        }
    }
    if (threadIdx.x == 0) bulk_wait_read<0>();
}

// bulk_warp: the same, but every WARP stages and writes its own 512-byte half of each row (lane 0 issues): no CTA barrier
template <int NT, int B>
__global__ void bulk_warp(double* out, long long nE, long long nP, int n_tiles, int H) {
    __shared__ __align__(128) double buf[NT / 32][2][2 * B][64];
    const long long blk = blockIdx.x;
    const int tile = blk % n_tiles;
    const long long col = blk / n_tiles;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const long long warp_e0 = (long long)tile * NT * 2 + warp * 64;
    long long valid = nE - warp_e0;
    if (valid <= 0) return;
    if (valid > 64) valid = 64;
    const unsigned row_bytes = (unsigned)valid * 8;
    const long long rows = nP * nE;
    double* base = out + col * 4 * rows + warp_e0;
    int slot = 0;
    for (int pol = 0; pol < 4; ++pol) {
        double* lo = base + pol * rows;
        double* hi = lo + (long long)H * nE;
        for (int i0 = 0; i0 < H; i0 += B) {
            const int nb = H - i0 < B ? H - i0 : B;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            for (int i = 0; i < nb; ++i) {
                *reinterpret_cast<double2*>(&buf[warp][slot][2 * i][lane * 2]) = make_double2(1.0 + i0 + i, 2.0);
                *reinterpret_cast<double2*>(&buf[warp][slot][2 * i + 1][lane * 2]) = make_double2(3.0, 4.0 + i0 + i);
            }
            fence_async();
            __syncwarp();
            if (lane == 0) {
                for (int i = 0; i < nb; ++i) {
                    bulk_store(lo + (long long)(i0 + i) * nE, &buf[warp][slot][2 * i][0], row_bytes);
                    bulk_store(hi + (long long)(i0 + i) * nE, &buf[warp][slot][2 * i + 1][0], row_bytes);
                }
                bulk_commit();
            }
            slot ^= 1;
        }
    }
    if (lane == 0) bulk_wait_read<0>();
}

template <typename F>
void timeit(const char* name, double bytes, F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
    }
    printf("%-28s %.3f ms  %.1f GB/s  (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int NT, int B> void run_bulk(double* buf, int nT) {
    const long long nE = 20000, nP = 721; const int H = 360;
    const int n_tiles = (nE + 2 * NT - 1) / (2 * NT);
    char name[64]; snprintf(name, sizeof(name), "bulk   NT=%d batch=%d", NT, B);
    timeit(name, 8.0 * nT * 4 * 2 * H * nE, [&] { bulk<NT, B><<<nT * n_tiles, NT>>>(buf, nE, nP, n_tiles, H); });
}
template <int NT, int B> void run_bulk_warp(double* buf, int nT) {
    const long long nE = 20000, nP = 721; const int H = 360;
    const int n_tiles = (nE + 2 * NT - 1) / (2 * NT);
    char name[64]; snprintf(name, sizeof(name), "bulk_warp NT=%d batch=%d", NT, B);
    timeit(name, 8.0 * nT * 4 * 2 * H * nE, [&] { bulk_warp<NT, B><<<nT * n_tiles, NT>>>(buf, nE, nP, n_tiles, H); });
}
template <int NT> void run_direct(double* buf, int nT) {
    const long long nE = 20000, nP = 721; const int H = 360;
    const int n_tiles = (nE + 2 * NT - 1) / (2 * NT);
    char name[64]; snprintf(name, sizeof(name), "direct NT=%d", NT);
    timeit(name, 8.0 * nT * 4 * 2 * H * nE, [&] { direct<NT><<<nT * n_tiles, NT>>>(buf, nE, nP, n_tiles, H); });
}

int main() {
    const int nT = 46;
    double* buf; cudaMalloc(&buf, 8ull * nT * 4 * 721 * 20000 + (1 << 20));
    run_direct<64>(buf, nT); run_direct<128>(buf, nT); run_direct<256>(buf, nT);
    run_bulk<64, 1>(buf, nT); run_bulk<64, 2>(buf, nT); run_bulk<64, 4>(buf, nT); run_bulk<64, 8>(buf, nT);
    run_bulk<128, 1>(buf, nT); run_bulk<128, 2>(buf, nT); run_bulk<128, 4>(buf, nT);
    run_bulk<256, 1>(buf, nT); run_bulk<256, 2>(buf, nT);
    run_bulk_warp<64, 2>(buf, nT); run_bulk_warp<64, 4>(buf, nT); run_bulk_warp<64, 8>(buf, nT);
    run_bulk_warp<128, 2>(buf, nT); run_bulk_warp<128, 4>(buf, nT);
    // check one value written by the bulk path
    double h[4]; cudaMemcpy(h, buf + 20000LL * 5, 32, cudaMemcpyDeviceToHost);
    printf("row 5: %.1f %.1f %.1f %.1f\n", h[0], h[1], h[2], h[3]);
    return 0;
}
