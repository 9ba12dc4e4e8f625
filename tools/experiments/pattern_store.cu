// Experiment: store-only kernels with the address pattern of yield_grid_kernel (no arithmetic):
// CTA width, thread-block clusters, and row-major sequential writing as the reference point.
#include <cstdio>
#include <cuda_runtime.h>
template <int NT>
__global__ void pattern(double* out, long long nE, long long nP, int n_tiles, int H) {
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
        __stcs((double2*)(base + pol * rows + 2LL * H * nE), make_double2(5.0, 6.0));
    }
}
template <int NT> void run(double* buf, int nT, int cluster, const char* name) {
    const long long nE = 20000, nP = 721; const int H = 360;
    int n_tiles = (nE + 2 * NT - 1) / (2 * NT);
    if (cluster > 1) n_tiles = (n_tiles + cluster - 1) / cluster * cluster;     // grid must be a multiple of the cluster
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(a);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(nT * n_tiles); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = 0; cfg.stream = 0;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = cluster > 1 ? 1 : 0;
        cudaLaunchKernelEx(&cfg, pattern<NT>, buf, nE, nP, n_tiles, H);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
    }
    const double bytes = 8.0 * nT * 4 * nP * nE;
    printf("%-8s cluster=%d nT=%d: %.3f ms  %.1f GB/s  (%s)\n", name, cluster, nT, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int nT = 46;
    double* buf; cudaMalloc(&buf, 8ull * nT * 4 * 721 * 20000 + (1 << 20));
    for (int cl : {1, 2, 4, 8}) { run<64>(buf, nT, cl, "NT=64"); }
    for (int cl : {1, 2, 4, 8}) { run<128>(buf, nT, cl, "NT=128"); }
    for (int cl : {1, 2, 4, 8}) { run<256>(buf, nT, cl, "NT=256"); }
    run<512>(buf, nT, 1, "NT=512"); run<1024>(buf, nT, 1, "NT=1024");
    return 0;
}
