// Experiment: cache policy of the stores for K2's address pattern (1 KB runs, rows 160 KB apart).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__device__ __forceinline__ void st2(double* p, double a, double b) {
    if (MODE == 0) *reinterpret_cast<double2*>(p) = make_double2(a, b);          // st.global (write-back)
    else if (MODE == 1) __stcs(reinterpret_cast<double2*>(p), make_double2(a, b));   // .cs streaming
    else if (MODE == 2) __stcg(reinterpret_cast<double2*>(p), make_double2(a, b));   // .cg
    else if (MODE == 3) __stwt(reinterpret_cast<double2*>(p), make_double2(a, b));   // .wt write-through
    else *reinterpret_cast<double2*>(p) = make_double2(a, b);
}
template <int NT, int MODE>
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
            st2<MODE>(lo, 1.0 + i, 2.0);
            st2<MODE>(hi, 3.0, 4.0 + i);
            lo += nE; hi += nE;
        }
        st2<MODE>(base + pol * rows + 2LL * H * nE, 5.0, 6.0);
    }
}
template <int NT, int MODE> void run(double* buf, int nT, const char* name) {
    const long long nE = 20000, nP = 721; const int H = 360;
    const int n_tiles = (nE + 2 * NT - 1) / (2 * NT);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(a);
        pattern<NT, MODE><<<nT * n_tiles, NT>>>(buf, nE, nP, n_tiles, H);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
    }
    printf("NT=%d %-12s: %.3f ms  %.1f GB/s (%s)\n", NT, name, best, 8.0 * nT * 4 * nP * nE / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int nT = 46;
    double* buf; cudaMalloc(&buf, 8ull * nT * 4 * 721 * 20000 + (1 << 20));
    run<64, 0>(buf, nT, "wb"); run<64, 1>(buf, nT, "cs"); run<64, 2>(buf, nT, "cg"); run<64, 3>(buf, nT, "wt");
    run<256, 0>(buf, nT, "wb"); run<256, 1>(buf, nT, "cs"); run<256, 2>(buf, nT, "cg"); run<256, 3>(buf, nT, "wt");
    return 0;
}
