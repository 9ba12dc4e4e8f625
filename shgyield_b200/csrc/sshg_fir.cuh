// Register-window FIR arithmetic of K1 (gauss_stream_kernel), kept in a header so that
// tests/hostmath_harness.cu can run exactly the same code on the CPU.
//
// A thread produces SR consecutive outputs o = base + q from the samples xw[t] = x[base - r + t],
// t = 0 .. 2r + SR - 1:   out[q] = sum_{j=0..2r} w_full[j] * xw[q + j].
// The loop runs over t; sample t feeds output q with tap j = t - q.  The SR taps a step needs are
// a sliding window over wext[k] = w_full[k - (SR-1)] (zero outside 0..2r) kept in registers,
// rotated statically (the t loop is unrolled by SR).
//
// XS: distance in doubles between consecutive samples of the window (1: a contiguous signal; the column kernel, which
// filters a non-contiguous axis of an N-D array, keeps its tile as [position along the axis][column] and passes the tile
// width).
#pragma once
#include <cuda_runtime.h>

#define SSHG_FIR_HD __host__ __device__ __forceinline__

namespace sshg {

constexpr int SR = 9;   // odd: a stride of 9 doubles between the windows of a warp is LDS-conflict-free

// One FIR step of the register window: sample v = xw[t] feeds output q with tap j = t - q.
// wr[k % SR] holds wext[k] = w_full[k - (SR-1)] for the current window k in [t, t + SR - 1].
// QLO / QHI (static) bound the outputs that own a tap at this step.
template <int XS, int U, int QLO, int QHI>
SSHG_FIR_HD void fir_step(const double* __restrict__ xw, const double* __restrict__ wext, int t0,
                                         double (&acc)[SR], double (&wr)[SR]) {
    const double v = xw[(t0 + U) * XS];
#pragma unroll
    for (int q = 0; q < SR; ++q)
        if (q >= QLO && q <= QHI) acc[q] = fma(wr[(U + 2 * SR - 1 - q) % SR], v, acc[q]);
    wr[U % SR] = wext[t0 + U + SR];                // tap entering the window (zero beyond the kernel)
}

// The last L full steps (t <= 2r) and the closing ramp t = 2r + s, s = 1 .. SR-1, where only the
// outputs q >= s still own a tap.  Everything is static once L = 2r - t0 + 1 is known.
template <int XS, int L, int U = 0>
SSHG_FIR_HD void fir_tail(const double* __restrict__ xw, const double* __restrict__ wext, int t0,
                                         double (&acc)[SR], double (&wr)[SR]) {
    if constexpr (U < L + SR - 1) {
        constexpr int S = U - L + 1;               // <= 0: full step
        fir_step<XS, U, (S > 0 ? S : 0), SR - 1>(xw, wext, t0, acc, wr);
        fir_tail<XS, L, U + 1>(xw, wext, t0, acc, wr);
    }
}

template <int XS, int U = 0>
SSHG_FIR_HD void fir_head(const double* __restrict__ xw, const double* __restrict__ wext,
                                         double (&acc)[SR], double (&wr)[SR]) {
    if constexpr (U < SR) {                        // opening ramp: t = U, outputs q <= U
        fir_step<XS, U, 0, U>(xw, wext, 0, acc, wr);
        fir_head<XS, U + 1>(xw, wext, acc, wr);
    }
}

template <int XS, int U = 0>
SSHG_FIR_HD void fir_full(const double* __restrict__ xw, const double* __restrict__ wext, int t0,
                                         double (&acc)[SR], double (&wr)[SR]) {
    if constexpr (U < SR) {
        fir_step<XS, U, 0, SR - 1>(xw, wext, t0, acc, wr);
        fir_full<XS, U + 1>(xw, wext, t0, acc, wr);
    }
}


// The whole window: acc[] must be zero and wr[q] = wext[q] on entry.
template <int XS = 1>
SSHG_FIR_HD void fir_window(const double* __restrict__ xw, const double* __restrict__ wext, int r, double (&acc)[SR],
                            double (&wr)[SR]) {
    if (2 * r >= SR - 1) {
        fir_head<XS>(xw, wext, acc, wr);           // t = 0 .. SR-1
        int t0 = SR;
        for (; t0 + SR - 1 <= 2 * r; t0 += SR) fir_full<XS>(xw, wext, t0, acc, wr);
        switch (2 * r - t0 + 1) {                  // full steps left before the closing ramp
            case 0: fir_tail<XS, 0>(xw, wext, t0, acc, wr); break;
            case 1: fir_tail<XS, 1>(xw, wext, t0, acc, wr); break;
            case 2: fir_tail<XS, 2>(xw, wext, t0, acc, wr); break;
            case 3: fir_tail<XS, 3>(xw, wext, t0, acc, wr); break;
            case 4: fir_tail<XS, 4>(xw, wext, t0, acc, wr); break;
            case 5: fir_tail<XS, 5>(xw, wext, t0, acc, wr); break;
            case 6: fir_tail<XS, 6>(xw, wext, t0, acc, wr); break;
            case 7: fir_tail<XS, 7>(xw, wext, t0, acc, wr); break;
            default: fir_tail<XS, 8>(xw, wext, t0, acc, wr); break;
        }
    } else {
        // very short kernels (radius < 4, sigma < 0.9 samples): both bounds checked at run time
        const int steps = 2 * r + SR;
        for (int t0 = 0; t0 < steps; t0 += SR) {
#pragma unroll
            for (int u = 0; u < SR; ++u) {
                const double v = xw[(t0 + u) * XS];
#pragma unroll
                for (int q = 0; q < SR; ++q) {
                    const int j = t0 + u - q;
                    if (j >= 0 && j <= 2 * r) acc[q] = fma(wr[(u + SR - 1 - q) % SR], v, acc[q]);
                }
                wr[u] = wext[t0 + u + SR];
            }
        }
    }
}

}  // namespace sshg
