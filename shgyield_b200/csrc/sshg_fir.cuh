// Register-window FIR arithmetic of K1 (gauss_stream_kernel), kept in a header so that
// tests/hostmath_harness.cu can run exactly the same code on the CPU.
//
// A thread produces SR consecutive outputs o = base + q from the samples xw[t] = x[base - r + t],
// t = 0 .. 2r + SR - 1:   out[q] = sum_{j=0..2r} w_full[j] * xw[q + j].
// The loop runs over t; sample t feeds output q with tap j = t - q.  The SR taps a step needs are
// a sliding window over wext[k] = w_full[k - (SR-1)] (zero outside 0..2r) kept in registers,
// rotated statically (the t loop is unrolled by SR).  Taps enter the window two at a time
// (one 128-bit broadcast load per two steps): the kernel keeps two copies of wext, one at a
// 16-byte aligned address (wA) and one 8 bytes further (wB), so that whichever parity the pair
// index has, one of the copies is aligned.
//
// Taps that do not exist (j < 0 or j > 2r) are skipped at compile time rather than multiplied by
// zero, so an Inf/NaN sample never reaches outputs beyond the true radius.
#pragma once
#include <cuda_runtime.h>

#define SSHG_FIR_HD __host__ __device__ __forceinline__

namespace sshg {

constexpr int SR = 13;  // outputs per thread; odd: a stride of 13 doubles between the windows of a warp is
                        // LDS-conflict-free, and (2r + SR) / SR staged samples are read per output

SSHG_FIR_HD void ld_pair(const double* p, double& a, double& b) {
#ifdef __CUDA_ARCH__
    const double2 v = *reinterpret_cast<const double2*>(p);     // 16-byte aligned by construction
    a = v.x;
    b = v.y;
#else
    a = p[0];
    b = p[1];
#endif
}

// One FIR step: sample v = xw[t0 + U] feeds output q with tap j = t - q, for QLO <= q <= QHI (static).
// wr[k % SR] holds wext[k] for the current window k in [t, t + SR - 1]; after the step the tap
// wext[t + SR] replaces the one that left.  `wsel` is the copy of wext whose element t0 + SR is
// 16-byte aligned; `pend` carries the second tap of a pair from an even step to the next odd one.
template <int U, int QLO, int QHI>
SSHG_FIR_HD void fir_step(const double* __restrict__ xw, const double* __restrict__ wsel, int t0, double (&acc)[SR],
                          double (&wr)[SR], double& pend) {
    const double v = xw[t0 + U];
#pragma unroll
    for (int q = 0; q < SR; ++q)
        if (q >= QLO && q <= QHI) acc[q] = fma(wr[(U + 2 * SR - 1 - q) % SR], v, acc[q]);
    if constexpr (U % 2 == 0) {
        double first;
        ld_pair(wsel + t0 + U + SR, first, pend);
        wr[U % SR] = first;
    } else {
        wr[U % SR] = pend;
    }
}

// The last L full steps (t <= 2r) and the closing ramp t = 2r + s, s = 1 .. SR-1, where only the
// outputs q >= s still own a tap.  Everything is static once L = 2r - t0 + 1 is known.
template <int L, int U = 0>
SSHG_FIR_HD void fir_tail(const double* __restrict__ xw, const double* __restrict__ wsel, int t0, double (&acc)[SR],
                          double (&wr)[SR], double& pend) {
    if constexpr (U < L + SR - 1) {
        constexpr int S = U - L + 1;               // <= 0: full step
        fir_step<U, (S > 0 ? S : 0), SR - 1>(xw, wsel, t0, acc, wr, pend);
        fir_tail<L, U + 1>(xw, wsel, t0, acc, wr, pend);
    }
}

template <int L = 0>
SSHG_FIR_HD void fir_tail_dispatch(int l, const double* __restrict__ xw, const double* __restrict__ wsel, int t0,
                                   double (&acc)[SR], double (&wr)[SR], double& pend) {
    if constexpr (L < SR) {
        if (l == L) fir_tail<L>(xw, wsel, t0, acc, wr, pend);
        else fir_tail_dispatch<L + 1>(l, xw, wsel, t0, acc, wr, pend);
    }
}

template <int U = 0>
SSHG_FIR_HD void fir_head(const double* __restrict__ xw, const double* __restrict__ wsel, double (&acc)[SR],
                          double (&wr)[SR], double& pend) {
    if constexpr (U < SR) {                        // opening ramp: t = U, outputs q <= U
        fir_step<U, 0, U>(xw, wsel, 0, acc, wr, pend);
        fir_head<U + 1>(xw, wsel, acc, wr, pend);
    }
}

template <int U = 0>
SSHG_FIR_HD void fir_full(const double* __restrict__ xw, const double* __restrict__ wsel, int t0, double (&acc)[SR],
                          double (&wr)[SR], double& pend) {
    if constexpr (U < SR) {
        fir_step<U, 0, SR - 1>(xw, wsel, t0, acc, wr, pend);
        fir_full<U + 1>(xw, wsel, t0, acc, wr, pend);
    }
}

// Number of tap slots the window may touch: wext[k], k < fir_ntap(r).
SSHG_FIR_HD int fir_ntap(int r) { return 2 * r + 2 + 3 * SR; }

// The whole window.  wA / wB: the two copies of wext (wA 16-byte aligned, wB 8 bytes further; on
// the host both may be the same array).  acc[] is overwritten.
SSHG_FIR_HD void fir_window(const double* __restrict__ xw, const double* __restrict__ wA,
                            const double* __restrict__ wB, int r, double (&acc)[SR]) {
    double wr[SR], pend = 0.0;
#pragma unroll
    for (int q = 0; q < SR; ++q) {
        acc[q] = 0.0;
        wr[q] = wA[q];
    }
    if (2 * r >= SR - 1) {
        // element t0 + SR of the chosen copy must be 16-byte aligned: even index -> wA, odd -> wB
        fir_head(xw, (SR & 1) ? wB : wA, acc, wr, pend);        // t = 0 .. SR-1
        int t0 = SR;
        for (; t0 + SR - 1 <= 2 * r; t0 += SR) fir_full(xw, ((t0 + SR) & 1) ? wB : wA, t0, acc, wr, pend);
        fir_tail_dispatch(2 * r - t0 + 1, xw, ((t0 + SR) & 1) ? wB : wA, t0, acc, wr, pend);
    } else {
        // very short kernels (radius < (SR-1)/2): both bounds checked at run time
        const int steps = 2 * r + SR;
        for (int t0 = 0; t0 < steps; t0 += SR) {
#pragma unroll
            for (int u = 0; u < SR; ++u) {
                const double v = xw[t0 + u];
#pragma unroll
                for (int q = 0; q < SR; ++q) {
                    const int j = t0 + u - q;
                    if (j >= 0 && j <= 2 * r) acc[q] = fma(wr[(u + SR - 1 - q) % SR], v, acc[q]);
                }
                wr[u] = wA[t0 + u + SR];
            }
        }
    }
}

}  // namespace sshg
