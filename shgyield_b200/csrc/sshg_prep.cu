// Per-energy preparation kernels: in-plane rotation of chi2 and cubic-spline resampling.
//
//  * sshg_rotate replaces rotate() of the reference (shgyield/shg.py:63-134).  The reference's
//    18 formulas are the rank-3 tensor rotation chi'_abc = R_ai R_bj R_ck chi_ijk about z by the
//    internal angle g = pi/2 - gamma (shg.py:65) written out on the 18 components with
//    chi_abc = chi_acb folded -- except for one sign (the chi_yyy contribution to chi_xxy,
//    shg.py:92).  Here the 18x18 real matrix is built from R on the host and applied per energy
//    on the device; `xxy_quirk` flips that one entry to reproduce the shipped behaviour.
//
//  * sshg_spline_resample replaces spline/splineC/splineEPS (shg.py:38-55, 423-424):
//    InterpolatedUnivariateSpline(x, y, ext=2) is FITPACK's interpolating cubic spline with
//    knots x[2:-2], i.e. the unique not-a-knot cubic interpolant.  It is computed here from its
//    second derivatives (tridiagonal system, parallel cyclic reduction, one CTA per signal) and
//    evaluated piecewise.
#include "sshg_internal.h"

#include <math.h>
#include <string.h>

namespace {

struct RotMatrix {
    double m[SSHG_NCHI][SSHG_NCHI];
};

// canonical rows: xxx xyy xzz xyz xxz xxy  yxx yyy yzz yyz yxz yxy  zxx zyy zzz zyz zxz zxy
const int kRowAxes[SSHG_NCHI][3] = {
    {0, 0, 0}, {0, 1, 1}, {0, 2, 2}, {0, 1, 2}, {0, 0, 2}, {0, 0, 1},
    {1, 0, 0}, {1, 1, 1}, {1, 2, 2}, {1, 1, 2}, {1, 0, 2}, {1, 0, 1},
    {2, 0, 0}, {2, 1, 1}, {2, 2, 2}, {2, 1, 2}, {2, 0, 2}, {2, 0, 1}};

int row_of(int a, int b, int c) {
    if (b > c) {
        int t = b;
        b = c;
        c = t;
    }
    for (int r = 0; r < SSHG_NCHI; ++r)
        if (kRowAxes[r][0] == a && kRowAxes[r][1] == b && kRowAxes[r][2] == c) return r;
    return -1;
}

__global__ void rotate_kernel(const double2* __restrict__ in, double2* __restrict__ out, long long nE,
                              const RotMatrix M) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nE) return;
    double2 v[SSHG_NCHI];
#pragma unroll
    for (int j = 0; j < SSHG_NCHI; ++j) v[j] = in[j * nE + e];
#pragma unroll 1
    for (int k = 0; k < SSHG_NCHI; ++k) {
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int j = 0; j < SSHG_NCHI; ++j) {
            const double m = M.m[k][j];
            if (m != 0.0) {                    // uniform branch: absent terms must not see Inf/NaN
                re = fma(m, v[j].x, re);
                im = fma(m, v[j].y, im);
            }
        }
        out[k * nE + e] = make_double2(re, im);
    }
}

// ---- spline -------------------------------------------------------------------------------
// Not-a-knot cubic spline through its second derivatives M: a tridiagonal, diagonally dominant system in
// M_1..M_{n-2}, solved by PARALLEL CYCLIC REDUCTION -- ceil(log2 m) levels, every unknown updated by every level, all
// threads of a CTA busy (the Thomas recurrences of round 1 ran on one thread: 0.36 ms per axis + 63 us per launch).
//
// Rows are kept normalised (diagonal 1):  A_i x_{i-s} + x_i + C_i x_{i+s} = D_i.  One level with stride s replaces
// row i by  row_i - A_i row_{i-s} - C_i row_{i+s},  renormalised:
//     beta = 1 - A_i C_{i-s} - C_i A_{i+s},   A' = -A_i A_{i-s} / beta,   C' = -C_i C_{i+s} / beta,
//     D'   = (D_i - A_i D_{i-s} - C_i D_{i+s}) / beta
// (neighbours outside 0..m-1 do not exist: their A_i / C_i is 0).  After L levels (2^L >= m) the rows are decoupled and
// x = D.  A, C and 1/beta depend on the axis only: `spline_pcr_factor_kernel` tabulates them once per axis (cached in
// the context while the same axis keeps coming), and a signal then costs 2 FMAs + 1 multiply per unknown and level.
//
// tab layout, m = n - 2, L levels:  ib0[m] | ih[n-1] | per level l: A[m] | C[m] | ib[m]      (ih[i] = 1 / (x[i+1] - x[i]))
__host__ __device__ inline int pcr_levels(long long m) {
    int L = 0;
    while ((1LL << L) < m) ++L;
    return L;
}
__host__ __device__ inline long long pcr_tab_doubles(long long n) { return (n - 2) + (n - 1) + 3LL * pcr_levels(n - 2) * (n - 2); }

__global__ void __launch_bounds__(1024) spline_pcr_factor_kernel(const double* __restrict__ x, double* __restrict__ tab,
                                                                 long long n) {
    const long long m = n - 2;
    const int L = pcr_levels(m);
    double* ib0 = tab;
    double* ih = tab + m;
    double* lev = ih + (n - 1);
    const double q0 = (x[1] - x[0]) / (x[2] - x[1]);
    const double q1 = (x[n - 1] - x[n - 2]) / (x[n - 2] - x[n - 3]);
    // level 0: the rows of the not-a-knot system, normalised by their diagonal
    for (long long i = threadIdx.x; i < m; i += blockDim.x) {
        const double hl = x[i + 1] - x[i], hr = x[i + 2] - x[i + 1];
        double lower = hl, diag = 2.0 * (hl + hr), upper = hr;
        if (i == 0) {                           // not-a-knot at x_1: M_0 = (1 + q0) M_1 - q0 M_2
            diag += hl * (1.0 + q0);
            upper -= hl * q0;
            lower = 0.0;
        }
        if (i == m - 1) {                       // not-a-knot at x_{n-2}
            diag += hr * (1.0 + q1);
            lower -= hr * q1;
            upper = 0.0;
        }
        if (m == 1) lower = 0.0;
        const double r = 1.0 / diag;
        ib0[i] = r;
        ih[i] = 1.0 / hl;
        if (i == m - 1) ih[i + 1] = 1.0 / hr;
        if (L > 0) {
            lev[i] = lower * r;                 // A of level 0
            lev[m + i] = upper * r;             // C of level 0
        }
    }
    for (int l = 0; l < L; ++l) {
        __syncthreads();                        // one CTA: the global writes of the previous level are visible
        const long long s = 1LL << l;
        double* A = lev + 3LL * l * m;
        double* C = A + m;
        double* ib = C + m;
        double* An = A + 3 * m;                 // level l + 1
        double* Cn = An + m;
        for (long long i = threadIdx.x; i < m; i += blockDim.x) {
            const double a = A[i], c = C[i];
            const bool lo = i - s >= 0, hi = i + s < m;
            const double beta = 1.0 - (lo ? a * C[i - s] : 0.0) - (hi ? c * A[i + s] : 0.0);
            const double r = 1.0 / beta;
            ib[i] = r;
            if (l + 1 < L) {
                An[i] = lo ? -(a * A[i - s]) * r : 0.0;
                Cn[i] = hi ? -(c * C[i + s]) * r : 0.0;
            }
        }
    }
}

// One CTA per signal; D ping-pongs between two buffers of m doubles: shared memory when it fits (`in_smem`), else the
// global scratch `dbuf` [rows][2][m].  y, M: [rows][n].
__global__ void __launch_bounds__(512) spline_pcr_solve_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                               const double* __restrict__ tab, double* __restrict__ M,
                                                               double* __restrict__ dbuf, long long n, int in_smem) {
    extern __shared__ double sp[];
    const long long m = n - 2;
    const int L = pcr_levels(m);
    const double* ib0 = tab;
    const double* ih = tab + m;
    const double* lev = ih + (n - 1);
    const double* yr = y + (long long)blockIdx.x * n;
    double* Mr = M + (long long)blockIdx.x * n;
    double* cur = in_smem ? sp : dbuf + (long long)blockIdx.x * 2 * m;
    double* nxt = cur + m;
    for (long long i = threadIdx.x; i < m; i += blockDim.x) {
        const double y0 = yr[i], y1 = yr[i + 1], y2 = yr[i + 2];
        cur[i] = 6.0 * ((y2 - y1) * ih[i + 1] - (y1 - y0) * ih[i]) * ib0[i];
    }
    for (int l = 0; l < L; ++l) {
        __syncthreads();
        const long long s = 1LL << l;
        const double* A = lev + 3LL * l * m;
        const double* C = A + m;
        const double* ib = C + m;
        for (long long i = threadIdx.x; i < m; i += blockDim.x) {
            double d = cur[i];
            if (i - s >= 0) d = fma(-A[i], cur[i - s], d);
            if (i + s < m) d = fma(-C[i], cur[i + s], d);
            nxt[i] = d * ib[i];
        }
        double* t = cur;
        cur = nxt;
        nxt = t;
    }
    __syncthreads();
    for (long long i = threadIdx.x; i < m; i += blockDim.x) Mr[i + 1] = cur[i];
    if (threadIdx.x == 0) {
        const double q0 = (x[1] - x[0]) / (x[2] - x[1]);
        const double q1 = (x[n - 1] - x[n - 2]) / (x[n - 2] - x[n - 3]);
        const double m1 = cur[0], m2 = m > 1 ? cur[1] : cur[0];
        const double ml = cur[m - 1], mk = m > 1 ? cur[m - 2] : cur[m - 1];
        Mr[0] = (1.0 + q0) * m1 - q0 * m2;
        Mr[n - 1] = (1.0 + q1) * ml - q1 * mk;
    }
}

__global__ void spline_eval_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                   const double* __restrict__ M, long long rows, long long n,
                                   const double* __restrict__ xq, long long nq, double* __restrict__ out,
                                   int* __restrict__ flag) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= rows * nq) return;
    const long long r = g / nq, k = g % nq;
    const double q = xq[k];
    if (!(q >= x[0] && q <= x[n - 1])) {        // also catches NaN
        if (r == 0) atomicOr(flag, 1);
        out[g] = nan("");
        return;
    }
    long long lo = 0, hi = n - 1;               // x[lo] <= q <= x[hi]
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (x[mid] <= q) lo = mid;
        else hi = mid;
    }
    const double h = x[lo + 1] - x[lo];
    const double a = x[lo + 1] - q, b = q - x[lo];
    const double M0 = M[r * n + lo], M1 = M[r * n + lo + 1];
    const double y0 = y[r * n + lo], y1 = y[r * n + lo + 1];
    out[g] = (M0 * a * a * a + M1 * b * b * b) / (6.0 * h) + (y0 / h - M0 * h / 6.0) * a + (y1 / h - M1 * h / 6.0) * b;
}

}  // namespace

static int spline_finish(sshg_ctx* ctx, void* dflag, void* dout, double* out, size_t out_bytes, int out_dev,
                         const char* who);

// FNV-1a over the bytes of a host axis (never 0): key of the cached spline factorisation.
static uint64_t axis_hash_of(const double* x, int64_t n) {
    uint64_t h = 1469598103934665603ull;
    const unsigned char* p = (const unsigned char*)x;
    for (size_t i = 0; i < (size_t)n * 8; ++i) h = (h ^ p[i]) * 1099511628211ull;
    return h ? h : 1;
}

extern "C" int sshg_rotate(sshg_ctx* ctx, const double* chi_in, double* chi_out, int64_t nE, double gamma_deg,
                           int xxy_quirk, int where) {
    SSHG_TRACE("sshg_rotate");
    SSHG_REQUIRE(ctx && chi_in && chi_out, "sshg_rotate: null argument");
    SSHG_REQUIRE(nE > 0, "sshg_rotate: nE must be positive");
    SSHG_REQUIRE(chi_in != chi_out, "sshg_rotate: in-place rotation is not supported");
    SSHG_CUDA(cudaSetDevice(ctx->device));
    const double deg = 3.141592653589793 / 180.0;
    const double g = 90.0 * deg - gamma_deg * deg;        // shg.py:65 with rotang = radians(gamma)
    const double s = sin(g), c = cos(g);
    const double R[3][3] = {{s, -c, 0.0}, {c, s, 0.0}, {0.0, 0.0, 1.0}};
    RotMatrix M;
    memset(&M, 0, sizeof(M));
    for (int k = 0; k < SSHG_NCHI; ++k) {
        const int a = kRowAxes[k][0], b = kRowAxes[k][1], cc = kRowAxes[k][2];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                for (int l = 0; l < 3; ++l) {
                    const double w = R[a][i] * R[b][j] * R[cc][l];
                    if (w != 0.0) M.m[k][row_of(i, j, l)] += w;
                }
    }
    if (xxy_quirk) M.m[5][7] = -M.m[5][7];                // chi_xxy <- chi_yyy, shg.py:92
    const int in_dev = (where & SSHG_IN_DEVICE) != 0, out_dev = (where & SSHG_OUT_DEVICE) != 0;
    const size_t bytes = (size_t)SSHG_NCHI * nE * 16;
    const void* din;
    if (int rc = sshg_stage_in(ctx, 10, chi_in, bytes, in_dev, &din)) return rc;
    void* dout = chi_out;
    if (!out_dev) {
        if (int rc = sshg_scratch(ctx, 11, bytes, &dout)) return rc;
    }
    SSHG_REQUIRE(((uintptr_t)din | (uintptr_t)dout) % 16 == 0, "sshg_rotate: device arrays must be 16-byte aligned");
    rotate_kernel<<<(unsigned)((nE + 127) / 128), 128, 0, ctx->stream>>>((const double2*)din, (double2*)dout, nE, M);
    SSHG_CUDA(cudaGetLastError());
    ctx->launches++;
    if (!out_dev) {
        SSHG_CUDA(cudaMemcpyAsync(chi_out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return SSHG_OK;
}

extern "C" int sshg_spline_resample(sshg_ctx* ctx, const double* x, const double* y, int64_t rows, int64_t n,
                                    const double* xq, int64_t nq, double* out, int where) {
    SSHG_TRACE("sshg_spline_resample");
    SSHG_REQUIRE(ctx && x && y && xq && out, "sshg_spline_resample: null argument");
    SSHG_REQUIRE(rows > 0 && nq > 0, "sshg_spline_resample: rows and nq must be positive");
    SSHG_REQUIRE(n >= 4, "sshg_spline_resample: a cubic interpolating spline needs at least 4 samples (got %lld)",
                 (long long)n);
    SSHG_CUDA(cudaSetDevice(ctx->device));
    const int in_dev = (where & SSHG_IN_DEVICE) != 0, out_dev = (where & SSHG_OUT_DEVICE) != 0;
    if (!in_dev) {
        for (int64_t i = 1; i < n; ++i)
            SSHG_REQUIRE(x[i] > x[i - 1], "sshg_spline_resample: x must be strictly increasing (x[%lld])", (long long)i);
    }
    const void *dx, *dy, *dq;
    if (int rc = sshg_stage_in(ctx, 10, x, (size_t)n * 8, in_dev, &dx)) return rc;
    if (int rc = sshg_stage_in(ctx, 11, y, (size_t)rows * n * 8, in_dev, &dy)) return rc;
    if (int rc = sshg_stage_in(ctx, 12, xq, (size_t)nq * 8, in_dev, &dq)) return rc;
    void *dflag, *dout = out;
    if (int rc = sshg_scratch(ctx, 15, 16, &dflag)) return rc;
    if (!out_dev) {
        if (int rc = sshg_scratch(ctx, 9, (size_t)rows * nq * 8, &dout)) return rc;
    }
    if (int rc = sshg_spline_device(ctx, (const double*)dx, (const double*)dy, rows, n, (const double*)dq, nq,
                                    (double*)dout, (int*)dflag, in_dev ? 0 : axis_hash_of(x, n))) return rc;
    return spline_finish(ctx, dflag, dout, out, (size_t)rows * nq * 8, out_dev, "sshg_spline_resample");
}

// Second derivatives + evaluation, all pointers on the device; *dflag is set when a query is out of range.
int sshg_spline_device(sshg_ctx* ctx, const double* dx, const double* dy, int64_t rows, int64_t n, const double* dq,
                       int64_t nq, double* dout, int* dflag, uint64_t axis_hash) {
    void *dM, *dtab;
    const long long m = n - 2;
    if (int rc = sshg_scratch(ctx, 13, (size_t)rows * n * 8, &dM)) return rc;
    if (int rc = sshg_scratch(ctx, SCR_SPLTAB, (size_t)pcr_tab_doubles(n) * 8, &dtab)) return rc;
    SSHG_CUDA(cudaMemsetAsync(dflag, 0, 4, ctx->stream));
    // the coefficient tables are reused while the same axis (same host bytes) keeps coming
    if (!(axis_hash != 0 && ctx->spl_hash == axis_hash && ctx->spl_n == n && ctx->spl_ptr == dtab)) {
        spline_pcr_factor_kernel<<<1, 1024, 0, ctx->stream>>>(dx, (double*)dtab, n);
        SSHG_CUDA(cudaGetLastError());
        ctx->launches++;
        ctx->spl_hash = axis_hash;
        ctx->spl_n = n;
        ctx->spl_ptr = dtab;
    }
    SSHG_REQUIRE(rows < (1LL << 31), "sshg_spline_resample: too many rows");
    const size_t solve_smem = (size_t)2 * m * 8;
    void* dbuf = nullptr;
    const int in_smem = solve_smem <= 200 * 1024;
    if (in_smem) {
        SSHG_CUDA(cudaFuncSetAttribute(spline_pcr_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    } else {
        if (int rc = sshg_scratch(ctx, 14, (size_t)rows * 2 * m * 8, &dbuf)) return rc;
    }
    spline_pcr_solve_kernel<<<(unsigned)rows, 512, in_smem ? solve_smem : 0, ctx->stream>>>(dx, dy, (const double*)dtab,
                                                                                          (double*)dM, (double*)dbuf, n,
                                                                                          in_smem);
    SSHG_CUDA(cudaGetLastError());
    const long long total = (long long)rows * nq;
    spline_eval_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(dx, dy, (const double*)dM, rows, n, dq,
                                                                               nq, dout, dflag);
    SSHG_CUDA(cudaGetLastError());
    ctx->launches += 2;
    return SSHG_OK;
}

// Reads the range flag (and the result when it goes to the host); one synchronisation.
static int spline_finish(sshg_ctx* ctx, void* dflag, void* dout, double* out, size_t out_bytes, int out_dev,
                         const char* who) {
    int flag = 0;
    SSHG_CUDA(cudaMemcpyAsync(&flag, dflag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (!out_dev) SSHG_CUDA(cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
    SSHG_REQUIRE(flag == 0, "%s: a query lies outside the sampled range [x[0], x[n-1]] (ext=2)", who);
    return SSHG_OK;
}

// broadC -> splineC -> evaluation in one call (reference shg.py:388-389, 423-424): the broadened
// rows stay on the device between K1 and the spline kernels.
extern "C" int sshg_broaden_resample(sshg_ctx* ctx, const double* x, const double* y, int64_t rows, int64_t n,
                                     double sigma, const double* xq, int64_t nq, double* out, int where) {
    SSHG_TRACE("sshg_broaden_resample");
    SSHG_REQUIRE(ctx && x && y && xq && out, "sshg_broaden_resample: null argument");
    SSHG_REQUIRE(rows > 0 && nq > 0, "sshg_broaden_resample: rows and nq must be positive");
    SSHG_REQUIRE(n >= 4, "sshg_broaden_resample: a cubic interpolating spline needs at least 4 samples (got %lld)",
                 (long long)n);
    SSHG_REQUIRE(sigma >= 0.0 && isfinite(sigma), "sshg_broaden_resample: sigma must be finite and >= 0 (got %g)", sigma);
    SSHG_CUDA(cudaSetDevice(ctx->device));
    const int in_dev = (where & SSHG_IN_DEVICE) != 0, out_dev = (where & SSHG_OUT_DEVICE) != 0;
    if (!in_dev) {
        for (int64_t i = 1; i < n; ++i)
            SSHG_REQUIRE(x[i] > x[i - 1], "sshg_broaden_resample: x must be strictly increasing (x[%lld])", (long long)i);
    }
    const void *dx, *dy, *dq;
    if (int rc = sshg_stage_in(ctx, 10, x, (size_t)n * 8, in_dev, &dx)) return rc;
    if (int rc = sshg_stage_in(ctx, 11, y, (size_t)rows * n * 8, in_dev, &dy)) return rc;
    if (int rc = sshg_stage_in(ctx, 12, xq, (size_t)nq * 8, in_dev, &dq)) return rc;
    if (sigma > 1e-15) {
        void* db;
        if (int rc = sshg_scratch(ctx, SCR_BROAD, (size_t)rows * n * 8, &db)) return rc;
        if (int rc = sshg_gauss_device(ctx, (const double*)dy, (double*)db, rows, n, n, 1, sigma)) return rc;
        dy = db;
    }
    void *dflag, *dout = out;
    if (int rc = sshg_scratch(ctx, 15, 16, &dflag)) return rc;
    if (!out_dev) {
        if (int rc = sshg_scratch(ctx, 9, (size_t)rows * nq * 8, &dout)) return rc;
    }
    if (int rc = sshg_spline_device(ctx, (const double*)dx, (const double*)dy, rows, n, (const double*)dq, nq,
                                    (double*)dout, (int*)dflag, in_dev ? 0 : axis_hash_of(x, n))) return rc;
    return spline_finish(ctx, dflag, dout, out, (size_t)rows * nq * 8, out_dev, "sshg_broaden_resample");
}
