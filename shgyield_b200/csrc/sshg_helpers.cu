// The reference's small public helpers as element-wise kernels (sm_100a, FP64 complex):
// wvec / frefs / frefp / ftrans / ftranp (shg.py:137-171) and mrc2w / mrc1w (shg.py:174-198).
// The fused yield kernel does not call these (it shares sub-expressions across them); they exist
// so that every function of the reference's module stays callable with device arithmetic.
#include "sshg_internal.h"
#include "sshg_math.cuh"

using namespace sshg;

namespace {

__global__ void fresnel_kernel(int kind, const double2* __restrict__ ei, const double2* __restrict__ ej,
                               const double* __restrict__ theta, double2* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double st = sin(theta[i]);
    const cplx a = mk(ei[i].x, ei[i].y);
    cplx r;
    if (kind == SSHG_F_WVEC) {
        r = wvec_(a, st);
    } else {
        const cplx b = mk(ej[i].x, ej[i].y);
        r = kind == SSHG_F_FREFS ? frefs_(a, b, st)
          : kind == SSHG_F_FREFP ? frefp_(a, b, st)
          : kind == SSHG_F_FTRANS ? ftrans_(a, b, st) : ftranp_(a, b, st);
    }
    out[i] = make_double2(r.re, r.im);
}

__global__ void multiref_kernel(int harmonic, const double* __restrict__ energy, const double2* __restrict__ eps_l,
                                const double2* __restrict__ f1, const double2* __restrict__ f2,
                                const double* __restrict__ theta, const double* __restrict__ thick,
                                double2* __restrict__ out, long long n, const Physics ph) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cplx r = multiref_(harmonic, energy[i], mk(eps_l[i].x, eps_l[i].y), mk(f1[i].x, f1[i].y),
                             mk(f2[i].x, f2[i].y), sin(theta[i]), thick[i], ph);
    out[i] = make_double2(r.re, r.im);
}

}  // namespace

extern "C" int sshg_fresnel(sshg_ctx* ctx, int kind, const double* epsi, const double* epsj, const double* theta_rad,
                            double* out, int64_t n, int where) {
    SSHG_REQUIRE(ctx && epsi && theta_rad && out, "sshg_fresnel: null argument");
    SSHG_REQUIRE(kind >= SSHG_F_WVEC && kind <= SSHG_F_FTRANP, "sshg_fresnel: unknown kind %d", kind);
    SSHG_REQUIRE(kind == SSHG_F_WVEC || epsj, "sshg_fresnel: epsj is required");
    SSHG_REQUIRE(n > 0, "sshg_fresnel: n must be positive");
    SSHG_CUDA(cudaSetDevice(ctx->device));
    const int in_dev = (where & SSHG_IN_DEVICE) != 0, out_dev = (where & SSHG_OUT_DEVICE) != 0;
    const void *di, *dj = nullptr, *dt;
    if (int rc = sshg_stage_in(ctx, 10, epsi, (size_t)n * 16, in_dev, &di)) return rc;
    if (epsj) {
        if (int rc = sshg_stage_in(ctx, 11, epsj, (size_t)n * 16, in_dev, &dj)) return rc;
    }
    if (int rc = sshg_stage_in(ctx, 12, theta_rad, (size_t)n * 8, in_dev, &dt)) return rc;
    void* dout = out;
    if (!out_dev) {
        if (int rc = sshg_scratch(ctx, 13, (size_t)n * 16, &dout)) return rc;
    }
    fresnel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(kind, (const double2*)di, (const double2*)dj,
                                                                       (const double*)dt, (double2*)dout, n);
    SSHG_CUDA(cudaGetLastError());
    ctx->launches++;
    if (!out_dev) {
        SSHG_CUDA(cudaMemcpyAsync(out, dout, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
        SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return SSHG_OK;
}

extern "C" int sshg_multiref(sshg_ctx* ctx, int harmonic, const double* energy_eV, const double* eps_l,
                             const double* fres1, const double* fres2, const double* theta_rad, const double* thick_nm,
                             const sshg_physics* phys, double* out, int64_t n, int where) {
    SSHG_REQUIRE(ctx && energy_eV && eps_l && fres1 && fres2 && theta_rad && thick_nm && phys && out,
                 "sshg_multiref: null argument");
    SSHG_REQUIRE(harmonic == 1 || harmonic == 2, "sshg_multiref: harmonic must be 1 or 2");
    SSHG_REQUIRE(n > 0, "sshg_multiref: n must be positive");
    SSHG_REQUIRE(phys->h_eVs > 0 && phys->c > 0, "sshg_multiref: h_eVs and c must be positive");
    SSHG_CUDA(cudaSetDevice(ctx->device));
    const int in_dev = (where & SSHG_IN_DEVICE) != 0, out_dev = (where & SSHG_OUT_DEVICE) != 0;
    const void *de, *dl, *d1, *d2, *dt, *dd;
    if (int rc = sshg_stage_in(ctx, 10, energy_eV, (size_t)n * 8, in_dev, &de)) return rc;
    if (int rc = sshg_stage_in(ctx, 11, eps_l, (size_t)n * 16, in_dev, &dl)) return rc;
    if (int rc = sshg_stage_in(ctx, 12, fres1, (size_t)n * 16, in_dev, &d1)) return rc;
    if (int rc = sshg_stage_in(ctx, 13, fres2, (size_t)n * 16, in_dev, &d2)) return rc;
    if (int rc = sshg_stage_in(ctx, 14, theta_rad, (size_t)n * 8, in_dev, &dt)) return rc;
    if (int rc = sshg_stage_in(ctx, 15, thick_nm, (size_t)n * 8, in_dev, &dd)) return rc;
    void* dout = out;
    if (!out_dev) {
        if (int rc = sshg_scratch(ctx, 9, (size_t)n * 16, &dout)) return rc;
    }
    Physics ph{};
    ph.kd = 1e-9 / (phys->h_eVs * phys->c);
    ph.mref = phys->mref;
    ph.sinc_normalized = phys->sinc_normalized;
    multiref_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        harmonic, (const double*)de, (const double2*)dl, (const double2*)d1, (const double2*)d2, (const double*)dt,
        (const double*)dd, (double2*)dout, n, ph);
    SSHG_CUDA(cudaGetLastError());
    ctx->launches++;
    if (!out_dev) {
        SSHG_CUDA(cudaMemcpyAsync(out, dout, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
        SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return SSHG_OK;
}
