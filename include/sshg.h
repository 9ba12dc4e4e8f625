/*
 * sshg.h -- C ABI of libsshg.so: SSHG yield evaluation on NVIDIA B200 (sm_100a), FP64.
 *
 * Drop-in boundary for ONE path of roguephysicist/SHGYield: `shgyield/shg.py::shgyield`
 * (reference shg.py:381-437) and the helpers it calls.  The reference has no FFI of its own
 * (it is 437 lines of NumPy/SciPy); each entry point below names the reference lines it
 * replaces.  A maintainer binds these with ctypes -- the stub is shown in INTEGRATION.md and
 * shipped as shgyield_b200/_lib.py.
 *
 * Conventions
 *   - plain pointers and sizes only; complex128 arrays are interleaved (re, im) doubles, i.e.
 *     the memory of a C-contiguous NumPy complex128 array;
 *   - every function returns 0 (SSHG_OK) or a negative sshg_status; the message of the last
 *     failure on the calling thread is sshg_last_error();
 *   - `where` says which pointers live on the device: SSHG_HOST (0): all host;
 *     SSHG_IN_DEVICE: inputs are device pointers; SSHG_OUT_DEVICE: `out` is a device pointer;
 *     SSHG_DEVICE = both.  Host calls are synchronous; when `out` is a device pointer the call
 *     is asynchronous and ordered on the context's stream (sshg_ctx_stream / sshg_ctx_sync);
 *   - the caller owns every buffer; the library never keeps a caller pointer after returning;
 *   - a context is bound to one device and is not thread-safe; use one per thread / rank.
 *   - there is NO CPU fallback: without a CUDA device sshg_ctx_create fails with SSHG_ECUDA.
 */
#ifndef SSHG_H_
#define SSHG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSHG_ABI_VERSION 2

#if defined(__GNUC__)
#define SSHG_API __attribute__((visibility("default")))
#else
#define SSHG_API
#endif

typedef enum {
    SSHG_OK = 0,
    SSHG_EINVAL = -1,   /* bad argument (Python shim raises ValueError)  */
    SSHG_ECUDA = -2,    /* CUDA runtime / launch failure (RuntimeError)    */
    SSHG_ENOMEM = -3    /* host or device allocation failed (MemoryError)  */
} sshg_status;

enum { SSHG_HOST = 0, SSHG_IN_DEVICE = 1, SSHG_OUT_DEVICE = 2, SSHG_DEVICE = 3,
       SSHG_SPECTRA_DEVICE = 4 /* sshg_yield_points_broadened: energy_eV, eps1w, eps2w, chi2 are device pointers (spectra
                                  prepared once and kept resident), the point arrays and `out` are host pointers */ };

/* polarisation order of every output: index 0..3 = pp, ps, sp, ss
 * (first letter = incoming 1w field, second = outgoing 2w field; reference shg.py:433-436) */
enum { SSHG_PP = 0, SSHG_PS = 1, SSHG_SP = 2, SSHG_SS = 3 };

/* canonical order of the 18 chi2 rows (reference shg.py:66-133):
 * xxx xyy xzz xyz xxz xxy  yxx yyy yzz yyz yxz yxy  zxx zyy zzz zyz zxz zxy */
#define SSHG_NCHI 18

typedef struct sshg_ctx sshg_ctx;

/* Physical constants and bug-compatibility switches shared by the yield entry points.
 * The reference reads the constants from scipy.constants at run time (shg.py:181,194,429-430);
 * the caller passes the same values so results do not drift with the CODATA edition. */
typedef struct {
    double h_eVs;            /* "Planck constant in eV s"            (shg.py:181,194) */
    double hbar_eVs;         /* "Planck constant over 2 pi in eV s"  (shg.py:429)     */
    double eps0;             /* vacuum permittivity, F/m             (shg.py:430)     */
    double c;                /* speed of light, m/s                  (shg.py:181,430) */
    int32_t mref;            /* 1: multiple reflections on (shg.py:179,192); 0: R^M = r^{lb} */
    int32_t sinc_normalized; /* 1: np.sinc = sin(pi x)/(pi x) as shipped (shg.py:183); 0: sin(x)/x */
    int32_t zzz_phi_quirk;   /* 1: chi_zzz term of r_pP times sin(phi)^3 as shipped (shg.py:269);
                                0: sin(theta)^3 (Eq. 50 of PRB 94, 115314)          */
    int32_t flags;           /* SSHG_FLAG_* bits                                            */
} sshg_physics;

/* sshg_physics.flags */
#define SSHG_FLAG_RADIANS 1  /* theta / phi of a point list are in radians (the reference's helpers) */
#define SSHG_FLAG_STRICT_BROADENING 2  /* sshg_yield_broadened: broaden every yield row itself (yield kernel, then the
                                          Gaussian FIR, slab by slab, ~4.5x slower) instead of broadening the 13
                                          azimuthal harmonics of a column and re-evaluating only the row segments
                                          next to an azimuthal node (the default) */

/* A rectangular sweep: energy x theta x phi x thickness (x 4 polarisations).
 * Replaces the reference's NumPy broadcast over rad_pp/ps/sp/ss + prefactor
 * (shg.py:201-378, 428-436 with sigma_out = 0).
 * Output layout: out[t - t0][d - d0][pol][p][e], double, e contiguous;
 * only the shard t in [t0,t1), d in [d0,d1) is computed and stored. */
typedef struct {
    int64_t nE, nT, nP, nD;
    const double* energy_eV;  /* [nE]                                                   */
    const double* theta_deg;  /* [nT]  angle of incidence                               */
    const double* phi_deg;    /* [nP]  azimuth                                          */
    const double* thick_nm;   /* [nD]  thin-layer thickness                             */
    const double* eps1w;      /* [3][nE][2]  eps(w)  of M1 (vacuum), M2 (layer), M3 (bulk) */
    const double* eps2w;      /* [3][nE][2]  eps(2w)                                    */
    const double* chi2;       /* [18][nE][2] chi2 rows in canonical order, ALREADY rotated */
    sshg_physics phys;
    int64_t t0, t1, d0, d1;   /* shard [t0, t1) x [d0, d1).  t1 = -1 (d1 = -1) selects the full range, and so does
                                 t0 = t1 = 0 (a zero-initialised struct).  Any other empty range (t0 == t1 > 0: the
                                 surplus ranks when there are more ranks than thetas) is legal: nothing is computed
                                 or stored and the call returns SSHG_OK. */
} sshg_problem;

/* A list of independent points (the general NumPy-broadcast shape of shgyield()):
 * point i uses spectra column eidx[i].  Output layout out[pol][i]. */
typedef struct {
    int64_t n;                /* number of points                                        */
    int64_t nE;               /* number of spectra columns                               */
    const double* energy_eV;  /* [nE]                                                   */
    const int64_t* eidx;      /* [n]   column of each point, 0 <= eidx < nE             */
    const double* theta_deg;  /* [n]                                                    */
    const double* phi_deg;    /* [n]                                                    */
    const double* thick_nm;   /* [n]                                                    */
    const double* eps1w;      /* [3][nE][2]                                             */
    const double* eps2w;      /* [3][nE][2]                                             */
    const double* chi2;       /* [18][nE][2], already rotated                           */
    sshg_physics phys;
} sshg_points;

/* ---- library / context ------------------------------------------------------------------- */
SSHG_API int sshg_abi_version(void);
SSHG_API const char* sshg_last_error(void);               /* thread-local, valid until the next call */
SSHG_API int sshg_device_count(int* count);
SSHG_API int sshg_ctx_create(int device, sshg_ctx** out); /* owns one stream + scratch               */
SSHG_API void sshg_ctx_destroy(sshg_ctx* ctx);
/* adopt a caller's cudaStream_t (e.g. torch's current stream); NULL selects the context's own
 * stream again; pass cudaStreamLegacy ((void*)1) to target the legacy default stream */
SSHG_API int sshg_ctx_set_stream(sshg_ctx* ctx, void* cuda_stream);
SSHG_API void* sshg_ctx_stream(sshg_ctx* ctx);
SSHG_API int sshg_ctx_sync(sshg_ctx* ctx);
SSHG_API int64_t sshg_ctx_launches(sshg_ctx* ctx);        /* kernels launched by this context so far */
/* Test and tuning switches of ONE context (nothing in the library reads the environment).  They select
 * between kernels that compute the same result, except "k3_fixup" = 0, which exists to test the error bound
 * of the harmonic sums and is documented as NOT holding the element-wise parity metric.
 * value -1 restores the library's own choice.  Keys:
 *   "k2_sparse"     0 | 1     dense / sparse-phi yield kernel                 (default: by the number of azimuths)
 *   "k2_nt"         64 | 256  CTA width of the dense yield kernel
 *   "k2_recurrence" 0 | 1     phase recurrence on uniform thickness axes      (default 1)
 *   "k3"            0 | 1     output broadening fused into the yield kernel   (default 1 where it applies)
 *   "k3_nt"         64 | 128  CTA width of the fused kernel
 *   "k3_fixup"      0 | 1     strict re-evaluation next to azimuthal nodes    (default 1)   TEST ONLY
 *   "nseg"          1 .. 16   phi segments per column
 *   "k1_generic"    0 | 1     generic instead of fixed-radius FIR kernels     (default 0)
 *   "k1_bulk"       0 | 1     FIR tiles moved by bulk copies (cp.async.bulk)  (default 1 where alignment allows)
 *   "pair_copy"     0 | 1     host-output sweeps into page-locked memory copy the sS rows of a (phi, phi + 180)
 *                             pair once and duplicate them on the host (same bits)   (default 1)
 *   "graphs"        0 | 1     CUDA-graph replay of small point-list calls      (default 1)
 *   "quads"         0 | 1     four rows (phi, 180 - phi, phi + 180, 360 - phi) from one evaluation on grids that
 *                             contain them             (default: fused kernel on short launches and radii > 32)
 * Unknown keys and values outside the listed sets fail with SSHG_EINVAL. */
SSHG_API int sshg_ctx_set_option(sshg_ctx* ctx, const char* key, int64_t value);
SSHG_API int sshg_ctx_get_option(sshg_ctx* ctx, const char* key, int64_t* value);
SSHG_API int sshg_host_alloc(void** ptr, int64_t bytes);  /* page-locked host memory for fast D2H    */
SSHG_API int sshg_host_free(void* ptr);
/* host array -> device memory obtained from sshg_device_alloc (synchronous): prepared spectra that stay resident */
SSHG_API int sshg_upload(sshg_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes);
/* 64-bit content hash of a host buffer (xxHash64 construction, ~10 GB/s): the key under which the Python layer
 * keeps prepared spectra, so that an array modified IN PLACE is noticed (reference to-do, shg.py:11-12) */
SSHG_API uint64_t sshg_hash64(const void* data, int64_t bytes, uint64_t seed);

/* ---- the gather of the yield cube without a collective: peer memory over NVLink ----------------
 * north_star's one exchange step is the gather of the theta slabs onto one GPU.  Instead of computing a
 * slab locally and sending it (ncclSend/ncclRecv), the destination rank allocates the whole cube with
 * sshg_device_alloc, exports it (sshg_ipc_export: 64 opaque bytes that travel to the other ranks of the
 * node by any means), every other rank maps it (sshg_ipc_open) and passes `mapped + slab offset` as the
 * device `out` of sshg_yield / sshg_yield_broadened: the yield kernel's 128-bit stores then go straight
 * over NVLink / NVSwitch into the destination's HBM and the gather disappears into the kernel.
 * (SURVEY.md 8e "B200-native alternative"; the reference itself is single-process NumPy.) */
#define SSHG_IPC_HANDLE_BYTES 64
SSHG_API int sshg_device_alloc(sshg_ctx* ctx, void** ptr, int64_t bytes);   /* cudaMalloc on the context's device */
SSHG_API int sshg_device_free(sshg_ctx* ctx, void* ptr);
SSHG_API int sshg_ipc_export(sshg_ctx* ctx, const void* dev_ptr, unsigned char handle[SSHG_IPC_HANDLE_BYTES]);
SSHG_API int sshg_ipc_open(sshg_ctx* ctx, const unsigned char handle[SSHG_IPC_HANDLE_BYTES], void** dev_ptr);
SSHG_API int sshg_ipc_close(sshg_ctx* ctx, void* dev_ptr);

/* ---- Gaussian broadening: replaces broad / broadC (shg.py:26-35), i.e.
 * scipy.ndimage.gaussian_filter along one axis: taps exp(-k^2/(2 sigma^2))/sum,
 * radius int(4 sigma + 0.5), 'reflect' boundary; sigma <= 1e-15 copies.
 * `rows` signals of `n` samples; sample j of row r is at in[r*row_stride + j*elem_stride]
 * (same addressing for out; in != out).  sigma is in SAMPLES (shg.py:13). */
SSHG_API int sshg_gauss1d(sshg_ctx* ctx, const double* in, double* out, int64_t rows, int64_t n,
                 int64_t row_stride, int64_t elem_stride, double sigma, int where);

/* The same along EVERY axis of a C-contiguous N-D array of extents shape[ndim] -- what broad() does to an N-D result
 * (scipy.ndimage.gaussian_filter(data, sigma), shg.py:26-28): one pass per axis on the device, the array crosses PCIe
 * once each way. */
SSHG_API int sshg_gauss_nd(sshg_ctx* ctx, const double* in, double* out, int32_t ndim, const int64_t* shape, double sigma,
                  int where);

/* ---- in-plane rotation of chi2: replaces rotate (shg.py:63-134).
 * chi_in / chi_out: [18][nE][2]; gamma_deg as passed to shgyield().  xxy_quirk = 1 keeps the
 * reference's sign on the chi_yyy contribution to chi_xxy (shg.py:92). */
SSHG_API int sshg_rotate(sshg_ctx* ctx, const double* chi_in, double* chi_out, int64_t nE,
                double gamma_deg, int xxy_quirk, int where);

/* ---- cubic interpolating (not-a-knot) spline resampling: replaces spline / splineC /
 * splineEPS (shg.py:38-55, 423-424), i.e. InterpolatedUnivariateSpline(x, y, ext=2)(xq).
 * y: [rows][n] sampled on the common axis x[n]; out: [rows][nq].  Queries outside
 * [x[0], x[n-1]] fail with SSHG_EINVAL (ext=2 raises ValueError). */
SSHG_API int sshg_spline_resample(sshg_ctx* ctx, const double* x, const double* y, int64_t rows, int64_t n,
                         const double* xq, int64_t nq, double* out, int where);

/* broadC -> splineC -> evaluation in one call (shg.py:388-389, 423-424): rows are broadened with
 * sigma (samples; <= 1e-15: none) and resampled at xq without leaving the device. */
SSHG_API int sshg_broaden_resample(sshg_ctx* ctx, const double* x, const double* y, int64_t rows, int64_t n,
                          double sigma, const double* xq, int64_t nq, double* out, int where);

/* ---- the fused yield kernel: Fresnel factors + three-layer multiple reflection + chi2
 * contraction + |Gamma r|^2 (shg.py:137-378, 428-436).  */
SSHG_API int sshg_yield(sshg_ctx* ctx, const sshg_problem* p, double* out, int where);
SSHG_API int sshg_yield_points(sshg_ctx* ctx, const sshg_points* p, double* out, int where);
/* The tail of shgyield() for a point list that is the C-order flattening of a NumPy broadcast of extents shape[ndim]
 * (prod(shape) = p->n): the four yields, then broad(., sigma_out) along every axis of the [shape] array
 * (shg.py:428-436); out[pol][n].  `where` = SSHG_HOST, SSHG_SPECTRA_DEVICE, or SSHG_DEVICE.  Host calls with resident
 * spectra (the interactive call shapes of the reference's GUI, gui.py:286-330) upload the point arrays in one copy
 * from page-locked staging memory and replay a CUDA graph (copy in, yield kernel, FIR passes, copy out) that is
 * recorded when a (shape, sigma_out, physics, spectra) combination comes the second time; option "graphs" = 0 disables. */
SSHG_API int sshg_yield_points_broadened(sshg_ctx* ctx, const sshg_points* p, double sigma_out, int32_t ndim,
                                const int64_t* shape, double* out, int where);
/* The same sweep with the output broadening of shgyield() (broad(., sigma_out) on every yield spectrum,
 * shg.py:26-28, 433-436; sigma_out in SAMPLES of the energy axis, 5.0 is the reference's default):
 * out[t][d][pol][p][:] = gaussian_filter(yield[t][d][pol][p][:], sigma_out) with scipy.ndimage's taps,
 * radius int(4 sigma + 0.5) and 'reflect' boundary.  Dense sweeps are broadened inside the yield kernel
 * (the FIR is applied to the 13 azimuthal harmonics of a column before the phi sweep; row segments whose
 * yield is next to an azimuthal node are re-evaluated row-wise by a second, usually idle, kernel: DESIGN.md K3),
 * so the cube is written once; sigma_out = 0 is sshg_yield. */
SSHG_API int sshg_yield_broadened(sshg_ctx* ctx, const sshg_problem* p, double sigma_out, double* out, int where);

/* ---- the reference's public helper functions, element-wise over n values ------------------
 * kind: wvec (shg.py:137-141, epsj unused), frefs / frefp (144-156), ftrans / ftranp (159-171);
 * epsi, epsj, out: complex [n][2]; theta in RADIANS as in the reference. */
enum { SSHG_F_WVEC = 0, SSHG_F_FREFS = 1, SSHG_F_FREFP = 2, SSHG_F_FTRANS = 3, SSHG_F_FTRANP = 4 };
SSHG_API int sshg_fresnel(sshg_ctx* ctx, int kind, const double* epsi, const double* epsj, const double* theta_rad,
                 double* out, int64_t n, int where);
/* mrc2w (harmonic = 2, shg.py:174-185) and mrc1w (harmonic = 1, shg.py:188-198): energy, theta_rad,
 * thick_nm real [n]; eps_l, fres1, fres2, out complex [n][2]. */
SSHG_API int sshg_multiref(sshg_ctx* ctx, int harmonic, const double* energy_eV, const double* eps_l,
                  const double* fres1, const double* fres2, const double* theta_rad, const double* thick_nm,
                  const sshg_physics* phys, double* out, int64_t n, int where);
/* rad_pp / rad_ps / rad_sp / rad_ss (shg.py:201-378): the complex amplitudes Gamma * r of a point
 * list, out complex [4][n][2]; honours SSHG_FLAG_RADIANS. */
SSHG_API int sshg_amplitudes_points(sshg_ctx* ctx, const sshg_points* p, double* out, int where);

/* The tail of shgyield() (shg.py:428-437) for LARGE NumPy broadcasts with outer-product structure -- the reference's
 * only way to express a sweep, e.g. energy[E], theta[T, 1], phi[P, 1, 1].  `p` describes the grid (host arrays, full
 * range); element (t, d, pol, phi, e) of the result is written to offset
 *   pol strides[2] + t strides[0] + d strides[1] + phi strides[3] + e        (doubles; strides[2] = prod(shape))
 * i.e. straight into the C-contiguous broadcast result of every polarisation, whose LAST axis must be the energy;
 * broad(., sigma_out) then runs over EVERY axis of each polarisation's array (scipy.ndimage.gaussian_filter, as the
 * reference does on N-D results) on the device, and the four arrays are copied to out_host[0..3] (pp, ps, sp, ss;
 * prod(shape) doubles each; pageable memory is written by several host threads). */
SSHG_API int sshg_yield_broadcast(sshg_ctx* ctx, const sshg_problem* p, const int64_t strides[4], double sigma_out,
                                  int32_t ndim, const int64_t* shape, double* const out_host[4]);
/* blocking device -> host copy of a large result (ordered after the context's stream); pageable destinations are
 * written by several host threads through page-locked staging buffers */
SSHG_API int sshg_download(sshg_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes);

/* ---- measurement helpers used by bench.py for the roofline denominators ---------------- */
/* streaming 128-bit stores over `bytes` of device memory; returns achieved GB/s (best of iters) */
SSHG_API int sshg_peak_store(sshg_ctx* ctx, int64_t bytes, int iters, double* gbs);
/* device-to-device copy of `bytes` by the bulk-copy engine (cp.async.bulk through shared memory, as K1 stages its
 * tiles); returns achieved GB/s, read + written (best of iters) */
SSHG_API int sshg_peak_copy(sshg_ctx* ctx, int64_t bytes, int iters, double* gbs);
/* register-resident DFMA loop on all SMs; returns achieved FP64 TFLOP/s (best of iters) */
SSHG_API int sshg_peak_dfma(sshg_ctx* ctx, int iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* SSHG_H_ */
