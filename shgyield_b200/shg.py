"""
Host-side mirror of the reference's `shgyield/shg.py` for the SSHG yield path.

Same entry point, argument meaning, return dict and error behaviour as the reference
(`shgyield(energy, eps_m1, eps_m2, eps_m3, chi2, theta, phi, thick, gamma=90, sigma_eps=0.0,
sigma_chi=0.0, sigma_out=5.0, mode='3-layer', mref=True)`, reference shg.py:381), but every
numerical stage runs in hand-written sm_100a CUDA kernels behind the C ABI of libsshg.so
(include/sshg.h):

    broad / broadC            -> sshg_gauss1d          (K1, shared-memory tiled FIR)
    spline / splineC / EPS    -> sshg_spline_resample  (not-a-knot cubic spline)
    rotate                    -> sshg_rotate
    wvec .. rad_ss, |.|^2     -> sshg_yield / sshg_yield_points (K2, fused)

The host only splits dicts, averages the (at most three) tensor components of a medium,
validates, and maps NumPy broadcast shapes onto the kernels' layouts.  There is no CPU
fallback: without libsshg.so and a B200 every compute call raises RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib

CHI2_KEYS = ("xxx", "xyy", "xzz", "xyz", "xxz", "xxy",
             "yxx", "yyy", "yzz", "yyz", "yxz", "yxy",
             "zxx", "zyy", "zzz", "zyz", "zxz", "zxy")          # reference shg.py:66-133
POLS = ("pp", "ps", "sp", "ss")

# grids with at least this many points go through the tiled grid kernel; smaller broadcasts
# use the one-thread-per-point kernel (launch-latency bound either way)
GRID_MIN_POINTS = 4096


_CONSTANTS = None


def physical_constants() -> dict:
    """The four constants the reference reads from scipy.constants at run time
    (reference shg.py:181,194,429-430); read from the same SciPy so results do not drift."""
    global _CONSTANTS
    if _CONSTANTS is None:
        from scipy import constants
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _CONSTANTS = {"h_eVs": constants.value("Planck constant in eV s"),
                          "hbar_eVs": constants.value("Planck constant over 2 pi in eV s"),
                          "eps0": constants.epsilon_0, "c": constants.c}
    return dict(_CONSTANTS)


def _physics(consts, mref=True, sinc_normalized=True, zzz_phi_quirk=True, flags=0) -> _lib.Physics:
    c = consts or physical_constants()
    return _lib.Physics(c["h_eVs"], c["hbar_eVs"], c["eps0"], c["c"], int(bool(mref)), int(bool(sinc_normalized)),
                        int(bool(zzz_phi_quirk)), int(flags))


# ---------------------------------------------------------------------------------------------
# K1: Gaussian broadening  (reference shg.py:26-35)
# ---------------------------------------------------------------------------------------------
def _gauss_nd(ctx, arr: np.ndarray, sigma: float) -> np.ndarray:
    """scipy.ndimage.gaussian_filter(arr, sigma) on a C-contiguous float64 array: every axis, one library call (the
    array is uploaded once, filtered axis by axis on the device and copied back once)."""
    out = np.empty_like(arr)
    shape = (C.c_int64 * max(arr.ndim, 1))(*arr.shape)
    _lib.check(ctx.lib.sshg_gauss_nd(ctx.handle, arr.ctypes.data, out.ctypes.data, arr.ndim, shape, float(sigma), _lib.HOST),
               ctx.lib)
    return out


def broad(data, sigma):
    """Gaussian broadening of a real array (sigma in samples, every axis, like the reference's
    scipy.ndimage.gaussian_filter call, shg.py:26-28)."""
    arr = np.asarray(data, dtype=np.float64)
    sigma = float(sigma)
    if arr.ndim == 0 or arr.size == 0 or not sigma > 1e-15:
        return arr.copy()                       # a 0-d input stays 0-d, like scipy's gaussian_filter
    arr = np.ascontiguousarray(arr)
    ctx = _lib.context()
    if arr.ndim == 1:
        out = np.empty_like(arr)
        _lib.check(ctx.lib.sshg_gauss1d(ctx.handle, arr.ctypes.data, out.ctypes.data, 1, arr.size, arr.size, 1, sigma,
                                        _lib.HOST), ctx.lib)
        return out
    return _gauss_nd(ctx, arr, sigma)


def broadC(data, sigma):
    """Gaussian broadening of a complex array, real and imaginary parts separately (shg.py:31-35)."""
    data = np.asarray(data)
    return broad(data.real, sigma) + 1j * broad(data.imag, sigma)


# ---------------------------------------------------------------------------------------------
# spline resampling  (reference shg.py:38-55)
# ---------------------------------------------------------------------------------------------
class _DeviceSpline:
    """Callable stand-in for InterpolatedUnivariateSpline(energy, data, ext=2)."""

    def __init__(self, data, energy):
        self.x = np.ascontiguousarray(energy, dtype=np.float64)
        self.y = np.ascontiguousarray(data, dtype=np.float64)
        if self.x.ndim != 1 or self.x.shape != self.y.shape:
            raise ValueError("x and y should have a same length")

    def __call__(self, xq):
        q = np.asarray(xq, dtype=np.float64)
        return _resample_rows(self.x, self.y[None, :], q.ravel())[0].reshape(q.shape)


def _resample_rows(x, rows, xq):
    """rows [R, n] sampled on x[n] -> [R, nq] at xq; ValueError outside the range (ext=2)."""
    ctx = _lib.context()
    x = np.ascontiguousarray(x, dtype=np.float64)
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    xq = np.ascontiguousarray(xq, dtype=np.float64)
    out = np.empty((rows.shape[0], xq.size))
    if xq.size == 0:
        return out
    _lib.check(ctx.lib.sshg_spline_resample(ctx.handle, x.ctypes.data, rows.ctypes.data, rows.shape[0], x.size,
                                            xq.ctypes.data, xq.size, out.ctypes.data, _lib.HOST), ctx.lib)
    return out


def spline(data, energy):
    return _DeviceSpline(data, energy)


def splineC(data, energy):
    data = np.asarray(data)
    return (_DeviceSpline(data.real, energy), _DeviceSpline(data.imag, energy))


def _resample_group(specs: dict, queries, sigma: float = 0.0):
    """{key: {'energy', 'data'}} -> {key: [complex array per query vector]}.  Every spectrum is
    broadened with `sigma` (samples) and resampled in ONE library call per distinct energy axis
    (sshg_broaden_resample: K1 and the spline kernels chained on the device); spectra that are the
    same object (symmetry aliases such as chi_yyz = chi_xxz) are computed once."""
    out = {k: [None] * len(queries) for k in specs}
    if not specs or not queries:
        return out
    ctx = _lib.context()
    groups = {}
    for k, v in specs.items():
        x = np.ascontiguousarray(v["energy"], dtype=np.float64)
        g = groups.setdefault((x.size, x.tobytes()), {"x": x, "uniq": {}, "keys": {}})
        slot = g["uniq"].setdefault(id(v["data"]), (len(g["uniq"]), v["data"]))[0]
        g["keys"][k] = slot
    sizes = [np.asarray(q).size for q in queries]
    allq = np.ascontiguousarray(np.concatenate([np.asarray(q, dtype=np.float64).ravel() for q in queries]))
    for g in groups.values():
        x = g["x"]
        rows = np.empty((2 * len(g["uniq"]), x.size))
        for slot, data in g["uniq"].values():
            d = np.asarray(data)
            if x.ndim != 1 or d.shape != x.shape:
                raise ValueError("x and y should have a same length")
            rows[2 * slot], rows[2 * slot + 1] = d.real, d.imag
        res = np.empty((rows.shape[0], allq.size))
        if allq.size:
            _lib.check(ctx.lib.sshg_broaden_resample(ctx.handle, x.ctypes.data, rows.ctypes.data, rows.shape[0], x.size,
                                                     float(sigma), allq.ctypes.data, allq.size, res.ctypes.data,
                                                     _lib.HOST), ctx.lib)
        off = 0
        for j, (q, sz) in enumerate(zip(queries, sizes)):
            for k, slot in g["keys"].items():
                out[k][j] = (res[2 * slot, off:off + sz] + 1j * res[2 * slot + 1, off:off + sz]).reshape(np.shape(q))
            off += sz
    return out


def splineEPS(data, energy):
    """Every epsilon component at E and at 2E (reference shg.py:50-55)."""
    energy = np.asarray(energy, dtype=np.float64)
    r = _resample_group(data, [energy, 2 * energy])
    return ({k: v[0] for k, v in r.items()}, {k: v[1] for k, v in r.items()})


def avgEPS(data):
    """Mean over the tensor components of one medium (reference shg.py:58-60); a medium that mixes
    numbers and spectra fails with ValueError exactly like the reference."""
    return np.mean(list(data.values()), axis=0)


# ---------------------------------------------------------------------------------------------
# chi2 rotation  (reference shg.py:63-134)
# ---------------------------------------------------------------------------------------------
def rotate(chi2, rotang, xxy_quirk=True):
    """In-plane rotation; `rotang` in radians as in the reference (which passes radians(gamma))."""
    shape = np.broadcast_shapes(*[np.shape(chi2[k]) for k in CHI2_KEYS])     # KeyError if a key is missing
    n = int(np.prod(shape, dtype=np.int64))
    packed = np.empty((18, n), dtype=np.complex128)
    for i, k in enumerate(CHI2_KEYS):
        packed[i] = np.broadcast_to(np.asarray(chi2[k], dtype=np.complex128), shape).ravel()
    out = _rotate_packed(packed, np.degrees(rotang), xxy_quirk)
    return {k: out[i].reshape(shape) for i, k in enumerate(CHI2_KEYS)}


def _rotate_packed(packed, gamma_deg, xxy_quirk=True):
    ctx = _lib.context()
    out = np.empty_like(packed)
    _lib.check(ctx.lib.sshg_rotate(ctx.handle, packed.ctypes.data, out.ctypes.data, packed.shape[1],
                                   float(gamma_deg), int(bool(xxy_quirk)), _lib.HOST), ctx.lib)
    return out


# ---------------------------------------------------------------------------------------------
# K2: the fused yield kernel
# ---------------------------------------------------------------------------------------------
def yield_points(energy, eidx, theta, phi, thick, eps1w, eps2w, chi2, mref=True, consts=None,
                 sinc_normalized=True, zzz_phi_quirk=True):
    """Independent points: point i uses spectra column eidx[i].  Returns float64 [4, n] (pp, ps, sp, ss)."""
    ctx = _lib.context()
    eidx = np.ascontiguousarray(eidx, dtype=np.int64)
    n = eidx.size
    a = dict(energy=np.ascontiguousarray(energy, dtype=np.float64),
             theta=np.ascontiguousarray(np.broadcast_to(theta, (n,)), dtype=np.float64),
             phi=np.ascontiguousarray(np.broadcast_to(phi, (n,)), dtype=np.float64),
             thick=np.ascontiguousarray(np.broadcast_to(thick, (n,)), dtype=np.float64),
             eps1w=np.ascontiguousarray(eps1w, dtype=np.complex128),
             eps2w=np.ascontiguousarray(eps2w, dtype=np.complex128),
             chi2=np.ascontiguousarray(chi2, dtype=np.complex128))
    nE = a["energy"].size
    if a["eps1w"].shape != (3, nE) or a["eps2w"].shape != (3, nE) or a["chi2"].shape != (18, nE):
        raise ValueError("eps1w/eps2w must be [3, nE] and chi2 [18, nE]")
    out = np.empty((4, n))
    if n == 0:
        return out
    p = _lib.Points(n=n, nE=nE, energy_eV=a["energy"].ctypes.data, eidx=eidx.ctypes.data,
                    theta_deg=a["theta"].ctypes.data, phi_deg=a["phi"].ctypes.data, thick_nm=a["thick"].ctypes.data,
                    eps1w=a["eps1w"].ctypes.data, eps2w=a["eps2w"].ctypes.data, chi2=a["chi2"].ctypes.data,
                    phys=_physics(consts, mref, sinc_normalized, zzz_phi_quirk))
    _lib.check(ctx.lib.sshg_yield_points(ctx.handle, C.byref(p), out.ctypes.data, _lib.HOST), ctx.lib)
    return out


def _prepare(energy, eps_m1, eps_m2, eps_m3, chi2, thick, gamma, sigma_eps, sigma_chi, mode):
    """Reference shg.py:387-426: broaden, resample at E / 2E, average, map the mode, rotate.
    Returns (eps1w [3,nE], eps2w [3,nE], chi_rot [18,nE], thick) for the flattened energies."""
    eps1w, eps2w, packed, thick = _resample_all(energy, eps_m1, eps_m2, eps_m3, chi2, thick, sigma_eps, sigma_chi, mode)
    chi_rot = _rotate_packed(packed, float(gamma)) if packed.shape[1] else packed
    return eps1w, eps2w, chi_rot, thick


def _resample_all(energy, eps_m1, eps_m2, eps_m3, chi2, thick, sigma_eps, sigma_chi, mode):
    """Everything of _prepare() that does not depend on gamma (shg.py:387-424): eps(w), eps(2w) [3, nE] and the
    un-rotated chi2 [18, nE]."""
    energy = np.asarray(energy, dtype=np.float64)
    nE = energy.size
    # all spectra of the three media: one broaden + resample call (queries E and 2E)
    specs = {(m, k): v for m, eps_m in enumerate((eps_m1, eps_m2, eps_m3)) for k, v in eps_m.items()
             if isinstance(v, dict)}
    rs = _resample_group(specs, [energy, 2 * energy], sigma_eps)
    media = []
    for m, eps_m in enumerate((eps_m1, eps_m2, eps_m3)):
        one = {k: rs[(m, k)][0] for k, v in eps_m.items() if isinstance(v, dict)}
        two = {k: rs[(m, k)][1] for k, v in eps_m.items() if isinstance(v, dict)}
        num = {k: v for k, v in eps_m.items() if not isinstance(v, dict)}
        one.update(num)
        two.update(num)
        media.append((one, two))
    if mode == "3-layer":
        e1 = [avgEPS(media[0][0]), avgEPS(media[1][0]), avgEPS(media[2][0])]
        e2 = [avgEPS(media[0][1]), avgEPS(media[1][1]), avgEPS(media[2][1])]
    elif mode == "fresnel" or mode == "2-layer":
        thick = 0
        e1 = [avgEPS(media[0][0]), avgEPS(media[2][0]), avgEPS(media[2][0])]
        e2 = [avgEPS(media[0][1]), avgEPS(media[0][1]), avgEPS(media[2][1])]
    else:
        # the reference fails here by accident (UnboundLocalError, shg.py:405-420); documented deviation
        raise ValueError("unknown mode %r (expected '3-layer', 'fresnel' or '2-layer')" % (mode,))
    eps1w = np.empty((3, nE), dtype=np.complex128)
    eps2w = np.empty((3, nE), dtype=np.complex128)
    for m in range(3):
        eps1w[m] = np.broadcast_to(np.asarray(e1[m], dtype=np.complex128), energy.shape).ravel()
        eps2w[m] = np.broadcast_to(np.asarray(e2[m], dtype=np.complex128), energy.shape).ravel()

    chi_num = {k: v for k, v in chi2.items() if not isinstance(v, dict)}
    chi_spec = {k: v for k, v in chi2.items() if isinstance(v, dict)}
    chi_new = {k: v[0] for k, v in _resample_group(chi_spec, [energy], sigma_chi).items()}
    chi_new.update(chi_num)
    packed = np.empty((18, nE), dtype=np.complex128)
    for i, k in enumerate(CHI2_KEYS):
        packed[i] = np.broadcast_to(np.asarray(chi_new[k], dtype=np.complex128), energy.shape).ravel()   # KeyError like the reference
    return eps1w, eps2w, packed, thick


# ---------------------------------------------------------------------------------------------
# Reuse of the prepared spectra between calls.  The reference re-runs broadening, splines and the rotation on
# every call and lists "avoid running unless changed" as a to-do (shg.py:11-12); its GUI calls shgyield() twice
# per slider tick with the same material (gui.py:286-330).  Here the broadened / resampled / averaged / rotated
# spectra of the last few distinct (material CONTENT, energy values, gamma, sigma_eps, sigma_chi, mode) combinations
# stay resident on the device.  The key is a 64-bit hash of the bytes of every array (sshg_hash64, ~10 GB/s), so an
# array modified in place is a different key: the cache is transparent and on by default.
# ---------------------------------------------------------------------------------------------
_PREPARE_CACHE = None           # OrderedDict: key -> _Prepared (created on first use); False: switched off
_PREPARE_CACHE_SIZE = 8
_RESAMPLE_CACHE: dict = {}      # key without gamma -> (eps1w, eps2w, un-rotated chi2, thick): host arrays


def cache_prepared(enable: bool = True, size: int = 8):
    """Switches the reuse of prepared spectra on (default) or off and sets how many distinct combinations are kept.
    Switching drops (and frees on the device) everything that is cached."""
    global _PREPARE_CACHE, _PREPARE_CACHE_SIZE
    if _PREPARE_CACHE:
        for ent in _PREPARE_CACHE.values():
            ent.release()
    for (cid, _), pool in list(_FREE_BLOCKS.items()):
        for ctx in list(_lib._contexts.values()):
            if id(ctx) == cid and ctx.handle:
                for ptr in pool:
                    ctx.lib.sshg_device_free(ctx.handle, C.c_void_p(ptr))
    _FREE_BLOCKS.clear()
    _PREPARE_CACHE = None if enable else False
    _PREPARE_CACHE_SIZE = max(1, int(size))
    _RESAMPLE_CACHE.clear()


_FREE_BLOCKS: dict = {}         # device blocks of evicted entries, by size: a miss reuses one instead of cudaMalloc / cudaFree


class _Prepared:
    """eps(w), eps(2w) [3, nE] and the rotated chi2 [18, nE] of one prepared combination: host arrays plus one resident
    device block  energy | eps1w | eps2w | chi2  (16-byte aligned parts)."""

    def __init__(self, e_flat, eps1w, eps2w, chi_rot):
        self.eps1w, self.eps2w, self.chi_rot = eps1w, eps2w, chi_rot
        self.nE = int(e_flat.size)
        self.ctx = None
        self.dev = 0
        if self.nE == 0:
            return
        try:
            ctx = _lib.context()
        except RuntimeError:
            return
        pad = (self.nE + 1) // 2 * 2
        nbytes = 8 * pad + 16 * 24 * self.nE
        pool = _FREE_BLOCKS.get((id(ctx), nbytes))
        if pool:
            ptr = C.c_void_p(pool.pop())
        else:
            ptr = C.c_void_p()
            _lib.check(ctx.lib.sshg_device_alloc(ctx.handle, C.byref(ptr), nbytes), ctx.lib)
        self.ctx, self.dev, self.nbytes = ctx, int(ptr.value), nbytes
        block = np.zeros(nbytes // 8)
        block[:self.nE] = e_flat
        block[pad:] = np.concatenate([eps1w.ravel(), eps2w.ravel(), chi_rot.ravel()]).view(np.float64)
        try:
            _lib.check(ctx.lib.sshg_upload(ctx.handle, ptr, block.ctypes.data, nbytes), ctx.lib)
        except Exception:
            self.release()
            raise
        self.d_energy = self.dev
        self.d_eps1w = self.dev + 8 * pad
        self.d_eps2w = self.d_eps1w + 16 * 3 * self.nE
        self.d_chi = self.d_eps2w + 16 * 3 * self.nE

    def release(self, keep: bool = False):
        """Frees the device block, or (keep) parks it for the next entry of the same size."""
        if self.dev and self.ctx is not None and self.ctx.handle:
            pool = _FREE_BLOCKS.setdefault((id(self.ctx), self.nbytes), [])
            if keep and len(pool) < 4:
                pool.append(self.dev)
            else:
                self.ctx.lib.sshg_device_free(self.ctx.handle, C.c_void_p(self.dev))
        self.dev = 0

    def __del__(self):
        try:
            self.release(keep=True)
        except Exception:
            pass


def _content_key(e_arr, mats, gamma, sigma_eps, sigma_chi, mode):
    """Hashable key of everything _prepare() depends on, by CONTENT.  Raises TypeError for inputs it cannot key."""
    lib = _lib.load()
    seen = {}

    def arr_key(a):
        k = seen.get(id(a))
        if k is None:
            b = np.ascontiguousarray(a)
            k = (b.dtype.str, b.shape, int(lib.sshg_hash64(b.ctypes.data, b.nbytes, 0)))
            seen[id(a)] = k
        return k

    parts = [arr_key(e_arr), float(gamma), float(sigma_eps), float(sigma_chi), str(mode)]
    for m in mats:
        row = []
        for name, v in m.items():
            if isinstance(v, dict):
                row.append((name, arr_key(v["energy"]), arr_key(v["data"])))
            elif isinstance(v, np.ndarray):
                row.append((name, arr_key(v)))
            else:
                row.append((name, complex(v)))
        parts.append(tuple(row))
    return tuple(parts)


def _prepare_cached(e_arr, eps_m1, eps_m2, eps_m3, chi2, thick, gamma, sigma_eps, sigma_chi, mode):
    """(_Prepared, effective thick) for this call; prepared spectra are reused by content (see above)."""
    global _PREPARE_CACHE
    key = None
    if _PREPARE_CACHE is not False:
        try:
            key = _content_key(e_arr, (eps_m1, eps_m2, eps_m3, chi2), gamma, sigma_eps, sigma_chi, mode)
        except (TypeError, ValueError, KeyError, AttributeError):
            key = None                          # unusual inputs: prepare afresh (and fail exactly like the reference)
    if key is not None:
        if _PREPARE_CACHE is None:
            from collections import OrderedDict
            _PREPARE_CACHE = OrderedDict()
        hit = _PREPARE_CACHE.get(key)
        if hit is not None:
            _PREPARE_CACHE.move_to_end(key)
            if mode in ("fresnel", "2-layer"):
                thick = 0                       # reference shg.py:414
            return hit, thick
    # a miss: when only gamma is new (the GUI's rotation slider) the broadened / resampled spectra are still there and
    # only the rotation is redone
    rkey = (key[0],) + key[2:] if key is not None else None
    res = _RESAMPLE_CACHE.get(rkey) if rkey is not None else None
    if res is None:
        res = _resample_all(e_arr, eps_m1, eps_m2, eps_m3, chi2, thick, sigma_eps, sigma_chi, mode)
        if rkey is not None:
            _RESAMPLE_CACHE[rkey] = res
            while len(_RESAMPLE_CACHE) > _PREPARE_CACHE_SIZE:
                _RESAMPLE_CACHE.pop(next(iter(_RESAMPLE_CACHE)))
    eps1w, eps2w, packed, thick_eff = res
    if mode not in ("fresnel", "2-layer"):
        thick_eff = thick                       # the cached value belongs to another call; only the mode overrides it
    chi_rot = _rotate_packed(packed, float(gamma)) if packed.shape[1] else packed
    ent = _Prepared(e_arr.ravel(), eps1w, eps2w, chi_rot)
    if key is not None:
        _PREPARE_CACHE[key] = ent
        while len(_PREPARE_CACHE) > _PREPARE_CACHE_SIZE:
            _PREPARE_CACHE.popitem(last=False)[1].release(keep=True)
    return ent, thick_eff


def _axes_of(shape, ndim):
    """Axes (in the ndim-dimensional broadcast result) along which an operand of `shape` varies."""
    pad = (1,) * (ndim - len(shape)) + tuple(shape)
    return tuple(i for i, s in enumerate(pad) if s != 1)


def _points_tail(prep, e_arr, t_arr, p_arr, d_arr, shape, n, sigma_out, mref, consts):
    """Yields of the flattened broadcast + broad(., sigma_out) along every axis in ONE library call
    (sshg_yield_points_broadened; resident spectra, packed upload, CUDA-graph replay).  Returns [4, n]."""
    ctx = _lib.context()
    eidx = np.ascontiguousarray(np.broadcast_to(np.arange(e_arr.size, dtype=np.int64).reshape(e_arr.shape), shape)).ravel()
    bc = lambda a: np.ascontiguousarray(np.broadcast_to(a, shape), dtype=np.float64).ravel()
    th, ph, dd = bc(t_arr), bc(p_arr), bc(d_arr)
    out = np.empty((4, n))
    if prep.dev:
        where = _lib.SPECTRA_DEVICE
        en, e1, e2, ch = prep.d_energy, prep.d_eps1w, prep.d_eps2w, prep.d_chi
    else:
        where = _lib.HOST
        e_flat = np.ascontiguousarray(e_arr, dtype=np.float64).ravel()
        en, e1, e2, ch = e_flat.ctypes.data, prep.eps1w.ctypes.data, prep.eps2w.ctypes.data, prep.chi_rot.ctypes.data
    p = _lib.Points(n=n, nE=e_arr.size, energy_eV=en, eidx=eidx.ctypes.data, theta_deg=th.ctypes.data,
                    phi_deg=ph.ctypes.data, thick_nm=dd.ctypes.data, eps1w=e1, eps2w=e2, chi2=ch,
                    phys=_physics(consts, mref))
    shp = (C.c_int64 * max(len(shape), 1))(*shape)
    _lib.check(ctx.lib.sshg_yield_points_broadened(ctx.handle, C.byref(p), float(sigma_out), len(shape), shp,
                                                   out.ctypes.data, where), ctx.lib)
    return out


def _broadcast_tail(prep, e_arr, t_arr, p_arr, d_arr, ax, shape, sigma_out, mref, consts):
    """Large outer-product broadcasts whose last axis is the energy: [pp, ps, sp, ss] arrays of `shape`, or None when the
    device cannot hold the intermediates (the caller then takes the slab-wise path)."""
    from .grid import _problem
    ctx = _lib.context()
    ndim = len(shape)
    elem = [1] * ndim                               # C-contiguous element strides of the result
    for i in range(ndim - 2, -1, -1):
        elem[i] = elem[i + 1] * shape[i + 1]
    total = int(np.prod(shape, dtype=np.int64))

    def axis_stride(axes):                          # an operand varies along at most one axis here, or is a scalar
        if len(axes) > 1:
            return None
        return elem[axes[0]] if axes else 0
    st = [axis_stride(ax[0]), axis_stride(ax[1]), total, axis_stride(ax[2])]       # theta, thick, pol, phi
    if any(v is None for v in st):
        return None                                 # an operand spread over several axes: not a 1-D axis of a grid
    keep = []
    p = _problem(e_arr.ravel(), t_arr.ravel(), p_arr.ravel(), d_arr.ravel(), prep.eps1w, prep.eps2w, prep.chi_rot,
                 _physics(consts, mref), None, None, keep)
    outs = [np.empty(shape) for _ in range(4)]
    strides = (C.c_int64 * 4)(*st)
    shp = (C.c_int64 * max(ndim, 1))(*shape)
    ptrs = (C.c_void_p * 4)(*[o.ctypes.data for o in outs])
    rc = ctx.lib.sshg_yield_broadcast(ctx.handle, C.byref(p), strides, float(sigma_out), ndim, shp, ptrs)
    if rc == _lib.SSHG_ENOMEM:
        return None
    _lib.check(rc, ctx.lib)
    return outs


def shgyield(energy, eps_m1, eps_m2, eps_m3, chi2, theta, phi, thick, gamma=90, sigma_eps=0.0, sigma_chi=0.0,
             sigma_out=5.0, mode="3-layer", mref=True):
    """
    SSHG yield R_pP, R_pS, R_sP, R_sS (cm^2/W), Eq. (38) of PRB 94, 115314 (2016).
    Drop-in for the reference's shg.shgyield (shg.py:381-437): angles in degrees, `thick` in nm,
    sigmas in samples; energy / theta / phi / thick broadcast like NumPy arrays.
    Returns {'energy': energy, 'phi': phi, 'pp', 'ps', 'sp', 'ss'} (key 'ps' = p in, S out).
    """
    consts = physical_constants()
    e_arr = np.asarray(energy, dtype=np.float64)
    prep, thick = _prepare_cached(e_arr, eps_m1, eps_m2, eps_m3, chi2, thick, gamma, sigma_eps, sigma_chi, mode)
    if not mref:
        thick = 0.0               # the reference never touches `thick` without multiple reflections (shg.py:179, 192),
                                  # so its shape does not enter the broadcast of the result either
    t_arr, p_arr, d_arr = (np.asarray(v, dtype=np.float64) for v in (theta, phi, thick))
    shape = np.broadcast_shapes(e_arr.shape, t_arr.shape, p_arr.shape, d_arr.shape)
    ndim = len(shape)
    n = int(np.prod(shape, dtype=np.int64))
    sigma_out = float(sigma_out)
    stack = None
    if n >= GRID_MIN_POINTS:
        ax = [_axes_of(a.shape, ndim) for a in (t_arr, d_arr, p_arr, e_arr)]       # cube order T, D, P, E
        flat = [a for axes in ax for a in axes]
        if len(set(flat)) == len(flat) and ax[3] == (ndim - 1,) and e_arr.size > 1:
            # outer-product structure with the energy as the last axis (the reference's sweep idiom: energy[E],
            # theta[T, 1], phi[P, 1, 1]): the kernels write straight into the layout of the result, the output
            # broadening of every axis runs on the device, four arrays come back (sshg_yield_broadcast)
            stack = _broadcast_tail(prep, e_arr, t_arr, p_arr, d_arr, ax, shape, sigma_out, mref, consts)
        if stack is None and len(set(flat)) == len(flat):                           # outer-product structure
            from .grid import yield_grid_host
            cube = yield_grid_host(e_arr.ravel(), t_arr.ravel(), p_arr.ravel(), d_arr.ravel(), prep.eps1w, prep.eps2w,
                                   prep.chi_rot, mref=mref, consts=consts)          # [T, D, 4, P, E]
            missing = [i for i in range(ndim) if i not in flat]                     # axes of extent 1
            order = flat + missing
            blocks = [tuple(shape[i] for i in axes) for axes in ax]
            res = []
            for k in range(4):
                r = cube[:, :, k].reshape(sum(blocks, ()) + (1,) * len(missing))
                res.append(np.ascontiguousarray(np.transpose(r, np.argsort(order))).reshape(shape))
            stack = np.ascontiguousarray(np.stack(res))                             # [4, *shape]
            # output broadening (reference shg.py:433-436): scalar sigma blurs every axis, like scipy's gaussian_filter
            if sigma_out > 1e-15 and ndim:
                ctx = _lib.context()
                stack = np.stack([_gauss_nd(ctx, np.ascontiguousarray(stack[k]), sigma_out) for k in range(4)])
    if stack is None:
        if n and ndim <= 8:
            stack = _points_tail(prep, e_arr, t_arr, p_arr, d_arr, shape, n, sigma_out, mref, consts).reshape((4,) + shape)
        elif n:                                     # more than 8 broadcast axes: yields, then the filter
            eidx = np.broadcast_to(np.arange(e_arr.size, dtype=np.int64).reshape(e_arr.shape), shape).ravel()
            bc = lambda a: np.broadcast_to(a, shape).ravel()
            y = yield_points(e_arr.ravel(), eidx, bc(t_arr), bc(p_arr), bc(d_arr), prep.eps1w, prep.eps2w, prep.chi_rot,
                             mref=mref, consts=consts)
            stack = np.stack([broad(y[k].reshape(shape), sigma_out) for k in range(4)])
        else:
            stack = np.empty((4,) + shape)
    out = {"energy": energy, "phi": phi}
    for k, pol in enumerate(POLS):
        # own array per polarisation (0-d stays a 0-d ndarray); the broadcast tail already made four of them
        out[pol] = stack[k] if isinstance(stack, list) else np.array(stack[k])
    return out


# ---------------------------------------------------------------------------------------------
# The reference's small public helpers (shg.py:137-198, 201-378), element-wise on the device.
# shgyield() does not use them -- the fused kernel shares sub-expressions across them -- but every
# function of the reference's module stays callable with the same argument meaning
# (theta / phi in RADIANS here, exactly like the reference's helpers).
# ---------------------------------------------------------------------------------------------
def _elementwise(fn, *args):
    """Broadcasts the arguments, flattens, calls fn(flat arrays, n) -> complex flat, restores the shape."""
    shape = np.broadcast_shapes(*[np.shape(a) for a in args])
    n = int(np.prod(shape, dtype=np.int64))
    out = np.empty(n, dtype=np.complex128)
    if n:
        fn([np.ascontiguousarray(np.broadcast_to(a, shape)).ravel() for a in args], n, out)
    res = out.reshape(shape)
    return res if shape else res[()]


def _fresnel(kind, epsi, epsj, theta):
    ctx = _lib.context()

    def run(flat, n, out):
        ei, ej, th = (np.ascontiguousarray(flat[0], dtype=np.complex128),
                      np.ascontiguousarray(flat[1], dtype=np.complex128),
                      np.ascontiguousarray(flat[2], dtype=np.float64))
        _lib.check(ctx.lib.sshg_fresnel(ctx.handle, kind, ei.ctypes.data, ej.ctypes.data, th.ctypes.data,
                                        out.ctypes.data, n, _lib.HOST), ctx.lib)
    return _elementwise(run, epsi, epsj, theta)


def wvec(eps, theta):
    """w = sqrt(eps - sin(theta)^2)  (reference shg.py:137-141)."""
    return _fresnel(0, eps, eps, theta)


def frefs(epsi, epsj, theta):
    """s-polarised reflection Fresnel factor (reference shg.py:144-148)."""
    return _fresnel(1, epsi, epsj, theta)


def frefp(epsi, epsj, theta):
    """p-polarised reflection Fresnel factor (reference shg.py:151-156)."""
    return _fresnel(2, epsi, epsj, theta)


def ftrans(epsi, epsj, theta):
    """s-polarised transmission Fresnel factor (reference shg.py:159-163)."""
    return _fresnel(3, epsi, epsj, theta)


def ftranp(epsi, epsj, theta):
    """p-polarised transmission Fresnel factor (reference shg.py:166-171)."""
    return _fresnel(4, epsi, epsj, theta)


def _multiref(harmonic, energy, eps0, fres1, fres2, theta, thick, mref):
    ctx = _lib.context()
    phys = _physics(None, mref)

    def run(flat, n, out):
        en, el, f1, f2, th, dd = flat
        arrs = [np.ascontiguousarray(en, dtype=np.float64), np.ascontiguousarray(el, dtype=np.complex128),
                np.ascontiguousarray(f1, dtype=np.complex128), np.ascontiguousarray(f2, dtype=np.complex128),
                np.ascontiguousarray(th, dtype=np.float64), np.ascontiguousarray(dd, dtype=np.float64)]
        _lib.check(ctx.lib.sshg_multiref(ctx.handle, harmonic, arrs[0].ctypes.data, arrs[1].ctypes.data,
                                         arrs[2].ctypes.data, arrs[3].ctypes.data, arrs[4].ctypes.data,
                                         arrs[5].ctypes.data, C.byref(phys), out.ctypes.data, n, _lib.HOST), ctx.lib)
    return _elementwise(run, energy, eps0, fres1, fres2, theta, thick)


def mrc2w(energy, eps0, fres1, fres2, theta, thick, mref):
    """2w multiple-reflection coefficient, Eq. (18) of PRB 94, 115314 (reference shg.py:174-185)."""
    return _multiref(2, energy, eps0, fres1, fres2, theta, thick, mref)


def mrc1w(energy, eps0, fres1, fres2, theta, thick, mref):
    """1w multiple-reflection coefficient, Eq. (21) of PRB 94, 115314 (reference shg.py:188-198)."""
    return _multiref(1, energy, eps0, fres1, fres2, theta, thick, mref)


def _amplitudes(energy, eps1w, eps2w, chi2, theta, phi, thick, mref):
    """Gamma * r for the four polarisations on the NumPy broadcast of all arguments: complex [4, ...]."""
    ctx = _lib.context()
    ops = [energy, theta, phi, thick] + [eps1w[k] for k in ("M1", "M2", "M3")] + \
          [eps2w[k] for k in ("M1", "M2", "M3")] + [chi2[k] for k in CHI2_KEYS]
    shape = np.broadcast_shapes(*[np.shape(o) for o in ops])
    n = int(np.prod(shape, dtype=np.int64))
    flat = lambda a, dt: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=dt), shape)).ravel()
    out = np.empty((4, n), dtype=np.complex128)
    if n:
        # The spectra (energy, 6 eps, 18 chi2) usually vary along ONE axis of the broadcast -- the energy axis -- while the
        # angles vary along the others: keep them as [nE] arrays plus an index per point instead of 25 arrays of n entries.
        ndim = len(shape)
        spec = [energy] + [eps1w[k] for k in ("M1", "M2", "M3")] + [eps2w[k] for k in ("M1", "M2", "M3")] + \
               [chi2[k] for k in CHI2_KEYS]
        spec_axes = sorted({ax for o in spec for ax in _axes_of(np.shape(o), ndim)})
        if len(spec_axes) <= 1:
            nE = shape[spec_axes[0]] if spec_axes else 1
            line = (nE,) if spec_axes else ()

            def along(o, dt):                       # the operand as a 1-D array along the spectral axis
                o = np.asarray(o, dtype=dt)
                return np.ascontiguousarray(np.broadcast_to(o.reshape(-1) if o.size > 1 else o.reshape(()), line)).reshape(nE)
            idx_shape = [1] * ndim
            if spec_axes:
                idx_shape[spec_axes[0]] = nE
            eidx = flat(np.arange(nE, dtype=np.int64).reshape(idx_shape), np.int64)
            a = dict(energy=along(energy, np.float64), eidx=eidx,
                     eps1w=np.stack([along(eps1w[k], np.complex128) for k in ("M1", "M2", "M3")]),
                     eps2w=np.stack([along(eps2w[k], np.complex128) for k in ("M1", "M2", "M3")]),
                     chi2=np.stack([along(chi2[k], np.complex128) for k in CHI2_KEYS]))
        else:                                       # spectra spread over several axes: one entry per point
            nE = n
            a = dict(energy=flat(energy, np.float64), eidx=np.arange(n, dtype=np.int64),
                     eps1w=np.stack([flat(eps1w[k], np.complex128) for k in ("M1", "M2", "M3")]),
                     eps2w=np.stack([flat(eps2w[k], np.complex128) for k in ("M1", "M2", "M3")]),
                     chi2=np.stack([flat(chi2[k], np.complex128) for k in CHI2_KEYS]))
        a.update(theta=flat(theta, np.float64), phi=flat(phi, np.float64), thick=flat(thick, np.float64))
        phys = _physics(None, mref)
        phys.flags = _lib.FLAG_RADIANS
        p = _lib.Points(n=n, nE=nE, energy_eV=a["energy"].ctypes.data, eidx=a["eidx"].ctypes.data,
                        theta_deg=a["theta"].ctypes.data, phi_deg=a["phi"].ctypes.data,
                        thick_nm=a["thick"].ctypes.data, eps1w=a["eps1w"].ctypes.data, eps2w=a["eps2w"].ctypes.data,
                        chi2=a["chi2"].ctypes.data, phys=phys)
        _lib.check(ctx.lib.sshg_amplitudes_points(ctx.handle, C.byref(p), out.ctypes.data, _lib.HOST), ctx.lib)
    res = out.reshape((4,) + shape)
    return res


def rad_pp(energy, eps1w, eps2w, chi2, theta, phi, thick, mref):
    """Gamma_pP r_pP, Eq. (50) of PRB 94, 115314 as shipped (reference shg.py:201-270)."""
    r = _amplitudes(energy, eps1w, eps2w, chi2, theta, phi, thick, mref)[0]
    return r if r.shape else r[()]


def rad_ps(energy, eps1w, eps2w, chi2, theta, phi, thick, mref):
    """Gamma_pS r_pS, Eq. (55) (reference shg.py:273-314)."""
    r = _amplitudes(energy, eps1w, eps2w, chi2, theta, phi, thick, mref)[1]
    return r if r.shape else r[()]


def rad_sp(energy, eps1w, eps2w, chi2, theta, phi, thick, mref):
    """Gamma_sP r_sP, Eq. (60) (reference shg.py:317-352)."""
    r = _amplitudes(energy, eps1w, eps2w, chi2, theta, phi, thick, mref)[2]
    return r if r.shape else r[()]


def rad_ss(energy, eps1w, eps2w, chi2, theta, phi, thick, mref):
    """Gamma_sS r_sS, Eq. (65) (reference shg.py:355-378)."""
    r = _amplitudes(energy, eps1w, eps2w, chi2, theta, phi, thick, mref)[3]
    return r if r.shape else r[()]
