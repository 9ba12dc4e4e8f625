"""
Shared definitions of the parity cases (used by tests/golden/make_golden.py, which runs
the UNMODIFIED reference in the build container, and by the tests, which run the oracle and
the CUDA path on the same inputs).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CHI2_KEYS = ("xxx", "xyy", "xzz", "xyz", "xxz", "xxy",
             "yxx", "yyy", "yzz", "yyz", "yxz", "yxy",
             "zxx", "zyy", "zzz", "zyz", "zxz", "zxy")
POLS = ("pp", "ps", "sp", "ss")

LSUPER, LSLAB = 180, 142.1885172213904          # reference main.py:32-33
BOHR_NM = 5.2918e-2                             # reference main.py:134


def load_spectra():
    return np.load(os.path.join(GOLDEN_DIR, "si111_spectra.npz"))


def si111_material(spectra=None):
    """The Si(111)(1x1):H material of the reference's main.py:29-75, rebuilt from the raw file
    columns stored in si111_spectra.npz (loadeps / loadshg arithmetic of main.py:14-22)."""
    z = spectra if spectra is not None else load_spectra()
    e = z["energy"]

    def eps(name, scale):
        re, im = z[name]
        return {"energy": e, "data": 1 + (4 * np.pi * ((re + 1j * im) * scale))}

    def shg(name, scale):
        r1, i1, r2, i2 = z[name]
        return {"energy": e, "data": scale * ((r1 + r2) + 1j * (i1 + i2))}

    m1 = {"xx": 1.0, "yy": 1.0, "zz": 1.0}
    m2 = {k: eps("SiH1x1-chi1-" + k, LSUPER / LSLAB) for k in ("xx", "yy", "zz")}
    m3 = {k: eps("SiBulk-chi1-" + k, 1) for k in ("xx", "yy", "zz")}
    chi2 = {k: shg("SiH1x1-chi2-" + k, 1e6 * 1e-24) for k in ("xxx", "xxz", "zxx", "zzz")}
    chi2["xyy"] = {"energy": e, "data": -1 * chi2["xxx"]["data"]}
    chi2["yxy"] = {"energy": e, "data": -1 * chi2["xxx"]["data"]}
    chi2["yyz"] = chi2["xxz"]
    chi2["zyy"] = chi2["zxx"]
    for k in ("xzz", "xyz", "xxy", "yxx", "yyy", "yzz", "yxz", "zyz", "zxz", "zxy"):
        chi2[k] = 0
    return m1, m2, m3, chi2


def full_tensor_material(spectra=None, seed=7):
    """Same media, but all 18 chi2 components non-zero (random complex mixtures of the four
    shipped spectra, a few plain numbers) so that every term of r_pP/pS/sP/sS and every row
    of `rotate` is exercised."""
    m1, m2, m3, chi2 = si111_material(spectra)
    rng = np.random.default_rng(seed)
    base = [chi2[k]["data"] for k in ("xxx", "xxz", "zxx", "zzz")]
    e = chi2["xxx"]["energy"]
    full = {}
    for i, k in enumerate(CHI2_KEYS):
        if i % 7 == 6:
            full[k] = complex(rng.normal(), rng.normal()) * 1e-19     # a scalar entry
        else:
            cw = rng.normal(size=4) + 1j * rng.normal(size=4)
            full[k] = {"energy": e, "data": sum(c * b for c, b in zip(cw, base))}
    return m1, m2, m3, full


def selftest_material(spectra=None):
    """Inputs of the reference's tests/self_test.py (legacy sS calculation), expressed for the
    shgyield() API: axis shifted by +0.01 eV so that the spline returns the exact samples the
    self-test picks by index stride."""
    z = spectra if spectra is not None else load_spectra()
    e = z["energy"] + 0.01
    epsl = z["test-SiH1x1-epsilon"]
    epsb = z["test-SiBulk-epsilon"]
    xxx = z["test-SiH1x1-chi2-xxx"]
    m1 = {"v": 1.0}
    m2 = {"l": {"energy": e, "data": epsl[0] + 1j * epsl[1]}}
    m3 = {"b": {"energy": e, "data": epsb[0] + 1j * epsb[1]}}
    x = {"energy": e, "data": xxx[0] + 1j * xxx[1]}
    chi2 = {k: 0 for k in CHI2_KEYS}
    chi2["xxx"] = x
    chi2["xyy"] = {"energy": e, "data": -1 * x["data"]}
    chi2["yxy"] = {"energy": e, "data": -1 * x["data"]}
    return m1, m2, m3, chi2


E_C1 = np.linspace(1.25, 2.50, 126)             # reference main.py:107

# (name, material, kwargs) -- every case is run through shg.shgyield of the reference.
API_CASES = [
    ("c1_d10", "si111", dict(energy=E_C1, theta=65, phi=30, thick=10, gamma=0, sigma_eps=0.0, sigma_chi=0.0, sigma_out=5)),
    ("c1_dslab", "si111", dict(energy=E_C1, theta=65, phi=30, thick=LSLAB * BOHR_NM, gamma=0, sigma_eps=0.0, sigma_chi=0.0, sigma_out=5)),
    ("c1_defaults", "si111", dict(energy=E_C1, theta=65, phi=30, thick=10)),
    ("c1_nobroad", "si111", dict(energy=E_C1, theta=65, phi=30, thick=10, gamma=0, sigma_out=0.0)),
    ("c1_sigmas", "si111", dict(energy=E_C1, theta=40, phi=75, thick=4.5, gamma=25, sigma_eps=3.0, sigma_chi=7.5, sigma_out=2.0)),
    ("c1_fresnel", "si111", dict(energy=E_C1, theta=65, phi=30, thick=10, gamma=0, sigma_out=5, mode="fresnel")),
    ("c1_2layer", "si111", dict(energy=E_C1, theta=30, phi=10, thick=10, gamma=0, sigma_out=0.0, mode="2-layer")),
    ("c1_nomref", "si111", dict(energy=E_C1, theta=65, phi=30, thick=10, gamma=0, sigma_out=5, mref=False)),
    ("full_g33", "full", dict(energy=E_C1, theta=55, phi=20, thick=12.5, gamma=33, sigma_out=0.0)),
    ("full_g90", "full", dict(energy=np.linspace(0.2, 9.9, 333), theta=12, phi=250, thick=3.0, gamma=90, sigma_out=1.5)),
    ("full_g0", "full", dict(energy=np.linspace(0.05, 4.0, 80), theta=77, phi=133, thick=60.0, gamma=0, sigma_eps=1.0, sigma_out=0.0)),
    ("polar", "si111", dict(energy=1.7, theta=65, phi=np.arange(0, 361), thick=10, gamma=0, sigma_out=0.0)),
    ("polar_full", "full", dict(energy=2.1, theta=45, phi=np.arange(0, 361), thick=10, gamma=40, sigma_out=0.0)),
    ("theta0", "full", dict(energy=E_C1, theta=0, phi=30, thick=10, gamma=0, sigma_out=0.0)),
    ("theta90", "full", dict(energy=E_C1, theta=90, phi=30, thick=10, gamma=0, sigma_out=0.0)),
    ("thick0", "full", dict(energy=E_C1, theta=65, phi=30, thick=0, gamma=0, sigma_out=0.0)),
    ("bcast3d", "full", dict(energy=np.linspace(0.5, 5.0, 40), theta=np.arange(0, 90, 9.0).reshape(10, 1),
                             phi=np.arange(0, 360, 30.0).reshape(12, 1, 1), thick=10, gamma=15, sigma_out=0.0)),
    ("bcast3d_blur", "si111", dict(energy=np.linspace(0.5, 5.0, 40), theta=np.arange(5, 85, 10.0).reshape(8, 1),
                                   phi=np.arange(0, 360, 45.0).reshape(8, 1, 1), thick=10, gamma=0, sigma_out=1.2)),
    ("bcast_thick", "full", dict(energy=np.linspace(0.5, 5.0, 25), theta=65, phi=30,
                                 thick=np.arange(1.0, 31.0).reshape(30, 1), gamma=0, sigma_out=0.0)),
    ("selftest_shape", "selftest", dict(energy=np.linspace(0.01, 1.00, 100), theta=65, phi=30, thick=10, gamma=0,
                                        sigma_eps=0.0, sigma_chi=0.0, sigma_out=0.0)),
    # edge shapes: empty energy axis, everything scalar (0-d result), a 2-D energy array blurred along both axes
    ("empty", "si111", dict(energy=np.array([]), theta=65, phi=30, thick=10, gamma=0, sigma_out=5)),
    ("scalar_all", "si111", dict(energy=1.5, theta=65, phi=30, thick=10, gamma=0, sigma_out=5)),
    ("energy2d_blur", "full", dict(energy=np.array([[1.5, 1.6, 1.7], [1.8, 1.9, 2.0]]), theta=65, phi=30, thick=10,
                                   gamma=10, sigma_out=0.7)),
    # a NaN sample in one spectrum poisons the whole (global) spline -> every output is NaN, no exception
    ("nan_sample", "si111_nan", dict(energy=E_C1, theta=65, phi=30, thick=10, gamma=0, sigma_out=0.0)),
    ("negative_energy_sign", "full", dict(energy=np.linspace(0.5, 3.0, 20), theta=110.0, phi=-75.0, thick=10, gamma=-20,
                                          sigma_out=0.0)),
]


def si111_nan_material(spectra=None):
    m1, m2, m3, chi2 = si111_material(spectra)
    m2 = {k: {"energy": v["energy"], "data": v["data"].copy()} for k, v in m2.items()}
    m2["xx"]["data"][150] = np.nan
    return m1, m2, m3, chi2


MATERIALS = {"si111": si111_material, "full": full_tensor_material, "selftest": selftest_material,
             "si111_nan": si111_nan_material}


# ---- a sweep with the output broadening, produced by the reference angle by angle (ref_blur_rows.npz) ----
BLUR_SWEEP = dict(energy=np.linspace(0.5, 5.0, 451), theta=np.array([0.0, 33.0, 65.0, 88.0]),
                  phi=np.array([0.0, 29.98, 30.0, 77.0, 180.0, 209.98, 210.0, 257.0, 359.5]))
BLUR_SWEEP_SIGMAS = (5.0, 1.5)
BLUR_SWEEP_THICK = 10.0
BLUR_SWEEP_GAMMA = {"si111": 0, "full": 33}


# ---- sampled points of the large grids (configs 3, 4, 5 of BASELINE.json) --------------------
from shgyield_b200.synthetic import config3_axes, config4_axes  # noqa: E402,F401  (one definition for tests and bench)


def sample_indices(shape, n, seed):
    """n random multi-indices of a grid of `shape` plus all its corners."""
    rng = np.random.default_rng(seed)
    idx = np.stack([rng.integers(0, s, n) for s in shape], axis=1)
    corners = np.array(np.meshgrid(*[[0, s - 1] for s in shape], indexing="ij")).reshape(len(shape), -1).T
    return np.unique(np.concatenate([idx, corners]), axis=0)


def parity_ok(a, b, rtol=1e-10, scale=None, rel_floor=1e-6):
    """SURVEY.md section 8c metric, per polarisation: |a-b| <= rtol*max|b| everywhere AND
    |a-b| <= rtol*|b| wherever |b| >= rel_floor*max|b| (1e-6; the fused output broadening K3, whose
    error is absolute -- a few 1e-15 of the column's azimuthal mean -- is held to 1e-4 and to that
    absolute bound).  NaN/Inf must coincide.
    `scale` (optional) = max |yield| over ALL four polarisations of the case: a polarisation
    that is a symmetry node over the whole array (e.g. R_pS of Si(111) at gamma=90, phi=30 is
    1e-52 against 1e-22) is pure rounding noise of an amplitude that cancels to 1e-16, i.e.
    1e-32 of `scale` in the yield; such arrays are compared with atol = rtol*1e-10*scale."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False, "shape %s vs %s" % (a.shape, b.shape)
    fin = np.isfinite(b)
    if not np.array_equal(np.isfinite(a), fin):
        return False, "non-finite pattern differs"
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return False, "NaN pattern differs"
    if not fin.any():
        return True, "all non-finite"
    af, bf = a[fin], b[fin]
    big = np.max(np.abs(bf))
    if scale is not None:
        big = max(big, 1e-10 * float(scale))
    err = np.abs(af - bf)
    if np.max(err) > rtol * big:
        return False, "abs err %.3e > %.1e * max %.3e" % (np.max(err), rtol, big)
    sel = np.abs(bf) >= rel_floor * big
    if sel.any():
        rel = np.max(err[sel] / np.abs(bf[sel]))
        if rel > rtol:
            return False, "rel err %.3e > %.1e" % (rel, rtol)
    return True, "ok"


def case_scale(gold: dict, prefix=""):
    """max finite |yield| over the four polarisations of a golden case."""
    m = 0.0
    for pol in POLS:
        v = np.asarray(gold[prefix + pol], dtype=np.float64)
        v = v[np.isfinite(v)]
        if v.size:
            m = max(m, float(np.max(np.abs(v))))
    return m


def rows_parity_ok(got, want, rtol=1e-10):
    """Row-wise form of the metric for arrays [..., rows, E] of one polarisation: every row that reaches 1e-6 of the
    array's maximum somewhere is held to `parity_ok` against ITS OWN maximum (element-wise rtol down to 1e-6 of the
    row's maximum) -- this is what catches a loss of relative accuracy on rows next to an azimuthal node.  Rows
    that stay below 1e-6 of the array's maximum everywhere (exact symmetry nodes) are outside the element-wise part
    of the metric: they must agree to rtol * array maximum and must not be negative."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    top = float(np.max(np.abs(want[np.isfinite(want)]))) if np.isfinite(want).any() else 0.0
    g2, w2 = got.reshape(-1, got.shape[-1]), want.reshape(-1, want.shape[-1])
    for r in range(w2.shape[0]):
        fin = np.isfinite(w2[r])
        rmax = float(np.max(np.abs(w2[r][fin]))) if fin.any() else 0.0
        if rmax >= 1e-6 * top:
            ok, msg = parity_ok(g2[r], w2[r], rtol=rtol)
            if not ok:
                return False, "row %d (max %.3e of array max %.3e): %s" % (r, rmax, top, msg)
        else:
            if not np.array_equal(np.isfinite(g2[r]), fin):
                return False, "row %d: non-finite pattern differs" % r
            if fin.any() and (np.max(np.abs(g2[r][fin] - w2[r][fin])) > rtol * top or np.min(g2[r][fin]) < 0.0):
                return False, "node row %d: abs err %.3e or negative" % (r, np.max(np.abs(g2[r][fin] - w2[r][fin])))
    return True, "ok"
