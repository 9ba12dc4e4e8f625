"""
TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

CPU (NumPy) restatement of the SSHG yield path of roguephysicist/SHGYield
(`shgyield/shg.py`, 437 lines of NumPy/SciPy) plus the two SciPy routines that
path reaches (`scipy.ndimage.gaussian_filter`, SciPy 1.18.1, and
`scipy.interpolate.InterpolatedUnivariateSpline(k=3, ext=2)`, FITPACK).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module, and only as the
checker / the timed CPU baseline.  The product (`shgyield_b200/`) never imports
it and has no CPU fallback.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks this file against
  * `tests/golden/si111_reference_out.npz`  (the reference's own golden
    `example/reference/si111-reference.out`, 126 x 4 yields),
  * `tests/golden/selftest_reference.npz`   (the `REFERENCE` array of the
    reference's `tests/self_test.py:115-149`),
  * `tests/golden/ref_api_cases.npz`, `ref_helpers.npz`, `ref_grid_samples.npz`,
    `ref_blur_rows.npz`                     (outputs of the unmodified reference
    imported from /root/reference by `tests/golden/make_golden.py`: shgyield()
    on 25 call shapes, the helper functions, sampled points of configs 3-5, and
    a sweep broadened along the energy axis angle by angle).

The restatement is table driven (one table row per term of the reference's
formulas) rather than a transcription; every table cites the lines it follows.
All `file:line` citations are relative to the reference checkout.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# constants: the reference reads them from scipy.constants at run time
# (shg.py:181,194,429-430).  SciPy 1.18.1 values, used when the caller does not
# pass its own.  (h, hbar are the legacy CODATA-2014 keys the reference names.)
# ----------------------------------------------------------------------------
SCIPY_1_18_CONSTANTS = {
    "h_eVs": 4.135667662e-15,       # "Planck constant in eV s"
    "hbar_eVs": 6.582119514e-16,    # "Planck constant over 2 pi in eV s"
    "eps0": 8.8541878188e-12,       # constants.epsilon_0 (CODATA 2022)
    "c": 299792458.0,               # constants.c
}

CHI2_KEYS = ("xxx", "xyy", "xzz", "xyz", "xxz", "xxy",
             "yxx", "yyy", "yzz", "yyz", "yxz", "yxy",
             "zxx", "zyy", "zzz", "zyz", "zxz", "zxy")     # shg.py:66-133


# ----------------------------------------------------------------------------
# Gaussian broadening: scipy.ndimage.gaussian_filter (order 0, truncate 4,
# mode 'reflect'), SciPy 1.18.1 scipy/ndimage/_filters.py:656-669 (taps),
# :688-754 (1-D filter) and :758-861 (N-D driver), C loop NI_Correlate1D
# (symmetric branch).  Called by shg.py:26-28 (broad) and :31-35 (broadC).
# ----------------------------------------------------------------------------
def gaussian_taps(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """Normalised taps w[-r..r]; r = int(truncate*sigma + 0.5)."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    x = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (sd * sd) * x ** 2)
    return w / w.sum()


def _reflect_index(idx: np.ndarray, n: int) -> np.ndarray:
    """'reflect' = half-sample symmetric (d c b a | a b c d | d c b a), repeated."""
    m = np.mod(idx, 2 * n)
    return np.where(m >= n, 2 * n - 1 - m, m)


def gaussian_filter1d(data, sigma: float, axis: int = -1) -> np.ndarray:
    data = np.asarray(data, dtype=np.float64)
    w = gaussian_taps(sigma)
    r = (len(w) - 1) // 2
    x = np.moveaxis(data, axis, -1)
    n = x.shape[-1]
    i = np.arange(n)
    out = x * w[r]                                   # centre tap first
    for j in range(r, 0, -1):                        # farthest pair first
        lo = x[..., _reflect_index(i - j, n)]
        hi = x[..., _reflect_index(i + j, n)]
        out = out + (lo + hi) * w[r - j]
    return np.moveaxis(out, -1, axis)


def gaussian_filter(data, sigma: float) -> np.ndarray:
    """Scalar sigma is applied along EVERY axis (SciPy _filters.py:843-861)."""
    out = np.array(data, dtype=np.float64, copy=True)
    if not sigma > 1e-15:
        return out
    for ax in range(out.ndim):
        out = gaussian_filter1d(out, sigma, axis=ax)
    return out


def broad(data, sigma):                              # shg.py:26-28
    return gaussian_filter(data, sigma)


def broadC(data, sigma):                             # shg.py:31-35
    data = np.asarray(data)
    return gaussian_filter(data.real, sigma) + 1j * gaussian_filter(data.imag, sigma)


# ----------------------------------------------------------------------------
# Spline resampling: InterpolatedUnivariateSpline(x, y, ext=2) is the cubic
# spline that interpolates every sample with FITPACK's knot choice
# t = x[2:-2] (+ 4-fold end knots), i.e. the *not-a-knot* cubic spline.  That
# interpolant is unique, so it is restated here through its second derivatives
# (tridiagonal solve).  Called by shg.py:38-55 and :423-424.
# ----------------------------------------------------------------------------
class NotAKnotSpline:
    def __init__(self, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        n = x.size
        if n < 4:
            raise ValueError("cubic interpolating spline needs at least 4 samples")
        h = np.diff(x)
        if not np.all(h > 0):
            raise ValueError("x must be strictly increasing")
        slope = np.diff(y) / h
        m = n - 2                                    # unknowns M_1 .. M_{n-2}
        lower = h[:-1].copy()                        # couples M_{i-1}
        diag = 2.0 * (h[:-1] + h[1:])
        upper = h[1:].copy()                         # couples M_{i+1}
        rhs = 6.0 * (slope[1:] - slope[:-1])
        # not-a-knot at x_1:  M_0 = (1 + h0/h1) M_1 - (h0/h1) M_2
        q0 = h[0] / h[1]
        diag[0] += h[0] * (1.0 + q0)
        upper[0] -= h[0] * q0
        # not-a-knot at x_{n-2}
        q1 = h[-1] / h[-2]
        diag[-1] += h[-1] * (1.0 + q1)
        lower[-1] -= h[-1] * q1
        # Thomas algorithm
        cp = np.empty(m)
        dp = np.empty(m)
        cp[0] = upper[0] / diag[0]
        dp[0] = rhs[0] / diag[0]
        for i in range(1, m):
            den = diag[i] - lower[i] * cp[i - 1]
            cp[i] = upper[i] / den
            dp[i] = (rhs[i] - lower[i] * dp[i - 1]) / den
        M = np.empty(n)
        M[m] = dp[m - 1]
        for i in range(m - 2, -1, -1):
            M[i + 1] = dp[i] - cp[i] * M[i + 2]
        M[0] = (1.0 + q0) * M[1] - q0 * M[2]
        M[-1] = (1.0 + q1) * M[-2] - q1 * M[-3]
        self.x, self.y, self.h, self.M = x, y, h, M

    def __call__(self, xq):
        xq = np.asarray(xq, dtype=np.float64)
        x, y, h, M = self.x, self.y, self.h, self.M
        if xq.size and (np.min(xq) < x[0] or np.max(xq) > x[-1]):
            raise ValueError("query outside the sampled range (ext=2)")   # shg.py:40
        i = np.clip(np.searchsorted(x, xq, side="right") - 1, 0, x.size - 2)
        hi = h[i]
        a = x[i + 1] - xq
        b = xq - x[i]
        return (M[i] * a ** 3 + M[i + 1] * b ** 3) / (6.0 * hi) \
            + (y[i] / hi - M[i] * hi / 6.0) * a + (y[i + 1] / hi - M[i + 1] * hi / 6.0) * b


def resample_complex(spec: dict, energy) -> np.ndarray:       # splineC + eval, shg.py:43-47,53-54
    re = NotAKnotSpline(spec["energy"], np.asarray(spec["data"]).real)(energy)
    im = NotAKnotSpline(spec["energy"], np.asarray(spec["data"]).imag)(energy)
    return re + 1j * im


# ----------------------------------------------------------------------------
# File loaders (main.py:14-22)
# ----------------------------------------------------------------------------
def loadeps(infile, scale):                          # main.py:14-17
    e, re, im = np.loadtxt(infile, unpack=True)
    return {"energy": e, "data": 1 + (4 * np.pi * ((re + 1j * im) * scale))}


def loadshg(infile, scale):                          # main.py:19-22
    e, r1, i1, r2, i2 = np.loadtxt(infile, unpack=True)
    return {"energy": e, "data": scale * ((r1 + r2) + 1j * (i1 + i2))}


# ----------------------------------------------------------------------------
# In-plane rotation of the 18 chi2 components, shg.py:63-134.
# Row = output key; entry = (multiplier, power of sin, power of cos, source key)
# with the internal angle g = pi/2 - rotang (shg.py:65).  'C2' marks the one
# -cos(2g) coefficient (shg.py:132).  NOTE the reference's 'xxy' row carries
# -sin*cos^2 on 'yyy' (shg.py:92) where a true rotation has +; kept.
# ----------------------------------------------------------------------------
_ROT = {
    "xxx": [(1, 3, 0, "xxx"), (1, 1, 2, "xyy"), (-2, 2, 1, "xxy"), (-1, 2, 1, "yxx"), (-1, 0, 3, "yyy"), (2, 1, 2, "yxy")],
    "xyy": [(1, 1, 2, "xxx"), (1, 3, 0, "xyy"), (2, 2, 1, "xxy"), (-1, 0, 3, "yxx"), (-1, 2, 1, "yyy"), (-2, 1, 2, "yxy")],
    "xzz": [(1, 1, 0, "xzz"), (-1, 0, 1, "yzz")],
    "xyz": [(1, 2, 0, "xyz"), (1, 1, 1, "xxz"), (-1, 1, 1, "yyz"), (-1, 0, 2, "yxz")],
    "xxz": [(-1, 1, 1, "xyz"), (1, 2, 0, "xxz"), (1, 0, 2, "yyz"), (-1, 1, 1, "yxz")],
    "xxy": [(1, 2, 1, "xxx"), (-1, 2, 1, "xyy"), (1, 3, 0, "xxy"), (-1, 1, 2, "xxy"), (-1, 1, 2, "yxx"), (-1, 1, 2, "yyy"), (1, 0, 3, "yxy"), (-1, 2, 1, "yxy")],
    "yxx": [(1, 2, 1, "xxx"), (1, 0, 3, "xyy"), (-2, 1, 2, "xxy"), (1, 3, 0, "yxx"), (1, 1, 2, "yyy"), (-2, 2, 1, "yxy")],
    "yyy": [(1, 0, 3, "xxx"), (1, 2, 1, "xyy"), (2, 1, 2, "xxy"), (1, 1, 2, "yxx"), (1, 3, 0, "yyy"), (2, 2, 1, "yxy")],
    "yzz": [(1, 0, 1, "xzz"), (1, 1, 0, "yzz")],
    "yyz": [(1, 1, 1, "xyz"), (1, 0, 2, "xxz"), (1, 2, 0, "yyz"), (1, 1, 1, "yxz")],
    "yxz": [(-1, 0, 2, "xyz"), (1, 1, 1, "xxz"), (-1, 1, 1, "yyz"), (1, 2, 0, "yxz")],
    "yxy": [(1, 1, 2, "xxx"), (-1, 1, 2, "xyy"), (-1, 0, 3, "xxy"), (1, 2, 1, "xxy"), (1, 2, 1, "yxx"), (-1, 2, 1, "yyy"), (1, 3, 0, "yxy"), (-1, 1, 2, "yxy")],
    "zxx": [(1, 2, 0, "zxx"), (1, 0, 2, "zyy"), (-2, 1, 1, "zxy")],
    "zyy": [(1, 0, 2, "zxx"), (1, 2, 0, "zyy"), (2, 1, 1, "zxy")],
    "zzz": [(1, 0, 0, "zzz")],
    "zyz": [(1, 1, 0, "zyz"), (1, 0, 1, "zxz")],
    "zxz": [(-1, 0, 1, "zyz"), (1, 1, 0, "zxz")],
    "zxy": [(1, 1, 1, "zxx"), (-1, 1, 1, "zyy"), ("C2", 0, 0, "zxy")],
}


def rotate(chi2: dict, rotang: float) -> dict:
    g = np.radians(90) - rotang
    s, c = np.sin(g), np.cos(g)
    out = {}
    for key, row in _ROT.items():
        acc = 0
        for mult, ps, pc, src in row:
            coef = -np.cos(2 * g) if mult == "C2" else mult * s ** ps * c ** pc
            acc = acc + coef * chi2[src]
        out[key] = acc
    return out


# ----------------------------------------------------------------------------
# Wave vectors, Fresnel factors, multiple-reflection coefficients
# ----------------------------------------------------------------------------
def wvec(eps, theta):                                # shg.py:137-141
    return np.sqrt(eps - np.sin(theta) ** 2)


def fresnel_r(pol, ei, ej, theta):                   # shg.py:144-156
    wi, wj = wvec(ei, theta), wvec(ej, theta)
    if pol == "s":
        return (wi - wj) / (wi + wj)
    return (wi * ej - wj * ei) / (wi * ej + wj * ei)


def fresnel_t(pol, ei, ej, theta):                   # shg.py:159-171
    wi, wj = wvec(ei, theta), wvec(ej, theta)
    if pol == "s":
        return 2 * wi / (wi + wj)
    return 2 * wi * np.sqrt(ei * ej) / (wi * ej + wj * ei)   # sqrt of the PRODUCT (shg.py:170)


def _sinc(x, normalized=True):
    """np.sinc = sin(pi x)/(pi x), 0 -> 1e-20 (numpy/lib/_function_base_impl.py:3849-3854);
    normalized=False is the sin(x)/x the reference's self_test.py:87 uses."""
    x = np.asanyarray(x)
    scale = np.pi if normalized else 1.0
    y = scale * np.where(x == 0, 1.0e-20, x)
    return np.sin(y) / y


def multiref_2w(energy, eps_l, r_lb, r_vl, theta, thick, mref, consts, sinc_normalized=True):   # shg.py:174-185
    if not mref:
        return r_lb
    delta = 8 * np.pi * ((energy * thick * 1e-9) / (consts["h_eVs"] * consts["c"])) * wvec(eps_l, theta)
    return (r_lb * np.exp(1j * (delta / 2))) / (1 + r_vl * r_lb * np.exp(1j * delta)) \
        * _sinc(delta / 2, sinc_normalized)


def multiref_1w(energy, eps_l, r_lb, r_vl, theta, thick, mref, consts):                          # shg.py:188-198
    if not mref:
        return r_lb
    varphi = 4 * np.pi * ((energy * thick * 1e-9) / (consts["h_eVs"] * consts["c"])) * wvec(eps_l, theta)
    return (r_lb * np.exp(1j * varphi)) / (1 + r_vl * r_lb * np.exp(1j * varphi))


# ----------------------------------------------------------------------------
# r_pP, r_pS, r_sP, r_sS -- one row per term of shg.py:218-269, :290-313,
# :334-351, :372-377.
#   row = (multiplier, out, inc, power of sin(theta), power of sin(phi),
#          power of cos(phi), chi key)
#   out: 'm' -> (1 - F2) * W_l      (x/y component of the 2w field, P out)
#        'p' -> (1 + F2)            (z component, P out)
#        '1' -> 1                   (S out)
#   inc: 'aa' -> (1 - f1)^2 w_l^2,  'ab' -> (1 + f1)(1 - f1) w_l,
#        'bb' -> (1 + f1)^2,        '1' -> 1 (s in)
# The last row of PP is the chi_zzz term that the reference multiplies by
# sin(phi)^3 (shg.py:269) where the theory has sin(theta)^3; `zzz_phi_quirk`
# selects which.
# ----------------------------------------------------------------------------
_PP = [
    (-1, "m", "aa", 0, 0, 3, "xxx"), (-2, "m", "aa", 0, 1, 2, "xxy"), (-2, "m", "ab", 1, 0, 2, "xxz"),
    (-1, "m", "aa", 0, 2, 1, "xyy"), (-2, "m", "ab", 1, 1, 1, "xyz"), (-1, "m", "bb", 2, 0, 1, "xzz"),
    (-1, "m", "aa", 0, 1, 2, "yxx"), (-2, "m", "aa", 0, 2, 1, "yxy"), (-2, "m", "ab", 1, 1, 1, "yxz"),
    (-1, "m", "aa", 0, 3, 0, "yyy"), (-2, "m", "ab", 1, 2, 0, "yyz"), (-1, "m", "bb", 2, 1, 0, "yzz"),
    (1, "p", "aa", 1, 0, 2, "zxx"), (2, "p", "ab", 2, 0, 1, "zxz"), (2, "p", "aa", 1, 1, 1, "zxy"),
    (1, "p", "aa", 1, 2, 0, "zyy"), (2, "p", "ab", 2, 1, 0, "zyz"),
]
_PP_ZZZ_QUIRK = (1, "p", "bb", 0, 3, 0, "zzz")       # shg.py:269 as shipped
_PP_ZZZ_THEORY = (1, "p", "bb", 3, 0, 0, "zzz")      # Eq. (50) of PRB 94, 115314
_PS = [
    (-1, "1", "aa", 0, 1, 2, "xxx"), (-2, "1", "aa", 0, 2, 1, "xxy"), (-2, "1", "ab", 1, 1, 1, "xxz"),
    (-1, "1", "aa", 0, 3, 0, "xyy"), (-2, "1", "ab", 1, 2, 0, "xyz"), (-1, "1", "bb", 2, 1, 0, "xzz"),
    (1, "1", "aa", 0, 0, 3, "yxx"), (2, "1", "aa", 0, 1, 2, "yxy"), (2, "1", "ab", 1, 0, 2, "yxz"),
    (1, "1", "aa", 0, 2, 1, "yyy"), (2, "1", "ab", 1, 1, 1, "yyz"), (1, "1", "bb", 2, 0, 1, "yzz"),
]
_SP = [
    (-1, "m", "1", 0, 2, 1, "xxx"), (2, "m", "1", 0, 1, 2, "xxy"), (-1, "m", "1", 0, 0, 3, "xyy"),
    (-1, "m", "1", 0, 3, 0, "yxx"), (2, "m", "1", 0, 2, 1, "yxy"), (-1, "m", "1", 0, 1, 2, "yyy"),
    (1, "p", "1", 1, 2, 0, "zxx"), (-2, "p", "1", 1, 1, 1, "zxy"), (1, "p", "1", 1, 0, 2, "zyy"),
]
_SS = [
    (-1, "1", "1", 0, 3, 0, "xxx"), (2, "1", "1", 0, 2, 1, "xxy"), (-1, "1", "1", 0, 1, 2, "xyy"),
    (1, "1", "1", 0, 2, 1, "yxx"), (1, "1", "1", 0, 0, 3, "yyy"), (-2, "1", "1", 0, 1, 2, "yxy"),
]


def _rad(pol, energy, eps1w, eps2w, chi2, theta, phi, thick, mref, consts,
         sinc_normalized=True, zzz_phi_quirk=True):
    """pre * r for pol in {'pp','ps','sp','ss'}; first letter = incoming (1w)
    polarisation, second = outgoing (2w): shg.py:201-216, 273-288, 317-332, 355-370."""
    pin, pout = pol[0], pol[1]
    e1, e2 = eps1w, eps2w
    F2 = multiref_2w(energy, e2["M2"], fresnel_r(pout, e2["M2"], e2["M3"], theta),
                     fresnel_r(pout, e2["M1"], e2["M2"], theta), theta, thick, mref, consts, sinc_normalized)
    f1 = multiref_1w(energy, e1["M2"], fresnel_r(pin, e1["M2"], e1["M3"], theta),
                     fresnel_r(pin, e1["M1"], e1["M2"], theta), theta, thick, mref, consts)
    if pout == "p":
        pre_out = fresnel_t("p", e2["M1"], e2["M2"], theta) / np.sqrt(e2["M2"])
    else:
        pre_out = fresnel_t("s", e2["M1"], e2["M2"], theta) * (1 + F2)
    if pin == "p":
        pre_in = fresnel_t("p", e1["M1"], e1["M2"], theta) / np.sqrt(e1["M2"])
    else:
        pre_in = fresnel_t("s", e1["M1"], e1["M2"], theta) * (1 + f1)
    pre = pre_out * pre_in ** 2
    W, w = wvec(e2["M2"], theta), wvec(e1["M2"], theta)
    outf = {"m": (1 - F2) * W, "p": (1 + F2), "1": 1.0}
    incf = {"aa": (1 - f1) ** 2 * w ** 2, "ab": (1 + f1) * (1 - f1) * w, "bb": (1 + f1) ** 2, "1": 1.0}
    rows = {"pp": _PP + [_PP_ZZZ_QUIRK if zzz_phi_quirk else _PP_ZZZ_THEORY],
            "ps": _PS, "sp": _SP, "ss": _SS}[pol]
    st, sp, cp = np.sin(theta), np.sin(phi), np.cos(phi)
    r = 0
    for mult, o, i, pt, ps, pc, key in rows:
        r = r + mult * outf[o] * incf[i] * st ** pt * sp ** ps * cp ** pc * chi2[key]
    return pre * r


def rad_pp(*a, **k): return _rad("pp", *a, **k)
def rad_ps(*a, **k): return _rad("ps", *a, **k)
def rad_sp(*a, **k): return _rad("sp", *a, **k)
def rad_ss(*a, **k): return _rad("ss", *a, **k)


def yield_points(energy, eps1w, eps2w, chi2_rot, theta_deg, phi_deg, thick, mref=True,
                 consts=None, sinc_normalized=True, zzz_phi_quirk=True) -> dict:
    """'L-kernel' oracle: the four un-broadened yields from arrays that are already
    resampled / averaged / rotated (shg.py:428-436 with sigma_out = 0).  Every
    argument broadcasts like NumPy."""
    consts = consts or SCIPY_1_18_CONSTANTS
    with np.errstate(all="ignore"):                  # shg.py:23
        pref = 1e4 * ((energy / consts["hbar_eVs"]) ** 2) / \
            (2 * consts["eps0"] * consts["c"] ** 3 * np.cos(np.radians(theta_deg)) ** 2)
        th, ph = np.radians(theta_deg), np.radians(phi_deg)
        out = {}
        for pol in ("pp", "ps", "sp", "ss"):
            z = _rad(pol, energy, eps1w, eps2w, chi2_rot, th, ph, thick, mref, consts,
                     sinc_normalized, zzz_phi_quirk)
            out[pol] = pref * np.absolute(1 / np.sqrt(eps1w["M2"]) * z) ** 2
    return out


# ----------------------------------------------------------------------------
# shgyield(): shg.py:381-437
# ----------------------------------------------------------------------------
def _medium(eps_m: dict, energy, sigma_eps):
    """shg.py:387-403 for one medium -> (dict @1w, dict @2w)."""
    num = {k: v for k, v in eps_m.items() if not isinstance(v, dict)}
    one, two = {}, {}
    for k, v in eps_m.items():
        if isinstance(v, dict):
            b = {"energy": v["energy"], "data": broadC(v["data"], sigma_eps)}
            one[k] = resample_complex(b, energy)
            two[k] = resample_complex(b, 2 * np.asarray(energy))
    one.update(num)
    two.update(num)
    return one, two


def avgEPS(d: dict):                                 # shg.py:58-60
    return np.mean(list(d.values()), axis=0)


def prepare(energy, eps_m1, eps_m2, eps_m3, chi2, thick, gamma=90, sigma_eps=0.0,
            sigma_chi=0.0, mode="3-layer"):
    """Everything of shg.py:387-426: returns (eps1w, eps2w, chi2_rot, thick)."""
    m1, m2, m3 = (_medium(m, energy, sigma_eps) for m in (eps_m1, eps_m2, eps_m3))
    if mode == "3-layer":                            # shg.py:405-411
        eps1w = {"M1": avgEPS(m1[0]), "M2": avgEPS(m2[0]), "M3": avgEPS(m3[0])}
        eps2w = {"M1": avgEPS(m1[1]), "M2": avgEPS(m2[1]), "M3": avgEPS(m3[1])}
    elif mode in ("fresnel", "2-layer"):             # shg.py:412-420
        thick = 0
        eps1w = {"M1": avgEPS(m1[0]), "M2": avgEPS(m3[0]), "M3": avgEPS(m3[0])}
        eps2w = {"M1": avgEPS(m1[1]), "M2": avgEPS(m1[1]), "M3": avgEPS(m3[1])}
    else:
        raise UnboundLocalError("mode %r: the reference leaves eps1w unbound" % (mode,))
    chi = {k: v for k, v in chi2.items() if not isinstance(v, dict)}
    for k, v in chi2.items():                        # shg.py:422-425
        if isinstance(v, dict):
            chi[k] = resample_complex({"energy": v["energy"], "data": broadC(v["data"], sigma_chi)}, energy)
    return eps1w, eps2w, rotate(chi, np.radians(gamma)), thick


def shgyield(energy, eps_m1, eps_m2, eps_m3, chi2, theta, phi, thick, gamma=90, sigma_eps=0.0,
             sigma_chi=0.0, sigma_out=5.0, mode="3-layer", mref=True, consts=None,
             sinc_normalized=True, zzz_phi_quirk=True):
    eps1w, eps2w, chi_rot, thick = prepare(energy, eps_m1, eps_m2, eps_m3, chi2, thick, gamma,
                                           sigma_eps, sigma_chi, mode)
    y = yield_points(energy, eps1w, eps2w, chi_rot, theta, phi, thick, mref, consts,
                     sinc_normalized, zzz_phi_quirk)
    out = {"energy": energy, "phi": phi}
    for pol in ("pp", "ps", "sp", "ss"):
        out[pol] = broad(y[pol], sigma_out)          # shg.py:433-436
    return out
