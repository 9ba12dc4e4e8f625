"""
Workloads of BASELINE.json for the SSHG yield path: the synthetic scale case and the
Si(111)(1x1):H sweeps.

`config5()` is the "synthetic scale" case BASELINE.json names: random 27-component
complex chi2, Lorentz-oscillator dielectric functions, 20k energies x 181 theta x
721 phi x 4 polarisations.  The arrays are fed at the kernel level (already
resampled / averaged / rotated), i.e. exactly what `rad_pp/ps/sp/ss` of the
reference (`shgyield/shg.py:201-378`) take.
"""
from __future__ import annotations

import os

import numpy as np

CHI2_KEYS = ("xxx", "xyy", "xzz", "xyz", "xxz", "xxy",
             "yxx", "yyy", "yzz", "yyz", "yxz", "yxy",
             "zxx", "zyy", "zzz", "zyz", "zxz", "zxy")
_AX = {"x": 0, "y": 1, "z": 2}


def lorentz_eps(omega, f, e0, g):
    """1 + sum_k f_k / (E_k^2 - w^2 - i g_k w); Im >= 0 for w >= 0."""
    w = np.asarray(omega, dtype=np.float64)[:, None]
    return 1.0 + np.sum(f / (e0 ** 2 - w ** 2 - 1j * g * w), axis=1)


def synthetic_spectra(n_energy: int, seed: int = 20250, e_min: float = 0.5, e_max: float = 5.0):
    """Returns (energy[nE], eps1w[3,nE], eps2w[3,nE], chi2[18,nE]) complex128, media order
    (vacuum, layer, bulk), chi2 rows in CHI2_KEYS order."""
    rng = np.random.default_rng(seed)
    energy = np.linspace(e_min, e_max, n_energy)
    eps1w = np.ones((3, n_energy), dtype=np.complex128)
    eps2w = np.ones((3, n_energy), dtype=np.complex128)
    for m in (1, 2):
        f = rng.uniform(20.0, 60.0, 4)
        e0 = rng.uniform(3.0, 6.0, 4)
        g = rng.uniform(0.1, 0.5, 4)
        eps1w[m] = lorentz_eps(energy, f, e0, g)
        eps2w[m] = lorentz_eps(2.0 * energy, f, e0, g)
    chi27 = (rng.standard_normal((3, 3, 3, n_energy)) + 1j * rng.standard_normal((3, 3, 3, n_energy))) * 1e-19
    chi2 = np.empty((18, n_energy), dtype=np.complex128)
    for i, key in enumerate(CHI2_KEYS):
        a, b, c = (_AX[t] for t in key)
        chi2[i] = 0.5 * (chi27[a, b, c] + chi27[a, c, b])
    return energy, eps1w, eps2w, chi2


def config5(n_energy: int = 20000, n_theta: int = 181, n_phi: int = 721, seed: int = 20250):
    energy, eps1w, eps2w, chi2 = synthetic_spectra(n_energy, seed)
    return {
        "energy": energy,
        "theta": np.linspace(0.0, 90.0, n_theta),
        "phi": np.linspace(0.0, 360.0, n_phi),
        "thick": np.array([10.0]),
        "eps1w": eps1w, "eps2w": eps2w, "chi2": chi2,
    }


# ---- Si(111)(1x1):H, the reference's example material (configs 1, 3 and 4 of BASELINE.json) --------------------
LSUPER, LSLAB = 180, 142.1885172213904          # reference main.py:32-33
EXAMPLE_SPECTRA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "example", "si111_spectra.npz")


def si111_material(path: str | None = None):
    """(eps_m1, eps_m2, eps_m3, chi2) of the reference's main.py:29-75 in the argument format of shgyield(), from the
    raw file columns in example/si111_spectra.npz (loadeps / loadshg arithmetic of main.py:14-22)."""
    z = np.load(path or EXAMPLE_SPECTRA)
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


def config3_axes():
    """BASELINE.json configs[2]: Si(111)(1x1):H angular map, 2000 E x 90 theta x 360 phi, d = 10 nm."""
    return dict(energy=np.linspace(0.005, 10.0, 2000), theta=np.arange(0, 90, dtype=np.float64),
                phi=np.arange(0, 360, dtype=np.float64), thick=np.array([10.0]), gamma=0)


def config4_axes():
    """BASELINE.json configs[3]: thin-film sweep, d = 1..100 nm x 90 theta over the spectrum (1000 E: the largest
    range whose 2w stays inside the 0-20 eV files), phi = 30 deg."""
    return dict(energy=0.01 * np.arange(1, 1001), theta=np.arange(0, 90, dtype=np.float64),
                phi=np.array([30.0]), thick=np.arange(1, 101, dtype=np.float64), gamma=0)
