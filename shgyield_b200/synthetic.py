"""
Synthetic workloads for the SSHG yield path.

`config5()` is the "synthetic scale" case BASELINE.json names: random 27-component
complex chi2, Lorentz-oscillator dielectric functions, 20k energies x 181 theta x
721 phi x 4 polarisations.  The arrays are fed at the kernel level (already
resampled / averaged / rotated), i.e. exactly what `rad_pp/ps/sp/ss` of the
reference (`shgyield/shg.py:201-378`) take.
"""
from __future__ import annotations

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
