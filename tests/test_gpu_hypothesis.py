"""Differential tests with generated inputs (hypothesis): random dielectric functions, chi2 tensors,
angles and thicknesses through the CUDA kernels (C ABI) against the CPU oracle."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import shg_oracle as orc
from tests import cases, util

pytestmark = pytest.mark.gpu
CONSTS = util.load_consts()
MORE = int(os.environ.get("SSHG_HYP_SCALE", "1"))       # SSHG_HYP_SCALE=8: a longer soak run of the same tests


def _spectra(rng, nE):
    energy = np.sort(rng.uniform(0.05, 9.0, nE))
    eps1w = np.empty((3, nE), dtype=np.complex128)
    eps2w = np.empty((3, nE), dtype=np.complex128)
    eps1w[0] = eps2w[0] = 1.0
    for m in (1, 2):
        # metals and dielectrics; a slightly negative Im (what broadening / splines can produce) exercises
        # the other side of the square-root branch cut
        eps1w[m] = rng.uniform(-20, 30, nE) + 1j * rng.uniform(-0.3, 25, nE)
        eps2w[m] = rng.uniform(-20, 30, nE) + 1j * rng.uniform(-0.3, 25, nE)
        eps1w[m][rng.integers(0, nE)] = rng.uniform(-20, 30) + 0j              # purely real (signed zero)
    chi = (rng.standard_normal((18, nE)) + 1j * rng.standard_normal((18, nE))) * 10.0 ** rng.uniform(-22, -17)
    chi[rng.integers(0, 18, 4)] = 0.0                                             # some symmetry zeros
    return energy, eps1w, eps2w, chi


def _oracle(energy, eps1w, eps2w, chi, theta, phi, thick, **kw):
    e1 = {"M%d" % (m + 1): eps1w[m] for m in range(3)}
    e2 = {"M%d" % (m + 1): eps2w[m] for m in range(3)}
    ch = {k: chi[i] for i, k in enumerate(cases.CHI2_KEYS)}
    return orc.yield_points(energy, e1, e2, ch, theta, phi, thick, consts=CONSTS, **kw)


@settings(max_examples=40 * MORE, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 700), mref=st.booleans(), sn=st.booleans(), zq=st.booleans())
def test_points_kernel_random(seed, n, mref, sn, zq):
    from shgyield_b200 import shg
    rng = np.random.default_rng(seed)
    nE = int(rng.integers(1, 40))
    energy, eps1w, eps2w, chi = _spectra(rng, nE)
    eidx = rng.integers(0, nE, n)
    theta, phi, thick = rng.uniform(0, 89.9, n), rng.uniform(-360, 720, n), rng.uniform(0, 120, n)
    got = shg.yield_points(energy, eidx, theta, phi, thick, eps1w, eps2w, chi, mref=mref, consts=CONSTS,
                           sinc_normalized=sn, zzz_phi_quirk=zq)
    want = _oracle(energy[eidx], eps1w[:, eidx], eps2w[:, eidx], chi[:, eidx], theta, phi, thick, mref=mref,
                   sinc_normalized=sn, zzz_phi_quirk=zq)
    scale = max(np.max(np.abs(want[p])) for p in cases.POLS)
    for k, pol in enumerate(cases.POLS):
        ok, msg = cases.parity_ok(got[k], want[pol], rtol=1e-10, scale=scale)
        assert ok, (seed, pol, msg)


@settings(max_examples=25 * MORE, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31 - 1), nE=st.integers(1, 1300), nT=st.integers(1, 5), nD=st.integers(1, 3),
       grid_kind=st.sampled_from(["paired", "paired_tail", "free", "single", "regular", "regular_closed"]),
       kernel=st.sampled_from(["dense", "dense_quads", "sparse"]))
def test_grid_kernel_random(seed, nE, nT, nD, grid_kind, kernel):
    from shgyield_b200 import grid
    with util.lib_options(k2_sparse=1 if kernel == "sparse" else 0, quads=1 if kernel == "dense_quads" else None):
        _grid_kernel_random(grid, seed, nE, nT, nD, grid_kind, kernel)


def _grid_kernel_random(grid, seed, nE, nT, nD, grid_kind, kernel):
    rng = np.random.default_rng(seed)
    energy, eps1w, eps2w, chi = _spectra(rng, nE)
    theta = np.sort(rng.uniform(0, 90, nT))
    thick = rng.uniform(0, 80, nD)
    h = int(rng.integers(1, 9))
    base = np.sort(rng.uniform(0, 180, h))
    m = int(rng.choice([4, 6, 8, 10, 12, 16, 18, 20, 24, 30, 36, 40]))     # regular grids over the full circle with steps that
                                                                           # are exact in binary: four rows per evaluation
    phi = {"paired": np.concatenate([base, base + 180.0]),
           "paired_tail": np.concatenate([base, base + 180.0, [360.0]]),
           "free": rng.uniform(0, 360, 2 * h + 1),
           "single": np.array([rng.uniform(0, 360)]),
           "regular": np.arange(m) * (360.0 / m),
           "regular_closed": np.linspace(0.0, 360.0, m + 1)}[grid_kind]
    got = grid.yield_grid_host(energy, theta, phi, thick, eps1w, eps2w, chi, consts=CONSTS)
    want = _oracle(energy, eps1w, eps2w, chi, theta.reshape(-1, 1, 1, 1), phi.reshape(1, 1, -1, 1), thick.reshape(1, -1, 1, 1))
    scale = max(np.max(np.abs(want[p])) for p in cases.POLS)
    for k, pol in enumerate(cases.POLS):
        w = np.broadcast_to(want[pol], (nT, nD, phi.size, nE))
        ok, msg = cases.parity_ok(got[:, :, k], w, rtol=1e-10, scale=scale)
        assert ok, (seed, grid_kind, kernel, pol, msg)


@settings(max_examples=20 * MORE, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31 - 1), nE=st.integers(1, 900), nT=st.integers(1, 3), nD=st.integers(1, 2),
       sigma=st.floats(0.13, 16.0), grid_kind=st.sampled_from(["paired", "paired_tail", "free", "regular", "regular_closed"]),
       tiles=st.sampled_from(["64", "128"]))
def test_broadened_sweep_random(seed, nE, nT, nD, sigma, grid_kind, tiles):
    """K3 (yields broadened along E inside the kernel, both tile widths) on generated media / tensors /
    grids against the oracle's broad(yield, sigma_out) -- absolute bound of the harmonic form plus the
    parity metric with the relative window from 1e-4 of the maximum (tests/test_gpu_blur.py)."""
    from shgyield_b200 import grid
    from tests.test_gpu_blur import _check
    rng = np.random.default_rng(seed)
    energy, eps1w, eps2w, chi = _spectra(rng, nE)
    theta = np.sort(rng.uniform(0, 90, nT))
    thick = rng.uniform(0, 80, nD)
    h = int(rng.integers(4, 12))
    base = np.sort(rng.uniform(0, 180, h))
    m = int(rng.choice([8, 10, 12, 16, 18, 20, 24, 30, 36, 40, 48]))
    phi = {"paired": np.concatenate([base, base + 180.0]),
           "paired_tail": np.concatenate([base, base + 180.0, [360.0]]),
           "free": rng.uniform(0, 360, 2 * h + 1),
           "regular": np.arange(m) * (360.0 / m),          # small launch: K3 takes four rows per evaluation
           "regular_closed": np.linspace(0.0, 360.0, m + 1)}[grid_kind]
    with util.lib_options(k3_nt=int(tiles), k2_sparse=0):  # fewer than 16 azimuths would otherwise take K2s -> K1
        got = grid.yield_grid_host(energy, theta, phi, thick, eps1w, eps2w, chi, consts=CONSTS, sigma_out=sigma)
    raw = _oracle(energy, eps1w, eps2w, chi, theta.reshape(-1, 1, 1, 1), phi.reshape(1, 1, -1, 1), thick.reshape(1, -1, 1, 1))
    want = np.stack([orc.gaussian_filter1d(np.broadcast_to(raw[p], (nT, nD, phi.size, nE)), sigma, axis=-1)
                     for p in cases.POLS], axis=2)
    _check(got, want, True, (seed, nE, sigma, grid_kind, tiles), col_bound=5e-12)


_SHAPES = [(), (1,), (5,), (3, 1), (2, 1, 1), (1, 4), (2, 3)]


@settings(max_examples=30 * MORE, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31 - 1), es=st.sampled_from([(), (7,), (40,), (2, 5)]), ts=st.sampled_from(_SHAPES),
       ps=st.sampled_from(_SHAPES), ds=st.sampled_from([(), (1,), (3, 1, 1)]), sigma_out=st.sampled_from([0.0, 0.7, 2.0]),
       mode=st.sampled_from(["3-layer", "2-layer", "fresnel"]), mref=st.booleans(), grid_path=st.booleans())
def test_api_broadcast_shapes_random(seed, es, ts, ps, ds, sigma_out, mode, mref, grid_path):
    """The drop-in shgyield() on generated NumPy-broadcast shapes of energy / theta / phi / thick (0-d to 3-d),
    modes, gamma and sigmas -- points kernel and grid kernel (threshold lowered), N-D output broadening --
    against the oracle's shgyield()."""
    from shgyield_b200 import shg
    rng = np.random.default_rng(seed)
    try:
        shape = np.broadcast_shapes(es, ts, ps, ds)
    except ValueError:
        return                                               # not broadcastable: nothing to compare
    m1, m2, m3, chi2 = cases.full_tensor_material(cases.load_spectra(), seed=seed % 7)
    energy = rng.uniform(0.3, 9.5, es) if es else float(rng.uniform(0.3, 9.5))
    theta = rng.uniform(0, 89, ts) if ts else float(rng.uniform(0, 89))
    phi = rng.uniform(-180, 540, ps) if ps else float(rng.uniform(-180, 540))
    thick = rng.uniform(0, 60, ds) if ds else float(rng.uniform(0, 60))
    kw = dict(gamma=float(rng.uniform(-90, 180)), sigma_eps=float(rng.choice([0.0, 1.5])), sigma_chi=float(rng.choice([0.0, 3.0])),
              sigma_out=sigma_out, mode=mode, mref=mref)
    old = shg.GRID_MIN_POINTS
    shg.GRID_MIN_POINTS = 1 if grid_path else old
    try:
        got = shg.shgyield(energy, m1, m2, m3, chi2, theta, phi, thick, **kw)
    finally:
        shg.GRID_MIN_POINTS = old
    want = orc.shgyield(energy, m1, m2, m3, chi2, theta, phi, thick, consts=CONSTS, **kw)
    scale = max(float(np.max(np.abs(want[p]))) for p in cases.POLS)
    for pol in cases.POLS:
        # (thick enters the broadcast only with mref and mode = 3-layer, exactly as in the reference)
        assert np.shape(got[pol]) == np.shape(want[pol]), (pol, np.shape(got[pol]), np.shape(want[pol]), shape)
        ok, msg = cases.parity_ok(got[pol], want[pol], rtol=1e-10, scale=scale)
        assert ok, (seed, es, ts, ps, ds, kw, grid_path, pol, msg)


@settings(max_examples=40 * MORE, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31 - 1), rows=st.integers(1, 6), half_n=st.integers(1, 4200), odd=st.booleans(),
       sigma=st.one_of(st.sampled_from([0.9, 1.0, 1.5, 2.0, 3.0, 4.5, 5.0, 6.5, 7.5, 8.0, 9.0, 10.0, 12.0, 14.0, 16.0]),
                       st.floats(0.3, 20.0)))
def test_gauss_rows_random(seed, rows, half_n, odd, sigma):
    """K1 on random row lists against scipy.ndimage: every radius family (fixed radius with bulk-copy or per-thread
    staging, generic register window, tiny radii), even and odd row lengths, rows shorter than the radius (repeated
    reflection), NaN / Inf samples confined to the radius; bulk-copy and per-thread staging bitwise equal."""
    import torch
    from scipy import ndimage
    from shgyield_b200 import grid
    rng = np.random.default_rng(seed)
    n = 2 * half_n + (1 if odd else 0)
    x = rng.normal(size=(rows, n)) * 10.0 ** rng.uniform(-3, 3)
    for _ in range(int(rng.integers(0, 3))):
        x[rng.integers(0, rows), rng.integers(0, n)] = rng.choice([np.nan, np.inf, -np.inf])
    d = torch.from_numpy(x).cuda()
    got = grid.broaden_energy_device(d, sigma).cpu().numpy()
    with util.lib_options(k1_bulk=0):
        ref = grid.broaden_energy_device(d, sigma).cpu().numpy()
    assert np.array_equal(got, ref, equal_nan=True), (seed, rows, n, sigma)
    with np.errstate(invalid="ignore"):
        want = np.stack([ndimage.gaussian_filter(row, sigma) for row in x])
    fin = np.isfinite(want)
    # Inf - Inf inside a window is NaN in both; +-Inf against NaN can differ with the summation order: finite pattern only
    assert np.array_equal(np.isfinite(got), fin), (seed, rows, n, sigma)
    scale = np.max(np.abs(x[np.isfinite(x)])) if np.isfinite(x).any() else 1.0
    np.testing.assert_allclose(got[fin], want[fin], rtol=0, atol=1e-14 * max(scale, 1e-300))
