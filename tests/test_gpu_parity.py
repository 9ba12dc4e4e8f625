"""
GPU parity tests (run with -m gpu on a B200): every call goes through the C ABI of libsshg.so.
The CPU oracle (oracle/shg_oracle.py) and the golden fixtures (tests/golden/, produced by the
unmodified reference) are the checkers.  Tolerance: FP64, relative 1e-10 in the metric of
tests/cases.py::parity_ok (north_star: "within a relative tolerance of 1e-10").
"""
import os

import numpy as np
import pytest

from oracle import shg_oracle as orc
from tests import cases, util

pytestmark = pytest.mark.gpu
RTOL = 1e-10
G = cases.GOLDEN_DIR


@pytest.fixture(scope="module")
def shg():
    from shgyield_b200 import shg as mod
    return mod


@pytest.fixture(scope="module")
def grid():
    from shgyield_b200 import grid as mod
    return mod


@pytest.fixture(scope="module")
def consts():
    return util.load_consts()


@pytest.fixture(scope="module")
def spectra():
    return cases.load_spectra()


def test_scipy_constants_match_fixture(shg, consts):
    assert shg.physical_constants() == consts


# ---- K1: Gaussian broadening vs the live SciPy ---------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 7, 126, 1024, 1025, 2001, 20000])
@pytest.mark.parametrize("sigma", [0.0, 0.2, 1.0, 2.0, 3.3, 3.5, 5.0, 6.0, 7.0, 8.0, 10.0, 50.0])
def test_gauss_matches_scipy(shg, n, sigma):
    """radius = int(4 sigma + 0.5): even radii 4..24, 28, 32, 40 take the fixed-radius kernel (taps in the
    constant bank), the others the generic register-window kernel."""
    from scipy import ndimage
    x = np.random.default_rng(n).normal(size=n)
    got = shg.broad(x, sigma)
    want = ndimage.gaussian_filter(x, sigma)
    np.testing.assert_allclose(got, want, rtol=0, atol=4e-15 * max(1.0, np.max(np.abs(x))))
    if sigma == 0.0:
        np.testing.assert_array_equal(got, x)


def test_gauss_huge_radius_and_nd(shg):
    from scipy import ndimage
    x = np.random.default_rng(1).normal(size=50)
    np.testing.assert_allclose(shg.broad(x, 7000.0), ndimage.gaussian_filter(x, 7000.0), rtol=0, atol=1e-13)
    y = np.random.default_rng(2).normal(size=(6, 1, 33, 5))
    np.testing.assert_allclose(shg.broad(y, 1.7), ndimage.gaussian_filter(y, 1.7), rtol=0, atol=1e-14)
    z = np.random.default_rng(3).normal(size=40) + 1j * np.random.default_rng(4).normal(size=40)
    want = ndimage.gaussian_filter(z.real, 2.5) + 1j * ndimage.gaussian_filter(z.imag, 2.5)
    np.testing.assert_allclose(shg.broadC(z, 2.5), want, rtol=0, atol=1e-14)


@pytest.mark.parametrize("shape", [(90, 7, 50), (361, 1, 126), (4, 100, 33), (75, 64), (200, 9), (3, 130, 8, 40), (700, 33), (150, 96), (40, 3, 70)])
@pytest.mark.parametrize("sigma", [0.4, 1.5, 5.0, 12.0, 30.0])
def test_gauss_nd_every_axis(shg, shape, sigma):
    """broad() of an N-D array filters every axis like scipy.ndimage.gaussian_filter: the last axis with the contiguous
    kernels, the others with the column kernel (32 neighbouring signals per CTA, register window down the column), very
    short axes and narrow trailing axes with the direct / strided kernels; a NaN poisons exactly its neighbourhood."""
    from scipy import ndimage
    x = np.random.default_rng(len(shape) * 1000 + shape[0]).normal(size=shape)
    got = shg.broad(x, sigma)
    np.testing.assert_allclose(got, ndimage.gaussian_filter(x, sigma), rtol=0, atol=2e-14)
    with util.lib_options(k1_bulk=0):              # tiles staged by per-thread loads instead of bulk copies: same bits
        assert np.array_equal(shg.broad(x, sigma), got)
    x[tuple(s // 2 for s in shape)] = np.nan
    got, want = shg.broad(x, sigma), ndimage.gaussian_filter(x, sigma)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=0, atol=2e-14)


def test_gauss_large_host_arrays_use_the_threaded_copies(shg):
    """Host arrays above 32 MB go up and come back through the page-locked staging buffers with several host threads
    (sshg_copy_in / sshg_copy_out): 144 MB in three chunks, the last one ragged; same numbers as SciPy."""
    from scipy import ndimage
    x = np.random.default_rng(5).normal(size=(3, 6_000_001))
    got = shg.broad(x[0], 2.0)
    np.testing.assert_allclose(got, ndimage.gaussian_filter(x[0], 2.0), rtol=0, atol=2e-14)
    got = shg.broad(x, 1.0)                        # both axes: the three-entry axis takes the direct kernel
    np.testing.assert_allclose(got, ndimage.gaussian_filter(x, 1.0), rtol=0, atol=2e-14)


@pytest.mark.parametrize("sigma", [2.0, 3.3, 5.0, 10.0])
def test_gauss_nan_stays_within_the_radius(shg, sigma):
    """An Inf/NaN sample poisons exactly the outputs within the radius, as in SciPy (taps that do not
    exist are skipped, not multiplied by zero); rows are independent."""
    from scipy import ndimage
    r = int(4 * sigma + 0.5)
    x = np.random.default_rng(3).normal(size=(3, 4000))
    x[1, 1700] = np.nan
    x[2, 5] = np.inf
    got, want = shg.broad(x, sigma), np.stack([ndimage.gaussian_filter(row, sigma) for row in x])
    got1 = np.stack([shg.broad(row, sigma) for row in x])           # 1-D calls: the energy axis only
    assert np.array_equal(np.isnan(got1[1]), np.abs(np.arange(4000) - 1700) <= r)
    assert np.array_equal(np.isfinite(got1), np.isfinite(want))
    np.testing.assert_allclose(got1[0], want[0], rtol=0, atol=1e-14)
    assert got.shape == x.shape


@pytest.mark.parametrize("n", [40, 80, 126, 128, 832, 834, 2912, 2914, 5824, 5850, 20000])
@pytest.mark.parametrize("sigma", [0.3, 0.7, 1.0, 3.3, 5.0, 6.5, 7.3, 8.0, 10.0, 12.0, 15.3, 16.0])
def test_gauss_bulk_copy_staging(shg, n, sigma):
    """The fixed-radius FIR kernels (every radius up to 64; odd radii stage one more sample on each side) move their
    tile + halo and their results with cp.async.bulk when rows are 16-byte aligned and of even length; the reflected ends of a row are filled from the staged samples.  Bitwise equal to the
    per-thread staging (option k1_bulk = 0) and equal to SciPy, over several rows, tiles that end exactly at / just
    before / just after the end of a row, and rows shorter than one tile; a NaN stays within the radius."""
    import torch
    from scipy import ndimage
    from shgyield_b200 import grid
    r = int(4 * sigma + 0.5)
    x = np.random.default_rng(n).normal(size=(7, n))
    x[3, n // 2] = np.nan
    x[5, 0] = np.inf
    x[6, n - 1] = -np.inf
    d = torch.from_numpy(x).cuda()
    with util.lib_options(k1_bulk=0):
        a = grid.broaden_energy_device(d, sigma).cpu().numpy()
    with util.lib_options(k1_bulk=1):
        b = grid.broaden_energy_device(d, sigma).cpu().numpy()
    c = grid.broaden_energy_device(d, sigma).cpu().numpy()                # the default
    assert np.array_equal(a, b, equal_nan=True) and np.array_equal(b, c, equal_nan=True)
    want = np.stack([ndimage.gaussian_filter(row, sigma) for row in x])
    assert np.array_equal(np.isfinite(b), np.isfinite(want)) and np.array_equal(np.isnan(b), np.isnan(want))
    ok = np.isfinite(want)
    np.testing.assert_allclose(b[ok], want[ok], rtol=0, atol=1e-14)
    if n >= 2 * r + 2:
        assert np.array_equal(np.isnan(b[3]), np.abs(np.arange(n) - n // 2) <= r)
    # rows of odd length (every other row only 8-byte aligned) take the per-thread staging
    e = grid.broaden_energy_device(d[:, 1:].contiguous(), sigma).cpu().numpy()
    w2 = np.stack([ndimage.gaussian_filter(row[1:], sigma) for row in x])
    ok2 = np.isfinite(w2)
    assert np.array_equal(np.isfinite(e), ok2)
    np.testing.assert_allclose(e[ok2], w2[ok2], rtol=0, atol=1e-14)


def test_gauss_matches_oracle_restatement(shg):
    x = np.random.default_rng(9).normal(size=(3, 300))
    np.testing.assert_allclose(shg.broad(x, 4.0), orc.gaussian_filter(x, 4.0), rtol=0, atol=1e-14)


# ---- spline resampling vs FITPACK ---------------------------------------------------------------
def test_spline_matches_fitpack(shg, spectra):
    from scipy.interpolate import InterpolatedUnivariateSpline
    e = spectra["energy"]
    q = np.concatenate([np.linspace(0.0, 20.0, 3777), e[:50], [e[-1]]])
    for name in ("SiH1x1-chi1-xx", "SiBulk-chi1-zz", "SiH1x1-chi2-zzz"):
        for row in spectra[name][:2]:
            want = InterpolatedUnivariateSpline(e, row, ext=2)(q)
            got = shg.spline(row, e)(q)
            np.testing.assert_allclose(got, want, rtol=0, atol=2e-13 * np.max(np.abs(row)))
    re, im = shg.splineC(spectra["SiH1x1-chi1-xx"][0] + 1j * spectra["SiH1x1-chi1-xx"][1], e)
    assert re(np.array([1.0])).shape == (1,) and im(2.0).shape == ()


def test_spline_nonuniform_axis(shg):
    from scipy.interpolate import InterpolatedUnivariateSpline
    rng = np.random.default_rng(0)
    x = np.cumsum(rng.uniform(0.05, 1.0, 57))
    y = rng.normal(size=57)
    q = rng.uniform(x[0], x[-1], 500)
    np.testing.assert_allclose(shg.spline(y, x)(q), InterpolatedUnivariateSpline(x, y, ext=2)(q), rtol=0, atol=1e-11)
    for n in (4, 5):
        np.testing.assert_allclose(shg.spline(y[:n], x[:n])(x[:n]), y[:n], rtol=0, atol=1e-13)


def test_spline_out_of_range_raises(shg, spectra):
    e = spectra["energy"]
    with pytest.raises(ValueError):
        shg.spline(spectra["SiH1x1-chi1-xx"][0], e)(np.array([1.0, 20.5]))
    with pytest.raises(ValueError):
        shg.spline(spectra["SiH1x1-chi1-xx"][0], e)(np.array([-0.1]))


# ---- rotate ---------------------------------------------------------------------------------------
def test_rotate_matches_reference(shg):
    z = np.load(os.path.join(G, "ref_helpers.npz"))
    chi = {k: z["rot_in"][i] for i, k in enumerate(cases.CHI2_KEYS)}
    for g in (0.0, 17.0, 90.0, 123.4, 270.0):
        r = shg.rotate(chi, np.radians(g))
        got = np.stack([r[k] for k in cases.CHI2_KEYS])
        np.testing.assert_allclose(got, z["rot_out_%g" % g], rtol=0, atol=2e-14)
    with pytest.raises(KeyError):
        shg.rotate({k: 1.0 for k in cases.CHI2_KEYS[:-1]}, 0.3)


def test_rotate_without_quirk_is_a_true_rotation(shg):
    """Without the xxy/yyy sign of the reference (shg.py:92) the map is a true rotation about z by
    `rotang`: rotating by +a and then by -a returns the input; with the quirk it does not."""
    rng = np.random.default_rng(5)
    chi = {k: rng.normal(size=3) + 1j * rng.normal(size=3) for k in cases.CHI2_KEYS}
    a = np.radians(37.0)
    back = shg.rotate(shg.rotate(chi, a, xxy_quirk=False), -a, xxy_quirk=False)
    for k in cases.CHI2_KEYS:
        np.testing.assert_allclose(back[k], chi[k], rtol=0, atol=1e-14)
    quirky = shg.rotate(shg.rotate(chi, a), -a)
    assert max(np.max(np.abs(quirky[k] - chi[k])) for k in cases.CHI2_KEYS) > 1e-3


# ---- K2 at kernel level -------------------------------------------------------------------------
def test_points_kernel_config5_samples(shg, consts):
    from shgyield_b200.synthetic import synthetic_spectra
    z = np.load(os.path.join(G, "ref_grid_samples.npz"))
    idx = z["c5_idx"]
    energy, eps1w, eps2w, chi = synthetic_spectra(20000)
    theta, phi = np.linspace(0.0, 90.0, 181), np.linspace(0.0, 360.0, 721)
    out = shg.yield_points(energy, idx[:, 3], theta[idx[:, 0]], phi[idx[:, 2]], 10.0, eps1w, eps2w, chi, consts=consts)
    sc = cases.case_scale(z, "c5_")
    for i, pol in enumerate(cases.POLS):
        ok, msg = cases.parity_ok(out[i], z["c5_" + pol], rtol=RTOL, scale=sc)
        assert ok, (pol, msg)


def _oracle_cube(cfg, consts, **kw):
    e1 = {"M%d" % (m + 1): cfg["eps1w"][m] for m in range(3)}
    e2 = {"M%d" % (m + 1): cfg["eps2w"][m] for m in range(3)}
    ch = {k: cfg["chi2"][i] for i, k in enumerate(cases.CHI2_KEYS)}
    T, D, P = cfg["theta"].size, cfg["thick"].size, cfg["phi"].size
    y = orc.yield_points(cfg["energy"], e1, e2, ch, cfg["theta"].reshape(T, 1, 1, 1), cfg["phi"].reshape(1, 1, P, 1),
                         cfg["thick"].reshape(1, D, 1, 1), consts=consts, **kw)
    return np.stack([np.broadcast_to(y[p], (T, D, P, cfg["energy"].size)) for p in cases.POLS], axis=2)


def _small_cfg(nE, theta, phi, thick, seed=1):
    from shgyield_b200.synthetic import synthetic_spectra
    energy, eps1w, eps2w, chi = synthetic_spectra(nE, seed=seed)
    return dict(energy=energy, theta=np.asarray(theta, float), phi=np.asarray(phi, float),
                thick=np.asarray(thick, float), eps1w=eps1w, eps2w=eps2w, chi2=chi)


GRID_SHAPES = [
    # nE, theta, phi, thick                                  what it exercises
    (1, [65.0], [30.0], [10.0]),                             # a single point
    (126, [65.0], [30.0], [10.0]),                           # config-1 shape
    (513, [0.0, 33.0, 90.0], np.arange(0, 360, 15.0), [10.0]),    # odd nE (64-bit stores), (phi, phi+180) pairs
    (600, [10.0, 80.0], np.linspace(0, 360, 41), [0.0, 7.5]),     # pairs + one un-paired row, two thicknesses
    (64, [45.0], np.linspace(0, 350, 7), [3.0]),             # no pairing possible
    (1030, [20.0, 40.0], [0.0, 180.0], [5.0]),               # two tiles, one pair
    (40, np.arange(0, 90, 9.0), [30.0], np.arange(1.0, 31.0)),    # config-4 shape (P = 1)
    (30, [55.0], np.linspace(0, 360, 1001), [12.0]),         # more than one phi chunk (500 pairs + 1)
]


@pytest.mark.parametrize("kernel", ["auto", "dense", "dense_quads", "sparse"])
@pytest.mark.parametrize("nE,theta,phi,thick", GRID_SHAPES, ids=[str(i) for i in range(len(GRID_SHAPES))])
def test_grid_kernel_matches_oracle(grid, consts, libopt, nE, theta, phi, thick, kernel):
    """Every shape through the dense kernel (phi / phi + 180 pairs; with the second mirror: four rows per evaluation)
    and the sparse-phi kernel."""
    if kernel != "auto":
        libopt(k2_sparse=1 if kernel == "sparse" else 0)
    if kernel == "dense_quads":
        libopt(quads=1)
    cfg = _small_cfg(nE, theta, phi, thick)
    got = grid.yield_grid_host(**cfg, consts=consts)
    want = _oracle_cube(cfg, consts)
    assert got.shape == want.shape
    for k, pol in enumerate(cases.POLS):
        ok, msg = cases.parity_ok(got[:, :, k], want[:, :, k], rtol=RTOL, scale=np.max(want))
        assert ok, (pol, msg)


@pytest.mark.parametrize("kernel", ["dense", "sparse"])
@pytest.mark.parametrize("kw", [dict(mref=False), dict(sinc_normalized=False), dict(zzz_phi_quirk=False)])
def test_grid_kernel_switches(grid, consts, libopt, kw, kernel):
    libopt(k2_sparse=1 if kernel == "sparse" else 0)
    cfg = _small_cfg(200, [15.0, 70.0], np.arange(0, 360, 20.0), [10.0, 55.0], seed=4)
    got = grid.yield_grid_host(**cfg, consts=consts, **kw)
    want = _oracle_cube(cfg, consts, **kw)
    for k, pol in enumerate(cases.POLS):
        ok, msg = cases.parity_ok(got[:, :, k], want[:, :, k], rtol=RTOL, scale=np.max(want))
        assert ok, (kw, pol, msg)


@pytest.mark.parametrize("kernel", ["dense", "sparse"])
def test_grid_shards_are_bitwise_slices(grid, consts, libopt, kernel):
    """Sharding changes no arithmetic: every shard equals the slice of the full cube bit for bit."""
    libopt(k2_sparse=1 if kernel == "sparse" else 0)
    cfg = _small_cfg(300, np.linspace(0, 90, 11), np.arange(0, 360, 10.0), [4.0, 9.0, 14.0], seed=2)
    full = grid.yield_grid_host(**cfg, consts=consts)
    for world in (2, 4, 8):
        for rank in range(world):
            lo, hi = grid.shard_bounds(11, world, rank)
            part = grid.yield_grid_host(**cfg, consts=consts, t_range=(lo, hi))
            assert np.array_equal(part, full[lo:hi])
    lo, hi = grid.shard_bounds(3, 2, 1)
    assert np.array_equal(grid.yield_grid_host(**cfg, consts=consts, d_range=(lo, hi)), full[:, lo:hi])
    # more ranks than thetas: the surplus ranks hold an empty shard, which is a legal no-op
    lo, hi = grid.shard_bounds(11, 16, 13)
    assert lo == hi == 11
    assert grid.yield_grid_host(**cfg, consts=consts, t_range=(lo, hi)).shape == (0, 3, 4, 36, 300)


def test_grid_device_and_host_paths_agree(grid, consts):
    """Device-resident output vs host output through the slab pipeline (several slabs, pinned)."""
    import torch
    cfg = _small_cfg(4096, np.linspace(5, 85, 40), np.arange(0, 360, 2.0), [10.0], seed=6)   # 944 MB cube
    dev = torch.device("cuda", 0)
    put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cube = grid.yield_grid_device(put(cfg["energy"]), put(cfg["theta"]), put(cfg["phi"]), put(cfg["thick"]),
                                  put(cfg["eps1w"]), put(cfg["eps2w"]), put(cfg["chi2"]), consts=consts)
    torch.cuda.synchronize()
    host = grid.yield_grid_host(**cfg, consts=consts, pinned=True)
    assert np.array_equal(cube.cpu().numpy(), host)
    # and the sampled oracle
    rng = np.random.default_rng(0)
    it, ip, ie = rng.integers(0, 40, 2000), rng.integers(0, 180, 2000), rng.integers(0, 4096, 2000)
    e1 = {"M%d" % (m + 1): cfg["eps1w"][m][ie] for m in range(3)}
    e2 = {"M%d" % (m + 1): cfg["eps2w"][m][ie] for m in range(3)}
    ch = {k: cfg["chi2"][i][ie] for i, k in enumerate(cases.CHI2_KEYS)}
    want = orc.yield_points(cfg["energy"][ie], e1, e2, ch, cfg["theta"][it], cfg["phi"][ip], 10.0, consts=consts)
    for k, pol in enumerate(cases.POLS):
        ok, msg = cases.parity_ok(host[it, 0, k, ip, ie], want[pol], rtol=RTOL, scale=max(np.max(want[p]) for p in cases.POLS))
        assert ok, (pol, msg)
    from shgyield_b200 import _lib
    _lib.pinned_free(host)


@pytest.mark.parametrize("sigma_out", [0.0, 5.0])
@pytest.mark.parametrize("nE,phi,thetas", [
    (300, np.linspace(0, 360, 721), 6),            # pairs + one un-paired row; several columns per slab
    (1024, np.arange(0, 360, 1.0), 3),             # pairs only
    (201, np.linspace(0, 360, 41), 4),             # odd nE, pairs + one un-paired row
    (128, np.linspace(0, 350, 17), 5),             # no pairing: nothing to duplicate
    (20000, np.arange(0, 360, 0.5), 3),            # 461 MB per column: one column per slab, three slabs
])
def test_host_output_copies_equal_rows_once(grid, consts, nE, phi, thetas, sigma_out):
    """Host-output sweeps into page-locked memory leave the second sS row of every (phi, phi + 180 deg) pair on the device
    and duplicate the first on the host (one eighth less PCIe traffic).  The rows are the same bits, so the result is
    bitwise what the plain copy (option pair_copy = 0), a pageable destination and the device path deliver -- with the
    pair form, the quad form and the fused / two-pass broadening."""
    import torch
    from shgyield_b200 import _lib
    cfg = _small_cfg(nE, np.linspace(3, 80, thetas), phi, [10.0], seed=11)
    shape = (thetas, 1, 4, phi.size, nE)
    ref = None
    variants = (dict(pair_copy=0), dict()) if nE > 1024 else (dict(pair_copy=0), dict(), dict(quads=1), dict(k3=0) if sigma_out else dict())
    for opts in variants:
        host = _lib.pinned_empty(shape)
        host[...] = -1.0
        with util.lib_options(**opts):
            grid.yield_grid_host(**cfg, consts=consts, sigma_out=sigma_out, out=host)
        if ref is None:
            ref = host.copy()
            H = phi.size // 2
            if np.array_equal(phi[H:2 * H], phi[:H] + 180.0):
                assert np.array_equal(ref[:, :, 3, :H], ref[:, :, 3, H:2 * H])       # the premise
        elif opts.get("quads") or opts.get("k3") == 0:
            # other item forms / the two-pass path round differently in the last bits; equal rows stay equal
            np.testing.assert_allclose(host, ref, rtol=1e-9, atol=1e-12 * np.max(ref))
            H = phi.size // 2
            if np.array_equal(phi[H:2 * H], phi[:H] + 180.0):
                assert np.array_equal(host[:, :, 3, :H], host[:, :, 3, H:2 * H])
        else:
            assert np.array_equal(host, ref), opts
        _lib.pinned_free(host)
    pageable = grid.yield_grid_host(**cfg, consts=consts, sigma_out=sigma_out)
    assert np.array_equal(pageable, ref)
    if nE <= 1024:
        dev = torch.device("cuda", 0)
        d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in cfg.items()}
        cube = grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"],
                                      consts=consts, sigma_out=sigma_out)
        assert np.array_equal(cube.cpu().numpy(), ref)


def test_host_output_equal_rows_on_a_symmetric_material(grid, shg, spectra, consts):
    """The same on Si(111) (exact azimuthal zeros in sS, strict rows in the broadened sweep): the clamped and the
    re-evaluated rows of a (phi, phi + 180 deg) pair are the same bits too, so the single copy changes nothing."""
    import torch
    from shgyield_b200 import _lib
    m1, m2, m3, chi2 = cases.si111_material(spectra)
    energy, theta, phi = np.linspace(0.5, 4.5, 600), np.arange(0, 90, 6.0), np.arange(0, 360, 1.0)
    eps1w, eps2w, chi_rot, thick = shg._prepare(energy, m1, m2, m3, chi2, 10.0, 0, 0.0, 0.0, "3-layer")
    args = (energy, theta, phi, np.atleast_1d(thick).astype(float), eps1w, eps2w, chi_rot)
    dev = torch.device("cuda", 0)
    dargs = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in args]
    for sigma_out in (0.0, 5.0):
        cube = grid.yield_grid_device(*dargs, consts=consts, sigma_out=sigma_out).cpu().numpy()
        host = _lib.pinned_empty(cube.shape)
        grid.yield_grid_host(*args, consts=consts, sigma_out=sigma_out, out=host)
        assert np.array_equal(host, cube), sigma_out
        assert np.array_equal(cube[:, :, 3, :180], cube[:, :, 3, 180:])
        _lib.pinned_free(host)


# ---- the drop-in API against the reference's outputs -------------------------------------------------
@pytest.mark.parametrize("name,mat,kw", cases.API_CASES, ids=[c[0] for c in cases.API_CASES])
def test_api_cases_match_reference(shg, spectra, name, mat, kw):
    gold = np.load(os.path.join(G, "ref_api_cases.npz"))
    m1, m2, m3, chi2 = cases.MATERIALS[mat](spectra)
    res = shg.shgyield(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, **kw)
    assert set(res) == {"energy", "phi", "pp", "ps", "sp", "ss"}
    assert res["energy"] is kw["energy"] and res["phi"] is kw["phi"]
    sc = cases.case_scale(gold, name + "_")
    for pol in cases.POLS:
        want = np.asarray(gold[name + "_" + pol])
        assert np.shape(res[pol]) == want.shape and res[pol].dtype == np.float64
        ok, msg = cases.parity_ok(res[pol], want, rtol=RTOL, scale=sc)
        assert ok, (name, pol, msg)


def test_api_grid_path_matches_points_path(shg, spectra, monkeypatch):
    """The broadcast call of the 'bcast3d' case through the tiled grid kernel (threshold lowered)."""
    gold = np.load(os.path.join(G, "ref_api_cases.npz"))
    name, mat, kw = next(c for c in cases.API_CASES if c[0] == "bcast3d")
    m1, m2, m3, chi2 = cases.MATERIALS[mat](spectra)
    monkeypatch.setattr(shg, "GRID_MIN_POINTS", 1)
    res = shg.shgyield(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, **kw)
    sc = cases.case_scale(gold, name + "_")
    for pol in cases.POLS:
        ok, msg = cases.parity_ok(res[pol], gold[name + "_" + pol], rtol=RTOL, scale=sc)
        assert ok, (pol, msg)
    name, mat, kw = next(c for c in cases.API_CASES if c[0] == "bcast_thick")
    m1, m2, m3, chi2 = cases.MATERIALS[mat](spectra)
    res = shg.shgyield(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, **kw)
    for pol in cases.POLS:
        ok, msg = cases.parity_ok(res[pol], gold[name + "_" + pol], rtol=RTOL, scale=cases.case_scale(gold, name + "_"))
        assert ok, (pol, msg)


@pytest.mark.parametrize("tail", ["device", "slabs"])
@pytest.mark.parametrize("name", ["bcast3d", "bcast3d_blur", "bcast_thick"])
def test_api_broadcast_tails_match_reference(shg, spectra, monkeypatch, name, tail):
    """Large outer-product broadcasts take sshg_yield_broadcast (kernels write the result layout, N-D output broadening
    on the device, four arrays come back); without device memory for it, or with another axis order, the slab-wise
    path with host-side transposes.  Both against the reference's outputs, incl. sigma_out blurring EVERY axis."""
    gold = np.load(os.path.join(G, "ref_api_cases.npz"))
    _, mat, kw = next(c for c in cases.API_CASES if c[0] == name)
    m1, m2, m3, chi2 = cases.MATERIALS[mat](spectra)
    monkeypatch.setattr(shg, "GRID_MIN_POINTS", 1)
    calls = []
    real = shg._broadcast_tail
    monkeypatch.setattr(shg, "_broadcast_tail", lambda *a: (calls.append(1), real(*a) if tail == "device" else None)[1])
    res = shg.shgyield(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, **kw)
    assert calls, "the energy is the last axis of these cases"
    for pol in cases.POLS:
        want = gold[name + "_" + pol]
        assert res[pol].shape == want.shape and res[pol].flags["C_CONTIGUOUS"] and res[pol].flags["OWNDATA"]
        ok, msg = cases.parity_ok(res[pol], want, rtol=RTOL, scale=cases.case_scale(gold, name + "_"))
        assert ok, (name, tail, pol, msg)


def test_api_broadcast_tail_large_and_other_axis_orders(shg, spectra, monkeypatch):
    """A 173 MB broadcast (threaded staged copies into four fresh arrays) is bitwise the transposed cube of
    shgyield_grid; with sigma_out the device tail agrees with the slab-wise one; an axis order with theta last, a
    thickness axis and scalar operands go through whichever tail applies and agree with the point-list path."""
    from shgyield_b200 import grid
    m1, m2, m3, chi2 = cases.si111_material(spectra)
    E, T, P = np.linspace(0.4, 4.9, 2000), np.arange(0, 90, 3.0), np.arange(0, 360, 4.0)
    kw = dict(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, thick=10.0, gamma=0)
    res = shg.shgyield(energy=E, theta=T[:, None], phi=P[:, None, None], sigma_out=0.0, **kw)
    cube = grid.shgyield_grid(E, m1, m2, m3, chi2, T, P, 10.0, gamma=0, device=False)["cube"]     # [T, 1, 4, P, E]
    for k, pol in enumerate(cases.POLS):
        assert np.array_equal(res[pol], np.transpose(cube[:, 0, k], (1, 0, 2))), pol
    blur = shg.shgyield(energy=E, theta=T[:, None], phi=P[:, None, None], sigma_out=1.5, **kw)
    with monkeypatch.context() as mp:
        mp.setattr(shg, "_broadcast_tail", lambda *a: None)
        slabs = shg.shgyield(energy=E, theta=T[:, None], phi=P[:, None, None], sigma_out=1.5, **kw)
    for pol in cases.POLS:
        np.testing.assert_allclose(blur[pol], slabs[pol], rtol=0, atol=1e-14 * np.max(slabs[pol]))
    # other structures (small, through the lowered threshold): thickness axis first, phi scalar; theta last
    monkeypatch.setattr(shg, "GRID_MIN_POINTS", 1)
    e, t, d = np.linspace(0.5, 4.5, 60), np.array([10.0, 40.0, 70.0]), np.arange(1.0, 9.0)
    shapes = [dict(energy=e, theta=t[:, None], phi=30.0, thick=d[:, None, None]),
              dict(energy=e[:, None], theta=t, phi=30.0, thick=12.0),
              dict(energy=e, theta=t[:, None, None], phi=np.array([0.0, 45.0])[:, None], thick=d[:, None, None, None])]
    for sh in shapes:
        for so in (0.0, 0.8):
            got = shg.shgyield(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, gamma=20, sigma_out=so, **sh)
            with monkeypatch.context() as mp:
                mp.setattr(shg, "GRID_MIN_POINTS", 10 ** 12)          # point-list path
                want = shg.shgyield(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, gamma=20, sigma_out=so, **sh)
            sc = max(np.max(want[p]) for p in cases.POLS)
            for pol in cases.POLS:
                ok, msg = cases.parity_ok(got[pol], want[pol], rtol=RTOL, scale=sc)
                assert ok and got[pol].shape == want[pol].shape, (sorted(sh), so, pol, msg)


def test_golden_si111_reference_out(shg, spectra):
    """The reference's own golden file example/reference/si111-reference.out (printed %.8e)."""
    gold = np.load(os.path.join(G, "si111_reference_out.npz"))
    m1, m2, m3, chi2 = cases.si111_material(spectra)
    res = shg.shgyield(cases.E_C1, m1, m2, m3, chi2, theta=65, phi=30, thick=10, gamma=0, sigma_eps=0.0,
                       sigma_chi=0.0, sigma_out=5)
    for pol in cases.POLS:
        np.testing.assert_allclose(res[pol], gold[pol], rtol=6e-9, atol=0)


def test_golden_selftest_reference(shg, spectra, consts):
    """REFERENCE of the reference's tests/self_test.py: legacy sS numbers (un-normalised sinc,
    CODATA-2014 eps0 hard-coded in self_test.py:104), tolerance of the self-test itself."""
    gold = np.load(os.path.join(G, "selftest_reference.npz"))["REFERENCE"]
    m1, m2, m3, chi2 = cases.selftest_material(spectra)
    energy = np.linspace(0.01, 1.00, 100)
    eps1w, eps2w, chi_rot, thick = shg._prepare(energy, m1, m2, m3, chi2, 10, 0, 0.0, 0.0, "3-layer")
    c = dict(consts, eps0=8.854187817620389e-12)
    y = shg.yield_points(energy, np.arange(100), 65.0, 30.0, thick, eps1w, eps2w, chi_rot, consts=c,
                         sinc_normalized=False)
    np.testing.assert_allclose(1e20 * y[3], gold, rtol=1e-6, atol=1e-12)


def test_error_behaviour(shg, spectra):
    m1, m2, m3, chi2 = cases.si111_material(spectra)
    kw = dict(theta=65, phi=30, thick=10, gamma=0)
    with pytest.raises(ValueError):                 # 2E leaves the 0-20 eV data (SciPy ext=2)
        shg.shgyield(np.linspace(9.0, 10.5, 10), m1, m2, m3, chi2, **kw)
    with pytest.raises(KeyError):                   # a chi2 key is missing (reference shg.py:66-133)
        shg.shgyield(cases.E_C1, m1, m2, m3, {k: v for k, v in chi2.items() if k != "zxy"}, **kw)
    with pytest.raises(ValueError):                 # unknown mode (documented deviation: reference -> UnboundLocalError)
        shg.shgyield(cases.E_C1, m1, m2, m3, chi2, mode="4-layer", **kw)
    with pytest.raises(ValueError):                 # a medium mixing numbers and spectra (np.mean fails, shg.py:60)
        shg.shgyield(cases.E_C1, m1, dict(m2, zz=1.0), m3, chi2, **kw)
    y = shg.shgyield(np.array([0.0, 1.5]), m1, m2, m3, chi2, sigma_out=0.0, **kw)      # E = 0 -> 0, no exception
    assert y["pp"][0] == 0.0 and np.isfinite(y["pp"][1])


# ---- the large grids of BASELINE.json -----------------------------------------------------------------
@pytest.mark.parametrize("tag", ["c3", "c4"])
def test_config3_config4_full_size(grid, spectra, tag):
    """Full-size sweeps of the Si(111)(1x1):H material, device resident, against sampled outputs of
    the reference plus size-independent properties."""
    import torch
    z = np.load(os.path.join(G, "ref_grid_samples.npz"))
    axes = cases.config3_axes() if tag == "c3" else cases.config4_axes()
    m1, m2, m3, chi2 = cases.si111_material(spectra)
    res = grid.shgyield_grid(axes["energy"], m1, m2, m3, chi2, axes["theta"], axes["phi"], axes["thick"],
                             gamma=axes["gamma"], sigma_out=0.0)
    torch.cuda.synchronize()
    cube = res["cube"]
    idx = torch.from_numpy(z[tag + "_idx"].astype(np.int64)).to(cube.device)
    sc = cases.case_scale(z, tag + "_")
    for k, pol in enumerate(cases.POLS):
        got = cube[idx[:, 0], idx[:, 1], k, idx[:, 2], idx[:, 3]].cpu().numpy()
        ok, msg = cases.parity_ok(got, z[tag + "_" + pol], rtol=RTOL, scale=sc)
        assert ok, (tag, pol, msg)
    assert bool(torch.isfinite(cube).all()) and float(cube.min()) >= 0.0
    if tag == "c3":
        # Si(111) C3v: R(phi) has period 120 deg; sS is symmetric under phi -> phi + 180
        ss = res["ss"][:, 0]
        scale = float(ss.max())
        assert float((ss[:, :120] - ss[:, 120:240]).abs().max()) <= 1e-10 * scale
        assert float((ss[:, :180] - ss[:, 180:]).abs().max()) <= 1e-10 * scale
    del cube, res
    torch.cuda.empty_cache()


def test_config5_full_size_in_slabs(grid, consts):
    """The 20k x 181 x 721 x 4 cube (83.5 GB) computed as theta slabs; every slab is checked against
    the sampled reference outputs that fall into it, and phi = 0 equals phi = 360."""
    import torch
    from shgyield_b200.synthetic import config5
    z = np.load(os.path.join(G, "ref_grid_samples.npz"))
    idx_all = z["c5_idx"].astype(np.int64)
    cfg = config5()
    dev = torch.device("cuda", 0)
    put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = {k: put(v) for k, v in cfg.items()}
    sc = cases.case_scale(z, "c5_")
    step = 16
    out = torch.empty((step, 1, 4, 721, 20000), dtype=torch.float64, device=dev)
    seen = 0
    for t0 in range(0, 181, step):
        t1 = min(181, t0 + step)
        view = out[: t1 - t0]
        grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"],
                               out=view, consts=consts, t_range=(t0, t1))
        torch.cuda.synchronize()
        sel = (idx_all[:, 0] >= t0) & (idx_all[:, 0] < t1)
        idx = torch.from_numpy(idx_all[sel]).to(dev)
        for k, pol in enumerate(cases.POLS):
            got = view[idx[:, 0] - t0, 0, k, idx[:, 2], idx[:, 3]].cpu().numpy()
            ok, msg = cases.parity_ok(got, z["c5_" + pol][sel], rtol=RTOL, scale=sc)
            assert ok, (t0, pol, msg)
        seen += int(sel.sum())
        top = float(view.max())
        assert float((view[:, :, :, 0] - view[:, :, :, 720]).abs().max()) <= 1e-10 * top
        assert bool(torch.isfinite(view).all())
    assert seen == idx_all.shape[0]
    assert float(out[180 - 176].abs().max()) == 0.0          # theta = 90 deg: exactly zero (w_v = 0)
    # linearity: r is linear in chi2, so doubling chi2 (exact in binary FP) quadruples every yield bit for bit
    grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"],
                           out=out, consts=consts, t_range=(40, 40 + step))
    out2 = torch.empty_like(out)
    grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], 2.0 * d["chi2"],
                           out=out2, consts=consts, t_range=(40, 40 + step))
    torch.cuda.synchronize()
    assert bool(torch.equal(out2, 4.0 * out))


def test_config5_full_theta_rows_vs_oracle(grid, consts):
    """Two complete theta rows of the 20k x 181 x 721 cube (2 x 57.7 M points, every phi and energy)
    against the CPU oracle, not just sampled points."""
    import torch
    from shgyield_b200.synthetic import config5
    cfg = config5()
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in cfg.items()}
    e1 = {"M%d" % (m + 1): cfg["eps1w"][m] for m in range(3)}
    e2 = {"M%d" % (m + 1): cfg["eps2w"][m] for m in range(3)}
    ch = {k: cfg["chi2"][i] for i, k in enumerate(cases.CHI2_KEYS)}
    for it in (65, 179):                                     # 32.5 deg and 89.5 deg (grazing: huge prefactor)
        got = grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"],
                                     consts=consts, t_range=(it, it + 1)).cpu().numpy()[0, 0]      # [4, P, E]
        worst = 0.0
        for p0 in range(0, 721, 103):
            phi = cfg["phi"][p0:p0 + 103]
            want = orc.yield_points(cfg["energy"], e1, e2, ch, cfg["theta"][it], phi[:, None], 10.0, True, consts)
            for k, pol in enumerate(cases.POLS):
                a, b = got[k, p0:p0 + 103], want[pol]
                worst = max(worst, float(np.max(np.abs(a - b)) / np.max(b)))
        assert worst < 1e-12, (it, worst)


def test_prepared_cache_is_transparent(shg, spectra):
    """The reuse of prepared spectra (on by default, keyed on CONTENT): same results bit for bit as without it, a hit
    when only angles / thickness / sigma_out change or when an equal copy of the material comes in, a miss when
    sigma_eps, gamma, the energies or the VALUES of a spectrum change -- including a change made in place."""
    from shgyield_b200 import _lib
    m1, m2, m3, chi2 = cases.full_tensor_material(spectra)
    chi2 = {k: ({"energy": v["energy"], "data": v["data"].copy()} if isinstance(v, dict) else v) for k, v in chi2.items()}
    kw = dict(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, thick=10, gamma=25, sigma_eps=2.0, sigma_chi=1.0)
    polar_phi = np.arange(0.0, 360.0, 45.0).reshape(8, 1)
    ctx = _lib.context()
    shg.cache_prepared(False)
    base = shg.shgyield(cases.E_C1, theta=65, phi=30, sigma_out=5, **kw)
    other = shg.shgyield(cases.E_C1, theta=40, phi=polar_phi, sigma_out=0.0, **kw)
    f0 = shg.shgyield(cases.E_C1, theta=65, phi=30, sigma_out=5, mode="fresnel", **kw)
    shg.cache_prepared(True)
    try:
        a = shg.shgyield(cases.E_C1, theta=65, phi=30, sigma_out=5, **kw)            # miss: fills the cache
        n0 = ctx.launches
        b = shg.shgyield(cases.E_C1, theta=65, phi=30, sigma_out=5, **kw)            # hit
        n1 = ctx.launches
        c = shg.shgyield(cases.E_C1, theta=40, phi=polar_phi, sigma_out=0.0, **kw)
        assert n1 - n0 == 2                                                          # yield kernel + output FIR only
        for pol in cases.POLS:
            assert np.array_equal(a[pol], base[pol]) and np.array_equal(b[pol], base[pol])
            assert np.array_equal(c[pol], other[pol])
        assert len(shg._PREPARE_CACHE) == 1
        shg.shgyield(cases.E_C1, theta=65, phi=30, sigma_out=5, **dict(kw, sigma_eps=3.0))
        shg.shgyield(cases.E_C1, theta=65, phi=30, sigma_out=5, **dict(kw, gamma=26))
        shg.shgyield(cases.E_C1[:-1], theta=65, phi=30, sigma_out=5, **kw)
        assert len(shg._PREPARE_CACHE) == 4
        # an equal COPY of the material is the same content: a hit
        copy = {k: ({"energy": v["energy"].copy(), "data": v["data"].copy()} if isinstance(v, dict) else v) for k, v in chi2.items()}
        shg.shgyield(cases.E_C1, theta=65, phi=30, sigma_out=5, **dict(kw, chi2=copy))
        assert len(shg._PREPARE_CACHE) == 4
        # a spectrum modified IN PLACE must miss, and the result must be the one of the modified material
        saved = chi2["xxx"]["data"][150:170].copy()
        chi2["xxx"]["data"][150:170] *= 1.5
        d = shg.shgyield(cases.E_C1, theta=65, phi=30, sigma_out=5, **kw)
        assert len(shg._PREPARE_CACHE) == 5
        shg.cache_prepared(False)
        d0 = shg.shgyield(cases.E_C1, theta=65, phi=30, sigma_out=5, **kw)
        shg.cache_prepared(True)
        assert any(not np.array_equal(d[pol], base[pol]) for pol in cases.POLS)
        for pol in cases.POLS:
            assert np.array_equal(d[pol], d0[pol])
        chi2["xxx"]["data"][150:170] = saved
        f = shg.shgyield(cases.E_C1, theta=65, phi=30, sigma_out=5, mode="fresnel", **kw)
        g = shg.shgyield(cases.E_C1, theta=65, phi=30, sigma_out=5, mode="fresnel", **kw)
        for pol in cases.POLS:
            assert np.array_equal(f[pol], g[pol]) and np.array_equal(f[pol], f0[pol])
    finally:
        shg.cache_prepared(True)                                                     # the default


def test_small_calls_replay_a_graph_with_the_same_bits(shg, spectra):
    """The reference GUI's two call shapes (gui.py:286-330: 361-phi polar plot, 126-energy spectrum with sigma_out):
    after the first call of a shape the library replays a recorded CUDA graph; results are bit for bit those of the
    kernel-by-kernel path (option "graphs" = 0) and match the reference-generated goldens."""
    from shgyield_b200 import _lib
    m1, m2, m3, chi2 = cases.si111_material(spectra)
    kw = dict(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, theta=65, thick=10, gamma=0)
    ctx = _lib.context()
    shapes = (dict(energy=1.7, phi=np.arange(0, 361), sigma_out=0.0), dict(energy=cases.E_C1, phi=30, sigma_out=5))
    with ctx.options(graphs=0):
        plain = [shg.shgyield(**s, **kw) for s in shapes]
    first = [shg.shgyield(**s, **kw) for s in shapes]                  # runs kernel by kernel; the combination is noted
    second = [shg.shgyield(**s, **kw) for s in shapes]                 # second sighting: kernel by kernel, then recorded
    n0 = ctx.launches
    again = [shg.shgyield(**s, **kw) for s in shapes]                  # replays
    assert ctx.launches - n0 == 1 + 2                                  # polar: yield kernel; spectrum: yield kernel + FIR
    moved = shg.shgyield(energy=cases.E_C1, phi=31.5, sigma_out=5, **dict(kw, theta=40))     # same shape, new angles: replay
    with ctx.options(graphs=0):
        moved0 = shg.shgyield(energy=cases.E_C1, phi=31.5, sigma_out=5, **dict(kw, theta=40))
    for pol in cases.POLS:
        for a, b, c, d in zip(plain, first, second, again):
            assert np.array_equal(a[pol], b[pol]) and np.array_equal(a[pol], c[pol]) and np.array_equal(a[pol], d[pol])
        assert np.array_equal(moved[pol], moved0[pol])
    gold = np.load(os.path.join(G, "ref_api_cases.npz"))
    for name, res in (("polar", again[0]), ("c1_d10", again[1])):
        sc = cases.case_scale(gold, name + "_")
        for pol in cases.POLS:
            ok, msg = cases.parity_ok(res[pol], gold[name + "_" + pol], rtol=RTOL, scale=sc)
            assert ok, (name, pol, msg)


def test_gauss_nd_matches_scipy(shg):
    """broad() of an N-D array = scipy.ndimage.gaussian_filter over every axis, one library call (middle axes are
    filtered in place on the device with outer x inner row addressing)."""
    from scipy import ndimage
    rng = np.random.default_rng(12)
    for shape, sigma in (((7, 5, 33), 1.3), ((4, 1, 9, 2), 0.8), ((3, 40), 5.0), ((2, 3, 4, 5, 6), 0.6), ((1, 1), 2.0)):
        x = rng.normal(size=shape)
        want = ndimage.gaussian_filter(x, sigma)
        got = shg.broad(x, sigma)
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) <= 4e-15 * max(1.0, np.max(np.abs(want))), (shape, sigma)


def test_gamma_only_changes_reuse_the_resampled_spectra(shg, spectra):
    """A call that differs from an earlier one only in gamma (the GUI's rotation slider) redoes the rotation alone; the
    result is bit for bit the one of a fresh preparation, also for the modes that override the thickness."""
    m1, m2, m3, chi2 = cases.full_tensor_material(spectra)
    kw = dict(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, theta=50, phi=20, sigma_eps=1.0, sigma_chi=2.0, sigma_out=3.0)
    for mode, thick in (("3-layer", 12.0), ("fresnel", 12.0), ("3-layer", np.array([[5.0], [25.0]]))):
        shg.cache_prepared(True)
        shg.shgyield(cases.E_C1, thick=thick, gamma=10, mode=mode, **kw)              # fills both levels
        got = shg.shgyield(cases.E_C1, thick=thick, gamma=37.5, mode=mode, **kw)      # new gamma: resampled spectra reused
        shg.cache_prepared(False)
        want = shg.shgyield(cases.E_C1, thick=thick, gamma=37.5, mode=mode, **kw)
        shg.cache_prepared(True)
        for pol in cases.POLS:
            assert got[pol].shape == want[pol].shape and np.array_equal(got[pol], want[pol]), (mode, pol)
