"""
GPU parity tests of the broadened sweep (sshg_yield_broadened, K3): the yield cube with the output
broadening of shgyield() (broad(., sigma_out) along the energy axis, reference shg.py:26-28, 433-436)
computed in one pass, against the CPU oracle (un-broadened yields -> scipy-style gaussian filter row by
row) and against the two-pass device path (K2 -> K1, library option "k3" = 0).

Tolerance: tests/cases.py::parity_ok with rtol 1e-10 in the strict metric (element-wise wherever the yield is
>= 1e-6 of the maximum).  The harmonic sum of K3 alone has an absolute error (a few 1e-15 of the column's
maximum over phi, checked here as < 5e-14); the segments where that is not enough -- values below 1e-3 of the
column's amplitude bound, i.e. next to an azimuthal node -- are flagged by K3's sweep and re-evaluated row-wise
inside the same CTA, so the product holds the strict metric (option "k3_fixup" = 0 shows the harmonic sums alone).
"""
import os

import numpy as np
import pytest

from oracle import shg_oracle as orc
from tests import cases, util
from tests.test_gpu_parity import _oracle_cube, _small_cfg

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope="module")
def grid():
    from shgyield_b200 import grid as mod
    return mod


@pytest.fixture(scope="module")
def consts():
    return util.load_consts()


def _blurred_oracle(cfg, consts, sigma, **kw):
    return orc.gaussian_filter1d(_oracle_cube(cfg, consts, **kw), sigma, axis=-1)


def _check(got, want, fused, what, col_bound=5e-14):
    """parity metric + K3's absolute bound.  col_bound: 5e-14 of the column's maximum over phi on the smooth
    spectra of these tests; the generated media of tests/test_gpu_hypothesis.py (gain media, |exponents| of
    several hundred) make the yield itself ill-conditioned -- there K2 and the oracle already differ by
    ~5e-14 relative -- and pass a looser bound."""
    assert got.shape == want.shape
    top = np.max(want)
    for k, pol in enumerate(cases.POLS):
        a, b = got[:, :, k], want[:, :, k]
        ok, msg = cases.parity_ok(a, b, rtol=RTOL, scale=top)      # K3 + fix-up and the two-pass path: the strict metric
        assert ok, (what, pol, msg)
        colmax = np.maximum(np.max(b, axis=2, keepdims=True), 1e-12 * top)       # max over phi of every column
        assert np.max(np.abs(a - b) / colmax) < col_bound, (what, pol)


BLUR_SHAPES = [
    # nE, theta, phi, thick
    (1, [65.0], np.arange(0, 360, 20.0), [10.0]),                 # one energy: every reflection lands on it
    (126, [65.0], np.arange(0, 360, 20.0), [10.0]),               # one tile, ragged
    (513, [0.0, 33.0, 90.0], np.arange(0, 360, 15.0), [10.0]),    # odd nE (64-bit stores), pairs, 5 tiles
    (600, [10.0, 80.0], np.linspace(0, 360, 41), [0.0, 7.5]),     # pairs + one un-paired row, two thicknesses
    (64, [45.0], np.linspace(0, 350, 17), [3.0]),                 # no pairing possible
    (30, [55.0], np.linspace(0, 360, 1001), [12.0]),              # several phi chunks, nE < radius
    (1030, [20.0, 40.0], np.arange(0, 360, 22.5), [5.0]),         # tile boundaries inside the axis
    (70, [45.0], np.arange(0, 360, 0.5), [7.0]),                  # 720 phi: quads / pairs over several table chunks at large radii
]


@pytest.mark.parametrize("path", ["fused", "fused64", "fused128", "fused_quads", "fused_pairs", "two_pass"])
@pytest.mark.parametrize("sigma", [0.3, 1.2, 5.0, 16.0, 23.0, 50.0])
@pytest.mark.parametrize("nE,theta,phi,thick", BLUR_SHAPES, ids=[str(i) for i in range(len(BLUR_SHAPES))])
def test_broadened_sweep_matches_oracle(grid, consts, libopt, nE, theta, phi, thick, sigma, path):
    """radius 1, 5, 20 (the default sigma_out = 5), 64 (two CTAs per SM up to there), 92 and 200 (the largest fused
    radius: sigma_out = 50, the end of the reference GUI's slider); K3 with the tile width
    the library picks and with 128- / 256-energy tiles forced."""
    libopt(k3=0 if path == "two_pass" else 1)
    if path in ("fused64", "fused128"):
        libopt(k3_nt=int(path[5:]))
    if path in ("fused_quads", "fused_pairs"):
        libopt(quads=1 if path == "fused_quads" else 0)
    cfg = _small_cfg(nE, theta, phi, thick, seed=3)
    got = grid.yield_grid_host(**cfg, consts=consts, sigma_out=sigma)
    _check(got, _blurred_oracle(cfg, consts, sigma), path != "two_pass", (nE, sigma, path))


def test_sparse_phi_sweeps_and_large_radii_take_the_two_pass_path(grid, consts):
    """P = 1 (config-4 shape) and sigma_out > 50: K2 / K2s -> K1 slab by slab; P = 1 forced through K3."""
    cfg = _small_cfg(300, np.arange(0, 90, 9.0), [30.0], np.arange(1.0, 31.0), seed=5)
    want = _blurred_oracle(cfg, consts, 5.0)
    _check(grid.yield_grid_host(**cfg, consts=consts, sigma_out=5.0), want, False, "sparse")
    with util.lib_options(k2_sparse=0):
        _check(grid.yield_grid_host(**cfg, consts=consts, sigma_out=5.0), want, True, "P=1 through K3")
    cfg = _small_cfg(400, [30.0, 60.0], np.arange(0, 360, 30.0), [10.0], seed=6)
    _check(grid.yield_grid_host(**cfg, consts=consts, sigma_out=20.0), _blurred_oracle(cfg, consts, 20.0), True, "r=80")
    _check(grid.yield_grid_host(**cfg, consts=consts, sigma_out=60.0), _blurred_oracle(cfg, consts, 60.0), False, "r=240")
    # a radius of 0 (sigma_out < 0.125) is the identity
    assert np.array_equal(grid.yield_grid_host(**cfg, consts=consts, sigma_out=0.1), grid.yield_grid_host(**cfg, consts=consts))


@pytest.mark.parametrize("kw", [dict(mref=False), dict(sinc_normalized=False), dict(zzz_phi_quirk=False)])
def test_broadened_sweep_switches(grid, consts, kw):
    cfg = _small_cfg(200, [15.0, 70.0], np.arange(0, 360, 20.0), [10.0, 55.0], seed=4)
    got = grid.yield_grid_host(**cfg, consts=consts, sigma_out=2.0, **kw)
    _check(got, _blurred_oracle(cfg, consts, 2.0, **kw), True, kw)


def test_broadened_shards_are_bitwise_slices(grid, consts):
    cfg = _small_cfg(300, np.linspace(0, 90, 11), np.arange(0, 360, 10.0), [4.0, 9.0, 14.0], seed=2)
    full = grid.yield_grid_host(**cfg, consts=consts, sigma_out=5.0)
    for world in (2, 8):
        for rank in range(world):
            lo, hi = grid.shard_bounds(11, world, rank)
            assert np.array_equal(grid.yield_grid_host(**cfg, consts=consts, sigma_out=5.0, t_range=(lo, hi)), full[lo:hi])
    lo, hi = grid.shard_bounds(3, 2, 1)
    assert np.array_equal(grid.yield_grid_host(**cfg, consts=consts, sigma_out=5.0, d_range=(lo, hi)), full[:, lo:hi])


def test_nan_sample_stays_within_the_radius(grid, consts):
    """A NaN in one spectrum sample: the reference's broadened yields are NaN exactly within the radius
    of that energy, for every angle and polarisation that uses the sample."""
    cfg = _small_cfg(400, [30.0], np.arange(0, 360, 30.0), [10.0], seed=8)
    cfg["chi2"] = cfg["chi2"].copy()
    cfg["chi2"][:, 200] = np.nan
    got = grid.yield_grid_host(**cfg, consts=consts, sigma_out=2.0)          # radius 8
    want = _blurred_oracle(cfg, consts, 2.0)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.isnan(got[0, 0, 0, 0, 192:209]).all() and not np.isnan(got[0, 0, 0, 0, :192]).any()


def test_broadened_device_host_and_two_pass_paths_agree(grid, consts, libopt):
    """Device-resident output, the host slab pipeline (several slabs) and the two-pass path on a 472 MB cube."""
    import torch
    cfg = _small_cfg(4096, np.linspace(5, 85, 20), np.arange(0, 360, 2.0), [10.0], seed=6)
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in cfg.items()}
    run = lambda: grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"],
                                         consts=consts, sigma_out=5.0)
    fused = run()
    host = grid.yield_grid_host(**cfg, consts=consts, sigma_out=5.0)
    assert np.array_equal(fused.cpu().numpy(), host)
    libopt(k3=0)
    two = run()
    torch.cuda.synchronize()
    top = float(two.max())
    colmax = two.amax(dim=3, keepdim=True).clamp_min(1e-12 * top)
    assert float(((fused - two).abs() / colmax).max()) < 5e-14
    assert float((fused - two).abs().max()) < 1e-13 * top


def test_api_sweep_with_output_broadening_matches_the_reference_loop(grid):
    """shgyield_grid(sigma_out=5) against the oracle's shgyield() called angle by angle (the only way the
    reference can broaden a sweep along the energy axis alone)."""
    m1, m2, m3, chi2 = cases.full_tensor_material(cases.load_spectra())
    energy = np.linspace(1.0, 3.0, 201)
    theta, phi = np.array([25.0, 65.0]), np.arange(0.0, 360.0, 40.0)
    res = grid.shgyield_grid(energy, m1, m2, m3, chi2, theta, phi, 10.0, gamma=35, sigma_out=5.0, device=False)
    for it, th in enumerate(theta):
        for ip, ph in enumerate(phi):
            ref = orc.shgyield(energy, m1, m2, m3, chi2, th, ph, 10.0, gamma=35, sigma_out=5.0)
            sc = max(np.max(ref[p]) for p in cases.POLS)
            for pol in cases.POLS:
                ok, msg = cases.parity_ok(res[pol][it, 0, ip], ref[pol], rtol=RTOL, scale=sc)
                assert ok, (th, ph, pol, msg)


def test_config5_shard_broadened_full_rows_vs_oracle(grid, consts):
    """A theta shard of the 20k x 181 x 721 cube with sigma_out = 5, complete (theta, phi) rows against the oracle."""
    import torch
    from shgyield_b200.synthetic import config5
    cfg = config5()
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in cfg.items()}
    t0, t1 = 60, 68
    cube = grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"],
                                  consts=consts, t_range=(t0, t1), sigma_out=5.0)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(cube).all())
    top = float(cube.max())
    assert float((cube[:, :, :, 0] - cube[:, :, :, 720]).abs().max()) <= 1e-12 * top      # phi = 0 == phi = 360
    e1 = {"M%d" % (m + 1): cfg["eps1w"][m] for m in range(3)}
    e2 = {"M%d" % (m + 1): cfg["eps2w"][m] for m in range(3)}
    ch = {k: cfg["chi2"][i] for i, k in enumerate(cases.CHI2_KEYS)}
    ips = np.array([0, 1, 77, 360, 361, 500, 719, 720])
    for it in (t0, t0 + 5, t1 - 1):
        raw = orc.yield_points(cfg["energy"], e1, e2, ch, cfg["theta"][it], cfg["phi"][ips][:, None], 10.0, True, consts)
        got = cube[it - t0, 0][:, torch.from_numpy(ips).to(dev)].cpu().numpy()            # [4, len(ips), E]
        for k, pol in enumerate(cases.POLS):
            want = orc.gaussian_filter1d(raw[pol], 5.0, axis=-1)
            assert np.max(np.abs(got[k] - want)) < 1e-13 * np.max(want), (it, pol)


@pytest.mark.parametrize("nseg", ["2", "3", "5", "16"])
@pytest.mark.parametrize("sigma", [0.0, 2.0])
def test_phi_segments_change_nothing(grid, consts, nseg, sigma):
    """Short launches cut the phi items of a column into segments, one CTA each (K2; K3 under the option "nseg"):
    uneven splits, more segments than items, pairs + un-paired tail -- bitwise the un-segmented result."""
    for phi in (np.arange(0, 360, 2.5), np.linspace(0, 360, 41), np.linspace(0, 350, 7)):
        cfg = _small_cfg(300, [20.0, 70.0], phi, [10.0], seed=9)
        with util.lib_options(nseg=1):
            one = grid.yield_grid_host(**cfg, consts=consts, sigma_out=sigma)
        with util.lib_options(nseg=int(nseg)):
            many = grid.yield_grid_host(**cfg, consts=consts, sigma_out=sigma)
        assert np.array_equal(one, many), (nseg, sigma, phi.size)


@pytest.mark.parametrize("path", ["fused", "two_pass"])
@pytest.mark.parametrize("mat", ["si111", "full"])
def test_broadened_sweep_vs_reference_golden(grid, libopt, mat, path):
    """shgyield_grid(sigma_out) against rows produced by the UNMODIFIED reference, shgyield() called once per
    (theta, phi) (tests/golden/ref_blur_rows.npz; phi includes points 0.02 deg from the C3v nodes of Si(111))."""
    libopt(k3=1 if path == "fused" else 0, k2_sparse=0)                  # 9 azimuths: K3 only when the dense path is forced
    z = np.load(os.path.join(cases.GOLDEN_DIR, "ref_blur_rows.npz"))
    m1, m2, m3, chi2 = cases.MATERIALS[mat](cases.load_spectra())
    for sigma in cases.BLUR_SWEEP_SIGMAS:
        res = grid.shgyield_grid(z["energy"], m1, m2, m3, chi2, z["theta"], z["phi"], cases.BLUR_SWEEP_THICK,
                                 gamma=cases.BLUR_SWEEP_GAMMA[mat], sigma_out=sigma, device=False)
        want = z["%s_sigma%g" % (mat, sigma)]                           # [T, 4, P, E]
        got = res["cube"][:, 0]                                         # [T, 4, P, E]
        top = float(np.max(want))
        for k, pol in enumerate(cases.POLS):
            ok, msg = cases.parity_ok(got[:, k], want[:, k], rtol=RTOL, scale=top)
            assert ok, (mat, sigma, path, pol, msg)
            for it in range(want.shape[0]):                     # row by row: the rows 0.02 deg from a node included
                ok, msg = cases.rows_parity_ok(got[it, k], want[it, k])
                assert ok, (mat, sigma, path, pol, "theta", it, msg)
            colmax = np.maximum(np.max(want[:, k], axis=1, keepdims=True), 1e-12 * top)
            assert np.max(np.abs(got[:, k] - want[:, k]) / colmax) < 5e-14, (mat, sigma, path, pol)


@pytest.mark.parametrize("tag", ["c3", "c4"])
def test_config3_config4_full_size_broadened(grid, tag):
    """Configs 3 and 4 of BASELINE.json at full size WITH the reference's default output broadening
    (sigma_out = 5): complete (theta, thickness, phi) rows against the oracle's shgyield() called for that
    angle -- including the exact C3v node rows of Si(111) -- plus the symmetries of the un-broadened cube,
    which the broadening along E preserves.  Config 3 takes K3, config 4 (P = 1) K2s -> K1."""
    import torch
    axes = cases.config3_axes() if tag == "c3" else cases.config4_axes()
    m1, m2, m3, chi2 = cases.si111_material(cases.load_spectra())
    res = grid.shgyield_grid(axes["energy"], m1, m2, m3, chi2, axes["theta"], axes["phi"], axes["thick"],
                             gamma=axes["gamma"], sigma_out=5.0)
    torch.cuda.synchronize()
    cube = res["cube"]                                                  # [T, D, 4, P, E]
    assert bool(torch.isfinite(cube).all())
    top = float(cube.max())
    assert float(cube.min()) >= -1e-13 * top                            # the harmonic sum may undershoot 0 by rounding
    # complete rows: per (theta, thickness) a set of azimuths that includes exact C3v nodes (0, 30, 90, 180 deg),
    # near-node rows (1, 91) and an off-node row (45), each against the oracle in the strict metric ROW BY ROW
    # (K3 flags the near-node segments and the fix-up kernel re-evaluates them); M = max over those rows of the
    # oracle, a lower bound of the column's maximum over phi, for K3's absolute bound.
    cols = {"c3": [(0, 0), (65, 0), (89, 0)], "c4": [(0, 0), (65, 9), (89, 99)]}[tag]
    phis = [0, 1, 30, 45, 90, 91, 180, 359] if tag == "c3" else [0]
    for it, idd in cols:
        ref = {ip: orc.shgyield(axes["energy"], m1, m2, m3, chi2, axes["theta"][it], axes["phi"][ip], axes["thick"][idd],
                                gamma=axes["gamma"], sigma_out=5.0) for ip in phis}
        sc4 = max(float(np.max(ref[ip][p])) for ip in phis for p in cases.POLS)
        for k, pol in enumerate(cases.POLS):
            want = np.stack([ref[ip][pol] for ip in phis])                              # [len(phis), E]
            got = cube[it, idd, k][torch.tensor(phis, device=cube.device)].cpu().numpy()
            if tag == "c3":
                M = np.maximum(want.max(axis=0, keepdims=True), 1e-30)
                assert np.max(np.abs(got - want) / M) < 1e-13, (tag, it, pol)
                assert np.max(np.abs(got - want)) <= RTOL * top
                ok, msg = cases.rows_parity_ok(got, want)
                assert ok, (tag, it, pol, msg)
            else:
                ok, msg = cases.parity_ok(got, want, rtol=RTOL, scale=sc4)      # sP / pS of Si(111) at phi = 30: pure node noise
                assert ok, (tag, it, idd, pol, msg)
    if tag == "c3":
        ss = res["ss"][:, 0]
        scale = float(ss.max())
        assert float((ss[:, :120] - ss[:, 120:240]).abs().max()) <= 1e-10 * scale
        assert float((ss[:, :180] - ss[:, 180:]).abs().max()) <= 1e-10 * scale
        # strict=True (SSHG_FLAG_STRICT_BROADENING: every row filtered itself) holds the element-wise metric row by
        # row, node rows included
        del cube, res
        res = grid.shgyield_grid(axes["energy"], m1, m2, m3, chi2, axes["theta"], axes["phi"], axes["thick"],
                                 gamma=axes["gamma"], sigma_out=5.0, strict=True)
        cube = res["cube"]
        for it, ip in ((0, 0), (0, 1), (65, 45), (65, 90), (89, 359)):
            ref = orc.shgyield(axes["energy"], m1, m2, m3, chi2, axes["theta"][it], axes["phi"][ip], axes["thick"][0],
                               gamma=axes["gamma"], sigma_out=5.0)
            sc = max(float(np.max(ref[p])) for p in cases.POLS)
            for k, pol in enumerate(cases.POLS):
                ok, msg = cases.parity_ok(cube[it, 0, k, ip].cpu().numpy(), ref[pol], rtol=RTOL, scale=sc)
                assert ok, ("strict", it, ip, pol, msg)
    del cube, res
    torch.cuda.empty_cache()


def test_k3_alone_has_the_absolute_bound_and_the_fixup_restores_the_strict_metric(grid, libopt):
    """Si(111) rows 0.02 deg from the C3v nodes: without the strict re-evaluation (option "k3_fixup" = 0) K3 holds its
    absolute bound; with it the near-node rows are re-evaluated row-wise and every row holds the strict metric."""
    z = np.load(os.path.join(cases.GOLDEN_DIR, "ref_blur_rows.npz"))
    m1, m2, m3, chi2 = cases.si111_material(cases.load_spectra())
    want = z["si111_sigma5"]                                            # [T, 4, P, E]
    top = float(np.max(want))
    run = lambda: grid.shgyield_grid(z["energy"], m1, m2, m3, chi2, z["theta"], z["phi"], cases.BLUR_SWEEP_THICK,
                                     gamma=0, sigma_out=5.0, device=False)["cube"][:, 0]
    libopt(k2_sparse=0)                                                  # 9 azimuths: force the dense path (K3)
    with util.lib_options(k3_fixup=0):
        alone = run()
    fixed = run()
    for k in range(4):
        colmax = np.maximum(np.max(want[:, k], axis=1, keepdims=True), 1e-12 * top)
        assert np.max(np.abs(alone[:, k] - want[:, k]) / colmax) < 5e-14            # K3 alone: the absolute bound
        assert np.max(np.abs(alone[:, k] - fixed[:, k]) / colmax) < 1e-13           # the fix-up moves values within it
        for it in range(want.shape[0]):
            ok, msg = cases.rows_parity_ok(fixed[it, k], want[it, k])               # ... and restores the strict metric
            assert ok, (k, it, msg)
    assert np.count_nonzero(alone != fixed) > 0                                      # near-node segments were re-evaluated
    assert float(np.min(fixed)) >= 0.0
