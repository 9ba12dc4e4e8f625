"""GPU parity of the reference's small public helpers (wvec .. mrc1w, rad_pp .. rad_ss) against
outputs of the unmodified reference (tests/golden/ref_helpers.npz) and the oracle."""
import os

import numpy as np
import pytest

from oracle import shg_oracle as orc
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shg():
    import shgyield.shg as mod          # the drop-in import path of the reference
    return mod


def test_fresnel_helpers_match_reference(shg):
    z = np.load(os.path.join(cases.GOLDEN_DIR, "ref_helpers.npz"))
    ei, ej, th = z["epsi"], z["epsj"], z["theta"]
    for name, args in (("wvec", (ei, th)), ("frefs", (ei, ej, th)), ("frefp", (ei, ej, th)),
                       ("ftrans", (ei, ej, th)), ("ftranp", (ei, ej, th))):
        got = getattr(shg, name)(*args)
        assert got.shape == z[name].shape and got.dtype == np.complex128
        np.testing.assert_allclose(got, z[name], rtol=1e-13, atol=0, err_msg=name)
    # branch cut / signed zero of the complex square root (eps = -3 +0j and -3 -0j)
    w = shg.wvec(np.array([complex(-3.0, 0.0), complex(-3.0, -0.0)]), 0.3)
    assert w[0].imag > 0 and w[1].imag < 0
    # scalar in -> scalar out, broadcasting like NumPy
    assert np.ndim(shg.wvec(11.7 + 0.1j, 0.5)) == 0
    assert shg.frefs(ei[:, None], ej[None, :4], 0.2).shape == (ei.size, 4)


def test_multiref_helpers_match_reference(shg):
    z = np.load(os.path.join(cases.GOLDEN_DIR, "ref_helpers.npz"))
    en, ej, th = z["energy"], z["epsj"], z["theta"]
    for d in (0.0, 3.0, 25.0):
        np.testing.assert_allclose(shg.mrc2w(en, ej, z["f1"], z["f2"], th, d, True), z["mrc2w_%g" % d], rtol=1e-12)
        np.testing.assert_allclose(shg.mrc1w(en, ej, z["f1"], z["f2"], th, d, True), z["mrc1w_%g" % d], rtol=1e-12)
    np.testing.assert_array_equal(shg.mrc2w(en, ej, z["f1"], z["f2"], th, 5.0, False), z["f1"])
    np.testing.assert_array_equal(shg.mrc1w(en, ej, z["f1"], z["f2"], th, 5.0, False), z["f1"])


def test_rad_amplitudes_match_oracle(shg):
    from shgyield_b200.synthetic import synthetic_spectra
    energy, eps1w, eps2w, chi = synthetic_spectra(50, seed=8)
    e1 = {"M%d" % (m + 1): eps1w[m] for m in range(3)}
    e2 = {"M%d" % (m + 1): eps2w[m] for m in range(3)}
    ch = {k: chi[i] for i, k in enumerate(cases.CHI2_KEYS)}
    theta = np.radians(np.array([[10.0], [65.0], [89.0]]))
    phi = np.radians(33.0)
    for mref in (True, False):
        for name in ("rad_pp", "rad_ps", "rad_sp", "rad_ss"):
            got = getattr(shg, name)(energy, e1, e2, ch, theta, phi, 10.0, mref)
            want = getattr(orc, name)(energy, e1, e2, ch, theta, phi, 10.0, mref, orc.SCIPY_1_18_CONSTANTS)
            assert got.shape == (3, 50)
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.max(np.abs(want)), err_msg=name)
    # spectra along the FIRST axis (compacted to [nE] + an index per point), and spectra spread over two axes (one entry
    # per point): the same numbers as the oracle's native broadcasting
    col = lambda d: {k: np.asarray(v)[:, None] for k, v in d.items()}
    th_row = np.radians(np.array([10.0, 65.0, 89.0]))
    got = shg.rad_pp(energy[:, None], col(e1), col(e2), col(ch), th_row, phi, 10.0, True)
    want = orc.rad_pp(energy[:, None], col(e1), col(e2), col(ch), th_row, phi, 10.0, True, orc.SCIPY_1_18_CONSTANTS)
    assert got.shape == (50, 3)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.max(np.abs(want)))
    ch2 = dict(ch, xxx=ch["xxx"] * np.array([[1.0], [2.0], [3.0]]))          # chi_xxx varies with the angle axis too
    got = shg.rad_sp(energy, e1, e2, ch2, theta, phi, 10.0, True)
    want = orc.rad_sp(energy, e1, e2, ch2, theta, phi, 10.0, True, orc.SCIPY_1_18_CONSTANTS)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.max(np.abs(want)))
    got = shg.rad_ss(1.7, {k: v[7] for k, v in e1.items()}, {k: v[7] for k, v in e2.items()}, {k: v[7] for k, v in ch.items()},
                     theta, np.radians(np.arange(0, 360, 45.0)), 10.0, True)   # scalar spectra, angles broadcast
    want = orc.rad_ss(1.7, {k: v[7] for k, v in e1.items()}, {k: v[7] for k, v in e2.items()},
                      {k: v[7] for k, v in ch.items()}, theta, np.radians(np.arange(0, 360, 45.0)), 10.0, True,
                      orc.SCIPY_1_18_CONSTANTS)
    assert got.shape == (3, 8)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.max(np.abs(want)))


def test_context_options_and_stream_handling():
    """sshg_ctx_set_option: unknown keys / values fail, -1 restores the library's choice; a device call on a
    temporary torch stream leaves the context on its own stream (the handle of a dropped stream is never kept)."""
    import torch
    from shgyield_b200 import _lib, grid
    from shgyield_b200.synthetic import config5
    ctx = _lib.context(0)
    with pytest.raises(ValueError):
        ctx.set_option("no_such_option", 1)
    with pytest.raises(ValueError):
        ctx.set_option("k3_nt", 96)
    with pytest.raises(ValueError):
        ctx.set_option("nseg", 17)
    assert ctx.get_option("k3") is None
    with ctx.options(k3=0, nseg=3):
        assert ctx.get_option("k3") == 0 and ctx.get_option("nseg") == 3
    assert ctx.get_option("k3") is None and ctx.get_option("nseg") is None

    own = ctx.lib.sshg_ctx_stream(ctx.handle)
    cfg = config5(n_energy=400, n_theta=3, n_phi=21)
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in cfg.items()}
    args = (d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"])
    ref = grid.yield_grid_device(*args)
    torch.cuda.synchronize()
    for _ in range(3):
        s = torch.cuda.Stream(dev)                      # a temporary stream, dropped after the call
        with torch.cuda.stream(s):
            got = grid.yield_grid_device(*args, sigma_out=0.0)
        s.synchronize()
        assert torch.equal(got, ref)
        del s
        assert ctx.lib.sshg_ctx_stream(ctx.handle) == own
        host = grid.yield_grid_host(**cfg)             # the host-array path keeps working on the own stream
        assert np.array_equal(host, ref.cpu().numpy())


def test_points_broadened_device_resident_and_bad_index():
    """sshg_yield_points_broadened with everything resident (where = SSHG_DEVICE): the same bits as the host call with
    resident spectra and as the plain host call, on a 2-D broadcast with the filter along both axes; an out-of-range
    column index in a device-resident list gives NaN for that point instead of an out-of-bounds read."""
    import ctypes as C
    import torch
    from shgyield_b200 import _lib
    from shgyield_b200.shg import _physics, physical_constants
    from shgyield_b200.synthetic import synthetic_spectra
    ctx = _lib.context(0)
    energy, eps1w, eps2w, chi = synthetic_spectra(40, seed=5)
    shape = (7, 40)
    n = 7 * 40
    theta = np.ascontiguousarray(np.broadcast_to(np.linspace(10, 70, 7)[:, None], shape)).ravel()
    phi = np.full(n, 33.0)
    thick = np.full(n, 12.0)
    eidx = np.ascontiguousarray(np.broadcast_to(np.arange(40, dtype=np.int64), shape)).ravel()
    phys = _physics(physical_constants(), True)
    shp = (C.c_int64 * 2)(*shape)

    def call(arrs, where, out_ptr):
        p = _lib.Points(n=n, nE=40, energy_eV=arrs[0], eidx=arrs[1], theta_deg=arrs[2], phi_deg=arrs[3], thick_nm=arrs[4],
                        eps1w=arrs[5], eps2w=arrs[6], chi2=arrs[7], phys=phys)
        _lib.check(ctx.lib.sshg_yield_points_broadened(ctx.handle, C.byref(p), 1.5, 2, shp, out_ptr, where), ctx.lib)

    host = [energy, eidx, theta, phi, thick, eps1w, eps2w, chi]
    out_h = np.empty((4, n))
    call([a.ctypes.data for a in host], _lib.HOST, out_h.ctypes.data)
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in host]
    out_s = np.empty((4, n))
    ptrs = [a.ctypes.data for a in host]
    for k in (0, 5, 6, 7):
        ptrs[k] = d[k].data_ptr()
    call(ptrs, _lib.SPECTRA_DEVICE, out_s.ctypes.data)
    out_d = torch.empty((4, n), dtype=torch.float64, device=dev)
    call([t.data_ptr() for t in d], _lib.DEVICE, out_d.data_ptr())
    ctx.sync()
    assert np.array_equal(out_h, out_s) and np.array_equal(out_h, out_d.cpu().numpy())
    # reference: the oracle's yields, scipy's filter over both axes
    e1 = {"M%d" % (m + 1): eps1w[m] for m in range(3)}
    e2 = {"M%d" % (m + 1): eps2w[m] for m in range(3)}
    ch = {k: chi[i] for i, k in enumerate(cases.CHI2_KEYS)}
    from tests import util
    want = orc.yield_points(energy, e1, e2, ch, np.linspace(10, 70, 7)[:, None], 33.0, 12.0, True, util.load_consts())
    for k, pol in enumerate(cases.POLS):
        w = orc.gaussian_filter1d(orc.gaussian_filter1d(np.broadcast_to(want[pol], shape), 1.5, axis=0), 1.5, axis=1)
        assert np.max(np.abs(out_h[k].reshape(shape) - w)) <= 1e-10 * np.max(np.abs(w)), pol
    # a bad index in a device-resident list: NaN for that point only (sigma_out = 0: no filter spreads it)
    bad = d[1].clone()
    bad[5] = 40
    p = _lib.Points(n=n, nE=40, energy_eV=d[0].data_ptr(), eidx=bad.data_ptr(), theta_deg=d[2].data_ptr(),
                    phi_deg=d[3].data_ptr(), thick_nm=d[4].data_ptr(), eps1w=d[5].data_ptr(), eps2w=d[6].data_ptr(),
                    chi2=d[7].data_ptr(), phys=phys)
    _lib.check(ctx.lib.sshg_yield_points_broadened(ctx.handle, C.byref(p), 0.0, 2, shp, out_d.data_ptr(), _lib.DEVICE), ctx.lib)
    ctx.sync()
    got = out_d.cpu().numpy()
    assert np.isnan(got[:, 5]).all() and np.count_nonzero(np.isnan(got)) == 4
    with pytest.raises(ValueError):                 # host lists are validated before anything is launched
        eb = eidx.copy()
        eb[3] = -1
        hb = [a.ctypes.data for a in host]
        hb[1] = eb.ctypes.data
        call(hb, _lib.HOST, out_h.ctypes.data)


def test_points_broadened_large_broadcast_takes_the_unstaged_path():
    """More than 2^20 points: no page-locked staging; the same bits as sshg_yield_points (sigma_out = 0) and a filter
    along the second axis that matches scipy's."""
    import ctypes as C
    from shgyield_b200 import _lib
    from shgyield_b200 import shg as host
    from shgyield_b200.synthetic import synthetic_spectra
    ctx = _lib.context(0)
    energy, eps1w, eps2w, chi = synthetic_spectra(24, seed=9)
    shape = (44000, 24)
    n = shape[0] * shape[1]
    assert n > 1 << 20
    rng = np.random.default_rng(3)
    theta = np.ascontiguousarray(np.broadcast_to(rng.uniform(0, 89, shape[0])[:, None], shape)).ravel()
    phi = np.ascontiguousarray(np.broadcast_to(rng.uniform(0, 360, shape[0])[:, None], shape)).ravel()
    thick = np.full(n, 8.0)
    eidx = np.ascontiguousarray(np.broadcast_to(np.arange(24, dtype=np.int64), shape)).ravel()
    want = host.yield_points(energy, eidx, theta, phi, thick, eps1w, eps2w, chi)
    phys = host._physics(host.physical_constants(), True)
    p = _lib.Points(n=n, nE=24, energy_eV=energy.ctypes.data, eidx=eidx.ctypes.data, theta_deg=theta.ctypes.data,
                    phi_deg=phi.ctypes.data, thick_nm=thick.ctypes.data, eps1w=eps1w.ctypes.data, eps2w=eps2w.ctypes.data,
                    chi2=chi.ctypes.data, phys=phys)
    shp = (C.c_int64 * 2)(*shape)
    got = np.empty((4, n))
    _lib.check(ctx.lib.sshg_yield_points_broadened(ctx.handle, C.byref(p), 0.0, 2, shp, got.ctypes.data, _lib.HOST), ctx.lib)
    assert np.array_equal(got, want)
    shp1 = (C.c_int64 * 2)(n // 24, 24)
    _lib.check(ctx.lib.sshg_yield_points_broadened(ctx.handle, C.byref(p), 1.0, 2, shp1, got.ctypes.data, _lib.HOST), ctx.lib)
    rows = want[2].reshape(shape)[:50]
    ref = orc.gaussian_filter1d(want[2].reshape(shape), 1.0, axis=1)[:50]
    ref = orc.gaussian_filter1d(np.pad(ref, ((0, 0), (0, 0))), 0.0, axis=0) if False else ref
    # the filter ran along BOTH axes; check the second axis on rows far from each other's influence is not possible, so
    # compare with the full two-axis filter of the first 200 rows (radius 4 along axis 0)
    full = orc.gaussian_filter1d(orc.gaussian_filter1d(want[2].reshape(shape)[:300], 1.0, axis=0), 1.0, axis=1)[:200]
    assert np.max(np.abs(got[2].reshape(shape)[:200] - full)) <= 1e-10 * np.max(np.abs(full))


def test_to_host_and_download_match_torch():
    """grid.to_host / sshg_download: threaded staged copies (above 32 MB) and the direct copy (below, and into
    page-locked memory) deliver the same bytes as torch's own copy, at sizes around the 64 MB staging chunks."""
    import torch
    from shgyield_b200 import _lib, grid
    ctx = _lib.context(0)
    for n in (1, 1000, (32 << 20) // 8 - 1, (32 << 20) // 8 + 3, (64 << 20) // 8, (64 << 20) // 8 + 1, 3 * (64 << 20) // 8 + 12345):
        t = torch.rand(n, dtype=torch.float64, device="cuda")
        assert np.array_equal(grid.to_host(t), t.cpu().numpy()), n
    t = torch.rand((5, 3, 1000, 999), dtype=torch.float64, device="cuda")
    assert np.array_equal(grid.to_host(t), t.cpu().numpy())
    pin = _lib.pinned_empty((t.numel(),))
    _lib.check(ctx.lib.sshg_download(ctx.handle, pin.ctypes.data, t.data_ptr(), pin.nbytes), ctx.lib)
    assert np.array_equal(pin, t.cpu().numpy().ravel())
    _lib.pinned_free(pin)
    assert grid.to_host(torch.empty((0, 4), dtype=torch.float64, device="cuda")).shape == (0, 4)


def test_yield_broadcast_rejects_bad_layouts():
    """sshg_yield_broadcast validates shape and layout before anything is launched: SSHG_EINVAL with a message."""
    import ctypes as C
    from shgyield_b200 import _lib, grid, shg
    from shgyield_b200.synthetic import synthetic_spectra
    ctx = _lib.context(0)
    energy, eps1w, eps2w, chi = synthetic_spectra(40, seed=1)
    theta, phi, thick = np.array([10.0, 20.0, 30.0]), np.arange(0, 360, 45.0), np.array([5.0])
    keep = []
    p = grid._problem(energy, theta, phi, thick, eps1w, eps2w, chi, shg._physics(None), None, None, keep)
    total = 8 * 3 * 40
    outs = [np.empty(total) for _ in range(4)]
    ptrs = (C.c_void_p * 4)(*[o.ctypes.data for o in outs])

    def call(strides, shape, sigma=0.0):
        st = (C.c_int64 * 4)(*strides)
        shp = (C.c_int64 * len(shape))(*shape)
        return ctx.lib.sshg_yield_broadcast(ctx.handle, C.byref(p), st, float(sigma), len(shape), shp, ptrs)
    good = (40, 0, total, 3 * 40)                                        # [phi][theta][E]
    assert call(good, (8, 3, 40)) == _lib.SSHG_OK
    ref = [o.copy() for o in outs]
    cube = grid.yield_grid_host(energy, theta, phi, thick, eps1w, eps2w, chi)        # [T, 1, 4, P, E]
    for k in range(4):
        assert np.array_equal(ref[k].reshape(8, 3, 40), np.transpose(cube[:, 0, k], (1, 0, 2)))
    assert call(good, (8, 3, 41)) == _lib.SSHG_EINVAL                     # shape does not match the grid
    assert call((40, 0, total - 1, 120), (8, 3, 40)) == _lib.SSHG_EINVAL   # polarisation stride
    assert call((40, 0, total, 121), (8, 3, 40)) == _lib.SSHG_EINVAL       # phi rows would run past the array
    assert call((-40, 0, total, 120), (8, 3, 40)) == _lib.SSHG_EINVAL
    assert call(good, (8, 3, 40), sigma=-1.0) == _lib.SSHG_EINVAL
    assert b"sigma_out" in ctx.lib.sshg_last_error()
