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
