"""
CPU check of the kernel physics: shgyield_b200/csrc/sshg_math.cuh is __host__ __device__, so
tests/hostmath_harness.cu (a TEST harness, not part of libsshg.so) runs the very same
column_setup / coeffs_* code on the CPU.  This validates the re-organised formulas against the
golden vectors without a GPU; the GPU tests then check the kernels themselves.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import shg_oracle as orc
from shgyield_b200 import _lib
from shgyield_b200.synthetic import synthetic_spectra
from tests import cases, util

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                                reason="nvcc not available")


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(util.REPO, "tests", "hostmath_harness.cu")
    outdir = os.path.join(util.REPO, "tests", "_build")
    os.makedirs(outdir, exist_ok=True)
    so = os.path.join(outdir, "hostmath.so")
    deps = [src, os.path.join(util.REPO, "shgyield_b200/csrc/sshg_math.cuh"),
            os.path.join(util.REPO, "shgyield_b200/csrc/sshg_fir.cuh"), os.path.join(util.REPO, "include/sshg.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        subprocess.run([nvcc, "-O2", "-std=c++17", "-w", "-Xcompiler", "-fPIC", "-shared",
                        "-I", os.path.join(util.REPO, "include"), "-o", so, src], check=True)
    lib = C.CDLL(so)
    for fn in (lib.hostmath_yield_points, lib.hostmath_sparse_points):
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(_lib.Points), C.c_void_p]
    return lib


def run_points(lib, energy, eidx, theta, phi, thick, eps1w, eps2w, chi, consts, mref=True,
               sinc_normalized=True, zzz_phi_quirk=True, sparse=False):
    n = eidx.size
    arrs = dict(energy=np.ascontiguousarray(energy, dtype=np.float64),
                eidx=np.ascontiguousarray(eidx, dtype=np.int64),
                theta=np.ascontiguousarray(np.broadcast_to(theta, (n,)), dtype=np.float64),
                phi=np.ascontiguousarray(np.broadcast_to(phi, (n,)), dtype=np.float64),
                thick=np.ascontiguousarray(np.broadcast_to(thick, (n,)), dtype=np.float64),
                eps1w=np.ascontiguousarray(eps1w, dtype=np.complex128),
                eps2w=np.ascontiguousarray(eps2w, dtype=np.complex128),
                chi=np.ascontiguousarray(chi, dtype=np.complex128))
    p = _lib.Points(n=n, nE=arrs["energy"].size, energy_eV=arrs["energy"].ctypes.data, eidx=arrs["eidx"].ctypes.data,
                    theta_deg=arrs["theta"].ctypes.data, phi_deg=arrs["phi"].ctypes.data,
                    thick_nm=arrs["thick"].ctypes.data, eps1w=arrs["eps1w"].ctypes.data,
                    eps2w=arrs["eps2w"].ctypes.data, chi2=arrs["chi"].ctypes.data,
                    phys=_lib.Physics(consts["h_eVs"], consts["hbar_eVs"], consts["eps0"], consts["c"],
                                      int(mref), int(sinc_normalized), int(zzz_phi_quirk), 0))
    out = np.empty((4, n))
    fn = lib.hostmath_sparse_points if sparse else lib.hostmath_yield_points
    assert fn(C.byref(p), out.ctypes.data) == 0
    return out


@pytest.mark.parametrize("sparse", [False, True], ids=["harmonic", "sparse"])
def test_config5_samples(harness, sparse):
    consts = util.load_consts()
    z = np.load(os.path.join(cases.GOLDEN_DIR, "ref_grid_samples.npz"))
    idx = z["c5_idx"]
    energy, eps1w, eps2w, chi = synthetic_spectra(20000)
    theta, phi = np.linspace(0.0, 90.0, 181), np.linspace(0.0, 360.0, 721)
    out = run_points(harness, energy, idx[:, 3], theta[idx[:, 0]], phi[idx[:, 2]], 10.0, eps1w, eps2w, chi, consts,
                     sparse=sparse)
    sc = cases.case_scale(z, "c5_")
    for i, pol in enumerate(cases.POLS):
        ok, msg = cases.parity_ok(out[i], z["c5_" + pol], rtol=1e-10, scale=sc)
        assert ok, (pol, msg)


@pytest.mark.parametrize("name,mat,kw", [c for c in cases.API_CASES if c[2].get("sigma_out", 5.0) == 0.0],
                         ids=[c[0] for c in cases.API_CASES if c[2].get("sigma_out", 5.0) == 0.0])
def test_api_cases_kernel_math(harness, name, mat, kw):
    """Oracle prepares (broaden, resample, average, rotate); the harness evaluates the yield."""
    consts = util.load_consts()
    gold = np.load(os.path.join(cases.GOLDEN_DIR, "ref_api_cases.npz"))
    m1, m2, m3, chi2 = cases.MATERIALS[mat](cases.load_spectra())
    kw = dict(kw)
    energy = np.atleast_1d(np.asarray(kw["energy"], dtype=np.float64))
    eps1w, eps2w, chi_rot, thick = orc.prepare(energy, m1, m2, m3, chi2, kw["thick"], kw.get("gamma", 90),
                                               kw.get("sigma_eps", 0.0), kw.get("sigma_chi", 0.0),
                                               kw.get("mode", "3-layer"))
    nE = energy.size
    shape = np.broadcast_shapes(np.shape(kw["energy"]), np.shape(kw["theta"]), np.shape(kw["phi"]), np.shape(thick))
    eidx = np.broadcast_to(np.arange(nE).reshape(np.shape(kw["energy"]) or (1,)), shape or (1,)).ravel()
    bc = lambda v: np.broadcast_to(np.asarray(v, dtype=np.float64), shape or (1,)).ravel()
    out = run_points(harness, energy, eidx, bc(kw["theta"]), bc(kw["phi"]), bc(thick), util.pack_eps(eps1w, nE),
                     util.pack_eps(eps2w, nE), util.pack_chi(chi_rot, nE), consts, mref=kw.get("mref", True))
    sc = cases.case_scale(gold, name + "_")
    for i, pol in enumerate(cases.POLS):
        want = np.asarray(gold[name + "_" + pol])
        ok, msg = cases.parity_ok(out[i].reshape(want.shape), want, rtol=1e-10, scale=sc)
        assert ok, (name, pol, msg)


def test_switches(harness):
    """sinc_normalized=0 and zzz_phi_quirk=0 against the oracle with the same switches."""
    consts = util.load_consts()
    energy, eps1w, eps2w, chi = synthetic_spectra(64, seed=3)
    rng = np.random.default_rng(0)
    n = 500
    eidx = rng.integers(0, 64, n)
    theta, phi, thick = rng.uniform(0, 89, n), rng.uniform(0, 360, n), rng.uniform(0, 50, n)
    e1 = {"M%d" % (m + 1): eps1w[m][eidx] for m in range(3)}
    e2 = {"M%d" % (m + 1): eps2w[m][eidx] for m in range(3)}
    ch = {k: chi[i][eidx] for i, k in enumerate(cases.CHI2_KEYS)}
    for sn, zq, mr in ((False, True, True), (True, False, True), (True, True, False), (False, False, False)):
        want = orc.yield_points(energy[eidx], e1, e2, ch, theta, phi, thick, mr, consts, sn, zq)
        for sparse in (False, True):
            out = run_points(harness, energy, eidx, theta, phi, thick, eps1w, eps2w, chi, consts, mr, sn, zq, sparse)
            for i, pol in enumerate(cases.POLS):
                ok, msg = cases.parity_ok(out[i], want[pol], rtol=1e-10)
                assert ok, (sn, zq, mr, sparse, pol, msg)


@pytest.mark.parametrize("sigma", [0.2, 0.6, 1.0, 2.0, 2.2, 2.3, 3.3, 4.4, 5.0, 5.6, 6.1, 7.7, 12.0, 50.0])
def test_fir_window_arithmetic(harness, sigma):
    """The register-window FIR of K1 (sshg_fir.cuh), run on the CPU, against SciPy -- every value of
    (2 radius) mod 9 and the short-kernel path; a NaN sample must not spread beyond the radius."""
    from scipy import ndimage
    harness.hostmath_fir.restype = C.c_int
    harness.hostmath_fir.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p]
    w = orc.gaussian_taps(sigma)
    r = (len(w) - 1) // 2
    taps = np.ascontiguousarray(w[r:])
    for n in (1, 5, 64, 257):
        x = np.random.default_rng(n).normal(size=n)
        out = np.empty(n)
        assert harness.hostmath_fir(x.ctypes.data, n, taps.ctypes.data, r, out.ctypes.data) == 0
        np.testing.assert_allclose(out, ndimage.gaussian_filter(x, sigma), rtol=0, atol=4e-15 * max(1.0, np.abs(x).max()))
    x = np.random.default_rng(1).normal(size=300)
    x[150] = np.nan
    out = np.empty(300)
    harness.hostmath_fir(x.ctypes.data, 300, taps.ctypes.data, r, out.ctypes.data)
    bad = np.isnan(out)
    assert np.array_equal(bad, np.abs(np.arange(300) - 150) <= r)


@pytest.mark.parametrize("sn", [True, False])
def test_phase_recurrence_on_uniform_thickness_grids(harness, sn):
    """PhaseWalk (exponentials advanced by a constant factor, sin from their difference, direct sine for
    small arguments) against the oracle on thickness grids from 0 / 1e-6 nm upwards."""
    consts = util.load_consts()
    harness.hostmath_sparse_walk.restype = C.c_int
    harness.hostmath_sparse_walk.argtypes = [C.POINTER(_lib.Points), C.c_longlong, C.c_double, C.c_double, C.c_double,
                                             C.c_double, C.c_longlong, C.c_void_p]
    energy, eps1w, eps2w, chi = synthetic_spectra(12, seed=5)
    arrs = [np.ascontiguousarray(a) for a in (energy, eps1w, eps2w, chi)]
    p = _lib.Points(n=1, nE=12, energy_eV=arrs[0].ctypes.data, eidx=0, theta_deg=0, phi_deg=0, thick_nm=0,
                    eps1w=arrs[1].ctypes.data, eps2w=arrs[2].ctypes.data, chi2=arrs[3].ctypes.data,
                    phys=_lib.Physics(consts["h_eVs"], consts["hbar_eVs"], consts["eps0"], consts["c"], 1, int(sn), 1, 0))
    for e, theta, phi, d0, step, nd in ((0, 65.0, 30.0, 1.0, 1.0, 100), (5, 20.0, 200.0, 0.0, 0.37, 257),
                                        (11, 80.0, 10.0, 1e-6, 2.5e-3, 64), (3, 45.0, 123.0, 500.0, -4.0, 100)):
        out = np.empty((4, nd))
        assert harness.hostmath_sparse_walk(C.byref(p), e, theta, phi, d0, step, nd, out.ctypes.data) == 0
        thick = d0 + step * np.arange(nd)
        e1 = {"M%d" % (m + 1): eps1w[m][e] for m in range(3)}
        e2 = {"M%d" % (m + 1): eps2w[m][e] for m in range(3)}
        ch = {k: chi[i][e] for i, k in enumerate(cases.CHI2_KEYS)}
        want = orc.yield_points(energy[e], e1, e2, ch, theta, phi, thick, True, consts, sn, True)
        for i, pol in enumerate(cases.POLS):
            np.testing.assert_allclose(out[i], want[pol], rtol=2e-12, atol=0, err_msg="%s e=%d" % (pol, e))


def _blur_column(harness, energy, eps1w, eps2w, chi, theta, thick, phi, sigma, consts, strict_nodes=True):
    harness.hostmath_blur_column.restype = C.c_int
    harness.hostmath_blur_column.argtypes = [C.POINTER(_lib.Points), C.c_double, C.c_double, C.c_void_p, C.c_longlong,
                                             C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    arrs = [np.ascontiguousarray(a) for a in (energy, eps1w, eps2w, chi)]
    nE = energy.size
    p = _lib.Points(n=1, nE=nE, energy_eV=arrs[0].ctypes.data, eidx=0, theta_deg=0, phi_deg=0, thick_nm=0,
                    eps1w=arrs[1].ctypes.data, eps2w=arrs[2].ctypes.data, chi2=arrs[3].ctypes.data,
                    phys=_lib.Physics(consts["h_eVs"], consts["hbar_eVs"], consts["eps0"], consts["c"], 1, 1, 1, 0))
    w = orc.gaussian_taps(sigma)
    r = (len(w) - 1) // 2
    taps = np.ascontiguousarray(w[r:])
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    out = np.empty((4, phi.size, nE))
    assert harness.hostmath_blur_column(C.byref(p), float(theta), float(thick), phi.ctypes.data, phi.size,
                                        taps.ctypes.data, r, int(strict_nodes), out.ctypes.data) == 0
    return out


@pytest.mark.parametrize("sigma", [0.3, 2.0, 5.0])
@pytest.mark.parametrize("material", ["synthetic", "si111"])
def test_harmonic_form_commutes_with_output_broadening(harness, material, sigma):
    """K3's algorithm on the CPU: |r(phi)|^2 expanded in 13 azimuthal harmonics per energy, the harmonic
    spectra broadened along the energy axis, then evaluated per phi -- against the oracle's
    broad(yield, sigma_out) row by row (shg.py:433-436).  Si(111) has exact symmetry nodes in phi."""
    consts = util.load_consts()
    if material == "synthetic":
        energy, eps1w, eps2w, chi = synthetic_spectra(90, seed=11)
    else:
        m1, m2, m3, chi2 = cases.si111_material(cases.load_spectra())
        energy = np.linspace(0.5, 5.0, 90)
        e1, e2, ch, _ = orc.prepare(energy, m1, m2, m3, chi2, 10.0, 0, 0.0, 0.0, "3-layer")
        eps1w, eps2w, chi = util.pack_eps(e1, 90), util.pack_eps(e2, 90), util.pack_chi(ch, 90)
    phi = np.concatenate([np.arange(0.0, 360.0, 7.5), [29.98, 30.0, 30.05, 123.456]])
    for theta, thick in ((65.0, 10.0), (0.0, 3.0), (89.0, 55.0)):
        d1 = {"M%d" % (m + 1): eps1w[m] for m in range(3)}
        d2 = {"M%d" % (m + 1): eps2w[m] for m in range(3)}
        ch = {k: chi[i] for i, k in enumerate(cases.CHI2_KEYS)}
        raw = orc.yield_points(energy, d1, d2, ch, theta, phi[:, None], thick, True, consts)
        want = {pol: orc.gaussian_filter1d(raw[pol], sigma, axis=-1) for pol in cases.POLS}
        sc = max(np.max(np.abs(v)) for v in want.values())
        # K3 alone: the error of the harmonic form is absolute, a few 1e-15 of the column's maximum over phi
        got = _blur_column(harness, energy, eps1w, eps2w, chi, theta, thick, phi, sigma, consts, strict_nodes=False)
        for i, pol in enumerate(cases.POLS):
            ok, msg = cases.parity_ok(got[i], want[pol], rtol=1e-10, scale=sc, rel_floor=1e-4)
            assert ok, (material, sigma, theta, pol, msg)
            colmax = np.maximum(want[pol].max(axis=0, keepdims=True), 1e-12 * sc)
            assert np.max(np.abs(got[i] - want[pol]) / colmax) < 5e-14, (material, sigma, theta, pol)
        # the product (K3 + re-evaluation of the flagged near-node segments): the strict metric, also row by row
        got = _blur_column(harness, energy, eps1w, eps2w, chi, theta, thick, phi, sigma, consts, strict_nodes=True)
        for i, pol in enumerate(cases.POLS):
            ok, msg = cases.parity_ok(got[i], want[pol], rtol=1e-10, scale=sc)
            assert ok, (material, sigma, theta, pol, msg)
            ok, msg = cases.rows_parity_ok(got[i], want[pol])
            assert ok, (material, sigma, theta, pol, msg)


@pytest.mark.parametrize("mat", ["si111", "full"])
def test_harmonic_form_vs_reference_broadened_rows(harness, mat):
    """K3's algorithm (CPU run of the kernel's own functions) against shgyield(sigma_out) of the UNMODIFIED
    reference, called once per (theta, phi): tests/golden/ref_blur_rows.npz."""
    consts = util.load_consts()
    z = np.load(os.path.join(cases.GOLDEN_DIR, "ref_blur_rows.npz"))
    E, TH, PH = z["energy"], z["theta"], z["phi"]
    m1, m2, m3, chi2 = cases.MATERIALS[mat](cases.load_spectra())
    e1, e2, ch, thick = orc.prepare(E, m1, m2, m3, chi2, cases.BLUR_SWEEP_THICK, cases.BLUR_SWEEP_GAMMA[mat], 0.0, 0.0,
                                    "3-layer")
    eps1w, eps2w, chi = util.pack_eps(e1, E.size), util.pack_eps(e2, E.size), util.pack_chi(ch, E.size)
    for sigma in cases.BLUR_SWEEP_SIGMAS:
        want = z["%s_sigma%g" % (mat, sigma)]                           # [T, 4, P, E]
        top = float(np.max(want))
        for it, th in enumerate(TH):
            got = _blur_column(harness, E, eps1w, eps2w, chi, th, thick, PH, sigma, consts)     # [4, P, E]
            for k, pol in enumerate(cases.POLS):
                ok, msg = cases.parity_ok(got[k], want[it, k], rtol=1e-10, scale=top)
                assert ok, (mat, sigma, th, pol, msg)
                ok, msg = cases.rows_parity_ok(got[k], want[it, k])      # row by row: rows 0.02 deg from a node included
                assert ok, (mat, sigma, th, pol, msg)
