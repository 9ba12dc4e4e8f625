#!/usr/bin/env python
"""
Generates the golden fixtures in this directory by running the UNMODIFIED reference
(roguephysicist/SHGYield, mounted read-only at /root/reference) in the build container.
The GPU box has no /root/reference: tests only read the .npz files written here.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [/root/reference]

Fixtures
  si111_spectra.npz          raw columns of the reference's example/ and tests/data spectra
  si111_reference_out.npz    example/reference/si111-reference.out (the reference's golden)
  selftest_reference.npz     REFERENCE array of tests/self_test.py:115-149
  ref_api_cases.npz          shg.shgyield() outputs for tests/cases.py::API_CASES
  ref_helpers.npz            rotate / wvec / fresnel / mrc outputs on random inputs
  ref_grid_samples.npz       sampled points of configs 3, 4 (API level) and 5 (kernel level)
  constants.npz              the four scipy.constants values the reference read here
  ref_blur_rows.npz          shg.shgyield(sigma_out > 0) called angle by angle on a sweep (`--blur` writes only this)
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
ONLY_BLUR = "--blur" in sys.argv
_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
REF = _pos[0] if _pos else "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(1, REPO)
sys.dont_write_bytecode = True
warnings.simplefilter("ignore")

import shgyield.shg as ref                                   # noqa: E402  (the reference)
assert os.path.abspath(ref.__file__).startswith(os.path.abspath(REF)), ref.__file__
from scipy import constants                                  # noqa: E402
from tests import cases                                      # noqa: E402
from shgyield_b200.synthetic import synthetic_spectra       # noqa: E402


def cols(path, skiprows=0):
    return np.loadtxt(path, unpack=True, skiprows=skiprows)


def blur_rows():
    """The reference's own output broadening on a sweep: shgyield() broadens every array it returns along ALL
    its axes, so a sweep broadened along the energy axis alone is shgyield() called once per (theta, phi)."""
    z = cases.load_spectra()
    out = {}
    for mat, gamma in (("si111", 0), ("full", 33)):
        m1, m2, m3, chi2 = cases.MATERIALS[mat](z)
        for k, (key, val) in enumerate(cases.BLUR_SWEEP.items()):
            out[key] = np.asarray(val, dtype=np.float64)
        E, TH, PH = cases.BLUR_SWEEP["energy"], cases.BLUR_SWEEP["theta"], cases.BLUR_SWEEP["phi"]
        for sigma in cases.BLUR_SWEEP_SIGMAS:
            cube = np.empty((len(TH), 4, len(PH), len(E)))
            for it, th in enumerate(TH):
                for ip, ph in enumerate(PH):
                    res = ref.shgyield(energy=E, eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, theta=th, phi=ph,
                                       thick=cases.BLUR_SWEEP_THICK, gamma=gamma, sigma_out=sigma)
                    for k, pol in enumerate(cases.POLS):
                        cube[it, k, ip] = res[pol]
            out["%s_sigma%g" % (mat, sigma)] = cube
    np.savez_compressed(os.path.join(HERE, "ref_blur_rows.npz"), **out)
    print("ref_blur_rows.npz", os.path.getsize(os.path.join(HERE, "ref_blur_rows.npz")), "bytes")


def main():
    if ONLY_BLUR:
        return blur_rows()
    # ---- spectra ---------------------------------------------------------------------------
    spectra = {}
    for name in ("SiBulk-chi1-xx", "SiBulk-chi1-yy", "SiBulk-chi1-zz",
                 "SiH1x1-chi1-xx", "SiH1x1-chi1-yy", "SiH1x1-chi1-zz"):
        d = cols(os.path.join(REF, "example/chi1-linear", name))
        spectra["energy"] = d[0]
        spectra[name] = d[1:3]
    for name in ("SiH1x1-chi2-xxx", "SiH1x1-chi2-xxz", "SiH1x1-chi2-zxx", "SiH1x1-chi2-zzz"):
        d = cols(os.path.join(REF, "example/chi2-nonlinear", name))
        assert np.array_equal(d[0], spectra["energy"])
        spectra[name] = d[1:5]
    for name in ("SiH1x1-epsilon", "SiBulk-epsilon", "SiH1x1-chi2-xxx"):
        d = cols(os.path.join(REF, "tests/data", name), skiprows=1)
        assert np.array_equal(d[0], spectra["energy"])
        spectra["test-" + name] = d[1:3]
    np.savez_compressed(os.path.join(HERE, "si111_spectra.npz"), **spectra)

    out = cols(os.path.join(REF, "example/reference/si111-reference.out"))
    np.savez_compressed(os.path.join(HERE, "si111_reference_out.npz"),
                        energy=out[0], pp=out[1], sp=out[2], ps=out[3], ss=out[4])

    # REFERENCE array: execute the reference's self-test module from its own root
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        glb = {"__name__": "self_test_golden"}
        exec(compile(open("tests/self_test.py").read(), "self_test.py", "exec"), glb)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "selftest_reference.npz"), REFERENCE=np.asarray(glb["REFERENCE"]))

    consts = dict(h_eVs=constants.value("Planck constant in eV s"),
                  hbar_eVs=constants.value("Planck constant over 2 pi in eV s"),
                  eps0=constants.epsilon_0, c=constants.c)
    np.savez(os.path.join(HERE, "constants.npz"), **consts)

    # ---- API cases -------------------------------------------------------------------------
    z = np.load(os.path.join(HERE, "si111_spectra.npz"))
    api = {}
    for name, mat, kw in cases.API_CASES:
        m1, m2, m3, chi2 = cases.MATERIALS[mat](z)
        res = ref.shgyield(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, **kw)
        for pol in cases.POLS:
            api[name + "_" + pol] = np.asarray(res[pol])
    np.savez_compressed(os.path.join(HERE, "ref_api_cases.npz"), **api)

    # ---- helper functions on random inputs --------------------------------------------------
    rng = np.random.default_rng(11)
    hp = {}
    n = 64
    chi = {k: rng.normal(size=n) + 1j * rng.normal(size=n) for k in cases.CHI2_KEYS}
    hp["rot_in"] = np.stack([chi[k] for k in cases.CHI2_KEYS])
    for g in (0.0, 17.0, 90.0, 123.4, 270.0):
        r = ref.rotate(chi, np.radians(g))
        hp["rot_out_%g" % g] = np.stack([r[k] for k in cases.CHI2_KEYS])
    ei = rng.normal(size=n) * 5 + 1j * np.abs(rng.normal(size=n)) * 3
    ej = rng.normal(size=n) * 5 + 1j * np.abs(rng.normal(size=n)) * 3
    ei[:4] = [1.0, 11.7 + 0j, -3.0 + 0j, complex(-3.0, -0.0)]           # real eps / branch cut / signed zero
    th = rng.uniform(0, np.pi / 2, n)
    en = rng.uniform(0.1, 9.0, n)
    hp.update(epsi=ei, epsj=ej, theta=th, energy=en)
    hp["wvec"] = ref.wvec(ei, th)
    hp["frefs"] = ref.frefs(ei, ej, th)
    hp["frefp"] = ref.frefp(ei, ej, th)
    hp["ftrans"] = ref.ftrans(ei, ej, th)
    hp["ftranp"] = ref.ftranp(ei, ej, th)
    f1, f2 = ref.frefp(ej, ei, th), ref.frefs(ei, ej, th)
    hp.update(f1=f1, f2=f2)
    for d in (0.0, 3.0, 25.0):
        hp["mrc2w_%g" % d] = ref.mrc2w(en, ej, f1, f2, th, d, True)
        hp["mrc1w_%g" % d] = ref.mrc1w(en, ej, f1, f2, th, d, True)
    np.savez_compressed(os.path.join(HERE, "ref_helpers.npz"), **hp)

    # ---- sampled points of the big grids ------------------------------------------------------
    gs = {}
    m1, m2, m3, chi2 = cases.si111_material(z)
    for tag, axes, seed in (("c3", cases.config3_axes(), 303), ("c4", cases.config4_axes(), 404)):
        shape = (axes["theta"].size, axes["thick"].size, axes["phi"].size, axes["energy"].size)
        idx = cases.sample_indices(shape, 20000, seed)
        res = ref.shgyield(energy=axes["energy"][idx[:, 3]], eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2,
                           theta=axes["theta"][idx[:, 0]], phi=axes["phi"][idx[:, 2]],
                           thick=axes["thick"][idx[:, 1]], gamma=axes["gamma"], sigma_out=0.0)
        gs[tag + "_idx"] = idx.astype(np.int32)
        for pol in cases.POLS:
            gs[tag + "_" + pol] = res[pol]

    # config 5 at kernel level: rad_* + prefactor (shg.py:428-436) straight on the synthetic arrays
    nE, nT, nP = 20000, 181, 721
    energy, eps1w, eps2w, chi = synthetic_spectra(nE)
    theta = np.linspace(0.0, 90.0, nT)
    phi = np.linspace(0.0, 360.0, nP)
    idx = cases.sample_indices((nT, 1, nP, nE), 30000, 505)
    ie, it, ip = idx[:, 3], idx[:, 0], idx[:, 2]
    e1 = {"M1": eps1w[0][ie], "M2": eps1w[1][ie], "M3": eps1w[2][ie]}
    e2 = {"M1": eps2w[0][ie], "M2": eps2w[1][ie], "M3": eps2w[2][ie]}
    ch = {k: chi[i][ie] for i, k in enumerate(cases.CHI2_KEYS)}
    E = energy[ie]
    pref = 1e4 * ((E / consts["hbar_eVs"]) ** 2) / (2 * consts["eps0"] * consts["c"] ** 3 * np.cos(np.radians(theta[it])) ** 2)
    gs["c5_idx"] = idx.astype(np.int32)
    for pol, fn in (("pp", ref.rad_pp), ("ps", ref.rad_ps), ("sp", ref.rad_sp), ("ss", ref.rad_ss)):
        zz = fn(E, e1, e2, ch, np.radians(theta[it]), np.radians(phi[ip]), 10.0, True)
        gs["c5_" + pol] = pref * np.absolute(1 / np.sqrt(e1["M2"]) * zz) ** 2
    np.savez_compressed(os.path.join(HERE, "ref_grid_samples.npz"), **gs)
    blur_rows()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-28s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
