"""Host logic of the file loaders and the (reconstructed) YAML command line -- no GPU needed."""
import os
import sys

import numpy as np
import pytest

from shgyield_b200 import io
from tests import cases, util

sys.path.insert(0, os.path.join(util.REPO, "example"))


@pytest.fixture(scope="module")
def example_dir(tmp_path_factory):
    import make_example_data
    d = tmp_path_factory.mktemp("example")
    make_example_data.write(str(d))
    import shutil
    shutil.copy(os.path.join(util.REPO, "example", "si111-input.yml"), d)
    return str(d)


def test_loaders_reproduce_main_py_arithmetic(example_dir):
    m1, m2, m3, chi2 = cases.si111_material()
    eps = io.loadeps(os.path.join(example_dir, "chi1-linear", "SiH1x1-chi1-xx"), cases.LSUPER / cases.LSLAB)
    np.testing.assert_allclose(eps["data"], m2["xx"]["data"], rtol=1e-6)       # files keep 7 digits
    np.testing.assert_array_equal(eps["energy"], m2["xx"]["energy"])
    shg = io.loadshg(os.path.join(example_dir, "chi2-nonlinear", "SiH1x1-chi2-zzz"), 1e6 * 1e-24)
    np.testing.assert_allclose(shg["data"], chi2["zzz"]["data"], rtol=1e-5, atol=1e-30)


def test_yaml_input_builds_the_reference_material(example_dir):
    cfg = io.load_input(os.path.join(example_dir, "si111-input.yml"))
    np.testing.assert_allclose(cfg["energy"], cases.E_C1, rtol=0, atol=1e-12)
    assert (cfg["theta"], cfg["phi"], cfg["thick"], cfg["gamma"]) == (65, 30, 10, 0)
    assert (cfg["sigma_eps"], cfg["sigma_chi"], cfg["sigma_out"], cfg["mode"], cfg["mref"]) == (0.0, 0.0, 5.0, "3-layer", True)
    assert cfg["eps_m1"] == {"xx": 1.0, "yy": 1.0, "zz": 1.0}
    assert sorted(cfg["eps_m2"]) == ["xx", "yy", "zz"] and sorted(cfg["eps_m3"]) == ["xx", "yy", "zz"]
    chi2 = cfg["chi2"]
    assert sorted(chi2) == sorted(io.CHI2_KEYS)
    np.testing.assert_array_equal(chi2["xyy"]["data"], -chi2["xxx"]["data"])
    np.testing.assert_array_equal(chi2["yxy"]["data"], -chi2["xxx"]["data"])
    assert chi2["yyz"] is chi2["xxz"] and chi2["zyy"] is chi2["zxx"]
    assert all(chi2[k] == 0 for k in ("xzz", "xyz", "xxy", "yxx", "yyy", "yzz", "yxz", "zyz", "zxz", "zxy"))


def test_chi2_alias_resolution_errors(tmp_path):
    with pytest.raises(ValueError):
        io.resolve_chi2({"xyy": "-yxy", "yxy": "-xyy"}, str(tmp_path))
    with pytest.raises(ValueError):
        io.resolve_chi2({"abc": 0}, str(tmp_path))
    r = io.resolve_chi2({"zzz": 2.5e-19, "xxx": 0}, str(tmp_path))
    assert r["zzz"] == 2.5e-19 and r["xxx"] == 0 and r["zxy"] == 0


def test_merged_chi1_file(tmp_path):
    e = np.linspace(0, 1, 6)
    cols = np.column_stack([e] + [np.full(6, v) for v in (1, 2, 3, 4, 5, 6)])
    p = tmp_path / "Si-chi1-xx_yy_zz"
    np.savetxt(p, cols)
    m = io._load_medium(str(p), 2.0, "")
    np.testing.assert_allclose(m["yy"]["data"], 1 + 4 * np.pi * (3 + 4j) * 2.0)


def test_save_spectrum_format(tmp_path):
    res = {"energy": np.array([1.25, 1.26]), "pp": np.array([9.16988599e-22, 1e-22]), "sp": np.array([2e-25, 3e-25]),
           "ps": np.array([7e-23, 8e-23]), "ss": np.array([2e-23, 3e-23])}
    p = tmp_path / "out.dat"
    io.save_spectrum(p, res)
    lines = open(p).read().splitlines()
    assert lines[0].startswith("# w(eV)  RpP") and "RsP" in lines[0]
    assert lines[1].split() == ["01.2500", "9.16988599e-22", "2.00000000e-25", "7.00000000e-23", "2.00000000e-23"]


@pytest.mark.gpu
def test_cli_reproduces_reference_golden(example_dir):
    """python shgyield.py si111-input.yml -> the reference's example/reference/si111-reference.out."""
    import subprocess
    out = os.path.join(example_dir, "yield.dat")
    subprocess.run([sys.executable, os.path.join(util.REPO, "shgyield.py"), os.path.join(example_dir, "si111-input.yml"),
                    "-o", out], check=True)
    got = np.loadtxt(out, unpack=True)
    gold = np.load(os.path.join(cases.GOLDEN_DIR, "si111_reference_out.npz"))
    np.testing.assert_allclose(got[0], gold["energy"], atol=5e-5)
    for col, pol in zip(got[1:], ("pp", "sp", "ps", "ss")):
        np.testing.assert_allclose(col, gold[pol], rtol=1e-5)        # text inputs carry 7 digits


@pytest.mark.gpu
def test_cli_sweep_writes_cube(example_dir):
    """theta / phi given as lists -> a sweep on the GPU, written as .npz; one column must equal the
    scalar run of the same input (sigma_out broadens along the energy axis only)."""
    import subprocess
    import yaml
    src = os.path.join(example_dir, "si111-input.yml")
    doc = yaml.safe_load(open(src))
    doc["parameters"]["theta"] = [55, 65]
    doc["parameters"]["phi"] = [0, 30, 90]
    swp = os.path.join(example_dir, "sweep.yml")
    yaml.safe_dump(doc, open(swp, "w"))
    out = os.path.join(example_dir, "sweep.npz")
    subprocess.run([sys.executable, os.path.join(util.REPO, "shgyield.py"), swp, "-o", out], check=True)
    z = np.load(out)
    assert z["pp"].shape == (2, 1, 3, 126)
    one = os.path.join(example_dir, "one.dat")
    subprocess.run([sys.executable, os.path.join(util.REPO, "shgyield.py"), src, "-o", one], check=True)
    col = np.loadtxt(one, unpack=True)
    for k, pol in zip((1, 2, 3, 4), ("pp", "sp", "ps", "ss")):
        np.testing.assert_allclose(z[pol][1, 0, 1], col[k], rtol=2e-8)       # text output keeps 9 digits
