"""Runs the ctypes stub printed in INTEGRATION.md (the binding a maintainer of the reference would
add) verbatim -- only the library path is substituted -- and checks it against the reference's output."""
import os
import re

import numpy as np
import pytest

from oracle import shg_oracle as orc
from tests import cases, util

pytestmark = pytest.mark.gpu


def test_integration_stub_reproduces_reference():
    from shgyield_b200 import _lib
    text = open(os.path.join(util.REPO, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(# shgyield/_sshg\.py.*?)```", text, re.S).group(1)
    assert 'C.CDLL("libsshg.so")' in code
    code = code.replace('C.CDLL("libsshg.so")', "C.CDLL(%r)" % _lib.LIB_PATH)
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    name, mat, kw = next(c for c in cases.API_CASES if c[0] == "c1_nobroad")
    m1, m2, m3, chi2 = cases.MATERIALS[mat](cases.load_spectra())
    # what shg.shgyield holds after rotate(): the oracle's host-side preparation (splines, averages, rotation)
    eps1w, eps2w, chi_rot, thick = orc.prepare(kw["energy"], m1, m2, m3, chi2, kw["thick"], kw["gamma"], 0.0, 0.0, "3-layer")
    cube = ns["yield_cube"](kw["energy"], eps1w, eps2w, chi_rot, kw["theta"], kw["phi"], thick)
    assert cube.shape == (1, 1, 4, 1, 126)
    gold = np.load(os.path.join(cases.GOLDEN_DIR, "ref_api_cases.npz"))
    sc = cases.case_scale(gold, name + "_")
    for k, pol in enumerate(cases.POLS):
        ok, msg = cases.parity_ok(cube[0, 0, k, 0], gold[name + "_" + pol], rtol=1e-10, scale=sc)
        assert ok, (pol, msg)
