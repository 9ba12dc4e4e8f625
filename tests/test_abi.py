"""CPU checks of the drop-in boundary: libsshg.so loads, exports every symbol include/sshg.h
declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from shgyield_b200 import _lib, build
from tests import util


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    hdr = open(os.path.join(util.REPO, "include", "sshg.h")).read()
    return sorted(set(re.findall(r"SSHG_API[^;(]*?\b(sshg_\w+)\s*\(", hdr)))


def test_header_and_binding_agree(lib):
    names = declared_symbols()
    assert len(names) >= 18
    assert sorted(_lib.SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None


def test_abi_version(lib):
    assert lib.sshg_abi_version() == _lib.ABI_VERSION
    hdr = open(os.path.join(util.REPO, "include", "sshg.h")).read()
    assert int(re.search(r"#define SSHG_ABI_VERSION (\d+)", hdr).group(1)) == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    # sshg_physics: 4 doubles + 4 int32; sshg_problem: 4 int64 + 7 pointers + physics + 4 int64
    assert C.sizeof(_lib.Physics) == 48
    assert C.sizeof(_lib.Problem) == 4 * 8 + 7 * 8 + 48 + 4 * 8
    assert C.sizeof(_lib.Points) == 2 * 8 + 8 * 8 + 48


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.sshg_ctx_create(0, C.byref(h))
    assert rc == _lib.SSHG_ECUDA and not h.value
    assert lib.sshg_last_error()
    with pytest.raises(RuntimeError):
        _lib.Context(0)
    from shgyield_b200 import shg
    with pytest.raises(RuntimeError):
        shg.broad(np.ones(8), 2.0)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load(str(tmp_path / "libsshg.so"))


def test_product_never_imports_oracle():
    """The product tree must not reference oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(util.REPO, "shgyield_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_library_reads_no_environment():
    """Tests and tuning drive the library through sshg_ctx_set_option; nothing in csrc/ may call getenv (a stray
    environment variable must not change results or speed of a drop-in library)."""
    csrc = os.path.join(util.REPO, "shgyield_b200", "csrc")
    for f in os.listdir(csrc):
        assert "getenv" not in open(os.path.join(csrc, f)).read(), f
