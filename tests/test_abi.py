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


def test_content_hash(lib):
    """sshg_hash64 (the key of the prepared-spectra cache): deterministic, sensitive to every byte and to the length,
    fine with unaligned tails; ~GB/s so that hashing a material costs tens of microseconds."""
    import time
    rng = np.random.default_rng(0)
    buf = rng.integers(0, 256, 100003, dtype=np.uint8)
    h = lambda a, n=None, seed=0: int(lib.sshg_hash64(a.ctypes.data, a.nbytes if n is None else n, seed))
    base = h(buf)
    assert base == h(buf.copy()) and base != h(buf, seed=1)
    seen = {base}
    for pos in (0, 1, 31, 32, 33, 50000, 100002):
        mod = buf.copy()
        mod[pos] ^= 1
        seen.add(h(mod))
    for n in (0, 1, 7, 8, 31, 32, 33, 100002):
        seen.add(h(buf, n))
    assert len(seen) == 1 + 7 + 8
    big = np.zeros(1 << 22, dtype=np.float64)
    t0 = time.perf_counter()
    h(big)
    assert big.nbytes / (time.perf_counter() - t0) > 1e9


def test_content_key_of_the_prepared_cache(lib):
    """The cache key sees values, not object identity: equal copies give the same key, an in-place change another."""
    from shgyield_b200 import shg
    e = np.linspace(0, 20, 2001)
    mk = lambda: {"xx": {"energy": e.copy(), "data": np.cos(e) + 1j * np.sin(e)}, "yy": 2.5, "zz": np.full(5, 1.0 + 2j)}
    m, chi = mk(), {"xxx": {"energy": e, "data": e * 1j}, "zzz": 0}
    key = lambda mm: shg._content_key(np.linspace(1, 2, 11), (mm, mm, mm, chi), 90, 0.0, 0.0, "3-layer")
    k0 = key(m)
    assert key(mk()) == k0
    m["xx"]["data"][100] += 1e-12
    assert key(m) != k0
    assert shg._content_key(np.linspace(1, 2, 11), (mk(), mk(), mk(), chi), 90, 0.0, 0.5, "3-layer") != k0
    with pytest.raises(TypeError):
        shg._content_key(np.linspace(1, 2, 11), ({"xx": None},), 90, 0.0, 0.0, "3-layer")
