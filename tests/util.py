"""Helpers shared by the tests: packing oracle-style dicts into the C-ABI array layout."""
from __future__ import annotations

import os

import numpy as np

from tests import cases

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_consts():
    z = np.load(os.path.join(cases.GOLDEN_DIR, "constants.npz"))
    return {k: float(z[k]) for k in z.files}


def pack_eps(eps: dict, n: int) -> np.ndarray:
    """{'M1','M2','M3'} (scalars or arrays) -> complex128 [3, n]."""
    out = np.empty((3, n), dtype=np.complex128)
    for i, k in enumerate(("M1", "M2", "M3")):
        out[i] = np.broadcast_to(np.asarray(eps[k], dtype=np.complex128), (n,))
    return out


def pack_chi(chi: dict, n: int) -> np.ndarray:
    out = np.empty((18, n), dtype=np.complex128)
    for i, k in enumerate(cases.CHI2_KEYS):
        out[i] = np.broadcast_to(np.asarray(chi[k], dtype=np.complex128), (n,))
    return out


def lib_options(**kv):
    """`with util.lib_options(k3=0, k2_sparse=1): ...` -- test / tuning switches of the process-wide context
    (sshg_ctx_set_option; the library reads no environment variables), restored on exit."""
    from shgyield_b200 import _lib
    return _lib.context().options(**kv)

