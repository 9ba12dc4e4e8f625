"""
ctypes binding of libsshg.so (C ABI: include/sshg.h).

The library is built in-tree by `shgyield_b200.build` and loaded lazily.  There is no CPU
fallback: if the library is missing, cannot be loaded, or no B200 is visible, the first compute
call raises RuntimeError.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsshg.so")

SSHG_OK, SSHG_EINVAL, SSHG_ECUDA, SSHG_ENOMEM = 0, -1, -2, -3
HOST, IN_DEVICE, OUT_DEVICE, DEVICE, SPECTRA_DEVICE = 0, 1, 2, 3, 4
ABI_VERSION = 2
FLAG_RADIANS = 1
FLAG_STRICT_BROADENING = 2
IPC_HANDLE_BYTES = 64

c_double_p = C.POINTER(C.c_double)


class Physics(C.Structure):
    _fields_ = [("h_eVs", C.c_double), ("hbar_eVs", C.c_double), ("eps0", C.c_double), ("c", C.c_double),
                ("mref", C.c_int32), ("sinc_normalized", C.c_int32), ("zzz_phi_quirk", C.c_int32),
                ("flags", C.c_int32)]


class Problem(C.Structure):
    _fields_ = [("nE", C.c_int64), ("nT", C.c_int64), ("nP", C.c_int64), ("nD", C.c_int64),
                ("energy_eV", C.c_void_p), ("theta_deg", C.c_void_p), ("phi_deg", C.c_void_p),
                ("thick_nm", C.c_void_p), ("eps1w", C.c_void_p), ("eps2w", C.c_void_p), ("chi2", C.c_void_p),
                ("phys", Physics),
                ("t0", C.c_int64), ("t1", C.c_int64), ("d0", C.c_int64), ("d1", C.c_int64)]


class Points(C.Structure):
    _fields_ = [("n", C.c_int64), ("nE", C.c_int64),
                ("energy_eV", C.c_void_p), ("eidx", C.c_void_p), ("theta_deg", C.c_void_p),
                ("phi_deg", C.c_void_p), ("thick_nm", C.c_void_p), ("eps1w", C.c_void_p),
                ("eps2w", C.c_void_p), ("chi2", C.c_void_p),
                ("phys", Physics)]


# every symbol include/sshg.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sshg_abi_version": (C.c_int, []),
    "sshg_last_error": (C.c_char_p, []),
    "sshg_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "sshg_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "sshg_ctx_destroy": (None, [C.c_void_p]),
    "sshg_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sshg_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "sshg_ctx_sync": (C.c_int, [C.c_void_p]),
    "sshg_ctx_launches": (C.c_int64, [C.c_void_p]),
    "sshg_ctx_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "sshg_ctx_get_option": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "sshg_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "sshg_host_free": (C.c_int, [C.c_void_p]),
    "sshg_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "sshg_hash64": (C.c_uint64, [C.c_void_p, C.c_int64, C.c_uint64]),
    "sshg_device_alloc": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int64]),
    "sshg_device_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sshg_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p]),
    "sshg_ipc_open": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "sshg_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sshg_gauss1d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                               C.c_double, C.c_int]),
    "sshg_gauss_nd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.c_int]),
    "sshg_rotate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_int]),
    "sshg_spline_resample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                       C.c_int64, C.c_void_p, C.c_int]),
    "sshg_broaden_resample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double,
                                        C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "sshg_yield": (C.c_int, [C.c_void_p, C.POINTER(Problem), C.c_void_p, C.c_int]),
    "sshg_yield_broadened": (C.c_int, [C.c_void_p, C.POINTER(Problem), C.c_double, C.c_void_p, C.c_int]),
    "sshg_yield_points": (C.c_int, [C.c_void_p, C.POINTER(Points), C.c_void_p, C.c_int]),
    "sshg_yield_points_broadened": (C.c_int, [C.c_void_p, C.POINTER(Points), C.c_double, C.c_int32, C.c_void_p, C.c_void_p,
                                              C.c_int]),
    "sshg_fresnel": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "sshg_multiref": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.POINTER(Physics), C.c_void_p, C.c_int64, C.c_int]),
    "sshg_amplitudes_points": (C.c_int, [C.c_void_p, C.POINTER(Points), C.c_void_p, C.c_int]),
    "sshg_yield_broadcast": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_double, C.c_int32,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_void_p)]),
    "sshg_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "sshg_peak_store": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, c_double_p]),
    "sshg_peak_copy": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, c_double_p]),
    "sshg_peak_dfma": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
}

_lib = None
_lock = threading.Lock()


def load(path: str | None = None):
    """Loads libsshg.so (no device needed) and types every exported symbol."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise RuntimeError(
                "libsshg.so not found at %s: build it with `python -m shgyield_b200.build` "
                "(there is no CPU fallback)" % p)
        lib = C.CDLL(p)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.sshg_abi_version() != ABI_VERSION:
            raise RuntimeError("libsshg.so ABI %d, binding expects %d" % (lib.sshg_abi_version(), ABI_VERSION))
        if path is None:
            _lib = lib
        return lib


def check(rc: int, lib=None):
    if rc == SSHG_OK:
        return
    msg = (lib or load()).sshg_last_error().decode("utf-8", "replace")
    if rc == SSHG_EINVAL:
        raise ValueError(msg)
    if rc == SSHG_ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


class Context:
    """One CUDA device + stream + scratch (sshg_ctx).  Not thread-safe; one per rank."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        check(self.lib.sshg_ctx_create(int(device), C.byref(h)), self.lib)
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.sshg_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        """Adopt a cudaStream_t handle (torch: stream.cuda_stream); None = the context's own stream.
        Handle 0 is the legacy default stream and is passed as cudaStreamLegacy (1)."""
        h = 0 if cuda_stream is None else (1 if cuda_stream == 0 else cuda_stream)
        check(self.lib.sshg_ctx_set_stream(self.handle, C.c_void_p(h)), self.lib)

    def sync(self):
        check(self.lib.sshg_ctx_sync(self.handle), self.lib)

    def set_option(self, key: str, value: int | None):
        """Test / tuning switch of this context (sshg_ctx_set_option, include/sshg.h); None = the library's choice."""
        check(self.lib.sshg_ctx_set_option(self.handle, key.encode(), -1 if value is None else int(value)), self.lib)

    def get_option(self, key: str) -> int | None:
        v = C.c_int64()
        check(self.lib.sshg_ctx_get_option(self.handle, key.encode(), C.byref(v)), self.lib)
        return None if v.value == -1 else int(v.value)

    @contextlib.contextmanager
    def options(self, **kv):
        """`with ctx.options(k3=0, k2_sparse=1): ...` -- sets the switches and restores the previous values."""
        old = {k: self.get_option(k) for k in kv}
        try:
            for k, v in kv.items():
                self.set_option(k, v)
            yield self
        finally:
            for k, v in old.items():
                self.set_option(k, v)

    @contextlib.contextmanager
    def on_stream(self, cuda_stream: int | None):
        """Queue the library calls of the block on a caller's stream, then return to the context's own stream.
        Both switches are ordered on the device (events), so the scratch buffers stay consistent and the context
        never keeps the handle of a stream the caller may destroy."""
        self.set_stream(cuda_stream)
        try:
            yield self
        finally:
            self.set_stream(None)

    @property
    def launches(self) -> int:
        return int(self.lib.sshg_ctx_launches(self.handle))


_contexts: dict = {}


def context(device: int | None = None) -> Context:
    """Process-wide context of a device (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    key = (os.getpid(), int(device))
    ctx = _contexts.get(key)
    if ctx is None or ctx.handle is None:
        ctx = _contexts[key] = Context(device)
    return ctx


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """Page-locked host array (cudaMallocHost) so that device-to-host copies run at PCIe speed."""
    lib = load()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    p = C.c_void_p()
    check(lib.sshg_host_alloc(C.byref(p), max(n, 1)), lib)
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)
    _pinned_keepalive[arr.ctypes.data] = (buf, _PinnedOwner(lib, p.value))
    return arr


class _PinnedOwner:
    def __init__(self, lib, addr):
        self.lib, self.addr = lib, addr

    def free(self):
        if self.addr:
            self.lib.sshg_host_free(C.c_void_p(self.addr))
            self.addr = 0


_pinned_keepalive: dict = {}


def pinned_free(arr: np.ndarray):
    ent = _pinned_keepalive.pop(arr.ctypes.data, None)
    if ent:
        ent[1].free()
