#!/usr/bin/env python
"""Where does the time of a large host-output call go?  First touch of a fresh NumPy array, page-locked allocation,
device-to-host copies into pageable / registered / page-locked memory (2 GB)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
nbytes = 2 << 30
dev = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda").fill_(1.5)
torch.cuda.synchronize()
cudart = C.CDLL("libcudart.so.12")
res = {}


def t(fn):
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, r


def d2h(dst_ptr):
    rc = cudart.cudaMemcpy(C.c_void_p(dst_ptr), C.c_void_p(dev.data_ptr()), C.c_size_t(nbytes), 2)
    assert rc == 0, rc


dt, a = t(lambda: np.empty(nbytes // 8))
res["np_empty_s"] = dt
dt, _ = t(lambda: d2h(a.ctypes.data))
res["d2h_fresh_pageable_s"] = dt
dt, _ = t(lambda: d2h(a.ctypes.data))
res["d2h_touched_pageable_s"] = dt
b = np.empty(nbytes // 8)
dt, _ = t(lambda: b.fill(0.0))
res["first_touch_fill_s"] = dt
dt, _ = t(lambda: cudart.cudaHostRegister(C.c_void_p(b.ctypes.data), C.c_size_t(nbytes), 0))
res["host_register_touched_s"] = dt
dt, _ = t(lambda: d2h(b.ctypes.data))
res["d2h_registered_s"] = dt
dt, _ = t(lambda: cudart.cudaHostUnregister(C.c_void_p(b.ctypes.data)))
res["host_unregister_s"] = dt
c = np.empty(nbytes // 8)
dt, _ = t(lambda: cudart.cudaHostRegister(C.c_void_p(c.ctypes.data), C.c_size_t(nbytes), 0))
res["host_register_fresh_s"] = dt
dt, _ = t(lambda: d2h(c.ctypes.data))
res["d2h_registered_fresh_s"] = dt
cudart.cudaHostUnregister(C.c_void_p(c.ctypes.data))
p = C.c_void_p()
dt, _ = t(lambda: cudart.cudaMallocHost(C.byref(p), C.c_size_t(nbytes)))
res["malloc_host_s"] = dt
dt, _ = t(lambda: d2h(p.value))
res["d2h_pinned_s"] = dt
dt, _ = t(lambda: cudart.cudaFreeHost(p))
res["free_host_s"] = dt
# parallel first touch with threads (numpy releases the GIL in fill on large slices)
from concurrent.futures import ThreadPoolExecutor
e = np.empty(nbytes // 8)
parts = np.array_split(e, 16)
with ThreadPoolExecutor(16) as ex:
    dt, _ = t(lambda: list(ex.map(lambda v: v.fill(0.0), parts)))
res["first_touch_16_threads_s"] = dt
res["cpus"] = len(os.sched_getaffinity(0))
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))
