#!/usr/bin/env python
"""Throughput of sshg_download (threaded staged device-to-host copy) into fresh and into reused pageable arrays, and
into page-locked memory, 2 GB."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from shgyield_b200 import _lib  # noqa: E402

if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
ctx = _lib.context(0)
n = (2 << 30) // 8
dev = torch.rand(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()


def dl(arr):
    t0 = time.perf_counter()
    _lib.check(ctx.lib.sshg_download(ctx.handle, arr.ctypes.data, dev.data_ptr(), arr.nbytes), ctx.lib)
    return time.perf_counter() - t0


dl(np.empty(1 << 23))                       # staging buffers allocated
res = {}
ts = []
for _ in range(3):
    a = np.empty(n)
    ts.append(dl(a))
    assert a[12345] == float(dev[12345]) and a[-1] == float(dev[-1])
    t0 = time.perf_counter()
    del a
    res["free_s"] = round(time.perf_counter() - t0, 4)
res["fresh_pageable_s"] = [round(t, 4) for t in ts]
a = np.empty(n)
dl(a)
res["reused_pageable_s"] = [round(dl(a), 4) for _ in range(3)]
p = _lib.pinned_empty((n,))
res["pinned_s"] = [round(dl(p), 4) for _ in range(3)]
res["gbs_fresh"] = round(2.147 / min(ts), 1)
res["lib"] = os.path.basename(_lib.LIB_PATH)
print(json.dumps(res))
