#!/usr/bin/env python
"""Driver for ncu: the column kernel (Gaussian FIR along a non-contiguous axis) on a [721, 64 x 4000] block, sigma 5."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from shgyield_b200 import _lib  # noqa: E402

ctx = _lib.context(0)
shape = (721, 256000)
x = torch.rand(shape, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
shp = (C.c_int64 * 2)(*shape)
with ctx.on_stream(torch.cuda.current_stream().cuda_stream):
    for _ in range(3):
        _lib.check(ctx.lib.sshg_gauss_nd(ctx.handle, x.data_ptr(), y.data_ptr(), 2, shp, 5.0, _lib.DEVICE), ctx.lib)
    torch.cuda.synchronize()
print("ok")
