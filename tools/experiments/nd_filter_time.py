import ctypes as C, time, numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
from shgyield_b200 import _lib
ctx = _lib.context(0)
def run(shape, sigma):
    x = torch.rand(shape, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    shp = (C.c_int64 * len(shape))(*shape)
    f = lambda: _lib.check(ctx.lib.sshg_gauss_nd(ctx.handle, x.data_ptr(), y.data_ptr(), len(shape), shp, float(sigma), _lib.DEVICE), ctx.lib)
    with ctx.on_stream(torch.cuda.current_stream().cuda_stream):
        f(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
    print(shape, sigma, "device ms", round(a.elapsed_time(b), 3), "GB", round(x.numel() * 8 / 2**30, 2), flush=True)
run((64, 721, 4000), 5.0)
run((4, 360, 90, 2000), 5.0)
run((64, 721 * 4000), 5.0)
run((64 * 721, 4000), 5.0)
run((64, 721, 4000), 1.0)
