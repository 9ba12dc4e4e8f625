#!/usr/bin/env python
"""Profiling driver for ncu: runs a few launches of one kernel on a representative workload.

    python tools/profile_case.py k2 [n_theta]    fused yield kernel, config-5 theta shard
    python tools/profile_case.py k3 [n_theta] [sigma_out]   broadened sweep (fused K3), config-5 theta shard
    python tools/profile_case.py k1 [sigma]      Gaussian FIR along E of a 4-theta config-5 slab
    python tools/profile_case.py c4              config 4 (thickness sweep, P = 1)
    python tools/profile_case.py c3              config 3 (angular map)
    python tools/profile_case.py small           the reference's small call shapes (timing of every C call)
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shgyield_b200 import _lib, grid, shg  # noqa: E402
from shgyield_b200.synthetic import config5  # noqa: E402
from tests import cases  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "k2"
dev = torch.device("cuda", 0)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
consts = shg.physical_constants()

if case in ("k2", "k1", "k3"):
    cfg = config5()
    d = {k: put(v) for k, v in cfg.items()}
    nth = int(sys.argv[2]) if case in ("k2", "k3") and len(sys.argv) > 2 else (37 if case == "k2" else 4)
    sigma_out = (float(sys.argv[3]) if len(sys.argv) > 3 else 5.0) if case == "k3" else 0.0
    out = torch.empty((nth, 1, 4, 721, 20000), dtype=torch.float64, device=dev)
    for _ in range(1 if case == "k1" else 3):
        grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"], out=out,
                               consts=consts, t_range=(60, 60 + nth), sigma_out=sigma_out)
    if case == "k1":
        sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
        for _ in range(3):
            res = grid.broaden_energy_device(out, sigma)
    torch.cuda.synchronize()
    print("ok", case, float(out[0, 0, 0, 0, :10].sum()))
elif case in ("c3", "c4"):
    axes = cases.config3_axes() if case == "c3" else cases.config4_axes()
    m1, m2, m3, chi2 = cases.si111_material()
    eps1w, eps2w, chi_rot, thick = shg._prepare(axes["energy"], m1, m2, m3, chi2, axes["thick"], 0, 0.0, 0.0, "3-layer")
    dd = [put(axes["energy"]), put(axes["theta"]), put(axes["phi"]), put(np.atleast_1d(thick).astype(float)), put(eps1w),
          put(eps2w), put(chi_rot)]
    for _ in range(3):
        out = grid.yield_grid_device(*dd, consts=consts)
    torch.cuda.synchronize()
    print("ok", case, tuple(out.shape))
else:
    import cProfile
    import pstats
    m1, m2, m3, chi2 = cases.si111_material()
    kw = dict(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, theta=65, thick=10, gamma=0)
    call = lambda: shg.shgyield(energy=cases.E_C1, phi=30, sigma_out=5, **kw)
    for _ in range(3):
        call()
    # wall time of every library call
    lib = _lib.context().lib
    names = ["sshg_broaden_resample", "sshg_rotate", "sshg_yield_points", "sshg_gauss1d", "sshg_yield"]
    acc = {n: [0, 0.0] for n in names}
    orig = {n: getattr(lib, n) for n in names}

    def wrap(n):
        f = orig[n]

        def g(*a):
            t0 = time.perf_counter()
            r = f(*a)
            acc[n][0] += 1
            acc[n][1] += time.perf_counter() - t0
            return r
        return g
    for n in names:
        setattr(lib, n, wrap(n))
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    total = time.perf_counter() - t0
    for n in names:
        setattr(lib, n, orig[n])
    print("shgyield(E=126) %.3f ms per call" % (1e3 * total / reps))
    for n, (c, t) in acc.items():
        if c:
            print("  %-24s %5.1f calls/shgyield  %8.3f ms per call  %8.3f ms per shgyield" % (n, c / reps, 1e3 * t / c, 1e3 * t / reps))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        call()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
