#!/usr/bin/env python
"""Latency of the reference GUI's two call shapes (gui.py:286-330) through the drop-in shgyield(), host in / host out:
the default path (prepared spectra reused by content, resident on the device, CUDA-graph replay), the same without
graphs, and cold calls (reuse switched off: K1, parallel-cyclic-reduction splines, rotation, yield, FIR every call).
One JSON object per line.

    python tools/small_calls.py [reps]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shgyield_b200 import _lib, shg  # noqa: E402
from shgyield_b200.synthetic import si111_material  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
m1, m2, m3, chi2 = si111_material()
kw = dict(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, theta=65, thick=142.1885172213904 * 5.2918e-2, gamma=0)
E = np.linspace(1.25, 2.50, 126)
shapes = {"gui_polar_361phi": dict(energy=1.7, phi=np.arange(0, 361), sigma_out=0),
          "gui_spectrum_126E_sigma_out5": dict(energy=E, phi=30, sigma_out=5),
          "spectrum_all_sigmas": dict(energy=E, phi=30, sigma_eps=3.0, sigma_chi=3.0, sigma_out=5)}
ctx = _lib.context(0)


def wall(fn, n):
    fn()
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return 1e3 * (time.perf_counter() - t0) / n


for name, s in shapes.items():
    call = lambda: shg.shgyield(**s, **kw)
    shg.cache_prepared(True)
    l0 = ctx.launches
    call()
    first_launches = ctx.launches - l0
    call()
    ms_default = wall(call, reps)
    l0 = ctx.launches
    call()
    hit_launches = ctx.launches - l0
    with ctx.options(graphs=0):
        ms_nograph = wall(call, reps)
    shg.cache_prepared(False)
    ms_cold = wall(call, max(20, reps // 5))
    l0 = ctx.launches
    call()
    cold_launches = ctx.launches - l0
    shg.cache_prepared(True)
    # a tick that changes gamma: prepared spectra miss (new rotation), spline tables of the axis stay cached
    g = [0.0]

    def tick():
        g[0] += 0.5
        shg.shgyield(**s, **dict(kw, gamma=g[0]))
    ms_gamma = wall(tick, max(20, reps // 5))
    print(json.dumps({"case": name, "ms_default": round(ms_default, 4), "ms_default_no_graph": round(ms_nograph, 4),
                      "ms_cold_no_reuse": round(ms_cold, 4), "ms_new_gamma_each_call": round(ms_gamma, 4),
                      "launches_first": first_launches, "launches_hit": hit_launches, "launches_cold": cold_launches}),
          flush=True)
