#!/usr/bin/env python
"""Wall time of the drop-in shgyield() on large NumPy-broadcast calls -- the reference's only way to express a sweep
(energy[E], theta[T, 1], phi[P, 1, 1]; SURVEY 8c 'L-grid') -- host arrays in, host arrays out.  One JSON line per case.

    python tools/broadcast_call.py [tag]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shgyield_b200 import grid, shg  # noqa: E402
from shgyield_b200.synthetic import si111_material  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else ""
m1, m2, m3, chi2 = si111_material()
kw = dict(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, thick=10.0, gamma=0)
cases_ = {"E126_T1_P361": (np.linspace(1.25, 2.5, 126), np.array([65.0]), np.arange(0, 361.0)),
          "E500_T30_P72": (np.linspace(0.5, 5.0, 500), np.arange(0, 90, 3.0), np.arange(0, 360, 5.0)),
          "config3_E2000_T90_P360": (np.linspace(0.005, 10.0, 2000), np.arange(0, 90, 1.0), np.arange(0, 360, 1.0))}
for name, (E, T, P) in cases_.items():
    pts = 4 * E.size * T.size * P.size
    for sigma_out in (0.0, 5.0):
        call = lambda: shg.shgyield(energy=E, theta=T[:, None], phi=P[:, None, None], sigma_out=sigma_out, **kw)
        call()
        t0 = time.perf_counter()
        reps = 3 if pts > 10 ** 7 else 30
        for _ in range(reps):
            r = call()
        dt = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        g = grid.shgyield_grid(E, m1, m2, m3, chi2, T, P, 10.0, gamma=0, sigma_out=sigma_out, device=False)
        dt_grid = time.perf_counter() - t0
        print(json.dumps({"tag": tag, "case": name, "sigma_out": sigma_out, "points": pts, "shape": list(r["pp"].shape),
                          "s_per_call": round(dt, 4), "points_per_s": round(pts / dt), "s_shgyield_grid_host": round(dt_grid, 4)}),
              flush=True)
