#!/usr/bin/env python
"""A/B of the host-output sweep into page-locked memory with and without the single copy of equal sS rows (option
pair_copy), same process, 32 thetas of config 5 per step (14.8 GB delivered)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from shgyield_b200 import _lib, grid, shg  # noqa: E402
from shgyield_b200.synthetic import config5  # noqa: E402

cfg = config5()
consts = shg.physical_constants()
nth = int(sys.argv[1]) if len(sys.argv) > 1 else 32
host = _lib.pinned_empty((nth, 1, 4, 721, 20000))
ctx = _lib.context(0)
step = lambda: grid.yield_grid_host(cfg["energy"], cfg["theta"], cfg["phi"], cfg["thick"], cfg["eps1w"], cfg["eps2w"],
                                    cfg["chi2"], consts=consts, t_range=(0, nth), out=host)
step()
for rnd in range(3):
    for pc in (0, 1):
        with ctx.options(pair_copy=pc):
            step()
            t0 = time.perf_counter()
            for _ in range(3):
                step()
            dt = (time.perf_counter() - t0) / 3
        print(json.dumps({"round": rnd, "pair_copy": pc, "s_per_step": round(dt, 4), "delivered_gbs": round(host.nbytes / dt / 1e9, 2),
                          "points_per_s": round(host.size / dt)}), flush=True)
