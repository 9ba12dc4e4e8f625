#!/usr/bin/env python
"""Timing of the broadened full config-5 cube (K3) and of K2 on the same box: python tools/k3_full_cube.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shgyield_b200 import grid, shg  # noqa: E402
from shgyield_b200.synthetic import config5  # noqa: E402
from tools.bench_configs import timed  # noqa: E402

dev = torch.device("cuda", 0)
cfg = config5()
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in cfg.items()}
consts = shg.physical_constants()
out = torch.empty((181, 1, 4, 721, 20000), dtype=torch.float64, device=dev)
args = [d[k] for k in ("energy", "theta", "phi", "thick", "eps1w", "eps2w", "chi2")]
for rep in range(2):
    for sigma in (0.0, 5.0):
        med, best = timed(lambda: grid.yield_grid_device(*args, out=out, consts=consts, sigma_out=sigma), 7, warm=4)
        print(json.dumps({"sigma_out": sigma, "ms_median": round(med, 3), "ms_best": round(best, 3),
                          "store_gbs": round(8e-9 * out.numel() / (med * 1e-3))}), flush=True)
