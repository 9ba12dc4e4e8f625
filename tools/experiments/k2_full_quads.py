#!/usr/bin/env python
"""K2 on the full config-5 cube (83.5 GB, one GPU, back to back under the power cap): pair items against quad items
(option quads = 1: 7.75 instead of 8.75 FP64 instructions per point, four row streams per thread).

    python tools/experiments/k2_full_quads.py [reps]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from shgyield_b200 import _lib, grid, shg  # noqa: E402
from shgyield_b200.synthetic import config5  # noqa: E402
from tools.bench_configs import timed  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 15
dev = torch.device("cuda", 0)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
consts = shg.physical_constants()
cfg = config5()
dd = {k: put(v) for k, v in cfg.items()}
c5 = [dd[k] for k in ("energy", "theta", "phi", "thick", "eps1w", "eps2w", "chi2")]
out = torch.empty((181, 1, 4, 721, 20000), dtype=torch.float64, device=dev)
ctx = _lib.context(0)
for rnd in range(3):
    for q in (0, 1):
        with ctx.options(quads=q):
            med, best = timed(lambda: grid.yield_grid_device(*c5, out=out, consts=consts), reps)
        print(json.dumps({"round": rnd, "quads": q, "ms_median": round(med, 3), "ms_best": round(best, 3),
                          "store_gbs": round(8e-9 * out.numel() / (med * 1e-3))}), flush=True)
