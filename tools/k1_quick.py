#!/usr/bin/env python
"""Quick timing of K1 (Gaussian FIR along the energy axis) on a cube slab: 8 thetas of config 5 (461 M samples, 7.4 GB
read + written), with the tiles staged by bulk copies (cp.async.bulk, the default) and by per-thread loads / stores
(option k1_bulk = 0), same process, same buffers.  One JSON line per sigma.

    python tools/k1_quick.py [tag] [nth]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shgyield_b200 import _lib, grid  # noqa: E402
from tools.bench_configs import timed  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else ""
nth = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
slab = torch.rand((nth, 1, 4, 721, 20000), dtype=torch.float64, device=dev)
n = slab.numel()
ctx = _lib.context(0)
for sigma in (0.7, 1.0, 2.0, 3.0, 3.3, 5.0, 6.0, 7.3, 8.0, 10.0, 12.0, 15.3, 16.0):
    row = {"tag": tag, "case": "k1_cube_sigma%g" % sigma, "samples": n, "taps": 2 * int(4 * sigma + 0.5) + 1}
    for name, opts in (("bulk", {}), ("per_thread", {"k1_bulk": 0})):
        with ctx.options(**opts):
            med, best = timed(lambda: grid.broaden_energy_device(slab, sigma), 9)
        row["ms_" + name] = round(med, 4)
        row["ms_best_" + name] = round(best, 4)
        row["rw_gbs_" + name] = round(16e-9 * n / (med * 1e-3))
        row["fp64_tflops_" + name] = round(2.0 * row["taps"] * n / (med * 1e-3) / 1e12, 2)
    print(json.dumps(row), flush=True)
