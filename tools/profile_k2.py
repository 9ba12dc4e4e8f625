#!/usr/bin/env python
"""Profiling driver for the fused yield kernel (K2): a theta shard of the config-5 sweep
(20000 E x 721 phi x 4 pol per theta), device resident, a few launches.  Used under ncu:

    python tools/profile_k2.py [n_theta=8] [launches=3]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shgyield_b200 import grid  # noqa: E402
from shgyield_b200.shg import physical_constants  # noqa: E402
from shgyield_b200.synthetic import config5  # noqa: E402

n_theta = int(sys.argv[1]) if len(sys.argv) > 1 else 8
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = config5()
dev = torch.device("cuda", 0)
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in cfg.items()}
out = torch.empty((n_theta, 1, 4, 721, 20000), dtype=torch.float64, device=dev)
consts = physical_constants()
t0 = 60
for _ in range(launches):
    grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"], out=out,
                           consts=consts, t_range=(t0, t0 + n_theta))
torch.cuda.synchronize()
print("ok", float(out.sum()))
