"""Tuning run: fused (K3) against two-pass output broadening at small sigma_out on a 23-theta shard of config 5."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from shgyield_b200 import _lib, grid, shg
from shgyield_b200.synthetic import config5
from tools.bench_configs import timed
dev = torch.device("cuda", 0)
cfg = config5()
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in cfg.items()}
consts = shg.physical_constants()
out = torch.empty((23, 1, 4, 721, 20000), dtype=torch.float64, device=dev)
args = [d[k] for k in ("energy", "theta", "phi", "thick", "eps1w", "eps2w", "chi2")]
for sigma in (0.3, 0.5, 0.75, 1.0, 1.5, 2.0):
    for k3 in ("1", "0"):
        with _lib.context(0).options(k3=int(k3)):
            med, best = timed(lambda: grid.yield_grid_device(*args, out=out, consts=consts, t_range=(0, 23), sigma_out=sigma), 5)
        print(json.dumps({"sigma": sigma, "radius": int(4*sigma+0.5), "path": "K3" if k3 == "1" else "two-pass", "ms": round(med, 3)}), flush=True)
