import json, os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from shgyield_b200 import _lib, grid, shg
from shgyield_b200.synthetic import config5
from tools.bench_configs import timed
dev = torch.device("cuda", 0)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
consts = shg.physical_constants()
ctx = _lib.context(0)
cfg = config5(); dd = {k: put(v) for k, v in cfg.items()}
c5 = [dd[k] for k in ("energy", "theta", "phi", "thick", "eps1w", "eps2w", "chi2")]
for name, t_range, shape in (("c5_23theta", (0, 23), (23, 1, 4, 721, 20000)), ("c5_46theta", (0, 46), (46, 1, 4, 721, 20000))):
    out = torch.empty(shape, dtype=torch.float64, device=dev)
    row = {"case": name}
    for rep in range(2):
        for fix in (1, 0):
            with ctx.options(k3_fixup=fix):
                med, best = timed(lambda: grid.yield_grid_device(*c5, out=out, consts=consts, t_range=t_range, sigma_out=5.0), 9)
            row["fixup%d_rep%d" % (fix, rep)] = round(med, 4)
    print(json.dumps(row), flush=True)
