import json, os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from shgyield_b200 import _lib, grid, shg
from shgyield_b200.synthetic import config5
from tests import cases
from tools.bench_configs import timed
dev = torch.device("cuda", 0)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
consts = shg.physical_constants()
ctx = _lib.context(0)
m1, m2, m3, chi2 = cases.si111_material()
axes = cases.config3_axes()
eps1w, eps2w, chi_rot, thick = shg._prepare(axes["energy"], m1, m2, m3, chi2, axes["thick"], axes["gamma"], 0.0, 0.0, "3-layer")
c3 = [put(axes["energy"]), put(axes["theta"]), put(axes["phi"]), put(np.atleast_1d(thick).astype(float)), put(eps1w), put(eps2w), put(chi_rot)]
cfg = config5(); dd = {k: put(v) for k, v in cfg.items()}
c5 = [dd[k] for k in ("energy", "theta", "phi", "thick", "eps1w", "eps2w", "chi2")]
for name, args, t_range, shape in (("config3", c3, None, (90, 1, 4, 360, 2000)), ("c5_23theta", c5, (0, 23), (23, 1, 4, 721, 20000))):
    out = torch.empty(shape, dtype=torch.float64, device=dev)
    for nt in (64, 128):
        for nseg in (1, 2):
            for fix in (1, 0):
                with ctx.options(k3_nt=nt, nseg=nseg, k3_fixup=fix):
                    med, best = timed(lambda: grid.yield_grid_device(*args, out=out, consts=consts, t_range=t_range, sigma_out=5.0), 9)
                print(json.dumps({"case": name, "k3_nt": nt, "nseg": nseg, "fixup": fix, "ms_median": round(med, 4), "ms_best": round(best, 4)}), flush=True)
