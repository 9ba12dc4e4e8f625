#!/usr/bin/env python
"""Times K3 of ONE build of the library with a given set of options (A/B of compile-time variants, one process each).

    python tools/experiments/k3_variant_run.py <libsshg.so> <tag> [key=value ...]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from shgyield_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
from shgyield_b200 import grid, shg  # noqa: E402
from shgyield_b200.synthetic import config5  # noqa: E402
from tests import cases  # noqa: E402
from tools.bench_configs import timed  # noqa: E402

tag = sys.argv[2]
opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in sys.argv[3:]}
sigmas = [float(s) for s in os.environ.get("SIGMAS", "5").split(",")]
dev = torch.device("cuda", 0)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
consts = shg.physical_constants()
m1, m2, m3, chi2 = cases.si111_material()
axes = cases.config3_axes()
eps1w, eps2w, chi_rot, thick = shg._prepare(axes["energy"], m1, m2, m3, chi2, axes["thick"], axes["gamma"], 0.0, 0.0, "3-layer")
c3 = [put(axes["energy"]), put(axes["theta"]), put(axes["phi"]), put(np.atleast_1d(thick).astype(float)), put(eps1w),
      put(eps2w), put(chi_rot)]
cfg = config5()
dd = {k: put(v) for k, v in cfg.items()}
c5 = [dd[k] for k in ("energy", "theta", "phi", "thick", "eps1w", "eps2w", "chi2")]
todo = [("config3", c3, None, (90, 1, 4, 360, 2000)), ("c5_8theta", c5, (0, 8), (8, 1, 4, 721, 20000)),
        ("c5_23theta", c5, (0, 23), (23, 1, 4, 721, 20000)), ("c5_46theta", c5, (0, 46), (46, 1, 4, 721, 20000))]
for name, args, t_range, shape in todo:
    out = torch.empty(shape, dtype=torch.float64, device=dev)
    for sigma in sigmas:
        with _lib.context(0).options(**opts):
            med, best = timed(lambda: grid.yield_grid_device(*args, out=out, consts=consts, t_range=t_range, sigma_out=sigma), 9)
        print(json.dumps({"tag": tag, "opts": opts, "case": name, "sigma_out": sigma, "ms_median": round(med, 4),
                          "ms_best": round(best, 4), "store_gbs": round(8e-9 * out.numel() / (med * 1e-3))}), flush=True)
    del out
