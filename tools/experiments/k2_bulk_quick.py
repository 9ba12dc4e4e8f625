#!/usr/bin/env python
"""A/B of the dense yield kernel's store path (library option "k2_bulk": 0 = per-thread streaming stores, 4 / 8 = runs
staged in shared memory and written by cp.async.bulk, 4 / 8 row pairs per batch) in ONE process on ONE box:
config 3, config-5 shards and the full 83.5 GB cube.  One JSON line per case.

    python tools/k2_quick.py [tag]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shgyield_b200 import _lib, grid, shg  # noqa: E402
from shgyield_b200.synthetic import config3_axes, config5, si111_material  # noqa: E402
from tools.bench_configs import timed  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else ""
dev = torch.device("cuda", 0)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
consts = shg.physical_constants()
ctx = _lib.context(0)
m1, m2, m3, chi2 = si111_material()
axes = config3_axes()
eps1w, eps2w, chi_rot, thick = shg._prepare(axes["energy"], m1, m2, m3, chi2, axes["thick"], axes["gamma"], 0.0, 0.0, "3-layer")
c3 = [put(axes["energy"]), put(axes["theta"]), put(axes["phi"]), put(np.atleast_1d(thick).astype(float)), put(eps1w),
      put(eps2w), put(chi_rot)]
cfg = config5()
dd = {k: put(v) for k, v in cfg.items()}
c5 = [dd[k] for k in ("energy", "theta", "phi", "thick", "eps1w", "eps2w", "chi2")]
cases_ = [("config3", c3, None, (90, 1, 4, 360, 2000)), ("c5_23theta", c5, (0, 23), (23, 1, 4, 721, 20000)),
          ("c5_46theta", c5, (0, 46), (46, 1, 4, 721, 20000)), ("c5_full_181theta", c5, None, (181, 1, 4, 721, 20000))]
for name, args, t_range, shape in cases_:
    out = torch.empty(shape, dtype=torch.float64, device=dev)
    ref = None
    row = {"tag": tag, "case": name}
    for bulk in (0, 4, 8, 0, 8):  # NOTE: option removed from the library after the experiment (git history)
        with ctx.options(k2_bulk=bulk):
            med, best = timed(lambda: grid.yield_grid_device(*args, out=out, consts=consts, t_range=t_range), 7)
        key = "bulk%d" % bulk + ("_again" if ("bulk%d_ms" % bulk) in row else "")
        row[key + "_ms"] = round(med, 4)
        row[key + "_gbs"] = round(8e-9 * out.numel() / (med * 1e-3))
        if ref is None:
            ref = out[:2].clone()
        else:
            row["bitwise_equal_to_direct"] = bool(row.get("bitwise_equal_to_direct", True) and torch.equal(out[:2], ref))
    print(json.dumps(row), flush=True)
    del out, ref
    torch.cuda.empty_cache()
