#!/usr/bin/env python
"""Config 3 (Si(111) angular map) with the output broadening, a few launches: driver for ncu launch lists."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shgyield_b200 import grid, shg  # noqa: E402
from tests import cases  # noqa: E402

dev = torch.device("cuda", 0)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
axes = cases.config3_axes()
m1, m2, m3, chi2 = cases.si111_material()
eps1w, eps2w, chi_rot, thick = shg._prepare(axes["energy"], m1, m2, m3, chi2, axes["thick"], 0, 0.0, 0.0, "3-layer")
dd = [put(axes["energy"]), put(axes["theta"]), put(axes["phi"]), put(np.atleast_1d(thick).astype(float)), put(eps1w),
      put(eps2w), put(chi_rot)]
sigma = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
for _ in range(4):
    out = grid.yield_grid_device(*dd, consts=shg.physical_constants(), sigma_out=sigma)
torch.cuda.synchronize()
print("ok", tuple(out.shape))
