"""
Drop-in import path of the reference (`import shgyield.shg as shg`): the SSHG yield entry points,
served by the B200 implementation in `shgyield_b200` (hand-written sm_100a CUDA behind libsshg.so).
"""
from shgyield_b200.shg import (  # noqa: F401
    CHI2_KEYS, avgEPS, broad, broadC, cache_prepared, frefp, frefs, ftranp, ftrans, mrc1w, mrc2w, physical_constants, rad_pp, rad_ps,
    rad_sp, rad_ss, rotate, shgyield, spline, splineC, splineEPS, wvec, yield_points,
)
from shgyield_b200.grid import shgyield_grid  # noqa: F401
