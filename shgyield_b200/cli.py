"""
`python shgyield.py <input.yml>` -- the command line the reference's README documents
(README.md:49-50).  Scalar theta / phi write the reference's text table (main.py:139-142);
lists of theta / phi / thickness run a sweep on the GPU and write the cube to an .npz file.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import io


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="shgyield.py", description="SSHG yield from chi1/chi2 spectra (B200, FP64)")
    ap.add_argument("input", help="YAML input file (see shgyield_b200/io.py::load_input)")
    ap.add_argument("-o", "--output", help="output file (overrides parameters.output)")
    args = ap.parse_args(argv)
    cfg = io.load_input(args.input)
    out = args.output or os.path.join(os.path.dirname(os.path.abspath(args.input)), cfg.pop("output"))
    cfg.pop("output", None)
    sweep = any(np.ndim(cfg[k]) > 0 for k in ("theta", "phi", "thick"))
    if not sweep:
        from .shg import shgyield
        res = shgyield(**cfg)
        io.save_spectrum(out, res)
        print("wrote %s (%d energies x 4 polarisations)" % (out, np.size(res["energy"])))
        return 0
    from .grid import shgyield_grid
    res = shgyield_grid(device=False, **cfg)
    if not out.endswith(".npz"):
        out = os.path.splitext(out)[0] + ".npz"
    np.savez(out, energy=res["energy"], theta=res["theta"], phi=res["phi"], thick=res["thick"],
             pp=res["pp"], ps=res["ps"], sp=res["sp"], ss=res["ss"])
    print("wrote %s (cube theta x thick x phi x energy = %s per polarisation)" % (out, res["pp"].shape))
    return 0


if __name__ == "__main__":
    sys.exit(main())
