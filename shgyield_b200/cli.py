"""
`python shgyield.py <input.yml>` -- the command line the reference's README documents
(README.md:49-50).  Scalar theta / phi write the reference's text table (main.py:139-142);
lists of theta / phi / thickness run a sweep on the GPU and write the cube to an .npz file; under
`python -m torch.distributed.run --nproc-per-node N shgyield.py <input.yml>` the sweep is sharded over N GPUs.
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
        cfg.pop("strict", None)
        res = shgyield(**cfg)
        io.save_spectrum(out, res)
        print("wrote %s (%d energies x 4 polarisations)" % (out, np.size(res["energy"])))
        return 0
    from .grid import shgyield_grid
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # launched by torch.distributed.run, one rank per GPU: theta (or thickness) shards, the cube gathered on
        # rank 0 by peer stores over NVLink (theta shards) or NCCL (thickness shards); rank 0 writes the file
        import torch
        import torch.distributed as dist
        rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        try:
            res = shgyield_grid(device=True, shard=(rank, world), gather="peer", **cfg)
            if rank == 0:
                from .grid import to_host
                cube = to_host(res["cube"])
                res = dict(res, cube=cube, **{pol: cube[:, :, k] for k, pol in enumerate(("pp", "ps", "sp", "ss"))})
            dist.barrier()
        finally:
            dist.destroy_process_group()
        if rank != 0:
            return 0
    else:
        res = shgyield_grid(device=False, **cfg)
    if not out.endswith(".npz"):
        out = os.path.splitext(out)[0] + ".npz"
    np.savez(out, energy=res["energy"], theta=res["theta"], phi=res["phi"], thick=res["thick"],
             pp=res["pp"], ps=res["ps"], sp=res["sp"], ss=res["ss"])
    print("wrote %s (cube theta x thick x phi x energy = %s per polarisation)" % (out, res["pp"].shape))
    return 0


if __name__ == "__main__":
    sys.exit(main())
