"""Worker of tests/test_gpu_multigpu.py (launched by torch.distributed.run, one rank per GPU):
theta-sharded sweep, bitwise comparison of every shard with the single-GPU result, NCCL gather."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shgyield_b200 import grid  # noqa: E402
from shgyield_b200.shg import physical_constants  # noqa: E402
from shgyield_b200.synthetic import config5  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = config5(n_energy=1000, n_theta=11, n_phi=73)
    consts = physical_constants()
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in cfg.items()}
    args = (d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"])
    lo, hi = grid.shard_bounds(11, world, rank)
    slab = grid.yield_grid_device(*args, consts=consts, t_range=(lo, hi))
    full_local = grid.yield_grid_device(*args, consts=consts)          # every rank also computes the whole cube
    torch.cuda.synchronize()
    ok = bool(torch.equal(slab, full_local[lo:hi]))                      # sharding changes no arithmetic
    gathered = grid.gather_slabs(slab, 11, 1, (rank, world, "theta"))
    if rank == 0:
        ok = ok and bool(torch.equal(gathered, full_local))
    else:
        ok = ok and gathered is None
    everywhere = grid.gather_slabs(slab, 11, 1, (rank, world, "theta"), all_ranks=True)
    ok = ok and bool(torch.equal(everywhere, full_local))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_OK" if int(flag.item()) == 1 else "MGPU_FAIL", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
