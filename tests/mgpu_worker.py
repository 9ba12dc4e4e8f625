"""Worker of tests/test_gpu_multigpu.py (launched by torch.distributed.run, one rank per GPU):
theta-sharded sweep, bitwise comparison of every shard with the single-GPU result, NCCL gather, and the
gather by peer stores (every rank's kernel writes into rank 0's cube over NVLink)."""
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
    # thickness sharding (fewer thetas than ranks): slabs are strided slices of the cube
    cfg2 = config5(n_energy=200, n_theta=1, n_phi=5)
    thick = np.linspace(1.0, 40.0, 9)
    d2 = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in cfg2.items()}
    d2["thick"] = torch.from_numpy(thick).to(dev)
    args2 = (d2["energy"], d2["theta"], d2["phi"], d2["thick"], d2["eps1w"], d2["eps2w"], d2["chi2"])
    assert grid.choose_shard_axis(1, 9, world) == "thick"
    lo2, hi2 = grid.shard_bounds(9, world, rank)
    slab2 = grid.yield_grid_device(*args2, consts=consts, d_range=(lo2, hi2))
    full2 = grid.yield_grid_device(*args2, consts=consts)
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(slab2, full2[:, lo2:hi2]))
    g2 = grid.gather_slabs(slab2, 1, 9, (rank, world, "thick"), all_ranks=True)
    ok = ok and bool(torch.equal(g2, full2))

    # the public sweep entry point with the reference's material arguments, sharded + gathered
    from tests import cases
    m1, m2, m3, chi2 = cases.si111_material()
    kw = dict(energy=np.linspace(0.5, 5.0, 300), eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2,
              theta=np.linspace(10, 80, 9), phi=np.arange(0, 360, 30.0), thick=[10.0], gamma=0, sigma_out=2.0)
    whole = grid.shgyield_grid(**kw)["cube"]
    part = grid.shgyield_grid(shard=(rank, world), gather=True, **kw)["cube"]
    torch.cuda.synchronize()
    if rank == 0:
        ok = ok and bool(torch.equal(part, whole))
    else:
        ok = ok and part is None

    # the gather without a collective: every rank's kernel stores its slab into rank 0's cube (CUDA IPC, NVLink)
    pc = grid.PeerCube((11, 1, 4, 73, 1000), dst=0)
    grid.yield_grid_device(*args, consts=consts, t_range=(lo, hi), out_ptr=pc.slab_ptr(lo))
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        ok = ok and bool(torch.equal(pc.tensor(), full_local))
    else:
        ok = ok and pc.tensor() is None
    pc.close()
    del pc
    part = grid.shgyield_grid(shard=(rank, world), gather="peer", **kw)["cube"]     # with the output broadening (K3)
    if rank == 0:
        ok = ok and bool(torch.equal(part, whole))
    else:
        ok = ok and part is None
    del part

    # more ranks than thetas: the surplus ranks hold an empty shard, skip the kernel, and still take part in
    # the collective construction of the peer cube, the barrier and the NCCL gather
    kw1 = dict(kw, theta=[65.0])
    assert grid.choose_shard_axis(1, 1, world) == "theta"
    whole1 = grid.shgyield_grid(**kw1)["cube"]
    for how in (True, "peer"):
        part = grid.shgyield_grid(shard=(rank, world), gather=how, **kw1)["cube"]
        torch.cuda.synchronize()
        ok = ok and (bool(torch.equal(part, whole1)) if rank == 0 else part is None)
    empty = grid.shgyield_grid(shard=(world - 1, world), **kw1)["cube"]
    ok = ok and tuple(empty.shape) == (0,) + tuple(whole1.shape[1:])

    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_OK" if int(flag.item()) == 1 else "MGPU_FAIL", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
