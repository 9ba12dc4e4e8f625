"""Host logic of the multi-GPU path on CPU: shard arithmetic and the gather of slabs over a
world_size-2 gloo process group (the data path itself has no collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shgyield_b200 import grid


def test_shard_bounds_cover_and_balance():
    for n, w in ((181, 8), (90, 8), (100, 8), (5, 8), (181, 1), (7, 2)):
        b = [grid.shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
    assert [hi - lo for lo, hi in (grid.shard_bounds(181, 8, r) for r in range(8))] == [23] * 5 + [22] * 3
    with pytest.raises(ValueError):
        grid.shard_bounds(10, 4, 4)


def test_choose_shard_axis():
    assert grid.choose_shard_axis(181, 1, 8) == "theta"
    assert grid.choose_shard_axis(90, 100, 8) == "theta"
    assert grid.choose_shard_axis(1, 100, 8) == "thick"
    assert grid.choose_shard_axis(1, 1, 8) == "theta"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, axis, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nT, nD, nP, nE = 7, 3, 4, 5
        full = torch.arange(nT * nD * 4 * nP * nE, dtype=torch.float64).reshape(nT, nD, 4, nP, nE)
        if axis == "theta":
            lo, hi = grid.shard_bounds(nT, world, rank)
            slab = full[lo:hi].clone()
        else:
            lo, hi = grid.shard_bounds(nD, world, rank)
            slab = full[:, lo:hi].clone()
        got = grid.gather_slabs(slab, nT, nD, (rank, world, axis))              # gather onto rank 0
        ok = bool(torch.equal(got, full)) if rank == 0 else got is None
        everywhere = grid.gather_slabs(slab, nT, nD, (rank, world, axis), all_ranks=True)
        pre = torch.zeros_like(full)
        into = grid.gather_slabs(slab, nT, nD, (rank, world, axis), dst=1, out=pre if rank == 1 else None)
        ok = ok and bool(torch.equal(everywhere, full)) and (bool(torch.equal(into, full)) if rank == 1 else into is None)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def _surplus_worker(rank, world, port, q):
    """world 3, two thetas: rank 2 holds an empty shard and still takes part in the gather."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nT, nD, nP, nE = 2, 1, 3, 4
        full = torch.arange(nT * nD * 4 * nP * nE, dtype=torch.float64).reshape(nT, nD, 4, nP, nE)
        lo, hi = grid.shard_bounds(nT, world, rank)
        slab = full[lo:hi].clone()
        got = grid.gather_slabs(slab, nT, nD, (rank, world, "theta"))
        ok = (bool(torch.equal(got, full)) if rank == 0 else got is None) and (hi - lo == 0) == (rank == 2)
        everywhere = grid.gather_slabs(slab, nT, nD, (rank, world, "theta"), all_ranks=True)
        q.put((rank, ok and bool(torch.equal(everywhere, full))))
    finally:
        dist.destroy_process_group()


def test_gather_slabs_gloo_more_ranks_than_thetas():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_surplus_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True), (2, True)]
    assert grid.choose_shard_axis(2, 1, 3) == "theta"


@pytest.mark.parametrize("axis", ["theta", "thick"])
def test_gather_slabs_gloo_world2(axis):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, axis, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
