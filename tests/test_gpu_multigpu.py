"""Multi-GPU path on real devices (needs >= 2 GPUs; skipped otherwise): one process per GPU over
NCCL, shards bitwise equal to the single-GPU cube, gather of the slabs."""
import os
import subprocess
import sys

import pytest

from tests import util

pytestmark = pytest.mark.gpu


def test_sharded_sweep_and_nccl_gather():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(util.REPO, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "MGPU_OK %d" % world in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
