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


def test_cli_sweep_sharded_over_gpus(tmp_path):
    """`torch.distributed.run ... shgyield.py sweep.yml`: theta shards on 2 GPUs, the cube gathered on rank 0 by
    peer stores; bitwise the single-process result."""
    import numpy as np
    import shutil
    import torch
    import yaml
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, os.path.join(util.REPO, "example"))
    import make_example_data
    d = str(tmp_path)
    make_example_data.write(d)
    doc = yaml.safe_load(open(os.path.join(util.REPO, "example", "si111-input.yml")))
    doc["parameters"]["theta"] = [15, 35, 55, 65, 75]
    doc["parameters"]["phi"] = list(range(0, 360, 20))
    swp = os.path.join(d, "sweep.yml")
    yaml.safe_dump(doc, open(swp, "w"))
    one, two = os.path.join(d, "one.npz"), os.path.join(d, "two.npz")
    cli = os.path.join(util.REPO, "shgyield.py")
    subprocess.run([sys.executable, cli, swp, "-o", one], check=True, timeout=300)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29537", cli, swp, "-o", two],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    a, b = np.load(one), np.load(two)
    assert a["pp"].shape == (5, 1, 18, 126)
    for k in ("pp", "ps", "sp", "ss", "theta", "phi", "energy"):
        assert np.array_equal(a[k], b[k]), k
