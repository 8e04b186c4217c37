"""Multi-GPU parity (needs >= 2 CUDA devices): contiguous index-range sharding with one NCCL
collective per sweep must reproduce the single-process oracle bit for bit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_sharded_sweeps_match_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import __graft_entry__
    __graft_entry__.build()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", "29617",
           os.path.join(ROOT, "tests", "_dist_worker.py")]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          timeout=900)
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert "dist worker ok" in proc.stdout
