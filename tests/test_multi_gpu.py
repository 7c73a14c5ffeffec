"""The NCCL / NVLink side of the data-parallel path on a box with at least two GPUs: runs tests/multi_gpu_check.py under
torchrun (one process per GPU).  Skipped on single-GPU boxes; the partitioning logic itself is covered on CPU with gloo
(tests/test_sharding.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_sharded_run_equals_unsharded_run_nccl_and_peer_totals():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "multi-GPU check ok" in res.stdout
