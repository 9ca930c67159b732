"""GPU, 2 ranks over NCCL (skipped with fewer than 2 devices): the sharded tile ranges + one
all-reduce reproduce the single-GPU results - edges and counts bit for bit, sums to 1e-13,
kriging slices identical (tools/mgpu_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_equal_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MGPU_CHECK PASS" in res.stdout
