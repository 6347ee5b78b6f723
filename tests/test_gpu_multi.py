"""-m gpu, needs 2 GPUs (skipped otherwise): the product kernels under a real NCCL process group -- sharded == unsharded."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_nccl_sharded_equals_unsharded():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "nccl_shard_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(p.stdout[-3000:], p.stderr[-3000:])
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
    assert "NCCL_SHARD_CHECK" in p.stdout
