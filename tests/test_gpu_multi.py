"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches tests/multi_gpu_check.py under
torchrun, one rank per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_operator_two_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
