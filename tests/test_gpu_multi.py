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


def _run_sweep(cmd_prefix):
    import json
    cmd = cmd_prefix + [os.path.join(ROOT, "scripts", "sweep_couplings.py"), "--J", "0.02,0.5,0.2,0.1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    return json.loads([line for line in out.stdout.splitlines() if line.startswith("{")][-1])


def test_coupling_sweep_single_gpu():
    """One System per grid point against the free-fermion closed form (reference scripts/computeTIinfinite.py:6-12)."""
    line = _run_sweep([sys.executable])
    assert line["n_gpus"] == 1 and len(line["points"]) == 4
    assert [p["J"] for p in line["points"]] == [0.02, 0.5, 0.2, 0.1]
    assert line["worst_error_vs_closed_form"] < 1e-7
    assert line["sweep_iterations_total"] > 0


def test_coupling_sweep_one_system_per_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    line = _run_sweep([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                       "--master-addr", "127.0.0.1", "--master-port", "29633"])
    assert line["n_gpus"] == 2 and len(line["seconds_per_rank"]) == 2
    assert [p["J"] for p in line["points"]] == [0.02, 0.5, 0.2, 0.1]
    assert line["worst_error_vs_closed_form"] < 1e-7
