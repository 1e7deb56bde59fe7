"""Multi-GPU parity: the distributed pipeline on 2 ranks (NCCL) against the single-GPU pipeline.
Skipped unless at least two GPUs are visible (the single-rank form of the same code path is
covered by tests/test_gpu_kernels.py::test_distributed_driver_single_rank_matches_oracle and the
schedule itself by the gloo tests in tests/test_dist_cpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "dist_check.py"), "2600", "1900", "80"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert "-> OK" in res.stdout
