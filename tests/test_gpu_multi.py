"""Multi-GPU parity on real devices (skipped on a single-GPU box; the host-side sharding logic and the algebra
of the exchange are covered on CPU by tests/test_multi_gpu_host.py): a 2-rank solve with motion priors, the
free interFrameRatio, GoodPosePrior blocks and free intrinsics must reproduce the single-GPU solve."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpu_solve_with_camera_only_blocks_matches_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("  OK") == 2 and "MISMATCH" not in r.stdout
