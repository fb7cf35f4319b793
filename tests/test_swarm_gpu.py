"""NCCL multi-robot path on real GPUs (needs >= 2 devices; skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_swarm_step_over_nccl_matches_per_robot_pools():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tools", "check_swarm.py")],
                         capture_output=True, text=True, timeout=600)
    assert "SWARM_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
