"""N-GPU == 1-GPU bit for bit.  Needs >= 2 visible GPUs; spawns
torchrun with two ranks (NCCL).  Skipped on single-GPU boxes — the CPU twin of this
invariant is tests/test_dist_cpu.py."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_sharded_run_is_bit_identical_on_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(ROOT / "scripts" / "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "FAIL" not in res.stdout
