"""Multi-GPU checks: need >= 2 visible GPUs (skipped on a single-GPU box).  Launches tools/dist_checks.py under torchrun:
SyncBatchNorm over NVLink peer memory and over NCCL vs an fp64 BatchNorm of the concatenated rows, DDP batch-dice gradients,
and all-reduced data-parallel gradients vs a single process on the batch of N patches."""
import os
import subprocess
import sys

import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_multi_gpu_checks_under_torchrun():
    n = min(torch.cuda.device_count(), 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(H.ROOT, "tools", "dist_checks.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    tail = "\n".join((p.stdout + p.stderr).splitlines()[-40:])
    assert p.returncode == 0 and "ALL CHECKS PASSED" in p.stdout, tail
