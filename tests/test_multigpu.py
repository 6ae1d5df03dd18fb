"""Needs >= 2 GPUs (skipped otherwise): bands rendered by different GPUs, assembled on
rank 0 by peer stores and by an NCCL gather, equal the single-GPU frame bit for bit."""
import os
import subprocess
import sys

import pytest

import helpers as h

pytestmark = pytest.mark.gpu


def test_two_gpu_bands_equal_single_gpu_frame():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(h.ROOT, "tests", "multigpu_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTIGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
