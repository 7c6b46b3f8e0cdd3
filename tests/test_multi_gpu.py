"""Fused peer gather on real GPUs (needs >= 2 devices on the box; skipped otherwise).  The host-side rank arithmetic
and the NCCL-style gather buffers are covered on CPU under gloo in tests/test_cabi_and_host.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_gather_is_bit_exact_across_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_peer_check.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0 and "peer gather ok" in out.stdout, out.stdout[-2000:]
