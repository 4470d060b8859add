"""Multi-GPU check of the hand-written NVLS all-reduce (csrc/collective.cu) against NCCL: needs >= 2 GPUs with
NVSwitch multicast on one node; skipped otherwise.  Runs tools/nvls_check.py under torchrun (bit-identical sums,
zero padding, no hang)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_nvls_allreduce_matches_nccl_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "tools", "nvls_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "trial 2: max abs diff vs NCCL 0.000e+00" in r.stdout or "rel L2" in r.stdout
