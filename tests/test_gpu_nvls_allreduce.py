"""Multi-GPU checks (need >= 2 GPUs on one node; skipped otherwise), run under torchrun:

  * tools/nvls_check.py      the hand-written NVLS all-reduce (csrc/collective.cu) against NCCL: the NVLS path must
                             actually be taken on an NVSwitch node, sums equal NCCL's, padding stays zero, no hang
  * tools/multiview_check.py SURVEY.md section 4 item 4: multiview_step with the REAL renderer, direct gradient sink
                             and the all-reduce on 2 ranks x 4 views against all 8 views rendered on one GPU
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torchrun(script, port, nproc=2, timeout=400):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


@pytest.mark.timeout(300)
def test_nvls_allreduce_matches_nccl_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    r = torchrun("nvls_check.py", 29571, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    if "no multicast address" in r.stdout:
        pytest.skip("this node has no NVSwitch multicast: " + r.stdout.splitlines()[0])
    assert "NVLS path: True" in r.stdout, r.stdout[-2000:]            # the hand-written collective really ran
    for trial in range(3):
        assert "trial %d: max abs diff vs NCCL 0.000e+00" % trial in r.stdout, r.stdout[-2000:]


@pytest.mark.timeout(500)
def test_multiview_step_two_ranks_equals_one_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    r = torchrun("multiview_check.py", 29573)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multiview_check ok" in r.stdout, r.stdout[-3000:]
