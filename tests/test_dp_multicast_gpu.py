"""The in-switch gradient all-reduce (csrc/collective.cu: multimem.ld_reduce / multimem.st over NVSwitch multicast
memory) against NCCL, launched as a 2-rank job. Needs two GPUs of one NVSwitch domain: skipped otherwise (the
1-GPU test box); scripts/gpu_multicast_check.py is the same check for any N under torchrun."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multimem_all_reduce_matches_nccl_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "gpu_multicast_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("bit-identical to rank 0: True") == 2
