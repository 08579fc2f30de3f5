"""Hardware test of the data-parallel step: a 2-rank job, each rank stepping on its half of one global
batch with the gradients all-reduced by the path training uses (csrc/collective.cu's multimem kernel when the box has
NVSwitch multicast, NCCL otherwise), must produce the gradients of ONE single-rank step on the concatenated batch,
bit-identical on both ranks. Needs two GPUs: skipped on the 1-GPU test box; `bench.py --gpus N` (N > 1) runs the same
check (DataParallel.self_check) before it times anything, so the driver's scaling run exercises it too."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dp_step_equals_single_rank_step_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "scripts", "gpu_dp_step_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("dp step == single-rank step: True") == 2


def test_memft_dp_step_equals_single_rank_step_two_ranks():
    """The same for the pre-training model (BASELINE config 4 is quoted on 8 x B200 DP)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29539", os.path.join(ROOT, "scripts", "gpu_memft_dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("memft dp step == single-rank step: True") == 2
