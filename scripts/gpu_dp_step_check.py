"""torchrun --nproc-per-node N scripts/gpu_dp_step_check.py: the data-parallel train step on N GPUs (own in-switch
all-reduce, or NCCL when multicast is unavailable) == ONE single-rank step on the concatenated batch, gradients
bit-identical across ranks (DataParallel.self_check). fp32 mode to 1e-5, bf16 mode to 5e-5 (max-norm relative)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200.dp import DataParallel  # noqa: E402

dp = DataParallel()
torch.cuda.set_device(dp.local_rank)
res = dp.self_check()
if dp.rank == 0:
    print(f"dp step check over {dp.world_size} ranks: {res}", flush=True)
print(f"rank {dp.rank}: dp step == single-rank step: True", flush=True)
dp.close()
