"""torchrun --nproc-per-node N scripts/gpu_memft_dp_check.py: the data-parallel gradient of the pre-training model
(memft.Model: per-rank forward / backward on its shard of one global batch, loss normalised by the GLOBAL valid-entry
counts, flat gradient all-reduced by NCCL) == the gradient of ONE single-rank step on the concatenated batch.
Dropout is switched off for the comparison (the ranks draw per-rank streams); fp32 mode, 2e-5 max-norm relative."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200 import memft as F  # noqa: E402
from vqa_transfer_externaldata_b200.dp import DataParallel  # noqa: E402

dp = DataParallel()
torch.cuda.set_device(dp.local_rank)
rank, world = dp.rank, dp.world_size
per = dict(B=6, K=12, n=5, Dv=64, D=128, L=128, W=24, A=200, T=5, Vq=50, Nws=20)
glob = dict(per, B=per["B"] * world)
gb = F.synthetic_batch(glob, seed=11)
cfg_g = F.make_config(glob, precision="fp32", keep_att=1.0, keep_joint=1.0)
params = F.xavier_params(cfg_g, seed=3)
sl = slice(rank * per["B"], (rank + 1) * per["B"])
shard = {k: v[sl] for k, v in gb.items()}
model = F.Model(shard, F.make_config(per, precision="fp32", keep_att=1.0, keep_joint=1.0), is_train=True, params=params)
model.train_step(apply_optimizer=False)          # forward + backward + all-reduce
g_dp = model.gradients()
model._dist = None                                # the single-rank reference runs without the collective
ref = F.Model(None, cfg_g, is_train=True, params=params)
ref._dist, ref.global_counts = None, None
ref.set_batch(gb)
ref.train_step(apply_optimizer=False)
g_ref = ref.gradients()
worst = 0.0
for k in g_ref:
    scale = max(np.abs(g_ref[k]).max(), 1e-20)
    worst = max(worst, float(np.abs(g_dp[k] - g_ref[k]).max() / scale))
ok = worst < 2e-5
if rank == 0:
    print(f"memft dp gradient check over {world} ranks: worst relative deviation {worst:.2e}", flush=True)
print(f"rank {rank}: memft dp step == single-rank step: {ok}", flush=True)
model.close()
ref.close()
dp.close()
sys.exit(0 if ok else 1)
