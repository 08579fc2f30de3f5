"""Per-phase time line of the two attention kernels at cfg1 (globaltimer stamps written by the kernels themselves).
   python scripts/gpu_attn_trace.py  -> medians over CTAs / samples, microseconds."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200.model import Model, make_synthetic_config  # noqa: E402

CFG1 = dict(B=512, K=36, Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=8192)
config, image_features, batch, _ = make_synthetic_config(CFG1, num_images=64)
model = Model(batch, config, is_train=True, image_features=image_features)
eng = model.engine
trace = torch.zeros(2 << 17, dtype=torch.int64, device="cuda")
for _ in range(3):
    model.train_step(batch)
eng.lib.vqa_internal_set_gru_trace(C.c_void_p(trace.data_ptr()))
model.train_step(batch)
torch.cuda.synchronize()
eng.lib.vqa_internal_set_gru_trace(C.c_void_p(0))
t = trace.cpu().numpy().astype(np.float64)
fw = t[65536:65536 + 148 * 8 * 8].reshape(148, 8, 8)
names = ["start->slab ready", "stats+merge+coeffs", "scores", "softmax", "pooling loop", "pool combine+write"]
print("== attention forward (persistent, 148 CTAs): per-sample phase medians, us")
valid = fw[:, :, 6] > 0
t0 = fw[valid][:, 0].min()
print(f"   kernel span {(fw[valid][:, 6].max() - t0) / 1e3:.1f} us; samples per CTA: {valid.sum(1).min()}..{valid.sum(1).max()}")
for s in range(4):
    v = valid[:, s]
    if not v.any():
        continue
    row = "  ".join(f"{n} {np.median(fw[v, s, i + 1] - fw[v, s, i]) / 1e3:5.2f}" for i, n in enumerate(names))
    print(f"   sample {s}: {row}   | total {np.median(fw[v, s, 6] - fw[v, s, 0]) / 1e3:5.2f}")
bw = t[98304:98304 + 512 * 8].reshape(512, 8)
names = ["dP->smem", "da = <V, dP>", "ds", "(slab wait)", "column pass (gates, T/U/V)", "combine + stats", "dz pass + store"]
print("== attention backward (one CTA per sample, 2 per SM): phase medians, us")
t0 = bw[:, 0].min()
print(f"   kernel span {(bw[:, 7].max() - t0) / 1e3:.1f} us; CTA life median {np.median(bw[:, 7] - bw[:, 0]) / 1e3:.2f} us")
print("   " + "  ".join(f"{n} {np.median(bw[:, i + 1] - bw[:, i]) / 1e3:5.2f}" for i, n in enumerate(names)))
