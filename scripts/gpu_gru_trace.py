"""Per-phase time line of the persistent GRU kernels (globaltimer stamps written by the kernel itself).
   python scripts/gpu_gru_trace.py   -> prints, per phase, medians over CTAs of:
   wait->first A tile, first tile->accumulator complete, accumulator->arrival, arrival->next phase start"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200 import synthetic as S  # noqa: E402
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
T = CFG1["T"]
WAVES = 2 if os.environ.get("VQA_GRU_WAVES") == "2" else 1
ncta = 128 // WAVES
for mode, name in ((0, "forward"), (1, "BPTT")):
    full = trace[mode << 17:(mode << 17) + ncta * WAVES * 2 * T * 8].cpu().numpy().reshape(ncta, WAVES, 2 * T, 8).astype(np.float64)
    t0 = full[full > 0].min()
    for wv in range(WAVES):
        tr = full[:, wv]
        print(f"== {name} wave {wv}: total {(full.max() - t0) / 1e3:.1f} us")
        print("phase  start(us)  wait->tile0  tile0->acc  acc->arrive  (medians over CTAs, us)   spread of arrive")
        for p in range(2 * T):
            a = tr[:, p, :]
            if a[:, 3].max() == 0:
                continue
            st = np.median(a[:, 0][a[:, 0] > 0]) if (a[:, 0] > 0).any() else np.nan

            def f(x, y):
                m = (a[:, x] > 0) & (a[:, y] > 0)
                return np.median((a[:, y] - a[:, x])[m]) / 1e3 if m.any() else float("nan")
            print(f"{p:4d}  {(st - t0) / 1e3:9.2f}  {f(0, 1):10.2f}  {f(1, 2):10.2f}  {f(2, 3):10.2f}      "
                  f"{(a[:, 3].max() - a[:, 3].min()) / 1e3:6.2f}   arrive@{(np.median(a[:, 3]) - t0) / 1e3:8.2f}")
