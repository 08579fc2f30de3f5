"""Is the train step bound by the host (Python + CUDA API calls) or by the device? Times N steps both ways."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200 import synthetic as S  # noqa: E402
from vqa_transfer_externaldata_b200.model import Model, make_synthetic_config  # noqa: E402

CFG1 = dict(B=512, K=36, Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=8192)
config, image_features, batch, _ = make_synthetic_config(CFG1, num_images=256)
model = Model(batch, config, is_train=True, image_features=image_features)
eng = model.engine
c = S.dims(**CFG1)
hb = [S.make_batch(c, 256, seed=1 + r) for r in range(2)]
pinned = [{k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in b.items()
           if k in ("image_idx", "q_intseq", "q_intseq_len", "answer_target")} for b in hb]
for _ in range(5):
    model.train_step(pinned[0], sync=True)
torch.cuda.synchronize()
N = 50
# (a) host time to ENQUEUE N steps with the device far behind? -> enqueue without any sync
t0 = time.perf_counter()
pend = None
for i in range(N):
    p, _, _ = model.train_step(pinned[i % 2], next_batch=pinned[(i + 1) % 2], sync=False)
    pend = p
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {N} steps: {(t1 - t0) / N * 1e3:.3f} ms/step host; until drained: {(t2 - t0) / N * 1e3:.3f} ms/step")
# (b) the same without prefetch / uploads (device-resident batch)
eng.stage_batch(pinned[0])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(N):
    eng.forward(seed=1, step=i, full_outputs=False)
    model.backward()
    eng.adam_step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"resident: enqueue {(t1 - t0) / N * 1e3:.3f} ms/step host; until drained: {(t2 - t0) / N * 1e3:.3f} ms/step")
