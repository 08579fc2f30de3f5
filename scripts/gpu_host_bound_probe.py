"""Is the train step bound by host enqueue time? Wall time of the enqueue loop (no synchronisation inside) against
the device time of the same steps, plus a per-call break-down of the host side."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200 import synthetic as S  # noqa: E402
from vqa_transfer_externaldata_b200.model import Model, make_synthetic_config  # noqa: E402

CFG1 = dict(B=512, K=36, Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=8192)
dev = torch.device("cuda:0")
c = S.dims(**CFG1)
n_img = 2048
bank = torch.randn(n_img, c["K"], c["Dv"], device=dev).abs_().mul_(0.5)
config, _, _, _ = make_synthetic_config(CFG1, precision="bf16", seed=4321, num_images=2)
feats = {"features": bank, "num_boxes": np.full(n_img, c["K"], np.int32), "max_box_num": c["K"], "vfeat_dim": c["Dv"]}
config.device = dev
hb = [S.make_batch(c, n_img, seed=1 + r) for r in range(4)]
db = [{k: torch.from_numpy(np.ascontiguousarray(b[k])).to(dev) for k in ("image_idx", "q_intseq", "q_intseq_len", "answer_target")} for b in hb]
model = Model(hb[0], config, is_train=True, image_features=feats)
eng = model.engine
T = {"stage": 0.0, "forward": 0.0, "prefetch": 0.0, "backward": 0.0, "adam": 0.0}


def step(i, acc=None):
    t0 = time.perf_counter()
    eng.stage_batch(db[i % 4])
    t1 = time.perf_counter()
    eng.forward(seed=1, step=i, full_outputs=False, defer_outputs=True)
    t2 = time.perf_counter()
    eng.prefetch_batch(db[(i + 1) % 4])
    t3 = time.perf_counter()
    eng.backward()
    t4 = time.perf_counter()
    eng.adam_step()
    t5 = time.perf_counter()
    if acc is not None:
        for k, v in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            acc[k] += v


for i in range(10):
    step(i)
torch.cuda.synchronize()
N = 100
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0 = time.perf_counter()
e0.record()
for i in range(N):
    step(i, T)
e1.record()
w1 = time.perf_counter()
torch.cuda.synchronize()
w2 = time.perf_counter()
print(f"host enqueue {1e3 * (w1 - w0) / N:.3f} ms/step; device {e0.elapsed_time(e1) / N:.3f} ms/step; "
      f"wall incl. drain {1e3 * (w2 - w0) / N:.3f} ms/step; launches/step {eng.launch_count() // 110}")
print("host per call (ms):", {k: round(1e3 * v / N, 3) for k, v in T.items()})

# ---- the end-to-end loop of bench.py (pinned host batches, next-batch prefetch, loss read back two steps later) ----
pinned = [{k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in b.items()
           if k in ("image_idx", "q_intseq", "q_intseq_len", "answer_target")} for b in hb]
pending = []
for i in range(8):
    p, _, _ = model.train_step(pinned[i % 4], next_batch=pinned[(i + 1) % 4], sync=False)
    pending.append(p)
for p in pending:
    p.get()
torch.cuda.synchronize()
pending = []
t_step = t_get = 0.0
e0.record()
w0 = time.perf_counter()
for i in range(N):
    a = time.perf_counter()
    p, _, _ = model.train_step(pinned[i % 4], next_batch=pinned[(i + 1) % 4], sync=False)
    b = time.perf_counter()
    pending.append(p)
    if len(pending) > 2:
        pending.pop(0).get()
    c2 = time.perf_counter()
    t_step += b - a
    t_get += c2 - b
e1.record()
w1 = time.perf_counter()
torch.cuda.synchronize()
print(f"e2e loop: host {1e3 * (w1 - w0) / N:.3f} ms/step (train_step call {1e3 * t_step / N:.3f}, loss get {1e3 * t_get / N:.3f}); "
      f"device {e0.elapsed_time(e1) / N:.3f} ms/step")


def loop(name, batches, nxt=True, get=True, sync_each=False):
    pend = []
    for i in range(6):
        p, _, _ = model.train_step(batches[i % 4], next_batch=batches[(i + 1) % 4] if nxt else None, sync=False)
    torch.cuda.synchronize()
    e0.record()
    for i in range(N):
        p, _, _ = model.train_step(batches[i % 4], next_batch=batches[(i + 1) % 4] if nxt else None, sync=False)
        pend.append(p)
        if get and len(pend) > 2:
            pend.pop(0).get()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: device {e0.elapsed_time(e1) / N:.3f} ms/step")


loop("pinned + prefetch + get", pinned)
loop("pinned + prefetch, no get", pinned, get=False)
loop("pinned, no prefetch, no get", pinned, nxt=False, get=False)
loop("device batches + prefetch + get", db)
loop("device batches + prefetch, no get", db, get=False)
loop("device batches, no prefetch, no get", db, nxt=False, get=False)
