"""Host-side throughput of the input pipeline: pure-Python mirror (input_ops.create) against the native
parser (input_native.create) on the same synthetic cfg1-shaped shards, first pass (parse) and cached replay.
   python scripts/input_pipeline_bench.py [n_samples]        (no GPU needed)"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200 import input_native as IN  # noqa: E402
from vqa_transfer_externaldata_b200 import input_ops as IO  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20480
A, B = 3000, 512
rng = np.random.default_rng(0)
samples = []
for i in range(n):
    t = int(rng.integers(3, 15))
    k = int(rng.integers(1, 4))
    samples.append({"qid": i, "image_id": b"COCO_train2014_%012d" % i, "image_idx": int(rng.integers(0, 4096)),
                    "q_intseq": rng.integers(1, 8192, size=t).tolist(), "answer_ids": rng.choice(A, size=k, replace=False).tolist(),
                    "answer_scores": rng.choice([0.3, 0.6, 0.9, 1.0], size=k).astype(np.float32).tolist()})
with tempfile.TemporaryDirectory() as d:
    IO.write_shards(d, "train", samples, A, num_shards=16)
    size = sum(os.path.getsize(os.path.join(d, "train", f)) for f in os.listdir(os.path.join(d, "train")))
    print(f"{n} samples, {size / n:.0f} bytes per record, batch {B}, host threads used: 1")
    for name, mod, kw in (("input_ops.create (pure Python)", IO, {}),
                          ("input_native.create (C parser, sparse targets)", IN, {"want_image_id": False}),
                          ("input_native.create (C parser, sparse targets + image_id strings)", IN, {}),
                          ("input_native.create (C parser, dense targets on the host)", IN, {"dense_target": True, "want_image_id": False})):
        t0 = time.perf_counter()
        it = mod.create(B, d, "train", is_train=True, seed=1, epochs=2, **kw)
        t1 = time.perf_counter()
        nb = (n + B - 1) // B
        for _ in range(nb):
            next(it)
        t2 = time.perf_counter()
        for _ in range(nb):
            next(it)
        t3 = time.perf_counter()
        print(f"{name}: open + index {t1 - t0:.3f} s; first pass {n / (t2 - t1):,.0f} samples/s "
              f"({(t2 - t1) / nb * 1e3:.2f} ms per batch); cached replay {n / max(t3 - t2, 1e-9):,.0f} samples/s")
