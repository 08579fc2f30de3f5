"""BASELINE config 4 (vlmap_memft bf_or_wordset_withatt_sp pre-training, bs 512, 5 + 5 entries per image, K 36, A 4000,
blanks <= 10 tokens) on one B200: train-step time with CUDA events, per-section split.
   python scripts/gpu_bench_memft.py [steps] [precision]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200 import memft as F  # noqa: E402

CFG4 = dict(B=512, K=36, n=5, Dv=2048, D=1024, L=1024, W=300, A=4000, T=10, Vq=8192, Nws=2000)


def synthetic(dims, seed=0):
    rng = np.random.default_rng(seed)
    B, K, n, T = dims["B"], dims["K"], dims["n"], dims["T"]
    batch = {"image_ft": (np.abs(rng.standard_normal((B, K, dims["Dv"]), dtype=np.float32)) * 0.5),
             "spatial_ft": rng.uniform(size=(B, K, 6)).astype(np.float32),
             "num_boxes": rng.integers(10, K + 1, size=B).astype(np.int32)}
    for kind in ("obj", "attr"):
        x0, y0 = rng.uniform(0, 0.5, size=(B, n)), rng.uniform(0, 0.5, size=(B, n))
        boxes = np.stack([x0, y0, x0 + rng.uniform(0.1, 0.5, size=(B, n)), y0 + rng.uniform(0.1, 0.5, size=(B, n))], axis=-1)
        ln = rng.integers(1, T + 1, size=(B, n)).astype(np.int32)
        blanks = rng.integers(1, dims["Vq"], size=(B, n, T)).astype(np.int32)
        blanks[np.arange(T)[None, None, :] >= ln[:, :, None]] = 0
        batch.update({f"{kind}_blank_fill/normal_boxes": boxes.astype(np.float32), f"{kind}_blank_fill/blanks": blanks,
                      f"{kind}_blank_fill/blanks_len": ln, f"{kind}_blank_fill/fills": rng.integers(0, dims["A"], size=(B, n)).astype(np.int32),
                      f"{kind}_blank_fill/num": rng.integers(1, n + 1, size=B).astype(np.int32),
                      f"{kind}_blank_fill/wordsets": rng.integers(0, dims["Nws"], size=(B, n)).astype(np.int32)})
    return batch


def init_params(cfg, seed=1):
    rng = np.random.default_rng(seed)
    p = {}
    for k, (shp, _) in F.FIELDS.items():
        s = shp(cfg)
        if k in ("wordset_map", "l_glove"):
            p[k] = rng.standard_normal(s, dtype=np.float32) * 0.4
        elif k.endswith("_gamma"):
            p[k] = np.ones(s, np.float32)
        elif k == "gru_gates_b":
            p[k] = np.ones(s, np.float32)
        elif len(s) == 2:
            lim = np.sqrt(6.0 / (s[0] + s[1]))
            p[k] = rng.uniform(-lim, lim, size=s).astype(np.float32)
        else:
            p[k] = np.zeros(s, np.float32)
    return p


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    cfg = F.make_config(CFG4, precision=precision)
    batch = synthetic(CFG4)
    model = F.Model(batch, cfg, is_train=True, params=init_params(cfg))
    for _ in range(3):
        model.train_step(sync=False)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    sec = np.zeros(4)
    t0 = time.perf_counter()
    for _ in range(steps):
        ev[0].record()
        model.forward(with_grad_seed=True)
        ev[1].record()
        model.backward()
        ev[2].record()
        model.adam_step()
        ev[3].record()
        model.global_step += 1
        torch.cuda.synchronize()
        sec += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3])]
    wall = (time.perf_counter() - t0) / steps * 1e3
    sec /= steps
    # the same without per-section synchronisation
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        model.train_step(sync=False)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    loss, report = model.fetch()
    E, R = 2 * CFG4["B"] * CFG4["n"], 4 * CFG4["B"] * CFG4["n"]
    c = cfg
    fwd = 2.0 * (E * c.Dv * c.L + E * c.W * c.L + R * c.L * c.L + R * c.L * 2 * c.L + R * 2 * c.L * c.A +
                 E * c.T * (c.W + c.L) * 3 * c.L)
    print(json.dumps({"config": "cfg4 vlmap_memft bf_or_wordset_withatt_sp train step", "precision": precision, "dims": CFG4,
                      "ms_per_step": ms, "images_per_s": CFG4["B"] / ms * 1e3, "entries_per_s": E / ms * 1e3,
                      "section_ms": {"forward": sec[0], "backward": sec[1], "clip_adam_refresh": sec[2], "sum": sec[3]},
                      "host_wall_ms_per_step_synced": wall, "gemm_tflop_per_step": 3 * fwd / 1e12,
                      "tflops": 3 * fwd / 1e9 / ms, "loss": loss, "finite": bool(np.isfinite(loss))}))


if __name__ == "__main__":
    main()
