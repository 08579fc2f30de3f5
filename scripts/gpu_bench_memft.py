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

CFG4 = F.CFG4
synthetic = F.synthetic_batch
init_params = F.xavier_params


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
