"""Throughput of BASELINE.json configs 2, 3 and 5 on one B200 (config 1 is bench.py's line; config 4 is the memft
pre-training graph, a 'next' row that is not built). Synthetic inputs of SURVEY 8d; bf16 mode; CUDA-event timing,
5 warm-up + 20 timed steps, inputs larger than L2 (device-resident feature bank indexed at random).

  cfg2  vqa/model_standard, B 512, K 36: train step (fwd + bwd + clip + Adam)      roofline 349 us / step
  cfg3  14 x 14 grid cells (K 196), B 256: train step                               roofline 599 us / step
  cfg5  inference sweep, K 100 padded boxes with 10..100 valid, forward only        roofline 0.59 us / sample

Writes one JSON object per line to stdout."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vqa_transfer_externaldata_b200 import synthetic as S  # noqa: E402
from vqa_transfer_externaldata_b200.engine import AnswerModelConfig, Engine  # noqa: E402


def make_engine(dims, variant, n_img, nbox_range=None, seed=0):
    c = S.dims(**dims)
    cfg = AnswerModelConfig(variant=variant, precision="bf16", **c)
    eng = Engine(cfg)
    g = torch.Generator(device=eng.device).manual_seed(99)
    bank = torch.randn(n_img, c["K"], c["Dv"], device=eng.device, generator=g).abs_().mul_(0.5)
    rng = np.random.default_rng(seed)
    if nbox_range is None:
        nb = np.full(n_img, c["K"], np.int32)
    else:
        nb = rng.integers(nbox_range[0], nbox_range[1] + 1, size=n_img).astype(np.int32)
        mask = torch.arange(c["K"], device=eng.device)[None, :] < torch.from_numpy(nb).to(eng.device)[:, None]
        bank *= mask[:, :, None]          # padded rows are zero in the adaptive 10-100 feature files
    eng.set_feature_bank(bank, nb)
    params, exist = S.init_params(c, seed=4321, variant=variant)
    is_obj, is_attr = S.make_answer_flags(c)
    eng.set_answer_masks(is_obj, is_attr, exist)
    eng.load_params(params)
    batches = [{k: torch.from_numpy(np.ascontiguousarray(v)).to(eng.device)
                for k, v in S.make_batch(c, n_img, seed=1234 + r).items()
                if k in ("image_idx", "q_intseq", "q_intseq_len", "answer_target")} for r in range(3)]
    return eng, c, batches


def timed(fn, steps=20, warmup=5):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def train_case(name, dims, variant, roof_us, n_img):
    eng, c, batches = make_engine(dims, variant, n_img)

    def step(i):
        eng.stage_batch(batches[i % len(batches)])
        eng.forward(seed=777, step=i, full_outputs=False)
        eng.backward()
        eng.adam_step(lr=1e-3, clip_norm=20.0)

    ms = timed(step)
    loss, _ = eng.read_scalars()
    print(json.dumps({"config": name, "mode": "train step fwd+bwd+clip+adam", "batch": c["B"], "K": c["K"],
                      "ms_per_step": ms, "samples_per_s": c["B"] / (ms * 1e-3), "roofline_us": roof_us,
                      "roofline_frac": roof_us * 1e-3 / ms, "loss_finite": bool(np.isfinite(loss))}), flush=True)
    eng.close()
    del eng
    torch.cuda.empty_cache()


def infer_case(B, n_img=2048):
    dims = dict(B=B, K=100, Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=8192)
    eng, c, batches = make_engine(dims, "vlmap_answer", n_img, nbox_range=(10, 100))

    def step(i):
        eng.stage_batch(batches[i % len(batches)])
        eng.forward(seed=777, step=i, full_outputs=True)

    ms = timed(step, steps=10, warmup=3)
    att = eng.outputs()["att_score"]
    ok = bool(torch.isfinite(att).all().item())
    print(json.dumps({"config": "cfg5", "mode": "inference forward (logits, attention, pred, report)", "batch": B,
                      "K": 100, "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "roofline_us_per_sample": 0.59,
                      "roofline_frac": 0.59e-3 * B / ms, "outputs_finite": ok}), flush=True)
    eng.close()
    del eng
    torch.cuda.empty_cache()


def main():
    torch.cuda.set_device(0)
    base = dict(Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=8192)
    train_case("cfg2", dict(B=512, K=36, **base), "standard", 349.0, 4096)
    train_case("cfg3", dict(B=256, K=196, **base), "vlmap_answer", 599.0, 1024)
    for B in (64, 256, 512, 1024, 2048, 4096, 8192):
        try:
            infer_case(B)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"config": "cfg5", "batch": B, "error": repr(e)[:300]}), flush=True)
            torch.cuda.empty_cache()
    return 0


if __name__ == "__main__":
    sys.exit(main())
