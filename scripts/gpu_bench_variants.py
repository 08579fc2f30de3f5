"""Train-step time of every model_type at cfg1 shapes (B 512, K 36, Dv 2048, A 3000, T 14) on one B200, bf16 mode.
The ent variant runs NUM_MARGINAL = 200 tiles per sample (102 400 rows through joint_fc and the answer head)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

from gpu_bench_configs import make_engine, timed  # noqa: E402
from vqa_transfer_externaldata_b200 import importer  # noqa: E402


def main():
    torch.cuda.set_device(0)
    dims = dict(B=512, K=36, Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=8192)
    for variant in importer.get_model_types():
        eng, c, batches = make_engine(dims, variant, 2048)

        def step(i):
            eng.stage_batch(batches[i % len(batches)])
            eng.forward(seed=777, step=i, full_outputs=False)
            eng.backward()
            eng.adam_step(lr=1e-3, clip_norm=20.0)

        ms = timed(step, steps=10, warmup=3)
        loss, _ = eng.read_scalars()
        print(json.dumps({"model_type": variant, "ms_per_step": ms, "samples_per_s": c["B"] / (ms * 1e-3),
                          "workspace_GB": eng.workspace.numel() / 1e9, "loss_finite": bool(np.isfinite(loss))}), flush=True)
        eng.close()
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
