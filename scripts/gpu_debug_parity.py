"""Debug aid (GPU box): per-tensor parity errors of the CUDA path and of an fp32 torch-CPU run, both
against the fp64 oracle, to separate kernel bugs from fp32 conditioning."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import answer_model_np as O
from oracle import answer_model_torch as OT
from parity_util import build_case, rel_err, rel_l2, run_both

MID = dict(B=48, K=36, Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=512)
SMALL = dict(B=16, K=12, Dv=256, D=128, L=128, A=200, T=6, W=20, Vq=50)

for name, dims, kw in (("SMALL", SMALL, {}), ("MID", MID, dict(num_images=40, batch=40, T=11)),
                       ("MID-noragged", MID, dict(num_images=40, ragged=False))):
    for prec in ("fp32", "bf16"):
        case = build_case(dims, precision=prec, seed=3, **kw)
        got, ref, ref_g = run_both(case)
        print(f"== {name} {prec}: loss {got['loss']:.6f} ref {ref['loss']:.6f}")
        print("   logit", f"{rel_err(got['logit'], ref['logit']):.2e}", "att", f"{rel_err(got['att_score'], ref['att_score']):.2e}",
              "pooled", f"{rel_err(got['pooled'], ref['pooled']):.2e}", "cond", f"{rel_err(got['condition'], ref['condition']):.2e}")
        print("   grads", {k: f"{rel_err(v, ref_g[k]):.1e}" for k, v in got["grads"].items()})
        print("   gr-l2", {k: f"{rel_l2(v, ref_g[k]):.1e}" for k, v in got["grads"].items()})
        if prec == "fp32":
            # a genuine fp32 implementation (torch CPU, fp32 GEMMs) against the same fp64 oracle
            eng = case["eng"]
            am, jm = eng.dropout_masks(777, 3)
            tp = {k: torch.tensor(v, dtype=torch.float32, requires_grad=True) for k, v in case["params"].items()}
            tb = {k: torch.tensor(v) for k, v in case["batch"].items()}
            to = OT.forward(tp, torch.tensor(case["feats"]), torch.tensor(case["nb"]), tb,
                            torch.tensor(case["m"]["train"], dtype=torch.float32),
                            att_mask=am.cpu().float(), joint_mask=jm.cpu().float())
            to["loss"].backward()
            print("   torch-fp32 vs fp64 oracle: logit", f"{rel_err(to['logit'].detach().numpy(), ref['logit']):.2e}")
            print("   torch-fp32 grads", {k: f"{rel_err(tp[k].grad.numpy(), ref_g[k]):.1e}" for k in got["grads"]})
        del case
        torch.cuda.empty_cache()
