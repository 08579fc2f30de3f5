import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import answer_model_np as O
from parity_util import build_case, rel_err, run_both

base = dict(B=16, K=12, Dv=256, D=128, L=128, A=200, T=6, W=20, Vq=50)
def run(tag, **over):
    kw = {k: over.pop(k) for k in list(over) if k in ("num_images", "batch", "T_", "ragged")}
    d = dict(base); d.update(over)
    case = build_case(d, precision="fp32", seed=3, **kw)
    got, ref, ref_g = run_both(case)
    nbox = case["nb"][case["batch"]["image_idx"]]
    print(tag, {k: f"{rel_err(got['grads'][k], ref_g[k]):.1e}" for k in ("v_w", "v_b", "v_gamma", "v_beta", "att_w", "qv_w")},
          "nbox", sorted(nbox.tolist())[:6], flush=True)
    return case, got, ref, ref_g

run("base        ")
run("K36         ", K=36)
run("D1024       ", D=1024)
run("K36 D1024   ", K=36, D=1024)
run("K36 D256    ", K=36, D=256)
run("K36 D512    ", K=36, D=512)
run("K36 D2048   ", K=36, D=2048)
run("K40 D1024   ", K=40, D=1024)
run("K32 D1024   ", K=32, D=1024)
run("K36 D1024 nr", K=36, D=1024, ragged=False)
# per-sample analysis on the failing config using the attention kernel alone
case, got, ref, ref_g = run("K36 D1024 B48", K=36, D=1024, B=48)
