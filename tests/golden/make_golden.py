"""Generates tests/golden/answer_model_small.npz: outputs of the fp64 oracle on seeded synthetic inputs.

The reference cannot run here (Python 2 + tensorflow-gpu 1.6), so these vectors pin the ORACLE
(against silent edits) and give the GPU tests a committed target; they are not reference outputs.
Inputs are regenerated from seeds by vqa_transfer_externaldata_b200.synthetic; dropout masks come from the
Philox restatement (oracle/philox_np.py) with (seed, step) below, i.e. the bits the CUDA path draws.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import answer_model_np as O  # noqa: E402
from oracle import philox_np as PH  # noqa: E402
from vqa_transfer_externaldata_b200 import synthetic as S  # noqa: E402

DIMS = dict(B=16, K=12, Dv=256, D=128, L=128, A=200, T=6, W=20, Vq=50)
SEED, STEP = 777, 3
CASES = {"vlmap_answer": 11, "standard": 12}   # variant -> data seed


def inputs(variant, seed):
    c = S.dims(**DIMS)
    params, exist = S.init_params(c, seed=seed, variant=variant, perturb=0.2)
    feats, nb = S.make_bank(c, num_images=24, seed=seed + 2, ragged_boxes=True)
    batch = S.make_batch(c, 24, seed=seed + 3)
    is_obj, is_attr = S.make_answer_flags(c)
    m = O.answer_masks(c["A"], c["num_train_answer"], is_obj, is_attr, exist)
    am = PH.keep_mask(c["B"] * c["K"] * c["D"], 0.8, SEED, STEP, PH.SITE_ATT).reshape(c["B"], c["K"], c["D"])
    jm = PH.keep_mask(c["B"] * c["J"], 0.5, SEED, STEP, PH.SITE_JOINT).reshape(c["B"], c["J"])
    return c, params, exist, feats, nb, batch, (is_obj, is_attr), m, am, jm


# the other members of the family (tests/golden/answer_model_variants.npz): data seed per model_type
VARIANT_CASES = {"vlmap_answer2": 21, "vlmap_answer_no_noise": 22, "vlmap_answer_noc": 23, "vlmap_answer_full": 24,
                 "vlmap_answer_vqa_all": 25, "vlmap_answer_vqa_all2": 26, "vlmap_answer_adapt": 27, "vlmap_answer_ent": 28}
NUM_MARGINAL = 5        # of the ent case (the reference's 200 would make the fixture's oracle run needlessly large)


def extras(variant, c):
    """Variant-specific random draws, from the same Philox restatement the device uses (sites 3 / 4 / 5)."""
    if variant in ("vlmap_answer_noc", "vlmap_answer_nocarch"):
        return {"joint_l_mask": PH.keep_mask(c["B"] * c["J"], 0.5, SEED, STEP, 3).reshape(c["B"], c["J"])}
    if variant == "vlmap_answer_full":
        return {"noise": PH.normal_draw(c["B"] * c["L"], SEED, STEP).reshape(c["B"], c["L"])}
    if variant == "vlmap_answer_ent":
        return {"num_marginal": NUM_MARGINAL,
                "ent_mask": PH.keep_mask(c["B"] * NUM_MARGINAL * c["J"], 0.5, SEED, STEP, 5).reshape(c["B"], NUM_MARGINAL, c["J"])}
    return {}


def run(variant, seed, operand_round=None):
    c, params, exist, feats, nb, batch, flags, m, am, jm = inputs(variant, seed)
    out, cache = O.forward(params, feats, nb, batch, m, variant=variant, att_mask=am, joint_mask=jm,
                           operand_round=operand_round, **extras(variant, c))
    g = O.backward(cache)
    return out, {f: g[f] for f in O.trainable_fields(variant)}


def summarise(out, g):
    """Compact fixture of one case: forward tensors in full (they are small), gradients as their l2 norm plus their
    first 48 entries."""
    blob = {"loss": np.float64(out["loss"]), "logit": out["logit"].astype(np.float32), "att_score": out["att_score"],
            "pred": out["pred"], "condition": out["condition"].astype(np.float32),
            "report_keys": np.array(sorted(out["report"])),
            "report": np.array([out["report"][k] for k in sorted(out["report"])])}
    for f, v in g.items():
        blob[f"grad_norm/{f}"] = np.float64(np.linalg.norm(v))
        blob[f"grad_head/{f}"] = np.asarray(v, np.float64).reshape(-1)[:48]
    return blob


def main():
    blob = {}
    for variant, seed in CASES.items():
        out, g = run(variant, seed)
        blob[f"{variant}/loss"] = np.float64(out["loss"])
        blob[f"{variant}/logit"] = out["logit"]
        blob[f"{variant}/att_score"] = out["att_score"]
        blob[f"{variant}/pred"] = out["pred"]
        blob[f"{variant}/condition"] = out["condition"]
        blob[f"{variant}/report"] = np.array([out["report"][k] for k in sorted(out["report"])])
        for f, v in g.items():
            blob[f"{variant}/grad/{f}"] = v.astype(np.float32) if v.size > 4096 else v
        blob[f"{variant}/att_mask_sum"] = np.int64(inputs(variant, seed)[8].sum())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "answer_model_small.npz")
    if "--variants-only" not in sys.argv:
        np.savez_compressed(path, **blob)
        print(path, os.path.getsize(path), "bytes")
    vblob = {}
    for variant, seed in VARIANT_CASES.items():
        out, g = run(variant, seed)
        vblob.update({f"{variant}/{k}": v for k, v in summarise(out, g).items()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "answer_model_variants.npz")
    np.savez_compressed(path, **vblob)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
