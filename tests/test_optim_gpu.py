"""clip_by_global_norm(20) + Adam on the device (vqa_adam_step_shadowed: one pass that also rewrites the bf16 operand
shadows of the updated weight matrices) against the oracle's restatement of optimize_loss (vqa/trainer.py:106-114),
and the refreshed shadows through a forward pass with the updated parameters."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import answer_model_np as O  # noqa: E402
from parity_util import build_case, rel_err  # noqa: E402

pytestmark = pytest.mark.gpu

SMALL = dict(B=16, K=12, Dv=256, D=128, L=128, A=200, T=6, W=20, Vq=50)


@pytest.mark.parametrize("variant,clip", [("vlmap_answer", 20.0), ("standard", 0.05), ("vlmap_answer_vqa_all", 20.0)])
def test_two_optimizer_steps_match_the_oracle(variant, clip):
    case = build_case(SMALL, variant=variant, precision="fp32", seed=31)
    eng = case["eng"]
    fields = O.trainable_fields(variant)
    p = {k: np.asarray(v, np.float64).copy() for k, v in case["params"].items()}
    m = {k: np.zeros_like(p[k]) for k in fields}
    v = {k: np.zeros_like(p[k]) for k in fields}
    eng.stage_batch(case["batch"])
    for t in (1, 2):
        eng.forward(seed=5, step=t)
        eng.backward()
        g = {f: eng.params.grad_views[f].detach().cpu().numpy().astype(np.float64) for f in fields}
        # the embedding gradient as TF holds it (IndexedSlices rows): the oracle's own dE for the current parameters
        att_mask, joint_mask = eng.dropout_masks(5, t)
        torch.cuda.synchronize()
        _, cache = O.forward(p, case["feats"], case["nb"], case["batch"], case["m"], variant=variant,
                             att_mask=att_mask.cpu().numpy(), joint_mask=joint_mask.cpu().numpy())
        inter = {}
        O.backward(cache, intermediates=inter)
        slice_ss = float((inter["dE"] ** 2).sum())
        dense_ss = float((g["embed"] ** 2).sum())
        assert slice_ss > dense_ss * 1.0001 or slice_ss < dense_ss * 0.9999   # repeated tokens: the two norms differ
        dev_slot = float(eng.params.grad_buf[eng.params.n_train].item())
        assert abs(dev_slot - slice_ss) <= 2e-3 * slice_ss               # fp32-mode gradients, 1e-4-class agreement
        eng.adam_step(lr=1e-3, clip_norm=clip)
        torch.cuda.synchronize()
        pt = {k: p[k] for k in fields}
        # the device's own slot value keeps the comparison of the update itself at round-off level
        gnorm = O.clip_adam_step(pt, g, m, v, t, clip=clip, slice_sumsq={"embed": dev_slot})
        p.update(pt)
        assert abs(eng.grad_norm.item() - gnorm) <= 1e-5 * gnorm
        if clip < 1.0:
            assert gnorm > clip   # the clip is active in this case
        for f in fields:
            got = eng.params.views[f].detach().cpu().numpy()
            assert np.abs(got - p[f]).max() <= 2e-6 * max(1.0, np.abs(p[f]).max()), (t, f)
    # frozen tensors untouched
    for f in set(case["params"]) - set(fields):
        assert np.array_equal(eng.params.views[f].detach().cpu().numpy(), case["params"][f])
    # the operand shadows follow the update: a forward with the device's parameters equals the oracle's forward with
    # the oracle's updated parameters
    eng.forward(seed=5, step=9)
    att_mask, joint_mask = eng.dropout_masks(5, 9)
    torch.cuda.synchronize()
    out, _ = O.forward(p, case["feats"], case["nb"], case["batch"], case["m"], variant=variant,
                       att_mask=att_mask.cpu().numpy(), joint_mask=joint_mask.cpu().numpy())
    live = case["m"]["exist"] > 0
    assert rel_err(eng.outputs()["logit"].cpu().numpy()[:, live], out["logit"][:, live]) < 1e-4
    loss, _ = eng.read_scalars()
    assert abs(loss - out["loss"]) / abs(out["loss"]) < 1e-4
