"""Groundwork for SURVEY 8 f2 (the vlmap pre-training graph, BASELINE config 4): the NumPy forward restatement and the
torch-autograd twin agree, and the twin's gradients match central finite differences of the NumPy forward. No CUDA
path exists for this graph yet; these tests pin the oracle the next round builds against."""
import numpy as np
import pytest

from oracle import memft_np as M

torch = pytest.importorskip("torch")
from oracle import memft_torch as MT  # noqa: E402

DIMS = dict(B=3, K=5, n=4, Dv=12, D=8, L=8, W=6, A=9, T=4, Vq=11, Nws=7)


def _setup(seed=0):
    p = M.init_params(DIMS, seed=seed)
    batch = M.make_batch(DIMS, seed=seed + 1)
    masks = M.make_masks(DIMS, seed=seed + 2)
    return p, batch, masks


def _to_torch(p, batch, masks, grad=True):
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=grad) for k, v in p.items()}
    tb = {k: torch.tensor(v, dtype=torch.float64 if np.asarray(v).dtype.kind == "f" else torch.int64) for k, v in batch.items()}
    tm = {k: torch.tensor(v) for k, v in masks.items()}
    return tp, tb, tm


def test_forward_matches_the_torch_twin():
    p, batch, masks = _setup()
    out = M.forward(p, batch, masks)
    tp, tb, tm = _to_torch(p, batch, masks, grad=False)
    loss, logits = MT.forward(tp, tb, tm)
    assert abs(loss.item() - out["loss"]) < 1e-12
    for k, v in logits.items():
        np.testing.assert_allclose(v.numpy(), out["logit"][k], rtol=1e-10, atol=1e-12)
    # the loss is the sum of the four branch losses (model...:70-73) and every report entry is there
    r = out["report"]
    assert abs(r["total_loss"] - sum(r[f"{k}_loss"] for k in ("obj_blank_fill", "attr_blank_fill", "obj_wordset", "attr_wordset"))) < 1e-12
    assert {f"{k}_{m}" for k in ("obj_blank_fill", "attr_blank_fill", "obj_wordset", "attr_wordset")
            for m in ("loss", "acc", "top_5_acc")} <= set(r)
    # masked attention: exact zeros beyond num_boxes, rows sum to one
    nb = np.repeat(batch["num_boxes"], DIMS["n"])
    for k in ("obj", "attr"):
        a = out["att"][k]
        assert all((a[i, nb[i]:] == 0).all() for i in range(len(nb))) and np.allclose(a.sum(1), 1.0)


def test_twin_gradients_match_finite_differences():
    p, batch, masks = _setup(seed=5)
    tp, tb, tm = _to_torch(p, batch, masks)
    loss, _ = MT.forward(tp, tb, tm)
    loss.backward()
    rng = np.random.default_rng(3)
    eps = 1e-6
    for name in M.PARAM_SHAPES:
        flat = p[name].reshape(-1)
        g = tp[name].grad.numpy().reshape(-1)
        for i in rng.choice(flat.size, size=min(4, flat.size), replace=False):
            old = flat[i]
            flat[i] = old + eps
            lp = M.forward(p, batch, masks)["loss"]
            flat[i] = old - eps
            lm = M.forward(p, batch, masks)["loss"]
            flat[i] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - g[i]) <= 1e-6 * max(1.0, abs(fd), abs(g[i])) + 2e-8, (name, i, fd, g[i])


def test_known_answers():
    # uniform logits: cross-entropy ln A; every label within the first TOP_K indices counts as a top-k hit (lower index first)
    logit = np.zeros((2, 3, 9))
    fills = np.array([[0, 4, 5], [8, 1, 2]])
    loss, acc, topk = M.n_way_classification_loss(logit, fills, np.array([3, 2]))
    assert abs(loss - np.log(9)) < 1e-12
    assert abs(acc - 1 / 5) < 1e-12            # argmax of equal logits = index 0: one of the five valid entries has label 0
    assert abs(topk - 3 / 5) < 1e-12           # labels 0, 4 (first row) and 1 (second row, entry 1) are among indices 0..4
    # the entries beyond `num` do not count
    l2, _, _ = M.n_way_classification_loss(np.where(np.arange(3)[None, :, None] >= 2, 50.0 * np.eye(9)[0], 0.0), fills, np.array([2, 2]))
    assert abs(l2 - np.log(9)) < 1e-12


def test_product_model_mirrors_the_oracle_parameter_set():
    """vqa_transfer_externaldata_b200.memft (the CUDA path's host side) declares the same variables with the same shapes
    as the oracle, under the reference's checkpoint names, and refuses to run without a GPU (no CPU fallback)."""
    from types import SimpleNamespace
    from vqa_transfer_externaldata_b200 import memft as F
    c = SimpleNamespace(**DIMS)
    assert list(F.FIELDS) == list(M.PARAM_SHAPES)
    for k, (shape, name) in F.FIELDS.items():
        assert tuple(shape(c)) == tuple(M.PARAM_SHAPES[k](DIMS)), k
        assert "/" in name
    assert F.TF_NAMES["cls_w"] == "classifier/fc/weights" and F.TF_NAMES["gru_gates_w"].startswith("encode_L_blank/")
    assert len(set(F.TF_NAMES.values())) == len(F.TF_NAMES)
    b = F.synthetic_batch(dict(DIMS, B=4), seed=3)
    assert set(b) == set(M.make_batch(dict(DIMS, B=4), seed=3))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            F.Model(None, F.make_config(dict(B=2, K=4, n=2, Dv=16, D=8, L=8, W=8, A=8, T=3, Vq=9, Nws=5)))
