"""Parity of the vlmap pre-training path (BASELINE config 4; vqa_transfer_externaldata_b200/memft.py over
include/vqa_memft.h) against the oracle: forward against the fp64 NumPy restatement (oracle/memft_np.py), gradients
against its torch-autograd twin (oracle/memft_torch.py), on identical inputs, weights and dropout masks (the masks the
kernels draw are exported and handed to the oracle).

Gates: fp32 mode <= 1e-4 relative (max-norm per tensor) on attention, pooled features, GRU state, logits, loss and
every gradient (measured: 1e-6 forward, 2e-5 gradients); top-1 / top-5 / masks exact. bf16 mode <= 2e-2 on the
forward tensors; its gradients are compared in relative L2 (<= 0.1): bf16 operands flip 0.2-0.5 % of the ReLU gates
against the plain fp64 oracle, each flip a 100 % error on that element (tests/test_oracle.py::test_relu_gate_flip_noise_model).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import answer_model_np as O  # noqa: E402
from oracle import memft_np as M  # noqa: E402
from oracle import memft_torch as MT  # noqa: E402

pytestmark = pytest.mark.gpu

DIMS = dict(B=6, K=12, n=5, Dv=64, D=128, L=128, W=24, A=200, T=5, Vq=50, Nws=20)
WIDE = dict(B=40, K=36, n=5, Dv=256, D=128, L=128, W=20, A=400, T=7, Vq=90, Nws=30)   # 400 sequences, 36 proposals


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def run_case(dims, precision, seed=0, train=True):
    from vqa_transfer_externaldata_b200 import memft as F
    p = {k: v.astype(np.float32).astype(np.float64) for k, v in M.init_params(dims, seed=seed).items()}
    batch = M.make_batch(dims, seed=seed + 1)
    for k in ("image_ft", "spatial_ft", "obj_blank_fill/normal_boxes", "attr_blank_fill/normal_boxes"):
        batch[k] = batch[k].astype(np.float32).astype(np.float64)
    cfg = F.make_config(dims, precision=precision)
    model = F.Model(batch, cfg, is_train=True, params=p, seed=1234 + seed)
    model.forward(dropout_step=3, with_grad_seed=train)
    masks = model.dropout_masks()
    ref = M.forward(p, batch, masks)
    loss, report = model.fetch()
    B, n = dims["B"], dims["n"]
    got = {"att": {"obj": model.mid_result["object_att_score"].cpu().numpy(), "attr": model.mid_result["attribute_att_score"].cpu().numpy()},
           "pooled": {"obj": model.mid_result["object_pooled_V_ft"].cpu().numpy(), "attr": model.mid_result["attribute_pooled_V_ft"].cpu().numpy()},
           "logit": {h: model.mid_result[h + "/logit"].cpu().numpy() for h in ("obj_blank_fill", "attr_blank_fill", "obj_wordset", "attr_wordset")},
           "loss": loss, "report": report, "q": model.buf.q.cpu().numpy()}
    err = {}
    for k in ("obj", "attr"):
        pooled_ref, att_ref = M.pooled_V_ft(p, batch, k, masks[f"att/{k}"])
        err[f"att/{k}"] = rel(got["att"][k], att_ref)
        err[f"pooled/{k}"] = rel(got["pooled"][k], pooled_ref)
        nb = np.repeat(batch["num_boxes"], n)
        assert all((got["att"][k][i, nb[i]:] == 0).all() for i in range(B * n)), "attention beyond num_boxes must be exactly zero"
    q_ref = []
    for k in ("obj", "attr"):
        blanks = np.asarray(batch[f"{k}_blank_fill/blanks"])
        E = p["l_glove"][blanks.reshape(B * n, -1)]
        q_ref.append(O.gru_fwd(E, np.asarray(batch[f"{k}_blank_fill/blanks_len"]).reshape(-1), p["gru_gates_w"], p["gru_gates_b"],
                               p["gru_cand_w"], p["gru_cand_b"])[0])
    err["q"] = rel(got["q"], np.concatenate(q_ref))
    for h, v in ref["logit"].items():
        err[f"logit/{h}"] = rel(got["logit"][h], v)
    err["loss"] = abs(loss - ref["loss"]) / abs(ref["loss"])
    grads = ref_g = None
    if train:
        model.backward()
        grads = model.gradients()
        tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
        tb = {k: torch.tensor(v, dtype=torch.float64 if np.asarray(v).dtype.kind == "f" else torch.int64) for k, v in batch.items()}
        tm = {k: torch.tensor(v) for k, v in masks.items()}
        tl, _ = MT.forward(tp, tb, tm)
        tl.backward()
        ref_g = {k: tp[k].grad.numpy() for k in p}
    model.close()
    return err, got, ref, grads, ref_g


def test_fp32_small():
    err, got, ref, grads, ref_g = run_case(DIMS, "fp32", seed=0)
    print({k: f"{v:.2e}" for k, v in err.items()})
    assert all(v < 1e-4 for v in err.values()), err
    for k, v in ref["report"].items():
        assert abs(got["report"][k] - v) <= 1e-4 * max(1.0, abs(v)), k     # accuracies: exact counts over the valid entries
    worst = {k: rel(grads[k], ref_g[k]) if np.abs(ref_g[k]).max() > 1e-12 else float(np.abs(grads[k]).max()) for k in ref_g}
    print({k: f"{v:.2e}" for k, v in worst.items()})
    bad = {k: v for k, v in worst.items() if not v < 1e-4}
    assert not bad, bad


def test_fp32_wide():
    """400 sequences (one 512-row wave of the recurrent kernels' fallback in fp32 mode), 36 proposals, ragged num_boxes / num."""
    err, got, ref, grads, ref_g = run_case(WIDE, "fp32", seed=3)
    assert all(v < 1e-4 for v in err.values()), err
    worst = {k: rel(grads[k], ref_g[k]) if np.abs(ref_g[k]).max() > 1e-12 else float(np.abs(grads[k]).max()) for k in ref_g}
    bad = {k: v for k, v in worst.items() if not v < 1e-4}
    assert not bad, bad


@pytest.mark.parametrize("dims", [DIMS, WIDE])
def test_bf16(dims):
    err, got, ref, grads, ref_g = run_case(dims, "bf16", seed=1)
    print({k: f"{v:.2e}" for k, v in err.items()})
    assert all(v < 2e-2 for v in err.values()), err
    for h in ("obj_blank_fill", "attr_blank_fill", "obj_wordset", "attr_wordset"):
        assert abs(got["report"][f"{h}_loss"] - ref["report"][f"{h}_loss"]) < 2e-2 * abs(ref["report"][f"{h}_loss"])
    worst = {k: rel_l2(grads[k], ref_g[k]) if np.abs(ref_g[k]).max() > 1e-12 else float(np.abs(grads[k]).max()) for k in ref_g}
    print({k: f"{v:.2e}" for k, v in worst.items()})
    bad = {k: v for k, v in worst.items() if not v < 0.1}
    assert not bad, bad


def test_top_k_tie_rule_and_masked_mean():
    """tf.nn.top_k / tf.argmax: the lower index wins a tie; loss and accuracies average over the valid entries only."""
    import ctypes as C
    from vqa_transfer_externaldata_b200 import lib as L
    lib = L.load()
    ops = C.c_void_p()
    L.check(lib.vqa_ops_create(C.byref(ops)))
    B, n, A = 2, 3, 16
    logit = np.zeros((B * n, A), np.float32)
    logit[0, [2, 5]] = 3.0            # tie between 2 and 5: argmax 2
    logit[1, :7] = 1.0                # seven-way tie: indices 0..4 are the top 5
    logit[2, 9] = 5.0
    logit[3, 1] = 2.0
    fills = np.array([5, 5, 9, 1, 0, 0], np.int32)
    num = np.array([3, 1], np.int32)  # rows 0, 1, 2 and 3 valid
    d = lambda a: torch.as_tensor(a).cuda()   # noqa: E731
    dl, df, dn = d(logit), d(fills), d(num)
    stats, rep = torch.zeros(B * n, 4, device="cuda"), torch.zeros(16, device="cuda")
    a = L.VqaSoftmaxCe(heads=1, B=B, n=n, A=A, top_k=5, logit=dl.data_ptr(), fills=df.data_ptr(), loss_scale=1.0,
                       stats=stats.data_ptr(), report=rep.data_ptr())
    a.num[0] = dn.data_ptr()
    L.check(lib.vqa_memft_softmax_ce(ops, C.byref(a), None))
    torch.cuda.synchronize()
    s = stats.cpu().numpy()
    assert s[:, 3].tolist() == [1, 1, 1, 1, 0, 0]
    assert s[:4, 1].tolist() == [0, 0, 1, 1]          # row 0: argmax is 2, not 5; row 1: argmax is 0
    assert s[:4, 2].tolist() == [1, 0, 1, 1]          # row 0: 5 is second; row 1: 5 is sixth of the tie
    ref_l, ref_a, ref_k = M.n_way_classification_loss(logit.reshape(B, n, A).astype(np.float64), fills.reshape(B, n), num)
    r = rep.cpu().numpy()
    assert abs(r[0] - ref_l) < 1e-6 and abs(r[1] - ref_a) < 1e-7 and abs(r[2] - ref_k) < 1e-7 and abs(r[3] - ref_l) < 1e-6
    lib.vqa_ops_destroy(ops)


def test_train_steps_reduce_the_loss_and_keep_checkpoint_names():
    from vqa_transfer_externaldata_b200 import memft as F
    p = M.init_params(DIMS, seed=4)
    batch = M.make_batch(DIMS, seed=5)
    model = F.Model(batch, F.make_config(DIMS, precision="bf16"), is_train=True, params=p)
    first = model.train_step()
    for _ in range(30):
        last = model.train_step()
    assert np.isfinite(first) and last < 0.7 * first, (first, last)
    sd = model.state_dict()
    assert {"spat_v_linear_v/fc/weights", "spat_att/compute/score/fc/weights", "encode_L_blank/rnn/gru_cell/gates/kernel",
            "classifier/fc/weights", "wordset_ft/LayerNorm/gamma", "L_GloVe/embed_map", "wordset_map/embed_map"} <= set(sd)
    assert sd["classifier/fc/weights"].shape == (2 * DIMS["L"], DIMS["A"])
    assert set(model.report) == {f"{h}_{m}" for h in F.HEADS for m in ("loss", "acc", "top_5_acc")} | {"total_loss"}
    model.close()
