"""ORACLE (test infrastructure, NOT product code): fp64 NumPy restatement of the FORWARD pass of the reference's vlmap
pre-training graph (BASELINE config 4) -- groundwork for the next round: the CUDA path for this graph does
not exist yet; its gradients come from the torch-autograd twin (oracle/memft_torch.py), which is checked against
central finite differences of this forward.

PARITY UNPINNED, like oracle/answer_model_np.py: TensorFlow 1.6 cannot run here; the op semantics are restated.

Reference followed: vlmap_memft/model_vlmap_bf_or_wordset_withatt_sp.py
  :56-74    build(): object / attribute V_ft, blank_fill and wordset branches; loss = sum of the four branch losses
  :323-365  build_object_V_ft (and :413-455 for attributes): spatial attention -- 6-d box features of the K proposals
            (spat_v_linear_v) against the 4 + 2-d key box of each of the n entries (spat_q_linear_v), Hadamard attention
            (scope spat_att, dropout 0.8, -inf mask beyond num_boxes), attended pooling of the RAW image features
  :505-556  build_object_blank_fill (:558-609 attributes): GRU over the blank's tokens (encode_L_blank), pooled_linear_l,
            q_linear_l, Hadamard joint_fc, dropout 0.5, classifier, masked softmax cross-entropy
  :367-411  build_object_wordset (:457-503 attributes): tanh(wordset_map[id]) -> wordset_ft (FC + LN + tanh) -> the same head
  :675-706  n_way_classification_loss: loss / top-1 / top-5 over the valid entries
The obj and attr branches share every variable (tf.AUTO_REUSE on the scope names). Every fc_layer input here is rank 3
([B, n, .] or [B*n, K, .]), so each LayerNorm runs over the whole [n, dim] / [K, dim] slab of a sample.
"""
import numpy as np

from . import answer_model_np as O

TOP_K = 5
PARAM_SHAPES = {   # field -> shape as a function of the dims dict (TF layout [in, out]); names = checkpoint scopes
    "wordset_map": lambda c: (c["Nws"], c["W"]),                     # wordset_map/embed_map
    "l_glove": lambda c: (c["Vq"], c["W"]),                          # L_GloVe/embed_map
    "sv_w": lambda c: (6, c["D"]), "sv_b": lambda c: (c["D"],), "sv_gamma": lambda c: (c["D"],), "sv_beta": lambda c: (c["D"],),   # spat_v_linear_v
    "sq_w": lambda c: (6, c["D"]), "sq_b": lambda c: (c["D"],), "sq_gamma": lambda c: (c["D"],), "sq_beta": lambda c: (c["D"],),   # spat_q_linear_v
    "att_w": lambda c: (c["D"], 1), "att_b": lambda c: (1,),         # spat_att/compute/score/fc
    "gru_gates_w": lambda c: (c["W"] + c["L"], 2 * c["L"]), "gru_gates_b": lambda c: (2 * c["L"],),   # encode_L_blank/rnn/gru_cell
    "gru_cand_w": lambda c: (c["W"] + c["L"], c["L"]), "gru_cand_b": lambda c: (c["L"],),
    "pl_w": lambda c: (c["Dv"], c["L"]), "pl_b": lambda c: (c["L"],), "pl_gamma": lambda c: (c["L"],), "pl_beta": lambda c: (c["L"],),
    "ql_w": lambda c: (c["L"], c["L"]), "ql_b": lambda c: (c["L"],), "ql_gamma": lambda c: (c["L"],), "ql_beta": lambda c: (c["L"],),
    "joint_w": lambda c: (c["L"], 2 * c["L"]), "joint_b": lambda c: (2 * c["L"],),
    "joint_gamma": lambda c: (2 * c["L"],), "joint_beta": lambda c: (2 * c["L"],),
    "cls_w": lambda c: (2 * c["L"], c["A"]), "cls_b": lambda c: (c["A"],),                            # classifier/fc
    "ws_w": lambda c: (c["W"], c["L"]), "ws_b": lambda c: (c["L"],), "ws_gamma": lambda c: (c["L"],), "ws_beta": lambda c: (c["L"],),   # wordset_ft
}


def init_params(c, seed=0, perturb=0.2):
    rng = np.random.default_rng(seed)
    p = {}
    for name, shp in PARAM_SHAPES.items():
        shape = shp(c)
        if name in ("wordset_map", "l_glove"):
            p[name] = rng.standard_normal(shape) * 0.4
        elif name.endswith("_gamma"):
            p[name] = 1.0 + perturb * rng.standard_normal(shape)
        elif name == "gru_gates_b":
            p[name] = 1.0 + perturb * rng.standard_normal(shape)
        elif len(shape) == 2:
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            p[name] = rng.uniform(-lim, lim, size=shape)
        else:
            p[name] = perturb * rng.standard_normal(shape)
    return p


def make_batch(c, seed=1):
    """Synthetic batch with the keys of vlmap_memft/datasets/dataset_vlmap.py (SURVEY 8d cfg4 shapes, small here)."""
    rng = np.random.default_rng(seed)
    B, K, n, T = c["B"], c["K"], c["n"], c["T"]
    nb = rng.integers(1, K + 1, size=B)
    nb[0] = K
    batch = {"image_ft": np.abs(rng.standard_normal((B, K, c["Dv"]))) * 0.5, "spatial_ft": rng.uniform(size=(B, K, 6)),
             "num_boxes": nb.astype(np.int32)}
    for kind in ("obj", "attr"):
        x0, y0 = rng.uniform(0, 0.5, size=(B, n)), rng.uniform(0, 0.5, size=(B, n))
        boxes = np.stack([x0, y0, x0 + rng.uniform(0.1, 0.5, size=(B, n)), y0 + rng.uniform(0.1, 0.5, size=(B, n))], axis=-1)
        ln = rng.integers(1, T + 1, size=(B, n)).astype(np.int32)
        blanks = rng.integers(1, c["Vq"], size=(B, n, T)).astype(np.int32)
        blanks[np.arange(T)[None, None, :] >= ln[:, :, None]] = 0
        num = rng.integers(1, n + 1, size=B).astype(np.int32)
        num[-1] = n
        batch.update({f"{kind}_blank_fill/normal_boxes": boxes, f"{kind}_blank_fill/blanks": blanks,
                      f"{kind}_blank_fill/blanks_len": ln, f"{kind}_blank_fill/fills": rng.integers(0, c["A"], size=(B, n)).astype(np.int32),
                      f"{kind}_blank_fill/num": num, f"{kind}_blank_fill/wordsets": rng.integers(0, c["Nws"], size=(B, n)).astype(np.int32)})
    return batch


def make_masks(c, seed=2):
    """0/1 keep masks of the six tf.nn.dropout sites: attention features of the two pooling branches, joint of the four heads."""
    rng = np.random.default_rng(seed)
    B, K, n, D, L = c["B"], c["K"], c["n"], c["D"], c["L"]
    m = {f"att/{k}": (rng.uniform(size=(B * n, K, D)) < 0.8).astype(np.float64) for k in ("obj", "attr")}
    m.update({f"joint/{k}": (rng.uniform(size=(B, n, 2 * L)) < 0.5).astype(np.float64)
              for k in ("obj_blank_fill", "attr_blank_fill", "obj_wordset", "attr_wordset")})
    return m


def fc_layer(x, w, b, gamma=None, beta=None, act=None):
    """modules.fc_layer: fully_connected over the last axis, optional layer_norm over all non-batch axes, activation."""
    z = x @ w + b
    if gamma is not None:
        z, _ = O.layer_norm_fwd(z, gamma, beta)
    if act == "relu":
        return np.maximum(z, 0.0)
    if act == "tanh":
        return np.tanh(z)
    return z


def pooled_V_ft(p, batch, kind, att_mask, keep_att=0.8):
    """build_object_V_ft / build_attribute_V_ft -> ([B, n, Dv] pooled features, [B*n, K] attention)."""
    V = np.asarray(batch["image_ft"], np.float64)
    B, K, Dv = V.shape
    boxes = np.asarray(batch[f"{kind}_blank_fill/normal_boxes"], np.float64)
    n = boxes.shape[1]
    Vt = np.repeat(V[:, None], n, axis=1).reshape(B * n, K, Dv)                       # tf.tile + reshape (:324-328)
    spat = np.repeat(np.asarray(batch["spatial_ft"], np.float64)[:, None], n, axis=1).reshape(B * n, K, 6)
    nb = np.repeat(np.asarray(batch["num_boxes"])[:, None], n, axis=1).reshape(-1)
    key = np.concatenate([boxes, (boxes[:, :, 2] - boxes[:, :, 0])[..., None], (boxes[:, :, 3] - boxes[:, :, 1])[..., None]], axis=-1)
    Hv = fc_layer(spat, p["sv_w"], p["sv_b"], p["sv_gamma"], p["sv_beta"], "relu")     # LN over [K, D] per (sample, entry)
    Hq = fc_layer(key, p["sq_w"], p["sq_b"], p["sq_gamma"], p["sq_beta"], "relu")      # LN over [n, D] per sample
    Hq = Hq.reshape(B * n, -1)
    F = Hv * Hq[:, None, :] * att_mask / keep_att                                     # hadamard_attention (modules.py:80-97)
    s = F @ p["att_w"].reshape(-1) + p["att_b"].reshape(())
    s = np.where(np.arange(K)[None, :] < nb[:, None], s, -np.inf)
    e = np.exp(s - s.max(axis=1, keepdims=True))
    a = e / e.sum(axis=1, keepdims=True)
    return np.einsum("bk,bkd->bd", a, Vt).reshape(B, n, Dv), a


def head(p, pooled, lang, joint_mask, keep_joint=0.5):
    """pooled_linear_l, q_linear_l, Hadamard, joint_fc, dropout, classifier on [B, n, .] inputs (:523-547)."""
    vl = fc_layer(pooled, p["pl_w"], p["pl_b"], p["pl_gamma"], p["pl_beta"], "relu")
    ll = fc_layer(lang, p["ql_w"], p["ql_b"], p["ql_gamma"], p["ql_beta"], "relu")
    j = fc_layer(vl * ll, p["joint_w"], p["joint_b"], p["joint_gamma"], p["joint_beta"], "relu") * joint_mask / keep_joint
    return fc_layer(j, p["cls_w"], p["cls_b"])


def n_way_classification_loss(logit, fills, num_valid, top_k=TOP_K):
    """softmax_cross_entropy_with_logits_v2 against one-hot fills, averaged over the valid entries; top-1 / top-k (:675-706)."""
    B, n, A = logit.shape
    z = logit - logit.max(axis=-1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(axis=-1, keepdims=True))
    ce = -np.take_along_axis(logp, np.asarray(fills)[..., None], axis=-1)[..., 0]
    mask = (np.arange(n)[None, :] < np.asarray(num_valid)[:, None]).astype(np.float64)
    loss = (ce * mask).sum() / mask.sum()
    pred = logit.argmax(axis=-1)
    acc = ((pred == fills) * mask).sum() / mask.sum()
    # tf.nn.top_k: the k largest, lower index first among equals
    order = np.argsort(-logit, axis=-1, kind="stable")[..., :top_k]
    topk = ((order == np.asarray(fills)[..., None]).any(axis=-1) * mask).sum() / mask.sum()
    return loss, acc, topk


def forward(p, batch, masks):
    """p: dict field -> fp64 array; masks: make_masks(). Returns dict with total loss, the report and the four logits."""
    p = {k: np.asarray(v, np.float64) for k, v in p.items()}
    out = {"report": {}, "logit": {}}
    total = 0.0
    pooled = {k: pooled_V_ft(p, batch, k, masks[f"att/{k}"]) for k in ("obj", "attr")}
    out["att"] = {k: v[1] for k, v in pooled.items()}
    for kind in ("obj", "attr"):
        blanks = np.asarray(batch[f"{kind}_blank_fill/blanks"])
        B, n, T = blanks.shape
        E = p["l_glove"][blanks.reshape(B * n, T)]                                   # embedding_lookup (:511-512)
        q, _ = O.gru_fwd(E, np.asarray(batch[f"{kind}_blank_fill/blanks_len"]).reshape(-1), p["gru_gates_w"],
                         p["gru_gates_b"], p["gru_cand_w"], p["gru_cand_b"])
        logit = head(p, pooled[kind][0], q.reshape(B, n, -1), masks[f"joint/{kind}_blank_fill"])
        loss, acc, topk = n_way_classification_loss(logit, batch[f"{kind}_blank_fill/fills"], batch[f"{kind}_blank_fill/num"])
        out["logit"][f"{kind}_blank_fill"] = logit
        out["report"].update({f"{kind}_blank_fill_loss": loss, f"{kind}_blank_fill_acc": acc,
                              f"{kind}_blank_fill_top_{TOP_K}_acc": topk})
        total += loss
    for kind in ("obj", "attr"):
        ws = np.tanh(p["wordset_map"][np.asarray(batch[f"{kind}_blank_fill/wordsets"])])   # (:373-374)
        ws_ft = fc_layer(ws, p["ws_w"], p["ws_b"], p["ws_gamma"], p["ws_beta"], "tanh")
        logit = head(p, pooled[kind][0], ws_ft, masks[f"joint/{kind}_wordset"])
        loss, acc, topk = n_way_classification_loss(logit, batch[f"{kind}_blank_fill/fills"], batch[f"{kind}_blank_fill/num"])
        out["logit"][f"{kind}_wordset"] = logit
        out["report"].update({f"{kind}_wordset_loss": loss, f"{kind}_wordset_acc": acc, f"{kind}_wordset_top_{TOP_K}_acc": topk})
        total += loss
    out["loss"] = total
    out["report"]["total_loss"] = total
    return out
