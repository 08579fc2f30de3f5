"""ORACLE twin (test infrastructure, NOT product code): the vlmap pre-training graph of oracle/memft_np.py written
independently with torch ops and differentiated by autograd -- the gradient reference for BASELINE config 4 (SURVEY 8 f2,
a 'next' row: no CUDA path yet). PARITY UNPINNED (TensorFlow was never executed).

Reference followed: vlmap_memft/model_vlmap_bf_or_wordset_withatt_sp.py:56-74, 323-609, 675-706; vlmap/modules.py:67-97,
23-39, 124-140, 630-650."""
import torch

from .answer_model_torch import gru_encode, layer_norm_all

TOP_K = 5


def fc(x, w, b, gamma=None, beta=None, act=None):
    z = torch.matmul(x, w) + b
    if gamma is not None:
        z = layer_norm_all(z, gamma, beta)
    if act == "relu":
        z = torch.relu(z)
    elif act == "tanh":
        z = torch.tanh(z)
    return z


def attend(p, batch, kind, att_mask, keep_att=0.8):
    V, spat, nb = batch["image_ft"], batch["spatial_ft"], batch["num_boxes"].long()
    boxes = batch[f"{kind}_blank_fill/normal_boxes"]
    B, K, Dv = V.shape
    n = boxes.shape[1]
    Vt = V.unsqueeze(1).expand(B, n, K, Dv).reshape(B * n, K, Dv)
    st = spat.unsqueeze(1).expand(B, n, K, 6).reshape(B * n, K, 6)
    nbt = nb.unsqueeze(1).expand(B, n).reshape(-1)
    key = torch.cat([boxes, (boxes[..., 2] - boxes[..., 0]).unsqueeze(-1), (boxes[..., 3] - boxes[..., 1]).unsqueeze(-1)], -1)
    Hv = fc(st, p["sv_w"], p["sv_b"], p["sv_gamma"], p["sv_beta"], "relu")
    Hq = fc(key, p["sq_w"], p["sq_b"], p["sq_gamma"], p["sq_beta"], "relu").reshape(B * n, -1)
    F = Hv * Hq.unsqueeze(1) * att_mask / keep_att
    s = torch.matmul(F, p["att_w"]).squeeze(-1) + p["att_b"]
    s = torch.where(torch.arange(K).unsqueeze(0) < nbt.unsqueeze(1), s, torch.full_like(s, float("-inf")))
    a = torch.softmax(s, dim=-1)
    return torch.bmm(a.unsqueeze(1), Vt).squeeze(1).reshape(B, n, Dv)


def head(p, pooled, lang, joint_mask, keep_joint=0.5):
    vl = fc(pooled, p["pl_w"], p["pl_b"], p["pl_gamma"], p["pl_beta"], "relu")
    ll = fc(lang, p["ql_w"], p["ql_b"], p["ql_gamma"], p["ql_beta"], "relu")
    j = fc(vl * ll, p["joint_w"], p["joint_b"], p["joint_gamma"], p["joint_beta"], "relu") * joint_mask / keep_joint
    return fc(j, p["cls_w"], p["cls_b"])


def masked_ce(logit, fills, num):
    B, n, A = logit.shape
    ce = torch.nn.functional.cross_entropy(logit.reshape(B * n, A), fills.long().reshape(-1), reduction="none").reshape(B, n)
    mask = (torch.arange(n).unsqueeze(0) < num.long().unsqueeze(1)).to(logit.dtype)
    return (ce * mask).sum() / mask.sum()


def forward(p, batch, masks):
    """p / batch / masks: dicts of torch tensors with the fields and keys of oracle/memft_np.py. Returns (loss, logits)."""
    pooled = {k: attend(p, batch, k, masks[f"att/{k}"]) for k in ("obj", "attr")}
    loss = 0.0
    logits = {}
    for kind in ("obj", "attr"):
        blanks = batch[f"{kind}_blank_fill/blanks"].long()
        B, n, T = blanks.shape
        E = p["l_glove"][blanks.reshape(B * n, T)]
        q = gru_encode(E, batch[f"{kind}_blank_fill/blanks_len"].reshape(-1), p["gru_gates_w"], p["gru_gates_b"],
                       p["gru_cand_w"], p["gru_cand_b"]).reshape(B, n, -1)
        logits[f"{kind}_blank_fill"] = head(p, pooled[kind], q, masks[f"joint/{kind}_blank_fill"])
        loss = loss + masked_ce(logits[f"{kind}_blank_fill"], batch[f"{kind}_blank_fill/fills"], batch[f"{kind}_blank_fill/num"])
    for kind in ("obj", "attr"):
        ws = torch.tanh(p["wordset_map"][batch[f"{kind}_blank_fill/wordsets"].long()])
        ws_ft = fc(ws, p["ws_w"], p["ws_b"], p["ws_gamma"], p["ws_beta"], "tanh")
        logits[f"{kind}_wordset"] = head(p, pooled[kind], ws_ft, masks[f"joint/{kind}_wordset"])
        loss = loss + masked_ce(logits[f"{kind}_wordset"], batch[f"{kind}_blank_fill/fills"], batch[f"{kind}_blank_fill/num"])
    return loss, logits
