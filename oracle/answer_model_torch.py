"""ORACLE twin (test infrastructure, NOT product code): PyTorch-CPU restatement of the same graph as
oracle/answer_model_np.py, written independently with torch ops and differentiated by autograd.

Two jobs:
  * cross-check of the NumPy oracle's hand-derived backward (tests/test_oracle.py);
  * the timed CPU baseline / `bench.py --impl reference` arm: the reference's own TF-1.6 CPU path
    cannot run in this image (no TensorFlow), so this port of the same graph (fp32, all host threads,
    oneDNN/MKL GEMMs) stands in for it and is labelled kind="port" everywhere.

PARITY UNPINNED (see answer_model_np.py header): TensorFlow was never executed.

Reference files followed: vqa/model_vlmap_answer.py:102-203, vqa/model_standard.py:193-285,
vlmap/modules.py:23-39, 67-97, 124-140, 630-650.
"""
import torch

LN_EPS = 1e-12


def layer_norm_all(z, gamma, beta):
    """tf.contrib.layers.layer_norm: moments over all non-batch axes, gamma/beta on the last axis."""
    dims = tuple(range(1, z.dim()))
    mu = z.mean(dim=dims, keepdim=True)
    var = (z - mu).pow(2).mean(dim=dims, keepdim=True)
    return (z - mu) * torch.rsqrt(var + LN_EPS) * gamma + beta


def fc_layer(x, w, b, gamma, beta):
    return torch.relu(layer_norm_all(torch.matmul(x, w) + b, gamma, beta))


def gru_encode(E, q_len, Wg, bg, Wc, bc):
    B, T, _ = E.shape
    L = Wc.shape[1]
    h = torch.zeros(B, L, dtype=E.dtype)
    for t in range(T):
        x = E[:, t]
        gates = torch.sigmoid(torch.cat([x, h], 1) @ Wg + bg)
        r, u = gates[:, :L], gates[:, L:]
        c = torch.tanh(torch.cat([x, r * h], 1) @ Wc + bc)
        hn = u * h + (1 - u) * c
        h = torch.where((t < q_len).unsqueeze(1), hn, h)
    return h


def forward(p, features, num_boxes, batch, train_mask, variant="vlmap_answer", keep_att=0.8,
            keep_joint=0.5, att_mask=None, joint_mask=None, joint_l_mask=None, noise=None, exist=None,
            num_marginal=200, ent_mask=None):
    """p: dict field -> torch tensor (requires_grad where wanted). Returns dict with loss, logit,
    att_score, pooled, condition, pred."""
    idx = batch["image_idx"].long()
    V = features[idx]
    nbox = num_boxes[idx].long()
    B, K, _ = V.shape
    Hv = fc_layer(V, p["v_w"], p["v_b"], p["v_gamma"], p["v_beta"])
    E = p["embed"][batch["q_intseq"].long()]
    q = gru_encode(E, batch["q_intseq_len"], p["gru_gates_w"], p["gru_gates_b"], p["gru_cand_w"],
                   p["gru_cand_b"])
    Hq = fc_layer(q, p["qv_w"], p["qv_b"], p["qv_gamma"], p["qv_beta"])
    F = Hv * Hq.unsqueeze(1)
    if att_mask is not None:
        F = F * att_mask
    F = F / keep_att
    s = torch.matmul(F, p["att_w"]).squeeze(-1) + p["att_b"]
    valid = torch.arange(K).unsqueeze(0) < nbox.unsqueeze(1)
    s = torch.where(valid, s, torch.full_like(s, float("-inf")))
    a = torch.softmax(s, dim=-1)
    Vp = V
    if variant == "vlmap_answer_adapt":        # vqa/model_vlmap_answer_adapt.py:132-142
        Vp = fc_layer(V, p["va_w"], p["va_b"], p["va_gamma"], p["va_beta"])
    P = torch.bmm(a.unsqueeze(1), Vp).squeeze(1)
    Hp = fc_layer(P, p["pl_w"], p["pl_b"], p["pl_gamma"], p["pl_beta"])
    ql_in, cond = q, q
    if variant == "vlmap_answer2":         # vqa/model_vlmap_answer2.py:127-131
        ql_in = torch.tanh(layer_norm_all(q @ p["qp_w"] + p["qp_b"], p["qp_gamma"], p["qp_beta"]))
        cond = ql_in
    elif variant == "vlmap_answer_no_noise":  # vqa/model_vlmap_answer_no_noise.py:122-125
        ql_in = q @ p["qp_w"] + p["qp_b"]
    elif variant == "vlmap_answer_full":      # vqa/model_vlmap_answer_full.py:124-134
        q_mean = q @ p["qp_w"] + p["qp_b"]
        q_lss = q @ p["qs_w"] + p["qs_b"]
        ql_in = q_mean + (noise if noise is not None else 0.0) * torch.sqrt(torch.exp(q_lss))
    Hl = fc_layer(ql_in, p["ql_w"], p["ql_b"], p["ql_gamma"], p["ql_beta"])
    if variant in ("vlmap_answer_noc", "vlmap_answer_nocarch"):   # vqa/model_vlmap_answer_noc.py:177-203
        Jv = fc_layer(Hp, p["joint_w"], p["joint_b"], p["joint_gamma"], p["joint_beta"])
        Jl = fc_layer(Hl, p["jl_w"], p["jl_b"], p["jl_gamma"], p["jl_beta"])
        if joint_mask is not None:
            Jv = Jv * joint_mask
        if joint_l_mask is not None:
            Jl = Jl * joint_l_mask
        logit = (Jv / keep_joint) @ p["ans_w"] + p["ans_b"] + (Jl / keep_joint) @ p["al_w"] + p["al_b"]
    else:
        Jn = fc_layer(Hp * Hl, p["joint_w"], p["joint_b"], p["joint_gamma"], p["joint_beta"])
        if joint_mask is not None:
            Jn = Jn * joint_mask
        Jd = Jn / keep_joint
        logit = Jd @ p["ans_w"] + p["ans_b"]
    target = batch["answer_target"].to(logit.dtype)
    BCE = lambda x: torch.nn.functional.binary_cross_entropy_with_logits(x, target, reduction="none")
    pred_logit = logit
    if variant in ("vlmap_answer_vqa_all", "vlmap_answer_vqa_all2"):
        # vqa/model_vlmap_answer_vqa_all.py:188-244 / _vqa_all2.py:188-243
        if variant == "vlmap_answer_vqa_all":
            mn = logit.min(dim=1, keepdim=True).values
            logit = logit * exist + mn * (1 - exist)
        tuned = Jd @ p["tw_w"] + p["tw_b"]
        if variant == "vlmap_answer_vqa_all":
            bce_train = (BCE(logit) + BCE(logit + tuned)) * train_mask
            bce = BCE(logit) + BCE(logit + tuned)
            pred_logit = logit + tuned
        else:
            bce_train = BCE(logit) * train_mask + BCE(tuned)
            bce = BCE(logit) + BCE(tuned)
            pred_logit = logit * (1 - train_mask) + tuned * train_mask
        logit = logit + tuned
    else:
        bce = BCE(logit)
        bce_train = bce * train_mask if variant != "standard" else bce
    loss = bce_train.sum(-1).mean()
    if variant == "vlmap_answer_ent":         # vqa/model_vlmap_answer_ent.py:193-213, 284-294
        M = num_marginal
        tile = Hp.detach().repeat(M, 1).reshape(-1, M, Hp.shape[1])       # tf.tile + tf.reshape, stop_gradient
        tj = fc_layer(tile * Hl.unsqueeze(1), p["joint_w"], p["joint_b"], p["joint_gamma"], p["joint_beta"])
        if ent_mask is not None:
            tj = tj * ent_mask
        tl = (tj / keep_joint) @ p["ans_w"] + p["ans_b"]
        sel = (exist * train_mask) > 0.5
        prob = torch.softmax(tl[:, :, sel], dim=-1)
        marg = prob.mean(dim=1)
        loss = loss + 0.1 * (marg * torch.log(marg + 1e-8)).sum(-1).mean()
    if variant == "vlmap_answer_full":        # latent_loss, weight 0.1 (:33, 217-223, 272-276)
        latent = -0.5 * (1 + q_lss - q_mean.pow(2) - torch.exp(q_lss)).sum(-1).mean()
        loss = loss + 0.1 * latent
    return {"loss": loss, "report_loss": bce.sum(-1).mean(), "logit": logit, "att_score": a, "pooled": P,
            "condition": cond, "pred": pred_logit.argmax(-1)}
