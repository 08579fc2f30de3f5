"""ORACLE (test infrastructure, NOT product code): fp64 NumPy restatement of the reference's VQA
answer-model hot path, forward and hand-derived backward.

PARITY UNPINNED: the reference ships no tests / golden tensors for this path and its arithmetic lives
in tensorflow-gpu==1.6.0 (requirements.txt:1), which is not vendored and cannot be installed here
(no TF, no h5py, no Python 2, no network). This file restates the TF-1.6 op semantics the reference's
call sites rely on; it is pinned only by (1) central finite differences on its own forward,
(2) an independently written PyTorch-autograd twin (oracle/answer_model_torch.py), and
(3) known-answer cases (tests/test_oracle.py). TensorFlow itself was never executed.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this package. The product path (vqa_transfer_externaldata_b200) never does.

Reference files followed (paths under /root/reference):
  vqa/model_vlmap_answer.py:102-288   graph, loss, metrics           (variant 'vlmap_answer')
  vqa/model_standard.py:193-376       same trunk, learned classifier  (variant 'standard')
  vqa/model_vlmap_answer2.py:127-131,164         q_L_ft2 = tanh(LN(FC(q))) feeds q_linear_l  ('vlmap_answer2')
  vqa/model_vlmap_answer_no_noise.py:122-125,157 q_L_mean = FC(q) feeds q_linear_l            ('vlmap_answer_no_noise')
  vqa/model_vlmap_answer_noc.py:177-203 (= _nocarch.py) two heads joint_v / joint_l, logits summed ('vlmap_answer_noc')
  vqa/model_vlmap_answer_full.py:124-134,166,217-223,272-276  q_L_mean + noise * sqrt(exp(q_L_log_sigma_sq)) feeds
                                      q_linear_l; loss += 0.1 * KL                                ('vlmap_answer_full')
  vqa/model_vlmap_answer_vqa_all.py:188-244   frozen head with min-logit fill of absent answers + TunedWordWeightAnswer
                                      on the SAME joint (:215-216), two BCE terms                 ('vlmap_answer_vqa_all')
  vqa/model_vlmap_answer_vqa_all2.py:188-243  same without the fill; BCE(tuned) unmasked; pred from
                                      logit * test_mask + tuned * train_mask                       ('vlmap_answer_vqa_all2')
  vqa/model_vlmap_answer_adapt.py:132-142     v_adapt = relu(LN(FC(V))) is what attention pools     ('vlmap_answer_adapt')
  vqa/model_vlmap_answer_ent.py:14-16,193-213,284-294  NUM_MARGINAL-way tiled joint head, marginal softmax over the
                                      train & existing answers, loss += 0.1 * negative entropy          ('vlmap_answer_ent')
  vlmap/modules.py:630-650            fc_layer = fully_connected -> layer_norm -> activation
  vlmap/modules.py:67-97              hadamard_attention
  vlmap/modules.py:23-39              attention_pooling
  vlmap/modules.py:124-140            encode_L (GRUCell + dynamic_rnn)
  vlmap/modules.py:589-627            WordWeightAnswer (a fully_connected with constant init)
  vqa/trainer.py:87-114               optimize_loss(Adam, clip_gradients=20.0)
"""
import numpy as np

LN_EPS = 1e-12  # tf.contrib.layers.layer_norm -> tf.nn.batch_normalization variance_epsilon

# C-struct field name -> TF checkpoint variable name (vlmap_answer); 'standard' nests the last four
# scopes under reasoning/ and calls the head reasoning/classifier (vqa/model_standard.py:251-275)
TF_NAMES = {
    "embed": "LearnGloVe/embed_map",
    "v_w": "v_linear_v/fc/weights", "v_b": "v_linear_v/fc/biases",
    "v_gamma": "v_linear_v/LayerNorm/gamma", "v_beta": "v_linear_v/LayerNorm/beta",
    "gru_gates_w": "encode_L/rnn/gru_cell/gates/kernel", "gru_gates_b": "encode_L/rnn/gru_cell/gates/bias",
    "gru_cand_w": "encode_L/rnn/gru_cell/candidate/kernel",
    "gru_cand_b": "encode_L/rnn/gru_cell/candidate/bias",
    "qv_w": "q_linear_v/fc/weights", "qv_b": "q_linear_v/fc/biases",
    "qv_gamma": "q_linear_v/LayerNorm/gamma", "qv_beta": "q_linear_v/LayerNorm/beta",
    "att_w": "hadamard_attention/compute/score/fc/weights",
    "att_b": "hadamard_attention/compute/score/fc/biases",
    "pl_w": "pooled_linear_l/fc/weights", "pl_b": "pooled_linear_l/fc/biases",
    "pl_gamma": "pooled_linear_l/LayerNorm/gamma", "pl_beta": "pooled_linear_l/LayerNorm/beta",
    "ql_w": "q_linear_l/fc/weights", "ql_b": "q_linear_l/fc/biases",
    "ql_gamma": "q_linear_l/LayerNorm/gamma", "ql_beta": "q_linear_l/LayerNorm/beta",
    "joint_w": "joint_fc/fc/weights", "joint_b": "joint_fc/fc/biases",
    "joint_gamma": "joint_fc/LayerNorm/gamma", "joint_beta": "joint_fc/LayerNorm/beta",
    "ans_w": "WordWeightAnswer/fc/weights", "ans_b": "WordWeightAnswer/fc/biases",
}
PARAM_FIELDS = list(TF_NAMES.keys())

# vqa/model_vlmap_answer.py:81-89 filter_train_vars: these top-level scopes are frozen
FROZEN_SCOPES_VLMAP_ANSWER = ("q_linear_l", "pooled_linear_l", "joint_fc", "WordWeightAnswer")


# extra question layer of the two variants (scope q_L_ft2 / q_L_mean: trained, never frozen)
NOC_FIELDS = ["jl_w", "jl_b", "jl_gamma", "jl_beta", "al_w", "al_b"]   # joint_l and WordWeightAnswerL (frozen)
TUNED_FIELDS = ["tw_w", "tw_b"]                     # TunedWordWeightAnswer (trained), model_vlmap_answer_vqa_all.py:215-219
ADAPT_FIELDS = ["va_w", "va_b", "va_gamma", "va_beta"]   # v_adapt (trained), model_vlmap_answer_adapt.py:132-135
EXTRA_FIELDS = {"vlmap_answer2": ["qp_w", "qp_b", "qp_gamma", "qp_beta"], "vlmap_answer_no_noise": ["qp_w", "qp_b"],
                "vlmap_answer_noc": NOC_FIELDS, "vlmap_answer_nocarch": NOC_FIELDS,
                "vlmap_answer_full": ["qp_w", "qp_b", "qs_w", "qs_b"],   # q_L_mean, q_L_log_sigma_sq (:124-131)
                "vlmap_answer_vqa_all": TUNED_FIELDS, "vlmap_answer_vqa_all2": TUNED_FIELDS,
                "vlmap_answer_adapt": ADAPT_FIELDS}
EXTRA_TF_NAMES = {
    "qs_w": "q_L_log_sigma_sq/fc/weights", "qs_b": "q_L_log_sigma_sq/fc/biases",
    "tw_w": "TunedWordWeightAnswer/fc/weights", "tw_b": "TunedWordWeightAnswer/fc/biases",
    "va_w": "v_adapt/fc/weights", "va_b": "v_adapt/fc/biases",
    "va_gamma": "v_adapt/LayerNorm/gamma", "va_beta": "v_adapt/LayerNorm/beta",
}
LATENT_LOSS_WEIGHT = 0.1   # model_vlmap_answer_full.py:33
W_ENTROPY = 0.1            # model_vlmap_answer_ent.py:14
NUM_MARGINAL = 200         # model_vlmap_answer_ent.py:16


def param_fields(variant):
    return PARAM_FIELDS + EXTRA_FIELDS.get(variant, [])


def trainable_fields(variant):
    """Fields optimize_loss receives as `variables` (vqa/trainer.py:99-114)."""
    if variant == "standard":  # vqa/model_standard.py:80-84: everything trains
        return list(PARAM_FIELDS)
    # vlmap_answer (:81-89), vlmap_answer2 (:69-78), vlmap_answer_no_noise (:66-74): same four frozen scopes
    base = [f for f in PARAM_FIELDS if TF_NAMES[f].split("/")[0] not in FROZEN_SCOPES_VLMAP_ANSWER]
    if variant in ("vlmap_answer_noc", "vlmap_answer_nocarch"):   # model_vlmap_answer_noc.py:78-88: both heads frozen
        return base
    return base + EXTRA_FIELDS.get(variant, [])


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def round_bf16(x):
    """Round-to-nearest-even to bfloat16 (returned as float64). Used by forward(operand_round=round_bf16) to
    restate the MIXED-PRECISION arithmetic the bf16 mode of the CUDA path specifies: every GEMM operand (and
    the stored pre-LN v-projection) is a bf16 value, everything else is fp32/fp64. With it the oracle makes
    the same ReLU gate decisions as the device, so bf16-mode gradients can be held to the 2e-2 gate."""
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(a.shape).astype(np.float64)


def _ident(x):
    return x


# ------------------------------------------------------------------------------------------------
# layers.layer_norm (vlmap/modules.py:646-647): statistics over ALL non-batch axes, gamma/beta on the
# last axis, biased variance via moments (two-pass), eps 1e-12.   SURVEY Q1
# ------------------------------------------------------------------------------------------------
def layer_norm_fwd(z, gamma, beta):
    axes = tuple(range(1, z.ndim))
    mu = z.mean(axis=axes, keepdims=True)
    var = ((z - mu) ** 2).mean(axis=axes, keepdims=True)
    rstd = 1.0 / np.sqrt(var + LN_EPS)
    xhat = (z - mu) * rstd
    return gamma * xhat + beta, (xhat, rstd)


def layer_norm_bwd(dy, gamma, cache):
    xhat, rstd = cache
    axes = tuple(range(1, dy.ndim))
    red = tuple(range(0, dy.ndim - 1))
    dgamma = (dy * xhat).sum(axis=red)
    dbeta = dy.sum(axis=red)
    dxhat = dy * gamma
    dz = rstd * (dxhat - dxhat.mean(axis=axes, keepdims=True)
                 - xhat * (dxhat * xhat).mean(axis=axes, keepdims=True))
    return dz, dgamma, dbeta


def fc_ln_relu_fwd(x, w, b, gamma, beta, q=_ident, qz=_ident):
    """modules.fc_layer(use_bias, use_ln, relu): rank>2 inputs contract the last axis.
    q rounds the GEMM input operand (w is passed in already rounded), qz the stored pre-LN output."""
    x = q(x)
    z = qz(x @ w + b)
    y, ln_cache = layer_norm_fwd(z, gamma, beta)
    return np.maximum(y, 0.0), (x, y, ln_cache)


def fc_ln_relu_bwd(dh, w, gamma, cache, need_dx=True, flip=None, gate=None):
    """flip: optional boolean array like y; True entries use the OPPOSITE ReLU gate. The parity tests use
    it to bound the effect of gates whose pre-activation is zero to working precision (|y| < tau): the
    backward pass is linear in the gates, so each undecidable gate contributes a fixed +-delta."""
    x, y, ln_cache = cache
    if gate is None:
        gate = y > 0
    else:  # ReLU gates decided elsewhere (the device's own decisions: the backward pass is linear in them)
        gate = np.asarray(gate, dtype=bool).reshape(y.shape)
    if flip is not None:
        gate = np.logical_xor(gate, flip)
    dy = dh * gate
    dz, dgamma, dbeta = layer_norm_bwd(dy, gamma, ln_cache)
    x2 = x.reshape(-1, x.shape[-1])
    dz2 = dz.reshape(-1, dz.shape[-1])
    dw = x2.T @ dz2
    db = dz2.sum(axis=0)
    dx = dz @ w.T if need_dx else None
    return dx, dw, db, dgamma, dbeta


# ------------------------------------------------------------------------------------------------
# GRU: tf.contrib.rnn.GRUCell under tf.nn.dynamic_rnn(sequence_length)  (vlmap/modules.py:131-135)
# gates = sigmoid([x, h] Wg + bg), split (r, u); c = tanh([x, r*h] Wc + bc); h' = u*h + (1-u)*c;
# for t >= len the state is copied through. Zero initial state.            SURVEY Q5
# ------------------------------------------------------------------------------------------------
def gru_fwd(E, q_len, Wg, bg, Wc, bc, q=_ident):
    """q rounds the matmul operands h and r*h (E, Wg, Wc arrive already rounded); the element-wise gate math
    always uses the unrounded state. In the CUDA path's bf16 mode the x-parts of the two pre-activations (x Wx + b,
    hoisted out of the recurrence as one product over all steps) are STORED as bf16: q rounds them too."""
    B, T, W = E.shape
    L = Wc.shape[1]
    h = np.zeros((B, L))
    steps = []
    for t in range(T):
        x = E[:, t, :]
        h_op = q(h)
        if q is _ident:
            g = np.concatenate([x, h_op], axis=1) @ Wg + bg
        else:
            g = q(x @ Wg[:W] + bg) + h_op @ Wg[W:]
        r, u = sigmoid(g[:, :L]), sigmoid(g[:, L:])
        rh_op = q(r * h)
        if q is _ident:
            c = np.tanh(np.concatenate([x, rh_op], axis=1) @ Wc + bc)
        else:
            c = np.tanh(q(x @ Wc[:W] + bc) + rh_op @ Wc[W:])
        hn = u * h + (1.0 - u) * c
        valid = (t < q_len)[:, None]
        steps.append((x, h, r, u, rh_op, c, valid, h_op))
        h = np.where(valid, hn, h)
    return h, steps


def gru_bwd(dq, steps, Wg, Wc, W):
    L = Wc.shape[1]
    dWg = np.zeros_like(Wg)
    dbg = np.zeros(Wg.shape[1])
    dWc = np.zeros_like(Wc)
    dbc = np.zeros(Wc.shape[1])
    dh = dq.copy()
    dE = []
    for (x, h, r, u, rh, c, valid, h_op) in reversed(steps):
        dhn = np.where(valid, dh, 0.0)        # gradient reaching h' (only valid steps used it)
        dh_prev = np.where(valid, 0.0, dh)    # copied-through state
        du = dhn * (h - c)
        dc = dhn * (1.0 - u)
        dh_prev = dh_prev + dhn * u
        dcp = dc * (1.0 - c * c)
        xc = np.concatenate([x, rh], axis=1)
        dWc += xc.T @ dcp
        dbc += dcp.sum(axis=0)
        dxc = dcp @ Wc.T
        dx = dxc[:, :W]
        drh = dxc[:, W:]
        dr = drh * h
        dh_prev = dh_prev + drh * r
        dg = np.concatenate([dr * r * (1.0 - r), du * u * (1.0 - u)], axis=1)
        xg = np.concatenate([x, h_op], axis=1)
        dWg += xg.T @ dg
        dbg += dg.sum(axis=0)
        dxg = dg @ Wg.T
        dx = dx + dxg[:, :W]
        dh_prev = dh_prev + dxg[:, W:]
        dE.append(dx)
        dh = dh_prev
    dE = np.stack(dE[::-1], axis=1)  # [B, T, W]
    return dE, dWg, dbg, dWc, dbc


# ------------------------------------------------------------------------------------------------
# loss + metrics (vqa/model_vlmap_answer.py:192-288)
# ------------------------------------------------------------------------------------------------
def bce_with_logits(x, z):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log(1 + exp(-|x|))"""
    return np.maximum(x, 0.0) - x * z + np.log1p(np.exp(-np.abs(x)))


def answer_masks(A, num_train_answer, is_object, is_attribute, answer_exist):
    tm = (np.arange(A) < num_train_answer).astype(np.float64)  # tf.sequence_mask, :39-42
    return {"train": tm, "test": 1.0 - tm, "obj": np.asarray(is_object, np.float64),
            "attr": np.asarray(is_attribute, np.float64), "exist": np.asarray(answer_exist, np.float64)}


def _normal(num, den):
    # tf.where(tf.equal(den, 0), den, num / den)
    return den if den == 0 else num / den


def metrics(logit, target, m, use_train_mask=True, loss_terms=None, pred_logit=None):
    """returns (train_loss, report dict, per-sample dict, pred).
    loss_terms: list of (logits, masked) -- the BCE terms that are summed (vqa_all: two terms, both train-masked,
    model_vlmap_answer_vqa_all.py:234-243; vqa_all2: untuned masked + tuned unmasked, _vqa_all2.py:231-239);
    default = one term on `logit`. pred_logit: what argmax runs on (default `logit`)."""
    B, A = logit.shape
    if loss_terms is None:
        loss_terms = [(logit, use_train_mask)]
    train_loss, report_loss = 0.0, 0.0
    for x, masked in loss_terms:
        loss = bce_with_logits(x, target)
        tmask = m["train"] if masked else np.ones(A)
        train_loss = train_loss + (loss * tmask).sum(axis=1).mean()
        report_loss = report_loss + loss.sum(axis=1).mean()
    pred = np.argmax(logit if pred_logit is None else pred_logit, axis=1).astype(np.int32)  # first maximal index (tf.argmax)
    oh = np.zeros((B, A))
    oh[np.arange(B), pred] = 1.0
    te, ob, at, ex, tm = m["test"], m["obj"], m["attr"], m["exist"], m["train"]
    ps = {
        "all_score": (oh * target).sum(1),
        "max_train_score": (target * tm).max(1),
        "test_obj_score": (oh * target * te * ob).sum(1),
        "test_obj_max_score": (target * te * ob).max(1),
        "test_attr_score": (oh * target * te * at).sum(1),
        "test_attr_max_score": (target * te * at).max(1),
    }
    acc = ps["all_score"].mean()
    exist_acc = (oh * target * ex).sum(1).mean()
    test_acc = (oh * target * te).sum(1).mean()
    test_obj_acc = ps["test_obj_score"].mean()
    test_attr_acc = ps["test_attr_score"].mean()
    train_exist_acc = (oh * target * ex * tm).sum(1).mean()
    max_exist = (target * ex).max(1).mean()
    max_train_exist = (target * ex * tm).max(1).mean()
    test_obj_max = ps["test_obj_max_score"].mean()
    test_attr_max = ps["test_attr_max_score"].mean()
    test_max = (target * te).max(1).mean()
    test_max_exist = (target * ex * te).max(1).mean()
    if use_train_mask:
        report = {"answer_train_loss": train_loss, "answer_report_loss": report_loss}
    else:  # model_standard reports a single loss (vqa/model_standard.py:283-285, 362)
        report = {"answer_train_loss": train_loss, "answer_report_loss": report_loss}
    report.update({
        "answer_acc": acc, "exist_acc": exist_acc, "test_acc": test_acc,
        "normal_test_acc": _normal(test_acc, test_max),
        "normal_test_object_acc": _normal(test_obj_acc, test_obj_max),
        "normal_test_attribute_acc": _normal(test_attr_acc, test_attr_max),
        "normal_exist_acc": _normal(exist_acc, max_exist),
        "normal_train_exist_acc": _normal(train_exist_acc, max_train_exist),
        "max_exist_acc": max_exist, "test_max_acc": test_max, "test_max_exist_acc": test_max_exist,
    })
    return train_loss, report, ps, pred


# ------------------------------------------------------------------------------------------------
# the graph
# ------------------------------------------------------------------------------------------------
GEMM_WEIGHTS = ("v_w", "gru_gates_w", "gru_cand_w", "qv_w", "pl_w", "ql_w", "joint_w", "ans_w", "qp_w", "jl_w", "al_w",
                "qs_w", "tw_w", "va_w")


def forward(p, features, num_boxes, batch, m, variant="vlmap_answer", keep_att=0.8, keep_joint=0.5,
            att_mask=None, joint_mask=None, operand_round=None, joint_l_mask=None, noise=None,
            num_marginal=NUM_MARGINAL, ent_mask=None, ent_tile=None):
    """Model.build() forward. p: dict field -> fp64 array (TF layout [in,out]).
    features [N,K,Dv], num_boxes [N]; batch: image_idx [B], q_intseq [B,T], q_intseq_len [B],
    answer_target [B,A]; att_mask [B,K,D] / joint_mask [B,J] are the 0/1 keep masks tf.nn.dropout
    would have drawn (None = keep everything but still scale by 1/keep as TF does with a mask of 1s).
    operand_round: None = the reference's plain arithmetic; round_bf16 = the CUDA path's bf16 mode (GEMM
    operands and the stored pre-LN v-projection rounded to bf16, see round_bf16).
    Returns (outputs dict, cache for backward)."""
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    q = operand_round or _ident
    p = {k: f64(v) for k, v in p.items()}
    if operand_round is not None:
        p.update({k: q(p[k]) for k in GEMM_WEIGHTS if k in p})
    idx = np.asarray(batch["image_idx"])
    V = q(f64(features)[idx])                                # model_vlmap_answer.py:110-117
    nbox = np.asarray(num_boxes)[idx].astype(np.int64)       # :118-119
    q_ids = np.asarray(batch["q_intseq"])
    q_len = np.asarray(batch["q_intseq_len"])
    target = f64(batch["answer_target"])
    B, K, Dv = V.shape
    W = p["embed"].shape[1]

    Hv, v_cache = fc_ln_relu_fwd(V, p["v_w"], p["v_b"], p["v_gamma"], p["v_beta"], qz=q)  # :126-129 (LN over K*D)
    E = q(p["embed"][q_ids])                                                         # :134
    qs, gru_steps = gru_fwd(E, q_len, p["gru_gates_w"], p["gru_gates_b"], p["gru_cand_w"], p["gru_cand_b"], q=q)
    Hq, q_cache = fc_ln_relu_fwd(qs, p["qv_w"], p["qv_b"], p["qv_gamma"], p["qv_beta"], q=q)  # :142-145

    # hadamard_attention (modules.py:80-97)
    D = Hv.shape[-1]
    am = np.ones((B, K, D)) if att_mask is None else f64(att_mask)
    F = Hv * Hq[:, None, :] * am / keep_att                  # tf.nn.dropout(score_feat, 0.8)
    s = F @ p["att_w"].reshape(D) + p["att_b"].reshape(())
    box_valid = np.arange(K)[None, :] < nbox[:, None]        # tf.sequence_mask
    s = np.where(box_valid, s, -np.inf)
    smax = s.max(axis=1, keepdims=True)
    e = np.exp(s - smax)
    a = e / e.sum(axis=1, keepdims=True)                     # exact zeros at masked slots
    va_cache, Vp = None, V
    if variant == "vlmap_answer_adapt":   # model_vlmap_answer_adapt.py:132-142: pool relu(LN(FC(V))) instead of V
        Vp, va_cache = fc_ln_relu_fwd(V, p["va_w"], p["va_b"], p["va_gamma"], p["va_beta"], qz=q)
        Vp = q(Vp)                       # the device stores v_adapt as the pooling kernel's operand plane
    P = np.einsum("bk,bkd->bd", a, Vp)                       # attention_pooling of the RAW features

    Hp, p_cache = fc_ln_relu_fwd(P, p["pl_w"], p["pl_b"], p["pl_gamma"], p["pl_beta"], q=q)   # :163-167
    # the variants put one more layer between the GRU state and q_linear_l
    qp_cache, cond = None, qs
    ql_in = qs
    if variant == "vlmap_answer2":        # q_L_ft2 = tanh(LN(q W2 + b2)), model_vlmap_answer2.py:127-130
        xq = q(qs)
        zq, ln_q = layer_norm_fwd(xq @ p["qp_w"] + p["qp_b"], p["qp_gamma"], p["qp_beta"])
        ql_in = np.tanh(zq)
        qp_cache = (xq, ql_in, ln_q)
        cond = ql_in                      # heavy_output['condition'] = q_L_ft2 (:131)
    elif variant == "vlmap_answer_no_noise":   # q_L_mean = q Wm + bm, model_vlmap_answer_no_noise.py:122-125
        xq = q(qs)
        ql_in = xq @ p["qp_w"] + p["qp_b"]
        qp_cache = (xq, None, None)
    elif variant == "vlmap_answer_full":       # model_vlmap_answer_full.py:124-134
        xq = q(qs)
        mean = xq @ p["qp_w"] + p["qp_b"]
        lss = xq @ p["qs_w"] + p["qs_b"]
        sigma = np.sqrt(np.exp(lss))
        nz = np.zeros_like(mean) if noise is None else f64(noise)
        ql_in = mean + nz * sigma
        qp_cache = (xq, mean, (lss, sigma, nz))
    Hl, l_cache = fc_ln_relu_fwd(ql_in, p["ql_w"], p["ql_b"], p["ql_gamma"], p["ql_beta"], q=q)  # :170-174
    noc = variant in ("vlmap_answer_noc", "vlmap_answer_nocarch")
    jl_cache, jlm, Jld = None, None, None
    if not noc:
        X = Hp * Hl
        Jn, j_cache = fc_ln_relu_fwd(X, p["joint_w"], p["joint_b"], p["joint_gamma"], p["joint_beta"], q=q)
        jm = np.ones_like(Jn) if joint_mask is None else f64(joint_mask)
        Jd = q(Jn * jm / keep_joint)                             # :180
        logit = Jd @ p["ans_w"] + p["ans_b"]                     # :183-185 / model_standard.py:272-275
    else:
        # model_vlmap_answer_noc.py:177-203: no Hadamard; each branch has its own joint layer, dropout and
        # word-weight head, and the two logits are added
        X = Hp
        Jn, j_cache = fc_ln_relu_fwd(Hp, p["joint_w"], p["joint_b"], p["joint_gamma"], p["joint_beta"], q=q)   # joint_v
        jm = np.ones_like(Jn) if joint_mask is None else f64(joint_mask)
        Jd = q(Jn * jm / keep_joint)
        Jln, jl_cache = fc_ln_relu_fwd(Hl, p["jl_w"], p["jl_b"], p["jl_gamma"], p["jl_beta"], q=q)             # joint_l
        jlm = np.ones_like(Jln) if joint_l_mask is None else f64(joint_l_mask)
        Jld = q(Jln * jlm / keep_joint)
        logit = (Jd @ p["ans_w"] + p["ans_b"]) + (Jld @ p["al_w"] + p["al_b"])

    use_tm = variant != "standard"
    tuned_cache = None
    if variant in ("vlmap_answer_vqa_all", "vlmap_answer_vqa_all2"):
        logit0 = logit
        ex = m["exist"]
        if variant == "vlmap_answer_vqa_all":   # _vqa_all.py:192-194: absent answers take the row minimum
            mn = logit0.min(axis=1, keepdims=True)
            L1 = logit0 * ex + mn * (1.0 - ex)
        else:
            L1 = logit0
        tuned = Jd @ p["tw_w"] + p["tw_b"]      # fc_layer(joint, ...) -- `joint`, not tuned_joint (:215-216)
        logit = L1 + tuned                      # output['logit'] (:225)
        if variant == "vlmap_answer_vqa_all":
            terms, pred_logit = [(L1, True), (logit, True)], logit
        else:
            terms = [(L1, True), (tuned, False)]
            pred_logit = L1 * m["test"] + tuned * m["train"]
        tuned_cache = (logit0, L1, tuned)
        train_loss, report, ps, pred = metrics(logit, target, m, loss_terms=terms, pred_logit=pred_logit)
    else:
        train_loss, report, ps, pred = metrics(logit, target, m, use_train_mask=use_tm)
    total_loss = train_loss
    ent_cache = None
    if variant == "vlmap_answer_ent":           # model_vlmap_answer_ent.py:193-213, 284-294
        M = int(num_marginal)
        # tf.reshape(tf.tile(stop_gradient(Hp), [M, 1]), [-1, M, L]): entry [b, m] is row (b*M + m) mod B of Hp
        tidx = (np.arange(B)[:, None] * M + np.arange(M)[None, :]) % B
        # tf.stop_gradient: the tile is a constant of the gradient. ent_tile lets the finite-difference check hold it at
        # its unperturbed value, which is what "no gradient through this path" means for a difference quotient.
        TP = Hp[tidx] if ent_tile is None else f64(ent_tile)              # [B, M, L], no gradient
        X2 = TP * Hl[:, None, :]
        Jn2, j2_cache = fc_ln_relu_fwd(X2, p["joint_w"], p["joint_b"], p["joint_gamma"], p["joint_beta"], q=q, qz=q)
        jm2 = np.ones_like(Jn2) if ent_mask is None else f64(ent_mask)    # LN over (M, J): SURVEY Q1
        Jd2 = q(Jn2 * jm2 / keep_joint)
        logit2 = Jd2 @ p["ans_w"] + p["ans_b"]                            # [B, M, A]
        sel = (m["exist"] * m["train"]) > 0.5                             # train_exist_answer_mask_bool
        ml = logit2[:, :, sel]
        e2 = np.exp(ml - ml.max(axis=-1, keepdims=True))
        prob = e2 / e2.sum(axis=-1, keepdims=True)
        marg = prob.mean(axis=1)                                          # [B, #selected]
        neg_ent = (marg * np.log(marg + 1e-8)).sum(axis=-1).mean()
        report["entropy"] = neg_ent
        report["weighted_entropy"] = W_ENTROPY * neg_ent
        total_loss = train_loss + W_ENTROPY * neg_ent
        ent_cache = dict(TP=TP, j2_cache=j2_cache, jm2=jm2, Jd2=Jd2, sel=sel, prob=prob, marg=marg, M=M)
    if variant == "vlmap_answer_full":          # _full.py:217-223, 272-276
        _, mean, (lss, _, _) = qp_cache
        latent = -0.5 * (1.0 + lss - mean ** 2 - np.exp(lss)).sum(axis=1).mean()
        report["latent_loss"] = latent
        report["train_latent_loss"] = LATENT_LOSS_WEIGHT * latent
        total_loss = train_loss + LATENT_LOSS_WEIGHT * latent
    out = {"loss": total_loss, "report": report, "att_score": a, "logit": logit, "pred": pred,
           "per_sample": ps, "condition": cond, "pooled": P}
    cache = dict(V=V, nbox=nbox, q_ids=q_ids, q_len=q_len, target=target, v_cache=v_cache, Hv=Hv,
                 gru_steps=gru_steps, q=qs, q_cache=q_cache, Hq=Hq, am=am, F=F, a=a, P=P, p_cache=p_cache,
                 Hp=Hp, l_cache=l_cache, Hl=Hl, X=X, j_cache=j_cache, jm=jm, Jd=Jd, logit=logit,
                 keep_att=keep_att, keep_joint=keep_joint, use_tm=use_tm, m=m, W=W, p=p, variant=variant,
                 qp_cache=qp_cache, jl_cache=jl_cache, jlm=jlm, Jld=Jld, noc=noc, tuned_cache=tuned_cache,
                 va_cache=va_cache, Vp=Vp, ent_cache=ent_cache)
    return out, cache


def backward(cache, loss_scale=1.0, intermediates=None, gate_flips=None, gate_override=None):
    """Gradients of loss_scale * train_loss w.r.t. every parameter (dict field -> array).
    Callers drop the frozen ones (trainable_fields). If `intermediates` is a dict it receives the
    activation gradients the per-kernel parity tests compare against (dP, dHq, dZv, dlogit)."""
    c = cache
    p = c["p"]
    gf = gate_flips or {}
    go = gate_override or {}
    B, A = c["logit"].shape
    tmask = c["m"]["train"] if c["use_tm"] else np.ones(A)
    g = {}
    if c.get("tuned_cache") is not None:
        logit0, L1, tuned = c["tuned_cache"]
        z, ex = c["target"], c["m"]["exist"]
        if c["variant"] == "vlmap_answer_vqa_all":
            d_tuned = (sigmoid(L1 + tuned) - z) * tmask / B * loss_scale
            dL1 = d_tuned + (sigmoid(L1) - z) * tmask / B * loss_scale
            # tf.reduce_min gradient: shared equally by the minimal entries (math_grad._MinOrMaxGrad)
            ind = (logit0 == logit0.min(axis=1, keepdims=True)).astype(np.float64)
            dmin = (dL1 * (1.0 - ex)).sum(axis=1, keepdims=True)
            dx = dL1 * ex + ind / ind.sum(axis=1, keepdims=True) * dmin
        else:
            dx = (sigmoid(L1) - z) * tmask / B * loss_scale
            d_tuned = (sigmoid(tuned) - z) / B * loss_scale
        g["tw_w"] = c["Jd"].T @ d_tuned
        g["tw_b"] = d_tuned.sum(0)
    else:
        dx = (sigmoid(c["logit"]) - c["target"]) * tmask / B * loss_scale
        d_tuned = None
    g["ans_w"] = c["Jd"].T @ dx
    g["ans_b"] = dx.sum(0)
    dJd = dx @ p["ans_w"].T
    if d_tuned is not None:
        dJd = dJd + d_tuned @ p["tw_w"].T
    dJn = dJd * c["jm"] / c["keep_joint"]
    dX, g["joint_w"], g["joint_b"], g["joint_gamma"], g["joint_beta"] = fc_ln_relu_bwd(
        dJn, p["joint_w"], p["joint_gamma"], c["j_cache"], flip=gf.get("joint"), gate=go.get("joint"))
    if not c.get("noc"):
        dHp, dHl = dX * c["Hl"], dX * c["Hp"]
        if c.get("ent_cache") is not None:
            # d(0.1 * mean_b sum_a marg log(marg + 1e-8)) back through the marginal softmax, the tiled head and the
            # broadcast Hl; the tiled Hp carries tf.stop_gradient
            ec = c["ent_cache"]
            marg, prob, M = ec["marg"], ec["prob"], ec["M"]
            dmarg = W_ENTROPY * loss_scale / B * (np.log(marg + 1e-8) + marg / (marg + 1e-8))
            dprob = np.broadcast_to(dmarg[:, None, :] / M, prob.shape)
            dml = prob * (dprob - (prob * dprob).sum(axis=-1, keepdims=True))
            dl2 = np.zeros(prob.shape[:2] + (A,))
            dl2[:, :, ec["sel"]] = dml
            g["ans_w"] = g["ans_w"] + ec["Jd2"].reshape(-1, ec["Jd2"].shape[-1]).T @ dl2.reshape(-1, A)
            g["ans_b"] = g["ans_b"] + dl2.sum(axis=(0, 1))
            dJn2 = (dl2 @ p["ans_w"].T) * ec["jm2"] / c["keep_joint"]
            dX2, gw, gb, gg, gbt = fc_ln_relu_bwd(dJn2, p["joint_w"], p["joint_gamma"], ec["j2_cache"],
                                                  flip=gf.get("joint2"), gate=go.get("joint2"))
            g["joint_w"] = g["joint_w"] + gw
            g["joint_b"] = g["joint_b"] + gb
            g["joint_gamma"] = g["joint_gamma"] + gg
            g["joint_beta"] = g["joint_beta"] + gbt
            dHl = dHl + (dX2 * ec["TP"]).sum(axis=1)
    else:
        dHp = dX
        g["al_w"] = c["Jld"].T @ dx
        g["al_b"] = dx.sum(0)
        dJln = (dx @ p["al_w"].T) * c["jlm"] / c["keep_joint"]
        dHl, g["jl_w"], g["jl_b"], g["jl_gamma"], g["jl_beta"] = fc_ln_relu_bwd(
            dJln, p["jl_w"], p["jl_gamma"], c["jl_cache"], flip=gf.get("jl"), gate=go.get("jl"))
    dP, g["pl_w"], g["pl_b"], g["pl_gamma"], g["pl_beta"] = fc_ln_relu_bwd(
        dHp, p["pl_w"], p["pl_gamma"], c["p_cache"], flip=gf.get("pl"), gate=go.get("pl"))
    dq, g["ql_w"], g["ql_b"], g["ql_gamma"], g["ql_beta"] = fc_ln_relu_bwd(
        dHl, p["ql_w"], p["ql_gamma"], c["l_cache"], flip=gf.get("ql"), gate=go.get("ql"))
    if c.get("qp_cache") is not None:     # back through q_L_ft2 / q_L_mean: dq is the gradient of ITS output so far
        xq, yq, ln_q = c["qp_cache"]
        if c["variant"] == "vlmap_answer2":
            dzq, g["qp_gamma"], g["qp_beta"] = layer_norm_bwd(dq * (1.0 - yq * yq), p["qp_gamma"], ln_q)
        elif c["variant"] == "vlmap_answer_full":
            # ql_in = mean + noise * exp(lss / 2); KL = -0.5 mean_b sum(1 + lss - mean^2 - exp(lss)), weight 0.1
            mean, (lss, sigma, nz) = yq, ln_q
            kw = LATENT_LOSS_WEIGHT * loss_scale / B
            dzq = dq + kw * mean
            dlss = dq * nz * sigma * 0.5 - 0.5 * kw * (1.0 - np.exp(lss))
            g["qs_w"] = xq.T @ dlss
            g["qs_b"] = dlss.sum(axis=0)
        else:
            dzq = dq
        g["qp_w"] = xq.T @ dzq
        g["qp_b"] = dzq.sum(axis=0)
        dq = dzq @ p["qp_w"].T
        if c["variant"] == "vlmap_answer_full":
            dq = dq + dlss @ p["qs_w"].T
    # attention pooling + softmax + score
    V, a = c["V"], c["a"]
    da = np.einsum("bkd,bd->bk", c.get("Vp", V), dP)
    if c.get("va_cache") is not None:   # adapt: the pooled tensor is a trained layer's output: d v_adapt = a (x) dP
        dVp = a[:, :, None] * dP[:, None, :]
        _, g["va_w"], g["va_b"], g["va_gamma"], g["va_beta"] = fc_ln_relu_bwd(
            dVp, p["va_w"], p["va_gamma"], c["va_cache"], need_dx=False, flip=gf.get("va"), gate=go.get("va"))
    ds = a * (da - (a * da).sum(axis=1, keepdims=True))      # masked slots: a = 0 -> ds = 0
    D = c["Hv"].shape[-1]
    w = p["att_w"].reshape(D)
    g["att_b"] = np.array([ds.sum()])
    g["att_w"] = np.einsum("bk,bkd->d", ds, c["F"]).reshape(D, 1)
    dF = ds[:, :, None] * w[None, None, :] * c["am"] / c["keep_att"]
    dHv = dF * c["Hq"][:, None, :]
    dHq = (dF * c["Hv"]).sum(axis=1)
    _, g["v_w"], g["v_b"], g["v_gamma"], g["v_beta"] = fc_ln_relu_bwd(
        dHv, p["v_w"], p["v_gamma"], c["v_cache"], need_dx=False, flip=gf.get("v"), gate=go.get("v"))   # V is data: no dV
    if intermediates is not None:
        _, y_v, ln_v = c["v_cache"]
        dZv, _, _ = layer_norm_bwd(dHv * (y_v > 0), p["v_gamma"], ln_v)
        intermediates.update(dlogit=dx, dP=dP, dHq=dHq, dZv=dZv, ds=ds, dHv=dHv)
    dq2, g["qv_w"], g["qv_b"], g["qv_gamma"], g["qv_beta"] = fc_ln_relu_bwd(
        dHq, p["qv_w"], p["qv_gamma"], c["q_cache"], flip=gf.get("qv"), gate=go.get("qv"))
    dq = dq + dq2
    dE, g["gru_gates_w"], g["gru_gates_b"], g["gru_cand_w"], g["gru_cand_b"] = gru_bwd(
        dq, c["gru_steps"], p["gru_gates_w"], p["gru_cand_w"], c["W"])
    demb = np.zeros_like(p["embed"])
    np.add.at(demb, c["q_ids"], dE)                          # gradient of embedding_lookup
    g["embed"] = demb
    if intermediates is not None:
        intermediates["dE"] = dE                             # the IndexedSlices values TF's global norm sees
    return g


RELU_LAYERS = {"v": "v_cache", "qv": "q_cache", "pl": "p_cache", "ql": "l_cache", "joint": "j_cache", "jl": "jl_cache",
               "va": "va_cache"}


def relu_near_ties(cache, tau):
    """Indices of ReLU pre-activations with |y| < tau, per layer: gates a working-precision run cannot
    be expected to reproduce."""
    return {name: (np.argwhere(np.abs(cache[key][1]) < tau) if cache.get(key) is not None else np.zeros((0, 2), int))
            for name, key in RELU_LAYERS.items()}


def v_layer_tie_delta(cache, dHv, idx):
    """Exact change of (v_w, v_b, v_gamma, v_beta) when the gate of v-projection element idx=(b,k,d) is
    flipped. V is data, so nothing else depends on that gate; only sample b's LayerNorm slab changes."""
    p = cache["p"]
    x, y, (xhat, rstd) = cache["v_cache"]
    b, k, d = (int(i) for i in idx)
    sign = -1.0 if y[b, k, d] > 0 else 1.0          # gate on -> off removes the term, off -> on adds it
    dy = np.zeros_like(y[b:b + 1])
    dy[0, k, d] = sign * dHv[b, k, d]
    dz, dgamma, dbeta = layer_norm_bwd(dy, p["v_gamma"], (xhat[b:b + 1], rstd[b:b + 1]))
    return {"v_w": x[b].T @ dz[0], "v_b": dz[0].sum(axis=0), "v_gamma": dgamma, "v_beta": dbeta}


# ------------------------------------------------------------------------------------------------
# optimizer (vqa/trainer.py:106-114): clip_by_global_norm(20) over train vars, then Adam
# ------------------------------------------------------------------------------------------------
def clip_adam_step(params, grads, m, v, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, clip=20.0, slice_sumsq=None):
    """params/grads/m/v: dict field -> array over the trainable set; t = 1-based step. In place.
    slice_sumsq: dict field -> sum of squares of that variable's gradient AS TF HOLDS IT. The gradient of
    tf.nn.embedding_lookup on a variable is an IndexedSlices (values [B*T, W], one row per token occurrence), and
    clip_ops.global_norm takes `t.values` of an IndexedSlices as they are -- duplicates NOT summed -- so optimize_loss'
    clip_by_global_norm(20.0) sees sum_{b,t} |dE[b,t]|^2 for LearnGloVe/embed_map, not the norm of the scattered dense
    gradient (vqa/trainer.py:106-114 -> tf.contrib.layers.optimize_loss -> clip_ops.clip_by_global_norm). The Adam update
    itself sums duplicate indices first (Optimizer._apply_sparse_duplicate_indices) and decays m, v of every row, i.e.
    it equals the dense update below."""
    ss = {k: float((g ** 2).sum()) for k, g in grads.items()}
    if slice_sumsq:
        ss.update({k: float(x) for k, x in slice_sumsq.items()})
    gnorm = np.sqrt(sum(ss.values()))
    scale = clip / max(gnorm, clip)
    lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    for k in grads:
        g = grads[k] * scale
        m[k] = beta1 * m[k] + (1.0 - beta1) * g
        v[k] = beta2 * v[k] + (1.0 - beta2) * g * g
        params[k] = params[k] - lr_t * m[k] / (np.sqrt(v[k]) + eps)
    return gnorm
