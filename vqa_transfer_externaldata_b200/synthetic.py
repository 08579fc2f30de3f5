"""Synthetic inputs and random-init weights of the reference's shapes: there is no
network, so no VQA v2 features / GloVe / pretrained vlmap checkpoints. Pure NumPy, seeded.

Initialisers follow TF-1.6 defaults at the reference's call sites:
  layers.fully_connected -> Xavier-uniform weights, zero biases   (vlmap/modules.py:634-641)
  layers.layer_norm      -> gamma 1, beta 0                       (vlmap/modules.py:647)
  rnn.GRUCell            -> Glorot-uniform kernels, gate bias 1.0, candidate bias 0 (modules.py:131)
  WordWeightAnswer       -> columns copied from exported class_weights by answer string, absent
                            answers: weight 0, bias -100          (vlmap/modules.py:600-614)
"""
import numpy as np

PARAM_SHAPES = {
    "embed": lambda c: (c["Vq"], c["W"]),
    "v_w": lambda c: (c["Dv"], c["D"]), "v_b": lambda c: (c["D"],),
    "v_gamma": lambda c: (c["D"],), "v_beta": lambda c: (c["D"],),
    "gru_gates_w": lambda c: (c["W"] + c["L"], 2 * c["L"]), "gru_gates_b": lambda c: (2 * c["L"],),
    "gru_cand_w": lambda c: (c["W"] + c["L"], c["L"]), "gru_cand_b": lambda c: (c["L"],),
    "qv_w": lambda c: (c["L"], c["D"]), "qv_b": lambda c: (c["D"],),
    "qv_gamma": lambda c: (c["D"],), "qv_beta": lambda c: (c["D"],),
    "att_w": lambda c: (c["D"], 1), "att_b": lambda c: (1,),
    # adapt pools the 1024-d v_adapt instead of the raw features (model_vlmap_answer_adapt.py:142-150)
    "pl_w": lambda c: (c["D"] if c.get("variant") == "vlmap_answer_adapt" else c["Dv"], c["L"]), "pl_b": lambda c: (c["L"],),
    "pl_gamma": lambda c: (c["L"],), "pl_beta": lambda c: (c["L"],),
    "ql_w": lambda c: (c["L"], c["L"]), "ql_b": lambda c: (c["L"],),
    "ql_gamma": lambda c: (c["L"],), "ql_beta": lambda c: (c["L"],),
    "joint_w": lambda c: (c["L"], c["J"]), "joint_b": lambda c: (c["J"],),
    "joint_gamma": lambda c: (c["J"],), "joint_beta": lambda c: (c["J"],),
    "ans_w": lambda c: (c["J"], c["A"]), "ans_b": lambda c: (c["A"],),
    # extra question layer of vlmap_answer2 (q_L_ft2) / vlmap_answer_no_noise (q_L_mean); V_DIM == L_DIM there
    "qp_w": lambda c: (c["L"], c["L"]), "qp_b": lambda c: (c["L"],),
    "qp_gamma": lambda c: (c["L"],), "qp_beta": lambda c: (c["L"],),
    # second branch of vlmap_answer_noc / nocarch: joint_l and WordWeightAnswerL
    "jl_w": lambda c: (c["L"], c["J"]), "jl_b": lambda c: (c["J"],),
    "jl_gamma": lambda c: (c["J"],), "jl_beta": lambda c: (c["J"],),
    "al_w": lambda c: (c["J"], c["A"]), "al_b": lambda c: (c["A"],),
    # vlmap_answer_full: q_L_log_sigma_sq next to q_L_mean (= qp_*)
    "qs_w": lambda c: (c["L"], c["L"]), "qs_b": lambda c: (c["L"],),
    # vlmap_answer_vqa_all / vqa_all2: TunedWordWeightAnswer
    "tw_w": lambda c: (c["J"], c["A"]), "tw_b": lambda c: (c["A"],),
    # vlmap_answer_adapt: v_adapt
    "va_w": lambda c: (c["Dv"], c["D"]), "va_b": lambda c: (c["D"],),
    "va_gamma": lambda c: (c["D"],), "va_beta": lambda c: (c["D"],),
}
_NOC = ("jl_w", "jl_b", "jl_gamma", "jl_beta", "al_w", "al_b")
EXTRA_FIELDS = {"vlmap_answer2": ("qp_w", "qp_b", "qp_gamma", "qp_beta"), "vlmap_answer_no_noise": ("qp_w", "qp_b"),
                "vlmap_answer_noc": _NOC, "vlmap_answer_nocarch": _NOC,
                "vlmap_answer_full": ("qp_w", "qp_b", "qs_w", "qs_b"),
                "vlmap_answer_vqa_all": ("tw_w", "tw_b"), "vlmap_answer_vqa_all2": ("tw_w", "tw_b"),
                "vlmap_answer_adapt": ("va_w", "va_b", "va_gamma", "va_beta")}
_EXTRA_PREFIXES = ("qp_", "jl_", "al_", "qs_", "tw_", "va_")


def dims(B=512, K=36, Dv=2048, D=1024, L=1024, J=None, A=3000, T=14, W=300, Vq=8192,
         num_train_answer=None):
    J = 2 * L if J is None else J
    nta = (A * 3) // 4 if num_train_answer is None else num_train_answer
    return dict(B=B, K=K, Dv=Dv, D=D, L=L, J=J, A=A, T=T, W=W, Vq=Vq, num_train_answer=nta)


def _xavier(rng, shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_params(c, seed=4321, variant="vlmap_answer", perturb=0.0, present_frac=0.8):
    """dict field -> fp32 array. perturb > 0 moves LN gamma/beta and biases off their init values so
    that parity tests exercise them (a trained checkpoint has non-trivial values there)."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, shp in PARAM_SHAPES.items():
        if name.startswith(_EXTRA_PREFIXES) and name not in EXTRA_FIELDS.get(variant, ()):
            continue
        shape = shp(dict(c, variant=variant))
        if name == "embed":
            p[name] = (rng.standard_normal(shape) * 0.4).astype(np.float32)  # GloVe-like scale
        elif name.endswith("_gamma"):
            p[name] = np.ones(shape, np.float32)
        elif name == "gru_gates_b":
            p[name] = np.ones(shape, np.float32)
        elif len(shape) == 2:
            p[name] = _xavier(rng, shape)
        else:
            p[name] = np.zeros(shape, np.float32)
    exist = np.ones(c["A"], np.float32)
    if variant != "standard":
        # synthetic exported word weights: class_weights [J, A'] with `present_frac` of answers present
        present = rng.uniform(size=c["A"]) < present_frac
        w = (rng.standard_normal((c["J"], c["A"])) * 0.05).astype(np.float32)
        b = (rng.standard_normal(c["A"]) * 0.5 - 3.0).astype(np.float32)
        w[:, ~present] = 0.0
        b[~present] = -100.0
        p["ans_w"], p["ans_b"] = w, b
        if "al_w" in p:   # the L head is remapped from l_class_weights of the same export: same absent answers
            wl = (rng.standard_normal((c["J"], c["A"])) * 0.05).astype(np.float32)
            bl = (rng.standard_normal(c["A"]) * 0.5).astype(np.float32)
            wl[:, ~present] = 0.0
            bl[~present] = -100.0
            p["al_w"], p["al_b"] = wl, bl
        exist = present.astype(np.float32)
    if perturb > 0:
        for name in p:
            if name.endswith("_gamma"):
                p[name] = (p[name] + perturb * rng.standard_normal(p[name].shape)).astype(np.float32)
            elif name.endswith(("_beta", "_b")) and name not in ("ans_b", "al_b"):
                p[name] = (p[name] + perturb * rng.standard_normal(p[name].shape)).astype(np.float32)
    return p, exist


def make_bank(c, num_images=64, seed=99, ragged_boxes=False):
    """image_features [N,K,Dv] (post-ReLU-like, non-negative) and num_boxes [N]."""
    rng = np.random.default_rng(seed)
    feats = (np.abs(rng.standard_normal((num_images, c["K"], c["Dv"]))) * 0.5).astype(np.float32)
    if ragged_boxes:
        nb = rng.integers(max(1, c["K"] // 10), c["K"] + 1, size=num_images).astype(np.int32)
        nb[0] = 1
        nb[-1] = c["K"]
        for i in range(num_images):  # padded rows are zero in the adaptive 10-100 feature files
            feats[i, nb[i]:] = 0.0
    else:
        nb = np.full(num_images, c["K"], np.int32)
    return feats, nb


def make_batch(c, num_images, seed=1234, batch=None, T=None):
    """batch dict keyed like vqa/datasets/input_ops_vqa_tf_record_memft.py:47-59"""
    rng = np.random.default_rng(seed)
    B = c["B"] if batch is None else batch
    T = c["T"] if T is None else T
    q_len = rng.integers(min(3, T), T + 1, size=B).astype(np.int32)
    q = rng.integers(1, c["Vq"], size=(B, T)).astype(np.int32)
    q[np.arange(T)[None, :] >= q_len[:, None]] = 0  # pad id 0
    target = np.zeros((B, c["A"]), np.float32)
    scores = np.array([0.3, 0.6, 0.9, 1.0], np.float32)  # get_score(), generator_tf_record_memft_genome.py:122-132
    for b in range(B):
        n = rng.integers(1, 4)
        ids = rng.choice(c["A"], size=n, replace=False)
        target[b, ids] = scores[rng.integers(0, 4, size=n)]
    return {
        "id": np.arange(B, dtype=np.int64),
        "image_idx": rng.integers(0, num_images, size=B).astype(np.int64),
        "q_intseq": q, "q_intseq_len": q_len, "answer_target": target,
    }


def make_answer_flags(c, seed=7):
    """is_object / is_attribute flags of answer_dict.pkl (construct_vocab_objattr_memft_genome.py:111-116)"""
    rng = np.random.default_rng(seed)
    kind = rng.integers(0, 3, size=c["A"])
    return (kind == 0).astype(np.float32), (kind == 1).astype(np.float32)
