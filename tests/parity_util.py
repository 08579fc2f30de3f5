"""Shared helpers of the GPU parity tests: run the CUDA path through the C ABI (Engine) and the NumPy
oracle on the same seeded inputs and the same dropout masks."""
import numpy as np

from oracle import answer_model_np as O
from vqa_transfer_externaldata_b200 import synthetic as S
from vqa_transfer_externaldata_b200.engine import AnswerModelConfig, Engine


def rel_err(x, ref):
    """||x - ref||_inf / max(||ref||_inf, tiny): the definition SURVEY 8d fixes for the parity gates."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-30))


def rel_l2(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-30))


def relu_tie_budget(cache, inter, ref_g, fields, tau, loss_scale=1.0, max_ties=96, layers=None,
                    gate_override=None):
    """Per-tensor elementwise budget sum_e |delta_e| over ReLU gates whose oracle pre-activation is within
    tau of zero. Gradients are linear in the gates (forward values are unaffected: relu(y) ~ 0 there), so a
    run that decides those gates differently lands within ref +- budget. Returns (budget dict, n_ties)."""
    ties = O.relu_near_ties(cache, tau)
    if layers is not None:  # gates of the other layers were taken from the device: nothing to bound there
        ties = {k: (v if k in layers else v[:0]) for k, v in ties.items()}
    budget = {f: np.zeros_like(ref_g[f]) for f in fields}
    n = sum(len(v) for v in ties.values())
    if n > max_ties:
        raise AssertionError(f"{n} near-tie ReLU gates at tau={tau}: too many to bound")
    for idx in ties["v"]:
        for f, dlt in O.v_layer_tie_delta(cache, inter["dHv"], idx).items():
            if f in budget:
                budget[f] += np.abs(dlt)
    for layer in ("qv", "pl", "ql", "joint", "jl", "va"):
        for idx in ties[layer]:
            flip = np.zeros(cache[O.RELU_LAYERS[layer]][1].shape, dtype=bool)
            flip[tuple(idx)] = True
            g2 = O.backward(cache, loss_scale=loss_scale, gate_flips={layer: flip}, gate_override=gate_override)
            for f in fields:
                budget[f] += np.abs(g2[f] - ref_g[f])
    return budget, n


def build_case(dims, variant="vlmap_answer", precision="fp32", seed=0, num_images=24, ragged=True,
               batch=None, T=None, perturb=0.2, keep_att=0.8, keep_joint=0.5, num_marginal=5):
    c = S.dims(**dims)
    cfg = AnswerModelConfig(variant=variant, precision=precision, keep_att=keep_att, keep_joint=keep_joint,
                            num_marginal=num_marginal, **c)
    params, exist = S.init_params(c, seed=seed, variant=variant, perturb=perturb)
    feats, nb = S.make_bank(c, num_images=num_images, seed=seed + 2, ragged_boxes=ragged)
    bt = S.make_batch(c, num_images, seed=seed + 3, batch=batch, T=T)
    is_obj, is_attr = S.make_answer_flags(c)
    eng = Engine(cfg)
    eng.set_feature_bank(feats, nb)
    eng.set_answer_masks(is_obj, is_attr, exist)
    eng.load_params(params)
    m = O.answer_masks(c["A"], c["num_train_answer"], is_obj, is_attr, exist)
    return dict(c=c, cfg=cfg, params=params, feats=feats, nb=nb, batch=bt, eng=eng, m=m)


def run_both(case, seed=777, step=3, loss_scale=1.0, emulate=None, device_gates=None, gate_tie=None):
    """emulate: None = restate the device's operand rounding when the case runs in bf16 mode (the oracle then
    makes the same ReLU gate decisions), False = always the reference's plain arithmetic.
    device_gates: take the ReLU gates of the 2-D heads from the device (every gate that differs from the oracle's own
    decision must be a near-tie: |pre-activation| < gate_tie). Default: with `emulate`. At B 512 even the fp32 mode
    meets a few dozen undecidable gates among 2.6 M, each of which would cost one more oracle backward to bound."""
    import torch
    eng, cfg = case["eng"], case["cfg"]
    eng.stage_batch(case["batch"])
    eng.forward(seed=seed, step=step)
    eng.backward(loss_scale=loss_scale)
    att_mask, joint_mask = eng.dropout_masks(seed, step)
    noc = cfg.variant in ("vlmap_answer_noc", "vlmap_answer_nocarch")
    jl_kw = {}
    if noc:  # the language branch draws its own dropout mask (a second tf.nn.dropout call)
        from vqa_transfer_externaldata_b200 import lib as L_
        jl_mask = eng.dropout_mask_site(L_.SITE_JOINT_L, seed, step)
        jl_kw = {"joint_l_mask": jl_mask.cpu().numpy()}
    if cfg.variant == "vlmap_answer_full":   # the reparameterisation noise the device drew for (seed, step)
        jl_kw = {"noise": eng.reparam_noise(seed, step).cpu().numpy()}
    if cfg.variant == "vlmap_answer_ent":    # the tiled joint's own dropout mask
        from vqa_transfer_externaldata_b200 import lib as L_
        jl_kw = {"num_marginal": cfg.num_marginal,
                 "ent_mask": eng.dropout_mask_site(L_.SITE_ENT, seed, step).cpu().numpy()}
    torch.cuda.synchronize()
    loss, report = eng.read_scalars()
    got = {"loss": loss, "report": report}
    got.update({k: v.detach().cpu().numpy() for k, v in eng.outputs().items()})
    got["condition"] = eng.o_condition[:eng.batch_size].cpu().numpy()
    got["pooled"] = eng.o_pooled[:eng.batch_size].cpu().numpy()
    got["grads"] = {f: g.detach().cpu().numpy() for f, g in eng.params.grad_views.items()}
    if emulate is None:
        emulate = cfg.precision == "bf16"
    out, cache = O.forward(case["params"], case["feats"], case["nb"], case["batch"], case["m"],
                           variant=cfg.variant, keep_att=cfg.keep_att, keep_joint=cfg.keep_joint,
                           att_mask=att_mask.cpu().numpy(), joint_mask=joint_mask.cpu().numpy(),
                           operand_round=O.round_bf16 if emulate else None, **jl_kw)
    if emulate:  # the reference's own arithmetic, for the forward gates of the north star
        case["plain_out"], _ = O.forward(case["params"], case["feats"], case["nb"], case["batch"], case["m"],
                                         variant=cfg.variant, keep_att=cfg.keep_att, keep_joint=cfg.keep_joint,
                                         att_mask=att_mask.cpu().numpy(), joint_mask=joint_mask.cpu().numpy(), **jl_kw)
    inter = {}
    gates = None
    if device_gates is None:
        device_gates = emulate
    if gate_tie is None:
        gate_tie = 2e-2 if emulate else 1e-4
    if device_gates:
        # bf16 mode: the backward pass is linear in the ReLU gates, and a gate whose pre-activation is zero to
        # bf16 working precision is decided by rounding noise. Take the four 2-D heads' gates from the DEVICE
        # (its saved post-ReLU activations), account for every gate that differs from the oracle's own
        # decision, and require those to be true near-ties.
        from vqa_transfer_externaldata_b200 import lib as L
        Bn = eng.batch_size
        hq = eng.peek_activation(L.ACT_HQ, torch.float32, (Bn, cfg.D)).cpu().numpy()
        hl = eng.peek_activation(L.ACT_HL, torch.float32, (Bn, cfg.L)).cpu().numpy()
        hp = eng.peek_activation(L.ACT_HP, torch.float32, (Bn, cfg.L)).cpu().numpy()
        jd = eng.peek_activation(L.ACT_JD, torch.bfloat16, (Bn, cfg.J)).float().cpu().numpy()
        gates = {"qv": hq > 0, "ql": hl > 0, "pl": hp > 0, "joint": (jd > 0) | (joint_mask.cpu().numpy() == 0)}
        drop = {"joint": joint_mask.cpu().numpy() == 0}
        if noc:
            jdl = eng.peek_activation(L.ACT_JDL, torch.bfloat16, (Bn, cfg.J)).float().cpu().numpy()
            gates["jl"] = (jdl > 0) | (jl_kw["joint_l_mask"] == 0)
            drop["jl"] = jl_kw["joint_l_mask"] == 0
        if cfg.variant == "vlmap_answer_adapt":   # v_adapt's gates from the device's own pooling operand
            va = eng.peek_activation(L.ACT_VA, torch.bfloat16, (Bn, cfg.K, cfg.D)).float().cpu().numpy()
            gates["va"] = va > 0
        n_diff, n_all = 0, 0
        for layer, gate in gates.items():
            y = cache[O.RELU_LAYERS[layer]][1]
            own = y > 0
            if layer in drop:
                own = own | drop[layer]
            diff = own != gate
            n_diff += int(diff.sum())
            n_all += diff.size
            # a gate the device decided differently must be a near-tie of the oracle (|y| small vs O(1) LN output)
            assert not diff.any() or np.abs(y[diff]).max() < gate_tie, (layer, np.abs(y[diff]).max())
        assert n_diff <= max(2, 2e-3 * n_all), (n_diff, n_all)
        case["gate_diffs"] = (n_diff, n_all)
    ref_g = O.backward(cache, loss_scale=loss_scale, intermediates=inter, gate_override=gates)
    case["oracle_cache"], case["oracle_inter"], case["loss_scale"] = cache, inter, loss_scale
    case["gate_override"] = gates
    return got, out, ref_g
