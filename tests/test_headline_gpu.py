"""Parity at the sizes the metric is quoted on (BASELINE.json configs[0]/[1]: B 512 x K 36 x Dv 2048, T 14, A 3000,
Vq 8192 -- vqa/trainer.py:335 batch 512) and at the top of the inference sweep (config 5, B 8192 x K 100), against the
fp64 NumPy oracle. The oracle finishes a 512-sample forward + backward at these layer sizes in ~15 s of CPU time, so
the whole batch is compared, not a subsample: 4 M-tiles of the head GEMMs, 4 row tiles x 32 CTAs of the recurrent
kernels, 3.5 samples per CTA of the persistent attention kernel are all inside the comparison.
Also: the >= 99.9 % top-1 gate of the north star on 4096 samples, and the range checks of the index inputs."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import answer_model_np as O  # noqa: E402
from parity_util import build_case, rel_err, run_both  # noqa: E402
from test_model_gpu import BF16_TOL, FP32_TOL, _check, _check_forward_plain  # noqa: E402

pytestmark = pytest.mark.gpu

CFG1 = dict(B=512, K=36, Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=8192)


def test_cfg1_exact_headline_shapes_bf16_forward_and_gradients():
    """cfg 1 exactly, bf16 mode (the mode bench.py times): forward tensors against the reference's plain arithmetic at
    2e-2, every gradient at 2e-2 (max-norm) against the oracle restating the mode's operand rounding, box masking exact."""
    case = build_case(CFG1, precision="bf16", seed=31, num_images=96)
    got, ref, ref_g = run_both(case)
    assert int(case["eng"].lib.vqa_gru_kernel_path()) & 1, "cfg1 must run the CTA-pair recurrent kernels"
    _check_forward_plain(case, got, BF16_TOL)
    worst = _check(case, got, ref, ref_g, BF16_TOL, exact_pred=False)
    print("cfg1 bf16 max-norm gradient errors:", {k: f"{v:.2e}" for k, v in worst.items()})
    n_diff, n_all = case["gate_diffs"]
    print(f"ReLU gates (2-D heads) decided differently from the oracle: {n_diff} of {n_all}")
    agree = float((got["pred"] == case["plain_out"]["pred"]).mean())
    print(f"top-1 agreement with the plain fp64 oracle on this batch: {agree:.4f}")


def test_cfg1_exact_headline_shapes_fp32_forward_and_gradients():
    """cfg 1 exactly in the reference-precision mode: logits, loss, attention, pooled, condition within 1e-4, argmax and
    box masking exact, every gradient within 1e-4 (max-norm) of the plain fp64 oracle. Among the 2.6 M ReLU gates of the
    2-D heads a few dozen have |pre-activation| < 1e-4 and are undecidable at fp32 working precision: those gates are
    taken from the device (each asserted to be such a near-tie); the v-projection's near-ties are bounded exactly by
    the oracle (parity_util.relu_tie_budget)."""
    case = build_case(CFG1, precision="fp32", seed=32, num_images=96)
    got, ref, ref_g = run_both(case, device_gates=True)
    worst = _check(case, got, ref, ref_g, FP32_TOL)
    print("cfg1 fp32 max-norm gradient errors:", {k: f"{v:.2e}" for k, v in worst.items()})
    print("ReLU gates (2-D heads) decided differently from the oracle: %d of %d" % case["gate_diffs"])


def test_top1_agreement_on_4096_samples():
    """north_star: >= 99.9 % top-1 answer agreement with the reference. Eight batches of 512 at cfg1 shapes, forward
    only, device pred against the plain fp64 oracle's argmax under the same dropout masks, in BOTH modes.
    fp32 mode: the raw agreement over all 4096 samples must be >= 99.9 %.
    bf16 mode: with random-init synthetic weights the reference's own best and second-best logits are closer than the
    mode's resolution for ~1 % of the samples (3000 near-i.i.d. logits: the top-2 margin is dense at zero; a trained
    head is peaked). A flip is only possible where that margin is below twice the logit error actually measured, so
    the gate is: argmax exact given the logits (every miss lies inside that margin) and >= 99.9 % agreement on the
    samples whose reference margin exceeds the mode's tolerance (2e-2 of the largest logit); the raw rate is printed."""
    from vqa_transfer_externaldata_b200 import synthetic as S
    cases = {p: build_case(CFG1, precision=p, seed=33, num_images=128) for p in ("bf16", "fp32")}
    c = cases["bf16"]["c"]
    live = cases["bf16"]["m"]["exist"] > 0
    stat = {p: dict(ok=0, n=0, miss_outside=0, ok_clear=0, n_clear=0, worst=0.0) for p in cases}
    for r in range(8):
        batch = S.make_batch(c, 128, seed=900 + r)
        out = None
        for p, case in cases.items():
            eng = case["eng"]
            eng.stage_batch(batch)
            eng.forward(seed=55, step=r)
            am, jm = eng.dropout_masks(55, r)
            torch.cuda.synchronize()
            pred = eng.outputs()["pred"].cpu().numpy()
            logit = eng.outputs()["logit"].cpu().numpy()
            if out is None:   # the masks depend on (seed, step) only: one oracle pass serves both modes
                out, _ = O.forward(case["params"], case["feats"], case["nb"], batch, case["m"], variant="vlmap_answer",
                                   keep_att=0.8, keep_joint=0.5, att_mask=am.cpu().numpy(), joint_mask=jm.cpu().numpy())
            st = stat[p]
            scale = np.abs(out["logit"][:, live]).max()
            err = np.abs(logit[:, live] - out["logit"][:, live]).max()
            st["worst"] = max(st["worst"], err / scale)
            ok = pred == out["pred"]
            srt = np.sort(out["logit"], axis=1)
            gap = srt[:, -1] - srt[:, -2]
            st["miss_outside"] += int((~ok & (gap > 2.0 * err)).sum())
            clear = gap > (BF16_TOL if p == "bf16" else FP32_TOL) * scale
            st["ok_clear"] += int((ok & clear).sum())
            st["n_clear"] += int(clear.sum())
            st["ok"] += int(ok.sum())
            st["n"] += ok.size
    for p, st in stat.items():
        print(f"{p}: top-1 agreement {st['ok']} of {st['n']} = {st['ok'] / st['n']:.5f}; on samples with a clear reference "
              f"margin {st['ok_clear']} of {st['n_clear']}; misses outside twice the logit error: {st['miss_outside']}; "
              f"worst logit error {st['worst']:.2e}")
        assert st["n"] >= 4096
        assert st["miss_outside"] == 0            # argmax is exact given the logits
        # (random-init heads are flat: about a third of the samples have a top-2 margin below 2e-2 of the largest logit)
        assert st["ok_clear"] >= 0.999 * st["n_clear"] and st["n_clear"] >= 0.5 * st["n"]
    assert stat["fp32"]["worst"] < FP32_TOL and stat["bf16"]["worst"] < BF16_TOL
    assert stat["fp32"]["ok"] >= 0.999 * stat["fp32"]["n"]


def test_cfg5_top_of_sweep_b8192_subsample():
    """BASELINE config 5 at the largest batch of the sweep (8192 x 100 padded boxes, 10..100 valid): the device runs the
    whole batch (the recurrent part on the CTA-pair kernels as eight two-wave launches of 1024 rows), the oracle a
    256-row subsample (rows are independent in the forward pass)."""
    dims = dict(B=8192, K=100, Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=8192)
    case = build_case(dims, precision="bf16", seed=34, num_images=192)
    eng = case["eng"]
    eng.stage_batch(case["batch"])
    eng.forward(seed=5, step=2)
    att_mask, joint_mask = eng.dropout_masks(5, 2)
    torch.cuda.synchronize()
    rows = np.sort(np.random.default_rng(1).choice(8192, size=256, replace=False))
    rows[0], rows[-1] = 0, 8191
    sub = {k: v[rows] for k, v in case["batch"].items()}
    out, _ = O.forward(case["params"], case["feats"], case["nb"], sub, case["m"], variant="vlmap_answer",
                       keep_att=0.8, keep_joint=0.5, att_mask=att_mask[rows].cpu().numpy(),
                       joint_mask=joint_mask[rows].cpu().numpy())
    got = {k: v[rows].detach().cpu().numpy() for k, v in eng.outputs().items() if k in ("logit", "att_score", "pred")}
    live = case["m"]["exist"] > 0
    assert rel_err(got["logit"][:, live], out["logit"][:, live]) < BF16_TOL
    assert rel_err(got["att_score"], out["att_score"]) < BF16_TOL
    nbox = case["nb"][sub["image_idx"]]
    for i in range(len(rows)):
        assert np.all(got["att_score"][i, nbox[i]:] == 0.0)
    srt = np.sort(out["logit"], axis=1)
    close = (srt[:, -1] - srt[:, -2]) < BF16_TOL * np.abs(out["logit"][:, live]).max()
    assert np.all((got["pred"] == out["pred"]) | close)
    assert int(eng.lib.vqa_gru_kernel_path()) & 1   # the CTA-pair kernels, not the single-CTA fallback


def test_out_of_range_indices_fail_loudly():
    """image_idx / token ids out of range: host batches raise before anything is enqueued (np.take / embedding_lookup
    raise in the reference), image_idx = -1 (parse_fn's default for a missing feature) wraps to the last image like
    np.take, and device-resident inputs are caught by the gather kernels (no out-of-bounds access; sticky counter)."""
    SMALL = dict(B=16, K=12, Dv=256, D=128, L=128, A=200, T=6, W=20, Vq=50)
    case = build_case(SMALL, precision="bf16", seed=35, num_images=24)
    eng = case["eng"]
    good = case["batch"]
    bad = dict(good, image_idx=good["image_idx"].copy())
    bad["image_idx"][3] = 24
    with pytest.raises(IndexError):
        eng.stage_batch(bad)
    bad = dict(good, q_intseq=good["q_intseq"].copy())
    bad["q_intseq"][2, 1] = 50
    with pytest.raises(IndexError):
        eng.stage_batch(bad)
    bad = dict(good, q_intseq_len=good["q_intseq_len"].copy())
    bad["q_intseq_len"][0] = 7
    with pytest.raises(ValueError):
        eng.stage_batch(bad)
    # -1 wraps to the last image
    wrap = dict(good, image_idx=good["image_idx"].copy())
    wrap["image_idx"][:] = -1
    last = dict(good, image_idx=np.full_like(good["image_idx"], 23))
    eng.stage_batch(wrap)
    eng.forward(seed=1, step=1)
    a = eng.outputs()["logit"].clone()
    eng.stage_batch(last)
    eng.forward(seed=1, step=1)
    assert torch.equal(a, eng.outputs()["logit"])
    eng.read_scalars()   # no error counted so far
    # device-resident inputs skip the host check: the kernels count and neutralise the offence
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in good.items() if k != "id"}
    dev["image_idx"][5] = 1000
    dev["q_intseq"][1, 0] = 12345
    eng.stage_batch(dev)
    eng.forward(seed=1, step=1)
    eng.backward()
    with pytest.raises(IndexError):
        eng.read_scalars()
    eng.stage_batch(good)
    eng.forward(seed=1, step=1)
    eng.read_scalars()   # the counter was cleared by the failed read
