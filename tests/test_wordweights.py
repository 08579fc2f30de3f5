"""File contracts around the path: WordWeightAnswer column remap (vlmap/modules.py:600-614), AnswerExistMask
(:575-586), the exporter's directory layout (vlmap_memft/export_word_weights.py:60-83), TF variable names."""
import numpy as np

from vqa_transfer_externaldata_b200 import wordweights as WW
from vqa_transfer_externaldata_b200.engine import frozen_fields, tf_name
from vqa_transfer_externaldata_b200 import importer


def test_word_weight_remap_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    exported_vocab = ["cat", "dog", "red", "blue", "tall"]
    cw = rng.standard_normal((8, 5)).astype(np.float32)
    cb = rng.standard_normal(5).astype(np.float32)
    WW.export_word_weights(str(tmp_path), cw, cb, exported_vocab)
    vqa_answers = {"vocab": ["blue", "zebra", "cat", "tall", "unknown"]}
    w, b = WW.word_weight_answer(8, vqa_answers, str(tmp_path))
    assert w.shape == (8, 5) and b.shape == (5,)
    np.testing.assert_array_equal(w[:, 0], cw[:, 3])
    np.testing.assert_array_equal(w[:, 2], cw[:, 0])
    np.testing.assert_array_equal(w[:, 3], cw[:, 4])
    assert b[0] == cb[3] and b[2] == cb[0]
    # absent answers: zero column, bias -100 exactly (modules.py:600-601)
    assert np.all(w[:, 1] == 0) and np.all(w[:, 4] == 0)
    assert b[1] == -100.0 and b[4] == -100.0
    np.testing.assert_array_equal(WW.answer_exist_mask(vqa_answers, str(tmp_path)), [1, 0, 1, 1, 0])
    np.testing.assert_array_equal(WW.answer_exist_mask(vqa_answers, None), np.ones(5))


def test_no_word_weight_dir_is_all_default():
    w, b = WW.word_weight_answer(4, {"vocab": ["a", "b"]}, None)
    assert np.all(w == 0) and np.all(b == -100.0)


def test_tf_names_and_frozen_sets():
    assert tf_name("ans_w", "vlmap_answer") == "WordWeightAnswer/fc/weights"
    assert tf_name("ans_w", "standard") == "reasoning/classifier/fc/weights"
    assert tf_name("joint_gamma", "standard") == "reasoning/joint_fc/LayerNorm/gamma"
    assert tf_name("gru_gates_w", "standard") == "encode_L/rnn/gru_cell/gates/kernel"
    fz = frozen_fields("vlmap_answer")
    assert {"pl_w", "ql_w", "joint_w", "ans_w", "ans_b", "joint_beta"} <= fz
    assert not ({"v_w", "embed", "gru_gates_w", "qv_w", "att_w"} & fz)
    assert frozen_fields("standard") == set()


def test_importer_names():
    assert importer.get_model_types() == ["vlmap_answer", "vlmap_answer2", "vlmap_answer_no_noise", "vlmap_answer_noc",
                                          "vlmap_answer_nocarch", "vlmap_answer_full", "vlmap_answer_vqa_all",
                                          "vlmap_answer_vqa_all2", "vlmap_answer_adapt", "vlmap_answer_ent", "standard"]
    assert importer.get_model_class("vlmap_answer_vqa_all2").MODEL_TYPE == "vlmap_answer_vqa_all2"
    assert importer.get_model_class("vlmap_answer_full").OLD_REPORT and importer.get_model_class("vlmap_answer_adapt").OLD_REPORT
    assert issubclass(importer.get_model_class("vlmap_answer_nocarch"), importer.get_model_class("vlmap_answer_noc"))
    from vqa_transfer_externaldata_b200 import model as M
    assert importer.get_model_class("vlmap_answer2") is M.Answer2Model
    assert importer.get_model_class("vlmap_answer_no_noise") is M.NoNoiseModel
    import pytest
    with pytest.raises(ValueError):
        importer.get_model_class("nope")
    assert importer.get_model_class("vlmap_answer_ent").MODEL_TYPE == "vlmap_answer_ent"
    with pytest.raises(NotImplementedError):
        importer.get_model_class("vlmap_only")


def test_learning_rate_schedule():
    """vqa/trainer.py:87-96: constant 0.001, or halved every 10 000 steps (staircase) with --lr_weight_decay."""
    from types import SimpleNamespace
    from vqa_transfer_externaldata_b200.model import Model
    m = Model.__new__(Model)
    m.config, m.global_step = SimpleNamespace(), 25000
    assert m.learning_rate() == 1e-3
    m.config = SimpleNamespace(learning_rate=2e-3, lr_weight_decay=True)
    assert m.learning_rate() == 2e-3 * 0.25
    m.global_step = 9999
    assert m.learning_rate() == 2e-3


def test_param_store_layout_on_cpu():
    """ParamStore: trainable tensors first (one contiguous all-reduce / Adam slice), 256-byte aligned starts, frozen after,
    TAIL slots past the gradients, views keyed by the reference's variable names."""
    import torch
    from vqa_transfer_externaldata_b200 import lib as L
    from vqa_transfer_externaldata_b200.engine import AnswerModelConfig, ParamStore, frozen_fields, tf_name
    for variant in ("vlmap_answer", "standard", "vlmap_answer_vqa_all", "vlmap_answer_adapt", "vlmap_answer_full"):
        cfg = AnswerModelConfig(B=8, K=4, Dv=64, D=32, L=32, J=64, A=24, T=5, W=12, Vq=30, num_train_answer=18, variant=variant)
        ps = ParamStore(cfg, torch.device("cpu"))
        assert set(ps.fields) == set(L.param_fields(variant))
        assert set(ps.frozen) == frozen_fields(variant) & set(ps.fields)
        end_train = max(ps.offsets[f][0] + ps.offsets[f][1] for f in ps.trainable)
        assert end_train <= ps.n_train and all(ps.offsets[f][0] >= ps.n_train for f in ps.frozen)
        assert all(ps.offsets[f][0] % 64 == 0 for f in ps.fields)                      # 64 floats = 256 bytes
        assert ps.grad_buf.numel() == ps.n_train + ps.TAIL and ps.grad.data_ptr() == ps.grad_buf.data_ptr()
        spans = sorted(ps.offsets.values())
        assert all(a[0] + a[1] <= b[0] for a, b in zip(spans, spans[1:]))              # no overlap
        names = ps.by_tf_name()
        assert len(names) == len(ps.fields) and tf_name("v_w", variant) in names
        if variant == "vlmap_answer_adapt":
            assert tuple(ps.views["pl_w"].shape) == (cfg.D, cfg.L) and "v_adapt/fc/weights" in names
        if variant == "vlmap_answer_vqa_all":
            assert "TunedWordWeightAnswer/fc/weights" in names and "tw_w" in ps.trainable
        # gradients of the late group (embedding + GRU) sit at the end of the trainable slice
        late = [f for f in ps.trainable if f in ("embed", "gru_gates_w", "gru_gates_b", "gru_cand_w", "gru_cand_b")]
        assert ps.trainable[-len(late):] == late
