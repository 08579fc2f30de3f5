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
