"""The reference-facing plugin interface on the GPU: importer.get_model_class(model_type)(batch, config, is_train,
image_features) for every model_type served, one train step each, and the dict keys a reference caller reads
(vqa/trainer.py:275-287, vqa/evaler.py:139-162) under the names that model_type uses in the reference."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from vqa_transfer_externaldata_b200 import importer  # noqa: E402
from vqa_transfer_externaldata_b200.model import make_synthetic_config  # noqa: E402

pytestmark = pytest.mark.gpu

SMALL = dict(B=16, K=12, Dv=256, D=128, L=128, A=200, T=6, W=20, Vq=50)
BASE_REPORT = {"answer_train_loss", "answer_report_loss", "answer_acc", "exist_acc", "test_acc", "normal_test_acc",
               "normal_test_object_acc", "normal_test_attribute_acc", "normal_exist_acc", "normal_train_exist_acc",
               "max_exist_acc", "test_max_acc", "test_max_exist_acc"}                 # model_vlmap_answer.py:275-288
OLD_REPORT = {"answer_train_loss", "answer_report_loss", "answer_accuracy", "exist_answer_accuracy",
              "test_answer_accuracy", "normal_test_answer_accuracy", "max_exist_answer_accuracy",
              "test_max_answer_accuracy", "test_max_exist_answer_accuracy"}          # model_vlmap_answer_no_noise.py:211-219
OUTPUT_KEYS = {"att_score", "logit", "pred", "all_score", "max_train_score", "test_obj_score", "test_obj_max_score",
               "test_attr_score", "test_attr_max_score"}                              # evaler.py:139-156


@pytest.mark.parametrize("model_type", importer.get_model_types())
def test_train_step_through_the_plugin_interface(model_type):
    config, feats, batch, _ = make_synthetic_config(SMALL, variant=model_type, precision="bf16", seed=5, num_images=16)
    config.num_marginal = 6
    cls = importer.get_model_class(model_type)
    model = cls(batch, config, is_train=True, image_features=feats)
    before = {k: v.copy() for k, v in model.state_dict().items()}
    loss, h2d, d2h = model.train_step()
    assert np.isfinite(loss) and h2d > 0 and d2h > 0
    want = OLD_REPORT if cls.OLD_REPORT else BASE_REPORT
    if model_type == "vlmap_answer_full":      # model_vlmap_answer_full.py:33-35, 221-223
        want = want | {"latent_loss", "train_latent_loss", "latent_loss_weight"}
        assert abs(model.losses["answer"] + model.losses["latent"] - loss) < 1e-4 * max(1.0, abs(loss))
        assert abs(model.report["train_latent_loss"] - 0.1 * model.report["latent_loss"]) < 1e-6
    if model_type == "vlmap_answer_ent":       # model_vlmap_answer_ent.py:290-294
        want = want | {"entropy", "weighted_entropy"}
        assert abs(model.losses["answer"] + model.losses["entropy"] - loss) < 1e-4 * max(1.0, abs(loss))
    assert set(model.report) == want
    assert OUTPUT_KEYS <= set(model.output)
    assert model.output["logit"].shape == (SMALL["B"], SMALL["A"])
    assert model.heavy_output["condition"].shape == (SMALL["B"], SMALL["L"])
    pooled_dim = SMALL["D"] if model_type == "vlmap_answer_adapt" else SMALL["Dv"]
    assert model.mid_result["pooled_V_ft"].shape == (SMALL["B"], pooled_dim)
    after = model.state_dict()
    names = list(after)
    trained = {n.split("/")[0] for n in model.filter_train_vars(names)}
    frozen = {n.split("/")[0] for n in names} - trained
    if model_type in ("vlmap_answer_noc", "vlmap_answer_nocarch"):   # model_vlmap_answer_noc.py:78-88
        assert {"q_linear_l", "pooled_linear_l", "joint_v", "joint_l", "WordWeightAnswerV", "WordWeightAnswerL"} <= frozen
    elif model_type != "standard":
        assert {"q_linear_l", "pooled_linear_l", "joint_fc", "WordWeightAnswer"} <= frozen
    changed = {n.split("/")[0] for n in names if not np.array_equal(before[n], after[n])}
    assert changed and changed <= trained, (changed - trained)
    if model_type in ("vlmap_answer_vqa_all", "vlmap_answer_vqa_all2"):
        # created by the reference, never reached by the loss (model_vlmap_answer_vqa_all.py:199-216): kept, unchanged
        assert "tuned_joint_fc/fc/weights" in after and "tuned_q_linear_l/LayerNorm/gamma" in after
        assert "TunedWordWeightAnswer" in changed
    if model_type == "vlmap_answer_adapt":
        assert "v_adapt" in changed and after["pooled_linear_l/fc/weights"].shape == (SMALL["D"], SMALL["L"])
    if model_type == "vlmap_answer_full":
        assert {"q_L_mean", "q_L_log_sigma_sq"} <= changed
    # a checkpoint round trip by variable name
    model.load_state_dict(after)


def test_tfrecord_batches_feed_the_model(tmp_path):
    """vqa/trainer.py: batch = input_ops.create(...); Model(batch, config, ...). A batch read back from TFRecord shards
    (input_ops.create) drives the CUDA path exactly like the dict it was written from."""
    from vqa_transfer_externaldata_b200 import input_ops as IO
    config, feats, batch, _ = make_synthetic_config(SMALL, variant="vlmap_answer", precision="bf16", seed=7, num_images=16)
    B, A = SMALL["B"], SMALL["A"]
    samples = []
    for i in range(B):
        ids = np.nonzero(batch["answer_target"][i])[0]
        samples.append({"qid": int(batch["id"][i]), "image_id": f"{i}".encode(), "image_idx": int(batch["image_idx"][i]),
                        "q_intseq": batch["q_intseq"][i, :batch["q_intseq_len"][i]],
                        "answer_ids": ids, "answer_scores": batch["answer_target"][i, ids]})
    IO.write_shards(str(tmp_path), "val", samples, A, num_shards=1)
    (rb,) = list(IO.create(B, str(tmp_path), "val", is_train=False))
    assert np.array_equal(rb["answer_target"], batch["answer_target"])
    assert np.array_equal(rb["q_intseq_len"], batch["q_intseq_len"])
    model = importer.get_model_class("vlmap_answer")(batch, config, is_train=False, image_features=feats)
    model.forward(batch, dropout_step=0)
    ref = model.output["logit"].clone()
    model.forward(rb, dropout_step=0)     # T of this batch = its longest question (padded_batch), not the configured maximum
    assert torch.allclose(model.output["logit"], ref, rtol=1e-5, atol=1e-5)
    loss_ref = float(model.fetch()[0])
    # the native reader (csrc/input_host.cu): same batch, the soft-score target as (row, id, score) triples that
    # vqa_densify_targets scatters on the device
    from vqa_transfer_externaldata_b200 import input_native as IN
    (nb,) = list(IN.create(B, str(tmp_path), "val", is_train=False))
    assert "answer_target" not in nb and np.array_equal(IN.densify(nb), batch["answer_target"])
    model.forward(nb, dropout_step=0)
    assert torch.allclose(model.output["logit"], ref, rtol=1e-5, atol=1e-5)
    assert abs(float(model.fetch()[0]) - loss_ref) <= 1e-6 * max(1.0, abs(loss_ref))
    got_target = model.engine.cur.d_target[:B * A].cpu().numpy().reshape(B, A)
    assert np.array_equal(got_target, batch["answer_target"])


def test_checkpoint_bundle_round_trip(tmp_path):
    """Model.save_checkpoint / load_checkpoint: TensorFlow's tensor-bundle files keyed by the reference's variable
    names (vqa/trainer.py:141-147, 173-186); a model restored from the bundle computes the same logits."""
    config, feats, batch, _ = make_synthetic_config(SMALL, variant="vlmap_answer_vqa_all", precision="bf16", seed=9, num_images=16)
    cls = importer.get_model_class("vlmap_answer_vqa_all")
    m1 = cls(batch, config, is_train=True, image_features=feats)
    m1.train_step()
    m1.train_step()
    prefix = str(tmp_path / "model-2")
    m1.save_checkpoint(prefix)
    from vqa_transfer_externaldata_b200 import tf_bundle
    names = set(tf_bundle.read_bundle(prefix))
    assert {"global_step", "LearnGloVe/embed_map", "encode_L/rnn/gru_cell/gates/kernel", "WordWeightAnswer/fc/weights",
            "TunedWordWeightAnswer/fc/biases", "tuned_joint_fc/LayerNorm/gamma"} <= names
    config2, _, _, _ = make_synthetic_config(SMALL, variant="vlmap_answer_vqa_all", precision="bf16", seed=77, num_images=16)
    config2.answer_exist_mask = config.answer_exist_mask              # same exported word-weight vocabulary
    m2 = cls(batch, config2, is_train=False, image_features=feats)     # different initial weights
    m2.load_checkpoint(prefix)
    assert m2.global_step == 2
    m1.forward(batch, dropout_step=5)
    m2.seed = m1.seed
    m2.forward(batch, dropout_step=5)                                   # same dropout draw
    assert torch.equal(m1.output["logit"], m2.output["logit"])
    # the optimizer slots travel with the checkpoint under the reference's names (tf.train.Saver of vqa/trainer.py:141-147)
    assert {"optimizer/TunedWordWeightAnswer/fc/weights/Adam", "optimizer/beta1_power",
            "optimizer/encode_L/rnn/gru_cell/gates/kernel/Adam_1"} <= names
    assert m2.engine.adam_t == 2 and torch.equal(m2.engine.params.adam_m, m1.engine.params.adam_m)
    # evaluation draws a fresh dropout mask per call (tf.nn.dropout has no train switch, one draw per session.run)
    m1.forward(batch)
    a = m1.output["logit"].clone()
    m1.forward(batch)
    assert not torch.equal(a, m1.output["logit"])


def test_pipelined_optimizer_tail_is_bit_identical(monkeypatch):
    """Model.train_step updates the embedding / GRU slice of the parameters on an auxiliary stream, under the first
    kernels of the next forward (vqa_set_optimizer_tail). Three steps with and without it must leave identical
    parameters, Adam moments and losses (bit for bit, except the embedding table and its moments, which follow the atomically
    scattered embedding gradient)."""
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("VQA_ADAM_TAIL", mode)
        config, feats, batch, _ = make_synthetic_config(SMALL, variant="vlmap_answer", precision="bf16", seed=9, num_images=16)
        model = importer.get_model_class("vlmap_answer")(batch, config, is_train=True, image_features=feats)
        losses = [model.train_step()[0] for _ in range(3)]
        res[mode] = (losses, model.state_dict(), model.optimizer_state_dict())
    # The first loss is bit-identical. From the second step on everything follows the embedding table, whose gradient is
    # a scatter-add of fp32 atomics: a moment that differs in its last bit can move the last bit of a parameter (seen once
    # in ~25 runs), so the comparison allows last-bit noise -- a stale weight (what this test guards against: the next
    # forward reading a parameter the auxiliary stream has not updated yet) is off by the whole update, ~1e-3 relative.
    assert res["1"][0][0] == res["0"][0][0]
    assert np.allclose(res["1"][0], res["0"][0], rtol=1e-6, atol=0)
    for k, v in res["0"][1].items():
        assert np.allclose(res["1"][1][k], v, rtol=5e-6, atol=5e-9), k
    for k, v in res["0"][2].items():
        a, b = np.asarray(res["1"][2][k]), np.asarray(v)
        assert np.allclose(a, b, rtol=1e-4 if "embed_map" in k else 2e-5, atol=1e-10), k
