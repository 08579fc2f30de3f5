"""`Model`: the host-side mirror of the reference's model classes for this path.

Same constructor signature and attributes as vqa/model_vlmap_answer.py:17-79 / vqa/model_standard.py:
    Model(batch, config, is_train=True, image_features=None)
    .loss .losses .report .output .mid_result .heavy_output .vocab .answer_dict
    filter_train_vars(vars), filter_transfer_vars(vars)
The reference builds a TF graph and evaluates it with session.run; here `forward()` / `backward()` /
`train_step()` run the CUDA path eagerly on the batch bound at construction (or one passed in).
"""
import os
import pickle
from types import SimpleNamespace

import numpy as np
import torch

from . import lib as L
from .engine import AnswerModelConfig, Engine, TF_NAMES, frozen_fields, tf_name
from . import wordweights

W_DIM = 300   # vqa/model_vlmap_answer.py:10-12
L_DIM = 1024
V_DIM = 1024


def _load_pickle(path):
    """vocab.pkl / answer_dict.pkl are Python-2 cPickle files (vqa/model_vlmap_answer.py:34-36)."""
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")


def get_dummy_data():
    """util.get_dummy_data (util/__init__.py:5-17): the --debug fixture, 500 x 36 x 2048 zeros, 1 box each."""
    bn, bs, dim = 500, 36, 2048
    return {"features": np.zeros([bn, bs, dim], np.float32), "spatials": np.zeros([bn, bs, 6], np.float32),
            "normal_boxes": np.zeros([bn, bs, 4], np.float32), "num_boxes": np.ones([bn], np.int32),
            "max_box_num": bs, "vfeat_dim": dim}


# report keys of the older family members (vqa/model_vlmap_answer_no_noise.py:211-219, _full.py:221-235,
# _adapt.py:212-220): nine scalars under longer names; the other four of the base model are not reported there
OLD_REPORT_NAMES = {
    "answer_train_loss": "answer_train_loss", "answer_report_loss": "answer_report_loss",
    "answer_acc": "answer_accuracy", "exist_acc": "exist_answer_accuracy", "test_acc": "test_answer_accuracy",
    "normal_test_acc": "normal_test_answer_accuracy", "max_exist_acc": "max_exist_answer_accuracy",
    "test_max_acc": "test_max_answer_accuracy", "test_max_exist_acc": "test_max_exist_answer_accuracy",
    "latent_loss": "latent_loss", "train_latent_loss": "train_latent_loss",
}


class Model(object):
    MODEL_TYPE = "vlmap_answer"
    OLD_REPORT = False     # True: report under the names of OLD_REPORT_NAMES
    # checkpoint variables the reference creates for this model_type that never reach the loss (kept so that
    # state_dict() / load_state_dict() round-trip a reference checkpoint): name -> shape as a function of the config
    DEAD_VARIABLES = {}

    def __init__(self, batch, config, is_train=True, image_features=None):
        self.batch = batch
        self.config = config
        self.image_dir = getattr(config, "image_dir", None)
        self.is_train = is_train
        self.word_weight_dir = getattr(config, "vlmap_word_weight_dir", None)
        self.losses, self.report, self.mid_result = {}, {}, {}
        self.output, self.heavy_output, self.vis_image = {}, {}, {}
        self.loss = None

        # vocab / answer_dict: pickles when paths are given (reference), objects otherwise (synthetic)
        self.vocab = getattr(config, "vocab", None)
        if self.vocab is None:
            self.vocab = _load_pickle(config.vocab_path)
        self.answer_dict = getattr(config, "answer_dict", None)
        if self.answer_dict is None:
            self.answer_dict = _load_pickle(os.path.join(config.tf_record_dir, "answer_dict.pkl"))
        self.num_answer = len(self.answer_dict["vocab"])
        self.num_train_answer = int(self.answer_dict["num_train_answer"])
        A = self.num_answer
        self.train_answer_mask = (np.arange(A) < self.num_train_answer).astype(np.float32)[None]
        self.test_answer_mask = 1.0 - self.train_answer_mask
        self.obj_answer_mask = np.asarray(self.answer_dict["is_object"], np.float32)[None]
        self.attr_answer_mask = np.asarray(self.answer_dict["is_attribute"], np.float32)[None]
        exist = getattr(config, "answer_exist_mask", None)  # synthetic configs carry it explicitly
        if exist is None:
            exist = wordweights.answer_exist_mask(self.answer_dict, self.word_weight_dir)
        self.answer_exist_mask = np.asarray(exist, np.float32)[None]

        # feature bank (vqa/model_vlmap_answer.py:54-77)
        if getattr(config, "debug", False):
            image_features = get_dummy_data()
        elif image_features is None:
            image_features = wordweights.load_feature_bank(config.vfeat_path)
        self.features = image_features["features"]
        self.num_boxes = image_features["num_boxes"]
        self.max_box_num = int(image_features["max_box_num"])
        self.vfeat_dim = int(image_features["vfeat_dim"])
        self.spatials = image_features.get("spatials")
        self.normal_boxes = image_features.get("normal_boxes")

        self.build()

    # ---- reference API: which variables train / transfer (by top-level scope) ------------------
    def filter_train_vars(self, trainable_vars):
        frozen_scopes = {tf_name(f, self.MODEL_TYPE).split("/")[0] for f in frozen_fields(self.MODEL_TYPE)}
        return [v for v in trainable_vars if _name(v).split("/")[0] not in frozen_scopes]

    TRANSFER_SCOPES = ("q_linear_l", "pooled_linear_l", "joint_fc")   # vqa/model_vlmap_answer.py:91-100 (and siblings)

    def filter_transfer_vars(self, all_vars):
        scopes = self.TRANSFER_SCOPES
        return [v for v in all_vars if _name(v).split("/")[0] in scopes]

    # ---- graph construction == engine construction ---------------------------------------------
    def build(self):
        cfgd = self.config
        T = int(getattr(cfgd, "max_q_len", 14))
        B = int(getattr(cfgd, "batch_size", 512))
        self.engine_config = AnswerModelConfig(
            B=B, K=self.max_box_num, Dv=self.vfeat_dim, D=int(getattr(cfgd, "v_dim", V_DIM)),
            L=int(getattr(cfgd, "l_dim", L_DIM)), J=2 * int(getattr(cfgd, "l_dim", L_DIM)),
            A=self.num_answer, T=T, W=int(getattr(cfgd, "w_dim", W_DIM)), Vq=len(self.vocab["vocab"]),
            num_train_answer=self.num_train_answer, variant=self.MODEL_TYPE,
            precision=getattr(cfgd, "precision", "bf16"), num_marginal=int(getattr(cfgd, "num_marginal", 200)))
        self.engine = Engine(self.engine_config, device=getattr(cfgd, "device", None))
        self.engine.set_feature_bank(self.features, self.num_boxes)
        self.engine.set_answer_masks(self.obj_answer_mask[0], self.attr_answer_mask[0], self.answer_exist_mask[0])
        params = getattr(cfgd, "init_params", None)
        if params is None:
            params = self.initial_params(seed=int(getattr(cfgd, "seed", 123)))
        self.engine.load_params(params)
        rng = np.random.default_rng(int(getattr(cfgd, "seed", 123)) + 99)
        self.dead_variables = {}
        for name, shape_fn in self.DEAD_VARIABLES.items():
            shape = shape_fn(self.engine_config)
            if name.endswith("weights"):
                lim = np.sqrt(6.0 / (shape[0] + shape[1]))
                self.dead_variables[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
            else:
                self.dead_variables[name] = (np.ones if name.endswith("gamma") else np.zeros)(shape, np.float32)
        self.seed = int(getattr(cfgd, "seed", 123))
        self.global_step = 0
        self.eval_draws = 0    # forward() calls so far outside train_step: each draws fresh dropout masks (see forward)
        self._dp = None
        return self.loss

    def initial_params(self, seed):
        """TF-default initialisers + the WordWeightAnswer remap from the exported word weights."""
        from .synthetic import init_params
        c = {k: getattr(self.engine_config, k) for k in
             ("B", "K", "Dv", "D", "L", "J", "A", "T", "W", "Vq", "num_train_answer")}
        p, _ = init_params(c, seed=seed, variant="standard")  # Xavier head; replaced below for vlmap_answer
        extra, _ = init_params(c, seed=seed + 1, variant=self.MODEL_TYPE)
        p.update({k: v for k, v in extra.items()
                  if k.startswith(("qp_", "jl_", "al_", "qs_", "tw_", "va_")) or (k == "pl_w" and self.MODEL_TYPE == "vlmap_answer_adapt")})   # layers of the variants
        glove = getattr(self.config, "glove_embed", None)
        if glove is not None:  # LearnGloVe (vlmap/modules.py:415-448): rows by vocabulary word
            p["embed"] = np.asarray(glove, np.float32)
        elif getattr(self.config, "vocab_path", None):   # a real vocabulary without its GloVe table: say so
            import warnings
            warnings.warn("config.glove_embed is not set: LearnGloVe/embed_map starts from a random table instead of the "
                          "GloVe rows the reference loads (vlmap/modules.py:415-448); restore a checkpoint or pass glove_embed")
        if self.MODEL_TYPE != "standard":
            if self.MODEL_TYPE in ("vlmap_answer_noc", "vlmap_answer_nocarch"):
                # vqa/model_vlmap_answer_noc.py:189-201: v_class_* / l_class_* of export_noc_word_weights.py:74-82
                p["ans_w"], p["ans_b"] = wordweights.word_weight_answer(
                    self.engine_config.J, self.answer_dict, self.word_weight_dir, "v_class_weights", "v_class_biases")
                p["al_w"], p["al_b"] = wordweights.word_weight_answer(
                    self.engine_config.J, self.answer_dict, self.word_weight_dir, "l_class_weights", "l_class_biases")
            else:
                w, b = wordweights.word_weight_answer(self.engine_config.J, self.answer_dict, self.word_weight_dir)
                p["ans_w"], p["ans_b"] = w, b
        return p

    # ---- checkpoint contract: tensors keyed by TF variable names ---------------------------------
    def state_dict(self):
        self.engine.sync_params()   # (a pipelined optimizer tail of the last train_step)
        sd = {k: v.detach().cpu().numpy().copy() for k, v in self.engine.params.by_tf_name().items()}
        sd.update({k: v.copy() for k, v in self.dead_variables.items()})
        return sd

    def load_state_dict(self, state, strict=True):
        self.engine.sync_params()
        fields = {tf_name(f, self.MODEL_TYPE): f for f in L.param_fields(self.MODEL_TYPE)}
        missing = [n for n in fields if n not in state]
        if strict and missing:
            raise KeyError(f"missing variables: {missing}")
        for name, f in fields.items():
            if name in state:
                self.engine.params.views[f].copy_(torch.as_tensor(np.asarray(state[name], np.float32)))
        for name in self.dead_variables:
            if name in state:
                self.dead_variables[name] = np.asarray(state[name], np.float32).copy()
        self.engine.prepare_params()

    def save_checkpoint(self, prefix):
        """tf.train.Saver.save equivalent (vqa/trainer.py:141-147): `<prefix>.index` + `<prefix>.data-00000-of-00001`
        in TensorFlow's tensor-bundle format, variables under their reference names plus `global_step`."""
        from . import tf_bundle
        sd = self.state_dict()
        sd["global_step"] = np.asarray(self.global_step, np.int64)
        sd.update(self.optimizer_state_dict())
        tf_bundle.write_bundle(prefix, sd)

    # optimize_loss(..., optimizer=AdamOptimizer, name='optimizer') (vqa/trainer.py:106-114) applies the gradients inside
    # variable_scope('optimizer'): tf.train.Saver stores the slots of every trained variable as `optimizer/<var>/Adam` (m)
    # and `optimizer/<var>/Adam_1` (v), plus `optimizer/beta1_power` and `optimizer/beta2_power`
    ADAM_SCOPE = "optimizer"

    def optimizer_state_dict(self):
        """Adam slots under the reference's checkpoint names (empty before the first optimizer step)."""
        self.engine.sync_params()
        e = self.engine
        ps = e.params
        if ps.adam_m is None:
            return {}
        out = {}
        for f in ps.trainable:
            o, n = ps.offsets[f]
            name = tf_name(f, self.MODEL_TYPE)
            shape = self.engine_config.shape(f)
            out[f"{self.ADAM_SCOPE}/{name}/Adam"] = ps.adam_m[o:o + n].view(shape).cpu().numpy().copy()
            out[f"{self.ADAM_SCOPE}/{name}/Adam_1"] = ps.adam_v[o:o + n].view(shape).cpu().numpy().copy()
        out[f"{self.ADAM_SCOPE}/beta1_power"] = np.asarray(0.9 ** e.adam_t, np.float32)
        out[f"{self.ADAM_SCOPE}/beta2_power"] = np.asarray(0.999 ** e.adam_t, np.float32)
        return out

    def load_optimizer_state_dict(self, state):
        """Restore m, v and the step count t (from beta1_power = 0.9^t) when the bundle carries them; returns whether it did."""
        e = self.engine
        ps = e.params
        key = f"{self.ADAM_SCOPE}/beta1_power"
        if key not in state:
            return False
        m = torch.zeros_like(ps.grad)
        v = torch.zeros_like(ps.grad)
        for f in ps.trainable:
            o, n = ps.offsets[f]
            name = tf_name(f, self.MODEL_TYPE)
            km, kv = f"{self.ADAM_SCOPE}/{name}/Adam", f"{self.ADAM_SCOPE}/{name}/Adam_1"
            if km not in state or kv not in state:
                return False
            m[o:o + n].copy_(torch.as_tensor(np.asarray(state[km], np.float32)).reshape(-1))
            v[o:o + n].copy_(torch.as_tensor(np.asarray(state[kv], np.float32)).reshape(-1))
        ps.adam_m, ps.adam_v = m, v
        b1p = float(np.asarray(state[key]))
        e.adam_t = int(round(np.log(b1p) / np.log(0.9))) if 0.0 < b1p < 1.0 else 0
        return True

    def load_checkpoint(self, prefix, strict=True):
        """Restore from a TensorFlow checkpoint bundle by variable name (vqa/trainer.py:173-186), Adam slots and beta
        powers included when present (a resumed run continues the bias correction instead of restarting it at t = 1);
        variables the path does not own are ignored."""
        from . import tf_bundle
        state = tf_bundle.read_bundle(prefix)
        if "global_step" in state:
            self.global_step = int(state["global_step"])
        self.load_state_dict(state, strict=strict)
        self.load_optimizer_state_dict(state)

    # ---- running the path --------------------------------------------------------------------------
    def attach_data_parallel(self, dp):
        self._dp = dp
        # Optional (VQA_DP_EARLY=1): take the non-GRU gradients before the BPTT so that their all-reduce runs under
        # it. Measured on 8 x B200 it LOSES (1.385 -> 1.439 ms/step): dWv leaves the concurrent weight-gradient
        # section for the critical path and the NCCL kernel delays the cooperative BPTT launch; so it is off.
        import os
        if dp is not None and dp.world_size > 1:
            dp.use_multicast_gradients(self.engine)   # in-switch all-reduce when NVSwitch multicast is available
        self.engine.set_early_gradients(dp is not None and dp.world_size > 1 and os.environ.get("VQA_DP_EARLY") == "1")

    def forward(self, batch=None, full_outputs=True, defer_outputs=False, dropout_step=None):
        """session.run([loss, report, output]) of the reference (vqa/evaler.py:118-123). Returns h2d bytes.
        tf.nn.dropout has no train switch: it fires in evaluation too, with a FRESH mask per session.run.
        So every call draws its masks from its own Philox step: the global step inside train_step, and outside it a
        counter of forward() calls placed above any reachable global step. dropout_step pins the draw (parity tests,
        reproducing a given evaluation)."""
        b = self.batch if batch is None else batch
        nbytes = self.engine.stage_batch(b)
        rank = self._dp.rank if self._dp is not None else 0
        if dropout_step is None:
            dropout_step = (1 << 40) + self.eval_draws
            self.eval_draws += 1
        self.engine.forward(seed=self.seed + 7919 * rank, step=int(dropout_step), full_outputs=full_outputs,
                            defer_outputs=defer_outputs)
        self._bind_outputs()
        return nbytes

    def backward(self):
        ws = self._dp.world_size if self._dp is not None else 1
        self.engine.backward(loss_scale=1.0 / ws)
        if self._dp is not None:
            self._dp.all_reduce_gradients(self.engine)

    def learning_rate(self):
        """The trainer's schedule (vqa/trainer.py:87-96): config.learning_rate (0.001), halved every 10 000 steps when
        config.lr_weight_decay is set (tf.train.exponential_decay, staircase=True, on the global step BEFORE this
        step's increment)."""
        lr = float(getattr(self.config, "learning_rate", 1e-3))
        if getattr(self.config, "lr_weight_decay", False):
            lr *= 0.5 ** (self.global_step // 10000)
        return lr

    def train_step(self, batch=None, lr=None, clip_norm=20.0, apply_optimizer=True, next_batch=None, sync=True):
        """run_train_step of vqa/trainer.py:275-287: forward + backward (+ all-reduce) + clip + Adam.
        Returns (loss, h2d_bytes, d2h_bytes); the loss read is the step's device->host copy.
        next_batch: host batch of the FOLLOWING step; its upload starts now on a copy stream and overlaps this
        step's kernels (the reference's tf.data pipeline prefetches the same way).
        sync=False: asynchronous dispatch -- `loss` is a handle whose .get() -> (loss, report) waits for this
        step only, so the host can enqueue step i+1 while the device runs step i."""
        h2d = self.forward(batch, full_outputs=False, defer_outputs=True,   # the backward below joins the loss kernels
                           dropout_step=self.global_step)
        if next_batch is not None:
            self.engine.prefetch_batch(next_batch)
        self.backward()
        if apply_optimizer:
            self.engine.adam_step(lr=self.learning_rate() if lr is None else lr, clip_norm=clip_norm, pipelined_tail=True)
        self.global_step += 1
        if not sync:
            pending = self.engine.read_scalars_async()
            return pending, h2d, pending.nbytes
        loss, report = self.engine.read_scalars()
        self._bind_report(loss, report)
        return loss, h2d, 4 * (1 + L.NUM_REPORT)

    def _bind_outputs(self):
        e = self.engine
        Bn = e.batch_size
        self.output = e.outputs()
        self.mid_result = {"att_score": self.output["att_score"], "logit": self.output["logit"],
                           "pred": self.output["pred"], "pooled_V_ft": e.o_pooled[:Bn]}
        self.heavy_output = {"condition": e.o_condition[:Bn]}

    def fetch(self):
        """Synchronise and return (loss, report) of the last forward."""
        self._bind_report(*self.engine.read_scalars())
        return self.loss, self.report

    def _bind_report(self, loss, report):
        """model.loss / model.losses / model.report under the key names this model_type uses in the reference."""
        self.loss = loss
        if "train_latent_loss" in report:   # vqa/model_vlmap_answer_full.py:220-221: two entries that sum to the loss
            self.losses["answer"] = report["answer_train_loss"]
            self.losses["latent"] = report["train_latent_loss"]
            report = dict(report, latent_loss_weight=0.1)
        elif "weighted_entropy" in report:  # vqa/model_vlmap_answer_ent.py:290-291
            self.losses["answer"] = report["answer_train_loss"]
            self.losses["entropy"] = report["weighted_entropy"]
        else:
            self.losses["answer"] = loss
        if self.OLD_REPORT:
            report = {OLD_REPORT_NAMES.get(k, k): v for k, v in report.items() if k in OLD_REPORT_NAMES or k == "latent_loss_weight"}
        self.report = report


class Answer2Model(Model):
    """vqa/model_vlmap_answer2.py: q_L_ft2 = tanh(LN(FC(q))) feeds q_linear_l and is heavy_output['condition']."""
    MODEL_TYPE = "vlmap_answer2"


class NoNoiseModel(Model):
    """vqa/model_vlmap_answer_no_noise.py: q_L_mean = FC(q) (linear) feeds q_linear_l."""
    MODEL_TYPE = "vlmap_answer_no_noise"
    OLD_REPORT = True


class FullModel(Model):
    """vqa/model_vlmap_answer_full.py: q_linear_l reads q_L_mean + N(0,1) * sqrt(exp(q_L_log_sigma_sq)); the KL latent
    loss enters model.loss with weight 0.1 and is reported as latent_loss / train_latent_loss."""
    MODEL_TYPE = "vlmap_answer_full"
    OLD_REPORT = True


class EntModel(Model):
    """vqa/model_vlmap_answer_ent.py: the base model plus the maximum-entropy regulariser (NUM_MARGINAL tiles of the
    joint head per sample, marginal softmax over the train & existing answers, loss += 0.1 * negative entropy)."""
    MODEL_TYPE = "vlmap_answer_ent"


class AdaptModel(Model):
    """vqa/model_vlmap_answer_adapt.py: attention pools v_adapt = relu(LN(FC(V))) (trained, 1024-d) instead of the raw
    features; pooled_linear_l/fc/weights is [V_DIM, L_DIM]."""
    MODEL_TYPE = "vlmap_answer_adapt"
    OLD_REPORT = True


class VqaAllModel(Model):
    """vqa/model_vlmap_answer_vqa_all.py: the frozen word-weight head (absent answers filled with the row minimum) plus
    the trained TunedWordWeightAnswer on the same joint; tuned_q_linear_l / tuned_joint_fc exist as variables but
    never reach the loss (:215-216 feeds `joint`)."""
    MODEL_TYPE = "vlmap_answer_vqa_all"
    DEAD_VARIABLES = {
        "tuned_q_linear_l/fc/weights": lambda c: (c.L, c.L), "tuned_q_linear_l/fc/biases": lambda c: (c.L,),
        "tuned_q_linear_l/LayerNorm/gamma": lambda c: (c.L,), "tuned_q_linear_l/LayerNorm/beta": lambda c: (c.L,),
        "tuned_joint_fc/fc/weights": lambda c: (c.L, c.J), "tuned_joint_fc/fc/biases": lambda c: (c.J,),
        "tuned_joint_fc/LayerNorm/gamma": lambda c: (c.J,), "tuned_joint_fc/LayerNorm/beta": lambda c: (c.J,),
    }


class VqaAll2Model(VqaAllModel):
    """vqa/model_vlmap_answer_vqa_all2.py: no fill; BCE(tuned) unmasked; pred = argmax(fixed * test + tuned * train)."""
    MODEL_TYPE = "vlmap_answer_vqa_all2"


class NocModel(Model):
    """vqa/model_vlmap_answer_noc.py (and the identical model_vlmap_answer_nocarch.py): no Hadamard fusion; a visual
    branch (pooled_linear_l -> joint_v -> WordWeightAnswerV) and a language branch (q_linear_l -> joint_l ->
    WordWeightAnswerL), logits added; heads initialised from v_/l_class_weights of export_noc_word_weights.py."""
    MODEL_TYPE = "vlmap_answer_noc"
    TRANSFER_SCOPES = ("q_linear_l", "pooled_linear_l", "joint_v", "joint_l")   # vqa/model_vlmap_answer_noc.py:90-101


class NocArchModel(NocModel):
    MODEL_TYPE = "vlmap_answer_nocarch"


class StandardModel(Model):
    """vqa/model_standard.py: same trunk, learned reasoning/classifier head, everything trainable."""
    MODEL_TYPE = "standard"
    TRANSFER_SCOPES = ("encode_L", "GloVe")   # vqa/model_standard.py:86-93


def _name(v):
    n = v if isinstance(v, str) else getattr(v, "name")
    return n.split(":")[0]


def make_synthetic_config(dims, variant="vlmap_answer", precision="bf16", seed=4321, num_images=64,
                          batch_size=None, ragged_boxes=False):
    """A reference-shaped `config` + `image_features` + `batch` built from synthetic data (no files)."""
    from . import synthetic as S
    c = S.dims(**dims)
    params, exist = S.init_params(c, seed=seed, variant=variant)
    feats, nb = S.make_bank(c, num_images=num_images, seed=seed + 1, ragged_boxes=ragged_boxes)
    is_obj, is_attr = S.make_answer_flags(c)
    vocab = {"vocab": [f"w{i}" for i in range(c["Vq"])]}
    vocab["dict"] = {w: i for i, w in enumerate(vocab["vocab"])}
    answers = [f"a{i}" for i in range(c["A"])]
    answer_dict = {"vocab": answers, "dict": {a: i for i, a in enumerate(answers)},
                   "num_train_answer": c["num_train_answer"], "is_object": is_obj, "is_attribute": is_attr}
    config = SimpleNamespace(
        vocab=vocab, answer_dict=answer_dict, vlmap_word_weight_dir=None, image_dir=None, debug=False,
        batch_size=c["B"] if batch_size is None else batch_size, max_q_len=c["T"], v_dim=c["D"], l_dim=c["L"],
        w_dim=c["W"], precision=precision, init_params=params, seed=seed, model_type=variant,
        answer_exist_mask=exist)
    image_features = {"features": feats, "num_boxes": nb, "max_box_num": c["K"], "vfeat_dim": c["Dv"],
                      "spatials": None, "normal_boxes": None}
    batch = S.make_batch(c, num_images, seed=seed + 2, batch=batch_size)
    return config, image_features, batch, exist
