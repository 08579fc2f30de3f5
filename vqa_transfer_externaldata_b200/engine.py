"""Host-side runtime over the C ABI: one `Engine` per GPU / process.

PyTorch is plumbing here (device memory, streams, pinned staging buffers, torch.distributed); every
kernel on the path is in libvqa_answer_b200.so and is reached through ctypes with raw pointers.
"""
import ctypes as C
import os
from dataclasses import dataclass, asdict

import numpy as np
import torch

from . import lib as L

# TF checkpoint variable names (vqa/model_vlmap_answer.py scopes via vlmap/modules.py:78-86,126,596,632-647)
TF_NAMES = {
    "embed": "LearnGloVe/embed_map",
    "v_w": "v_linear_v/fc/weights", "v_b": "v_linear_v/fc/biases",
    "v_gamma": "v_linear_v/LayerNorm/gamma", "v_beta": "v_linear_v/LayerNorm/beta",
    "gru_gates_w": "encode_L/rnn/gru_cell/gates/kernel", "gru_gates_b": "encode_L/rnn/gru_cell/gates/bias",
    "gru_cand_w": "encode_L/rnn/gru_cell/candidate/kernel", "gru_cand_b": "encode_L/rnn/gru_cell/candidate/bias",
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
_REASONING = ("pl_", "ql_", "joint_")


_QP_SCOPE = {"vlmap_answer2": "q_L_ft2", "vlmap_answer_no_noise": "q_L_mean"}   # model_vlmap_answer2.py:127-130 / _no_noise.py:122-125
_QP_LEAF = {"qp_w": "fc/weights", "qp_b": "fc/biases", "qp_gamma": "LayerNorm/gamma", "qp_beta": "LayerNorm/beta"}


_NOC_NAMES = {   # vqa/model_vlmap_answer_noc.py:177-203
    "joint_w": "joint_v/fc/weights", "joint_b": "joint_v/fc/biases",
    "joint_gamma": "joint_v/LayerNorm/gamma", "joint_beta": "joint_v/LayerNorm/beta",
    "jl_w": "joint_l/fc/weights", "jl_b": "joint_l/fc/biases",
    "jl_gamma": "joint_l/LayerNorm/gamma", "jl_beta": "joint_l/LayerNorm/beta",
    "ans_w": "WordWeightAnswerV/fc/weights", "ans_b": "WordWeightAnswerV/fc/biases",
    "al_w": "WordWeightAnswerL/fc/weights", "al_b": "WordWeightAnswerL/fc/biases",
}
_NOC = ("vlmap_answer_noc", "vlmap_answer_nocarch")
_QP_SCOPE["vlmap_answer_full"] = "q_L_mean"     # vqa/model_vlmap_answer_full.py:124-131
_EXTRA_NAMES = {
    "qs_w": "q_L_log_sigma_sq/fc/weights", "qs_b": "q_L_log_sigma_sq/fc/biases",
    "tw_w": "TunedWordWeightAnswer/fc/weights", "tw_b": "TunedWordWeightAnswer/fc/biases",   # _vqa_all.py:215-219
    "va_w": "v_adapt/fc/weights", "va_b": "v_adapt/fc/biases",                               # _adapt.py:132-135
    "va_gamma": "v_adapt/LayerNorm/gamma", "va_beta": "v_adapt/LayerNorm/beta",
}


def tf_name(field, variant):
    """Checkpoint variable name of a parameter field for a model_type."""
    if variant in _NOC and field in _NOC_NAMES:
        return _NOC_NAMES[field]
    if field in _EXTRA_NAMES:
        return _EXTRA_NAMES[field]
    if field in _QP_LEAF:
        return _QP_SCOPE[variant] + "/" + _QP_LEAF[field]
    if variant == "standard":  # vqa/model_standard.py:251-275
        if field.startswith(_REASONING):
            return "reasoning/" + TF_NAMES[field]
        if field == "ans_w":
            return "reasoning/classifier/fc/weights"
        if field == "ans_b":
            return "reasoning/classifier/fc/biases"
    return TF_NAMES[field]


@dataclass
class AnswerModelConfig:
    B: int = 512
    K: int = 36
    Dv: int = 2048
    D: int = 1024          # V_DIM
    L: int = 1024          # L_DIM
    J: int = 2048
    A: int = 3000
    T: int = 14
    W: int = 300           # W_DIM
    Vq: int = 8192
    num_train_answer: int = 2250
    variant: str = "vlmap_answer"   # vqa/importer.py model_type: 'vlmap_answer' | 'standard'
    precision: str = "bf16"         # 'bf16' | 'fp32'
    keep_att: float = 0.8
    keep_joint: float = 0.5
    num_marginal: int = 200         # NUM_MARGINAL of the ent variant (vqa/model_vlmap_answer_ent.py:16)

    def shape(self, field):
        c = asdict(self)
        from .synthetic import PARAM_SHAPES
        return PARAM_SHAPES[field](c)

    def to_c(self):
        return L.VqaConfig(
            B=self.B, K=self.K, Dv=self.Dv, D=self.D, L=self.L, J=self.J, A=self.A, T=self.T, W=self.W,
            Vq=self.Vq, num_train_answer=self.num_train_answer,
            variant=L.VARIANTS[self.variant],
            precision={"bf16": L.PREC_BF16, "fp32": L.PREC_FP32}[self.precision],
            keep_att=self.keep_att, keep_joint=self.keep_joint, num_marginal=self.num_marginal)


def frozen_fields(variant):
    """filter_train_vars: vqa/model_vlmap_answer.py:81-89 drops q_linear_l, pooled_linear_l, joint_fc,
    WordWeightAnswer; vqa/model_standard.py:80-84 trains everything."""
    if variant == "standard":
        return set()
    # vlmap_answer, vlmap_answer2 (model_vlmap_answer2.py:69-78) and vlmap_answer_no_noise (:66-74) freeze the same
    # four scopes; their extra question layer trains
    fz = {f for f in L.PARAM_FIELDS if TF_NAMES[f].split("/")[0] in
          ("q_linear_l", "pooled_linear_l", "joint_fc", "WordWeightAnswer")}
    if variant in _NOC:   # model_vlmap_answer_noc.py:78-88: joint_v, joint_l, WordWeightAnswerV/L frozen as well
        fz |= {"jl_w", "jl_b", "jl_gamma", "jl_beta", "al_w", "al_b"}
    return fz


def _align(n, a=64):
    return (n + a - 1) // a * a


LATE_GRAD_FIELDS = ("embed", "gru_gates_w", "gru_gates_b", "gru_cand_w", "gru_cand_b")


class ParamStore:
    """fp32 master parameters in TF layout inside ONE flat device buffer: trainable tensors first (so the
    gradient all-reduce and the fused clip+Adam step run over one contiguous slice), frozen after."""
    TAIL = 64

    def __init__(self, cfg, device):
        self.cfg = cfg
        frozen = frozen_fields(cfg.variant)
        self.fields = L.param_fields(cfg.variant)
        trainable = [f for f in self.fields if f not in frozen]
        # gradients complete BEFORE the GRU's BPTT first (everything but the embedding and the GRU): with early
        # gradients enabled that slice is all-reduced under the recurrent kernels (dp.DataParallel)
        early = [f for f in trainable if f not in LATE_GRAD_FIELDS]
        self.trainable = early + [f for f in trainable if f in LATE_GRAD_FIELDS]
        self.frozen = [f for f in self.fields if f in frozen]
        self.offsets, off = {}, 0
        self.n_early = 0
        for f in self.trainable:
            n = int(np.prod(cfg.shape(f)))
            self.offsets[f] = (off, n)
            off += _align(n)  # 256-byte aligned starts (float4 / TMA friendly)
            if f in early:
                self.n_early = off
        self.n_train = off
        for f in self.frozen:
            n = int(np.prod(cfg.shape(f)))
            self.offsets[f] = (off, n)
            off += _align(n)
        self.n_total = off
        self.flat = torch.zeros(self.n_total, dtype=torch.float32, device=device)
        # TAIL floats past the gradients travel with them through the data-parallel all-reduce: slot 0 holds the sum of
        # squares of the embedding gradient's IndexedSlices rows (vqa_set_embedding_slice_norm)
        self.grad_buf = torch.zeros(self.n_train + self.TAIL, dtype=torch.float32, device=device)
        self.grad = self.grad_buf[:self.n_train]
        self.adam_m = None
        self.adam_v = None
        self.views = {f: self.flat[o:o + n].view(cfg.shape(f)) for f, (o, n) in self.offsets.items()}
        self.grad_views = {f: self.grad[self.offsets[f][0]:self.offsets[f][0] + self.offsets[f][1]]
                           .view(cfg.shape(f)) for f in self.trainable}

    def rebind_grad(self, new_grad):
        """Move the flat gradient buffer (e.g. into symmetric memory for the multicast all-reduce)."""
        assert new_grad.numel() >= self.n_train + self.TAIL and new_grad.dtype == torch.float32
        new_grad[:self.n_train + self.TAIL].zero_()
        self.grad_buf = new_grad[:self.n_train + self.TAIL]
        self.grad = new_grad[:self.n_train]
        self.grad_views = {f: self.grad[self.offsets[f][0]:self.offsets[f][0] + self.offsets[f][1]]
                           .view(self.cfg.shape(f)) for f in self.trainable}

    def load(self, params):
        """params: dict field -> array-like (numpy / torch), fp32, TF layout."""
        for f in self.fields:
            t = torch.as_tensor(np.asarray(params[f], dtype=np.float32))
            if tuple(t.shape) != tuple(self.cfg.shape(f)):
                raise ValueError(f"{f}: shape {tuple(t.shape)} != {tuple(self.cfg.shape(f))}")
            self.views[f].copy_(t)

    def by_tf_name(self):
        return {tf_name(f, self.cfg.variant): self.views[f] for f in self.fields}

    def c_params(self):
        p = L.VqaParams()
        for f in self.fields:
            setattr(p, f, self.views[f].data_ptr())
        return p

    def c_grads(self):
        g = L.VqaParams()
        for f in self.fields:
            setattr(g, f, self.grad_views[f].data_ptr() if f in self.grad_views else None)
        return g


class PendingScalars:
    _pool = {}

    def __init__(self, eng):
        slots = PendingScalars._pool.setdefault(id(eng), [])
        self.buf = slots.pop() if slots else torch.zeros(1 + L.NUM_REPORT).pin_memory()
        self._slots = slots
        # ONE device->host copy of [loss | report], on the read-back stream: the compute stream only records an event
        # (a copy enqueued on the compute stream itself costs it ~10 us of copy-engine latency per step)
        eng.sync_outputs()
        main = torch.cuda.current_stream(eng.device)
        eng.d2h_stream.wait_stream(main)
        with torch.cuda.stream(eng.d2h_stream):
            self.buf.copy_(eng.o_scalars, non_blocking=True)
            self.event = torch.cuda.Event()
            self.event.record(eng.d2h_stream)
        # the scalars are double-buffered: only the forward after next rewrites this buffer, and it waits for this copy
        eng._last_d2h[eng._oi] = self.event
        self.nbytes = 4 * (1 + L.NUM_REPORT)
        self._keys, self._all = eng.report_keys, eng._all_report_keys

    def get(self):
        self.event.synchronize()
        vals = self.buf.tolist()
        self._slots.append(self.buf)
        full = dict(zip(self._all, vals[1:]))
        return vals[0], {k: full[k] for k in self._keys}


class BatchSet:
    """One set of static batch buffers: pinned host staging + device (fixed addresses)."""

    def __init__(self, cfg, dev):
        self.d_image_idx = torch.zeros(cfg.B, dtype=torch.int64, device=dev)
        self.d_q = torch.zeros(cfg.B * cfg.T, dtype=torch.int32, device=dev)
        self.d_qlen = torch.zeros(cfg.B, dtype=torch.int32, device=dev)
        self.d_target = torch.zeros(cfg.B * cfg.A, dtype=torch.float32, device=dev)
        self.h_image_idx = torch.zeros(cfg.B, dtype=torch.int64).pin_memory()
        self.h_q = torch.zeros(cfg.B * cfg.T, dtype=torch.int32).pin_memory()
        self.h_qlen = torch.zeros(cfg.B, dtype=torch.int32).pin_memory()
        self.h_target = torch.zeros(cfg.B * cfg.A, dtype=torch.float32).pin_memory()
        # sparse soft-score targets (input_native batches): (row, answer id, score) triples, densified on the device
        self.ans_cap = 16 * cfg.B
        self.d_ans = torch.zeros(3 * self.ans_cap, dtype=torch.int32, device=dev)
        self.h_ans = torch.zeros(3 * self.ans_cap, dtype=torch.int32).pin_memory()
        self.batch_size, self.q_len_max = 0, cfg.T
        self.ready = torch.cuda.Event()
        self.staged = None   # event after the last H2D copies out of the pinned staging buffers of this set
        self.source = None   # the host batch object this set was filled from (prefetch bookkeeping)


class Engine:
    """Owns the C handle, its workspace, the parameter store, the feature bank and static batch buffers."""

    def __init__(self, cfg, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("vqa_transfer_externaldata_b200 needs a CUDA (sm_100a) device; there is no CPU path")
        self.lib = L.load()
        self.cfg = cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        torch.cuda.set_device(self.device)
        self._cfg_c = cfg.to_c()
        self.h = C.c_void_p()
        L.check(self.lib.vqa_create(C.byref(self._cfg_c), C.byref(self.h)))
        nbytes = C.c_uint64()
        L.check(self.lib.vqa_workspace_bytes(self.h, C.byref(nbytes)))
        self.workspace = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=self.device)
        base = (self.workspace.data_ptr() + 255) // 256 * 256
        L.check(self.lib.vqa_set_workspace(self.h, C.c_void_p(base), nbytes.value))
        self.params = ParamStore(cfg, self.device)
        self._p = self.params.c_params()
        self._g = self.params.c_grads()
        # clip norm with the embedding gradient as TF holds it (IndexedSlices rows); VQA_DENSE_CLIP_NORM=1 = dense norm
        self.slice_norm = "embed" in self.params.trainable and os.environ.get("VQA_DENSE_CLIP_NORM", "0") != "1"
        self._bind_slice_slot()
        dev = self.device
        # two sets of static batch buffers (addresses stay fixed -> capturable): while a step computes on one
        # set, the next batch is uploaded into the other on a copy stream (prefetch_batch)
        self._sets = [BatchSet(cfg, dev), BatchSet(cfg, dev)]
        self._cur = 0
        self.copy_stream = torch.cuda.Stream(device=dev)
        # outputs
        # [loss | report], read back with one copy; two buffers used alternately, so that a step never waits for the
        # read-back of the step before it (that copy only starts when the previous step has finished)
        self._o_scalars = [torch.zeros(1 + L.NUM_REPORT, device=dev) for _ in range(2)]
        self._oi = 0
        self.d2h_stream = torch.cuda.Stream(device=dev)
        self._last_d2h = [None, None]
        self.o_att = torch.zeros(cfg.B, cfg.K, device=dev)
        self.o_logit = torch.zeros(cfg.B, cfg.A, device=dev)
        self.o_pred = torch.zeros(cfg.B, dtype=torch.int32, device=dev)
        self.o_per_sample = torch.zeros(len(L.PER_SAMPLE_KEYS) * cfg.B, device=dev)
        self.o_condition = torch.zeros(cfg.B, cfg.L, device=dev)
        # adapt pools the D-wide v_adapt (vqa/model_vlmap_answer_adapt.py:142)
        self.o_pooled = torch.zeros(cfg.B, cfg.D if cfg.variant == "vlmap_answer_adapt" else cfg.Dv, device=dev)
        self.h_scalars = torch.zeros(1 + L.NUM_REPORT).pin_memory()
        # report keys this model_type fills: the 13 common ones (+ the latent losses of the full variant)
        self._all_report_keys = L.REPORT_KEYS + L.EXTRA_REPORT_KEYS
        self.report_keys = L.REPORT_KEYS + L.VARIANT_REPORT_KEYS.get(cfg.variant, [])
        self.grad_norm = torch.zeros(1, device=dev)
        self._outs_all = [L.VqaOutputs(
            loss=sc[:1].data_ptr(), report=sc[1:].data_ptr(), att_score=self.o_att.data_ptr(),
            logit=self.o_logit.data_ptr(), pred=self.o_pred.data_ptr(),
            per_sample=self.o_per_sample.data_ptr(), condition=self.o_condition.data_ptr(),
            pooled=self.o_pooled.data_ptr()) for sc in self._o_scalars]
        self._outs_min_all = [L.VqaOutputs(loss=sc[:1].data_ptr(), report=sc[1:].data_ptr()) for sc in self._o_scalars]
        self.bank = None
        self.masks = None
        self.adam_t = 0
        self._adam_tail = 0
        self._deferred = False
        self.early_gradients = False
        # cross-step feature prefetch (vqa_prefetch_features): the next batch's gather runs under this step's BPTT, as
        # 2-CTA clusters on the ten TPCs the cooperative recurrent grid leaves idle; bit-identical results
        # (tests/test_model_gpu.py), ~1 % faster (1.032 vs 1.04 ms/step). VQA_PREFETCH_FEATURES=0 turns it off.
        self.prefetch_features = os.environ.get("VQA_PREFETCH_FEATURES", "1") != "0"
        self.validate_inputs = os.environ.get("VQA_VALIDATE_INPUTS", "1") != "0"   # host-side range checks in _fill

    def close(self):
        if self.h:
            self.lib.vqa_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- residents ------------------------------------------------------------------------------
    def set_feature_bank(self, features, num_boxes):
        """image_features [N,K,Dv] fp32 + num_boxes [N] -> HBM once (the reference keeps them in host RAM
        and ships 151 MB per step, vqa/model_vlmap_answer.py:57-70,110-117)."""
        f = torch.as_tensor(features)
        if f.dim() != 3 or f.shape[1] != self.cfg.K or f.shape[2] != self.cfg.Dv:
            raise ValueError(f"feature bank shape {tuple(f.shape)} != [N, {self.cfg.K}, {self.cfg.Dv}]")
        self.bank_features = f.to(self.device, dtype=torch.float32).contiguous()
        self.bank_num_boxes = torch.as_tensor(np.asarray(num_boxes)).to(self.device, dtype=torch.int32).contiguous()
        # bf16 mode: a one-off bf16 copy of the bank (the very conversion the per-step gather would do, done once), so
        # the gather reads half the bytes. VQA_BANK_BF16=0 keeps the fp32-only layout.
        self.bank_bf16 = None
        if self.cfg.precision == "bf16" and os.environ.get("VQA_BANK_BF16", "1") != "0":
            n = self.bank_features.numel()
            self.bank_bf16 = torch.empty(self.bank_features.shape, dtype=torch.bfloat16, device=self.device)
            cols = self.cfg.K * self.cfg.Dv
            L.check(self.lib.vqa_split_bf16(self.h, self.bank_features.data_ptr(), n // cols, cols, cols,
                                            self.bank_bf16.data_ptr(), None, cols, self._stream()))
        self.bank = L.VqaFeatureBank(features=self.bank_features.data_ptr(),
                                     num_boxes=self.bank_num_boxes.data_ptr(),
                                     num_images=self.bank_features.shape[0],
                                     features_bf16=self.bank_bf16.data_ptr() if self.bank_bf16 is not None else None)

    def set_answer_masks(self, is_object, is_attribute, answer_exist):
        dev = self.device
        self._m = [torch.as_tensor(np.asarray(x, dtype=np.float32)).to(dev).contiguous()
                   for x in (is_object, is_attribute, answer_exist)]
        for m in self._m:
            if m.numel() != self.cfg.A:
                raise ValueError("answer mask length != A")
        self.masks = L.VqaAnswerMasks(is_object=self._m[0].data_ptr(), is_attribute=self._m[1].data_ptr(),
                                      answer_exist=self._m[2].data_ptr())

    def load_params(self, params):
        self.params.load(params)
        self.prepare_params()

    def prepare_params(self, trainable_only=False):
        """Refresh the bf16 operand shadows; after an optimizer step only the trainable weights changed."""
        p = self._p
        if trainable_only:
            p = L.VqaParams()
            for f in self.params.trainable:
                setattr(p, f, self.params.views[f].data_ptr())
        L.check(self.lib.vqa_prepare_params(self.h, C.byref(p), self._stream()))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- batch ----------------------------------------------------------------------------------
    @property
    def cur(self):
        return self._sets[self._cur]

    o_scalars = property(lambda self: self._o_scalars[self._oi])   # of the most recent forward
    o_loss = property(lambda self: self._o_scalars[self._oi][:1])
    o_report = property(lambda self: self._o_scalars[self._oi][1:])

    # the active set's buffers under their historical names
    d_image_idx = property(lambda self: self.cur.d_image_idx)
    d_q = property(lambda self: self.cur.d_q)
    d_qlen = property(lambda self: self.cur.d_qlen)
    d_target = property(lambda self: self.cur.d_target)
    h_image_idx = property(lambda self: self.cur.h_image_idx)
    h_q = property(lambda self: self.cur.h_q)
    h_qlen = property(lambda self: self.cur.h_qlen)
    h_target = property(lambda self: self.cur.h_target)

    @property
    def batch_size(self):
        return self.cur.batch_size

    @batch_size.setter
    def batch_size(self, v):
        self.cur.batch_size = v

    @property
    def q_len_max(self):
        return self.cur.q_len_max

    @q_len_max.setter
    def q_len_max(self, v):
        self.cur.q_len_max = v

    def _fill(self, bs, batch):
        """Host batch dict (keys of input_ops_vqa_tf_record_memft.py:47-59) -> device buffers of set `bs`,
        asynchronously on the current stream. Values may be NumPy arrays (staged through pinned memory) or
        pinned torch tensors of the right dtype (copied directly). Returns the h2d byte count."""
        cfg = self.cfg
        q = batch["q_intseq"]
        Bn, T = int(q.shape[0]), int(q.shape[1])
        if Bn > cfg.B or T > cfg.T:
            raise ValueError(f"batch [{Bn}, T={T}] exceeds config (B={cfg.B}, T={cfg.T})")
        sparse = "answer_target" not in batch and "answer_sparse" in batch
        if not sparse and tuple(batch["answer_target"].shape) != (Bn, cfg.A):
            raise ValueError(f"answer_target shape {tuple(batch['answer_target'].shape)} != ({Bn}, {cfg.A})")
        items = [("image_idx", bs.h_image_idx, bs.d_image_idx, torch.int64, np.int64, Bn),
                 ("q_intseq", bs.h_q, bs.d_q, torch.int32, np.int32, Bn * T),
                 ("q_intseq_len", bs.h_qlen, bs.d_qlen, torch.int32, np.int32, Bn)]
        if not sparse:
            items.append(("answer_target", bs.h_target, bs.d_target, torch.float32, np.float32, Bn * cfg.A))
        nbytes = 0
        used_staging = False
        if sparse:
            # tf.sparse_to_dense on the device (vqa_densify_targets): the step uploads the triples, not [B, A] floats
            rows, ids, scores = (np.asarray(x) for x in batch["answer_sparse"])
            n = int(rows.shape[0])
            if n > bs.ans_cap:
                raise ValueError(f"{n} answer triples exceed the staging capacity {bs.ans_cap}")
            if n and (int(rows.min()) < 0 or int(rows.max()) >= Bn or int(ids.min()) < 0 or int(ids.max()) >= cfg.A):
                raise IndexError("answer_sparse: row / answer id out of range")
            if bs.staged is not None:
                bs.staged.synchronize()
            used_staging = True
            cap = bs.ans_cap
            h = bs.h_ans.numpy()
            h[:n] = rows.astype(np.int32, copy=False)
            h[cap:cap + n] = ids.astype(np.int32, copy=False)
            h[2 * cap:2 * cap + n] = scores.astype(np.float32, copy=False).view(np.int32)
            for k in range(3):
                bs.d_ans[k * cap:k * cap + n].copy_(bs.h_ans[k * cap:k * cap + n], non_blocking=True)
            d0 = bs.d_ans.data_ptr()
            L.check(self.lib.vqa_densify_targets(C.c_void_p(d0), C.c_void_p(d0 + 4 * cap), C.c_void_p(d0 + 8 * cap),
                                                 C.c_int32(n), C.c_int32(Bn), C.c_int32(cfg.A),
                                                 C.c_void_p(bs.d_target.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
            nbytes += 12 * n
        for key, hbuf, dbuf, tdt, ndt, n in items:
            v = batch[key]
            direct = isinstance(v, torch.Tensor) and v.dtype == tdt and v.is_contiguous() and (v.is_pinned() or v.is_cuda)
            if self.validate_inputs and not (direct and v.is_cuda):
                self._validate(key, v, Bn, T)   # host inputs are range-checked here; device-resident ones by the kernels
            if direct:
                src = v.view(-1)   # pinned host memory, or already on the device (device-resident input pipelines)
            else:
                if not used_staging and bs.staged is not None:
                    # the previous upload out of this set's pinned staging buffers may still be queued behind kernels
                    # (train_step(sync=False) lets the host run ahead): rewriting them now would change THAT step's data
                    bs.staged.synchronize()
                used_staging = True
                hbuf[:n].copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(v, dtype=ndt)).reshape(-1)))
                src = hbuf[:n]
            dbuf[:n].copy_(src, non_blocking=True)
            nbytes += n * dbuf.element_size()
        if used_staging:
            if bs.staged is None:
                bs.staged = torch.cuda.Event()
            bs.staged.record(torch.cuda.current_stream(self.device))
        bs.batch_size, bs.q_len_max = Bn, T
        return nbytes

    def _validate(self, key, v, Bn, T):
        """Range checks of a HOST batch field, with the reference's error behaviour: tf.nn.embedding_lookup raises on
        ids outside [0, Vq), np.take on an image_idx outside [-N, N) (negative ones wrap: parse_fn's default for a
        missing feature is -1, vqa/datasets/input_ops_vqa_tf_record_memft.py:28-46), dynamic_rnn needs
        0 <= len <= T. Device-resident inputs are checked by the kernels instead (vqa_input_error_count)."""
        if key == "answer_target":
            return
        a = v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        if a.size == 0:
            return
        lo, hi = int(a.min()), int(a.max())
        if key == "image_idx":
            n = int(self.bank.num_images) if self.bank is not None else None
            if n is not None and (lo < -n or hi >= n):
                raise IndexError(f"image_idx out of range: [{lo}, {hi}] for a feature bank of {n} images")
        elif key == "q_intseq":
            if lo < 0 or hi >= self.cfg.Vq:
                raise IndexError(f"q_intseq token id out of range: [{lo}, {hi}] for a vocabulary of {self.cfg.Vq}")
        elif key == "q_intseq_len":
            if lo < 0 or hi > T:
                raise ValueError(f"q_intseq_len out of range: [{lo}, {hi}] for padded length {T}")

    def check_input_errors(self):
        """Raise if a gather kernel met an out-of-range image_idx / token id since the last check (device-resident
        inputs; host inputs never get that far). Synchronises with the device."""
        n = C.c_uint32()
        L.check(self.lib.vqa_input_error_count(C.byref(n), 1))
        if n.value:
            raise IndexError(f"{n.value} out-of-range image_idx / q_intseq entries reached the device gathers "
                             "(replaced by index 0; results of those samples are meaningless)")

    def stage_batch(self, batch):
        """Upload `batch` into the active set on the current stream (or adopt it if prefetch_batch already
        put this very object into the other set). Returns h2d bytes."""
        other = self._sets[1 - self._cur]
        if other.source is not None and other.source is batch:
            other.source = None
            self._cur = 1 - self._cur
            torch.cuda.current_stream(self.device).wait_event(other.ready)
            return other.nbytes
        return self._fill(self.cur, batch)

    def prefetch_batch(self, batch):
        """Start uploading the NEXT batch into the inactive set on the copy stream; the following
        stage_batch(batch) with the same object adopts it. The inactive set must not be in use: its last
        reader (the step before the current one) has been enqueued on the compute stream before this call."""
        other = self._sets[1 - self._cur]
        self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            other.nbytes = self._fill(other, batch)
            if self.bank is not None and self.prefetch_features:
                # the feature gather depends on no parameter: register the NEXT batch; this step's backward launches
                # its gather next to the weight-gradient GEMMs (the event recorded here orders it after the upload)
                b = L.VqaBatch(batch_size=other.batch_size, q_len_max=other.q_len_max,
                               image_idx=other.d_image_idx.data_ptr(), q_intseq=other.d_q.data_ptr(),
                               q_intseq_len=other.d_qlen.data_ptr(), answer_target=other.d_target.data_ptr())
                L.check(self.lib.vqa_prefetch_features(self.h, C.byref(self.bank), C.byref(b),
                                                       C.c_void_p(self.copy_stream.cuda_stream)))
            other.ready.record(self.copy_stream)
        other.source = batch

    def upload_staged(self, Bn, T):
        A = self.cfg.A
        bs = self.cur
        bs.d_image_idx[:Bn].copy_(bs.h_image_idx[:Bn], non_blocking=True)
        bs.d_q[:Bn * T].copy_(bs.h_q[:Bn * T], non_blocking=True)
        bs.d_qlen[:Bn].copy_(bs.h_qlen[:Bn], non_blocking=True)
        bs.d_target[:Bn * A].copy_(bs.h_target[:Bn * A], non_blocking=True)
        bs.batch_size, bs.q_len_max = Bn, T
        return Bn * 8 + Bn * T * 4 + Bn * 4 + Bn * A * 4

    def _c_batch(self):
        bs = self.cur
        return L.VqaBatch(batch_size=bs.batch_size, q_len_max=bs.q_len_max,
                          image_idx=bs.d_image_idx.data_ptr(), q_intseq=bs.d_q.data_ptr(),
                          q_intseq_len=bs.d_qlen.data_ptr(), answer_target=bs.d_target.data_ptr())

    # ---- the path -------------------------------------------------------------------------------
    def forward(self, seed=0, step=0, full_outputs=True, defer_outputs=False):
        """defer_outputs: a training step (a backward pass follows): the loss / metrics kernels run on an auxiliary
        stream beside the start of the backward pass, which joins them (vqa_set_deferred_outputs)."""
        if self.bank is None or self.masks is None:
            raise RuntimeError("set_feature_bank() and set_answer_masks() first")
        if defer_outputs != self._deferred:
            L.check(self.lib.vqa_set_deferred_outputs(self.h, int(defer_outputs)))
            self._deferred = defer_outputs
        b = self._c_batch()
        self._oi ^= 1                    # this step's loss / report go to the other scalar buffer
        outs = (self._outs_all if full_outputs else self._outs_min_all)[self._oi]
        if self._last_d2h[self._oi] is not None:   # the read-back that used this buffer two steps ago
            torch.cuda.current_stream(self.device).wait_event(self._last_d2h[self._oi])
            self._last_d2h[self._oi] = None
        L.check(self.lib.vqa_forward(self.h, C.byref(self._p), C.byref(self.bank), C.byref(b),
                                     C.byref(self.masks), C.c_uint64(seed), C.c_uint64(step), C.byref(outs),
                                     self._stream()))

    def backward(self, loss_scale=1.0):
        b = self._c_batch()
        L.check(self.lib.vqa_backward(self.h, C.byref(self._p), C.byref(b), C.byref(self._g),
                                      C.c_float(loss_scale), self._stream()))

    def _bind_slice_slot(self):
        slot = self.params.grad_buf[self.params.n_train:].data_ptr() if self.slice_norm else None
        L.check(self.lib.vqa_set_embedding_slice_norm(self.h, slot))

    def rebind_gradients(self, new_grad):
        """Use `new_grad` (>= n_train fp32 elements) as the flat gradient buffer from now on."""
        self.params.rebind_grad(new_grad)
        self._g = self.params.c_grads()
        self._bind_slice_slot()

    def set_early_gradients(self, enable=True):
        """Data-parallel overlap: produce the non-GRU gradients before the BPTT (vqa_set_early_gradients)."""
        L.check(self.lib.vqa_set_early_gradients(self.h, int(bool(enable))))
        self.early_gradients = bool(enable)

    def stream_wait_early_gradients(self, stream):
        L.check(self.lib.vqa_stream_wait_early_gradients(self.h, C.c_void_p(stream.cuda_stream)))

    def adam_step(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, clip_norm=20.0, pipelined_tail=False):
        """optimize_loss(Adam, clip_gradients=20.0) over the trainable slice (vqa/trainer.py:106-114).
        pipelined_tail (training loops): the embedding / GRU slice of the buffer is updated on an auxiliary stream, under
        the first kernels of the next forward (vqa_set_optimizer_tail); read parameters through sync_params() then."""
        ps = self.params
        if ps.adam_m is None:
            ps.adam_m = torch.zeros_like(ps.grad)
            ps.adam_v = torch.zeros_like(ps.grad)
        self.adam_t += 1
        tail = ps.n_early if (pipelined_tail and 0 < ps.n_early < ps.n_train and os.environ.get("VQA_ADAM_TAIL", "1") != "0") else 0
        if tail != self._adam_tail:
            L.check(self.lib.vqa_set_optimizer_tail(self.h, tail))
            self._adam_tail = tail
        # one pass: clip + Adam + the bf16 operand shadows of the updated weight matrices (+ the GRU repack)
        L.check(self.lib.vqa_adam_step_shadowed(self.h, C.byref(self._p), ps.flat.data_ptr(), ps.grad.data_ptr(),
                                                ps.adam_m.data_ptr(), ps.adam_v.data_ptr(), ps.n_train, lr, beta1,
                                                beta2, eps, clip_norm, self.adam_t, self.grad_norm.data_ptr(),
                                                self._stream()))

    def sync_params(self):
        """Order the current stream after a pipelined optimizer tail (no-op otherwise): call before reading parameters."""
        L.check(self.lib.vqa_sync_params(self.h, self._stream()))

    def dropout_masks(self, seed, step, batch=None):
        Bn = self.batch_size if batch is None else batch
        att = torch.empty(Bn, self.cfg.K, self.cfg.D, dtype=torch.uint8, device=self.device)
        joint = torch.empty(Bn, self.cfg.J, dtype=torch.uint8, device=self.device)
        L.check(self.lib.vqa_dropout_masks(self.h, Bn, C.c_uint64(seed), C.c_uint64(step), att.data_ptr(),
                                           joint.data_ptr(), self._stream()))
        return att, joint

    def dropout_mask_site(self, site, seed, step, batch=None):
        """0 / 1 keep mask of one dropout site (L.SITE_*) for (seed, step)."""
        Bn = self.batch_size if batch is None else batch
        shape = ((Bn, self.cfg.K, self.cfg.D) if site == L.SITE_ATT else
                 (Bn, self.cfg.num_marginal, self.cfg.J) if site == L.SITE_ENT else (Bn, self.cfg.J))
        out = torch.empty(*shape, dtype=torch.uint8, device=self.device)
        L.check(self.lib.vqa_dropout_mask_site(self.h, site, Bn, C.c_uint64(seed), C.c_uint64(step), out.data_ptr(),
                                               self._stream()))
        return out

    def reparam_noise(self, seed, step, batch=None):
        """The N(0, 1) draw [batch, L] the full variant's forward uses for (seed, step) (vqa_reparam_noise)."""
        Bn = self.batch_size if batch is None else batch
        out = torch.empty(Bn, self.cfg.L, dtype=torch.float32, device=self.device)
        L.check(self.lib.vqa_reparam_noise(self.h, Bn, C.c_uint64(seed), C.c_uint64(step), out.data_ptr(), self._stream()))
        return out

    def peek_activation(self, which, dtype, shape):
        """Copy of an activation the last forward saved in the workspace (vqa_peek_activation)."""
        ptr, nbytes = C.c_void_p(), C.c_uint64()
        L.check(self.lib.vqa_peek_activation(self.h, which, C.byref(ptr), C.byref(nbytes)))
        off = ptr.value - self.workspace.data_ptr()
        return self.workspace[off:off + nbytes.value].view(dtype).view(shape).clone()

    def sync_outputs(self):
        """Join deferred outputs of the last forward into the current stream (no-op otherwise)."""
        L.check(self.lib.vqa_sync_outputs(self.h, self._stream()))

    def read_scalars(self):
        """D2H of loss + report (pinned, synchronises the current stream)."""
        self.sync_outputs()
        self.h_scalars.copy_(self.o_scalars, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self.check_input_errors()
        vals = self.h_scalars.tolist()
        full = dict(zip(self._all_report_keys, vals[1:]))
        return vals[0], {k: full[k] for k in self.report_keys}

    def read_scalars_async(self):
        """Enqueue the D2H of loss + report into a fresh pinned slot and return a handle; handle.get() waits for
        THAT copy only, so the host can run one step ahead of the device (asynchronous dispatch)."""
        return PendingScalars(self)

    def outputs(self):
        Bn = self.batch_size
        ps = self.o_per_sample[:len(L.PER_SAMPLE_KEYS) * Bn].view(len(L.PER_SAMPLE_KEYS), Bn)
        out = {"att_score": self.o_att[:Bn], "logit": self.o_logit[:Bn], "pred": self.o_pred[:Bn]}
        out.update({k: ps[i] for i, k in enumerate(L.PER_SAMPLE_KEYS)})
        return out

    def launch_count(self):
        return int(self.lib.vqa_launch_count())
