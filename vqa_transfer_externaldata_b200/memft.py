"""The vlmap pre-training graph (BASELINE config 4) on the B200 operator library.

Mirror of vlmap_memft/model_vlmap_bf_or_wordset_withatt_sp.py: `Model(batch, config, is_train)` with `.loss`,
`.losses`, `.report`, `.mid_result` and the checkpoint variable names of the reference (TF_NAMES). The reference's Python
builds a TensorFlow graph out of library operators; this Python enqueues the same operators of libvqa_answer_b200.so
(include/vqa_memft.h) on one CUDA stream -- tcgen05 GEMMs, the CTA-pair GRU kernels, slab LayerNorm, the spatial
attention block, softmax cross-entropy with top-k -- forward and hand-derived backward, then global-norm clip + Adam
(vlmap_memft/trainer.py:126-150). There is no CPU path and no autograd: every gradient is a kernel of the library.

Row layouts: "entries" E = 2 * B * n rows ordered [kind (obj, attr)][image][entry]; "head rows" R = 4 * B * n ordered
[head (obj_blank_fill, attr_blank_fill, obj_wordset, attr_wordset)][image][entry]. The object and attribute branches
share every variable (tf.AUTO_REUSE), so each layer is ONE GEMM over all its rows; pooled_linear_l is evaluated once
per kind and shared by the blank-fill and wordset heads (the reference evaluates it twice on the same input).
"""
import ctypes as C
import os
from types import SimpleNamespace

import numpy as np
import torch

from . import lib as L

TOP_K = 5
KINDS = ("obj", "attr")
HEADS = ("obj_blank_fill", "attr_blank_fill", "obj_wordset", "attr_wordset")
SITE_ATT0, SITE_JOINT0 = 16, 18     # dropout sites: 16, 17 attention features (obj, attr); 18..21 joint of the four heads

# field -> (shape as a function of the config, checkpoint variable name of the reference)
FIELDS = {
    "wordset_map": (lambda c: (c.Nws, c.W), "wordset_map/embed_map"),
    "l_glove": (lambda c: (c.Vq, c.W), "L_GloVe/embed_map"),
    "sv_w": (lambda c: (6, c.D), "spat_v_linear_v/fc/weights"), "sv_b": (lambda c: (c.D,), "spat_v_linear_v/fc/biases"),
    "sv_gamma": (lambda c: (c.D,), "spat_v_linear_v/LayerNorm/gamma"), "sv_beta": (lambda c: (c.D,), "spat_v_linear_v/LayerNorm/beta"),
    "sq_w": (lambda c: (6, c.D), "spat_q_linear_v/fc/weights"), "sq_b": (lambda c: (c.D,), "spat_q_linear_v/fc/biases"),
    "sq_gamma": (lambda c: (c.D,), "spat_q_linear_v/LayerNorm/gamma"), "sq_beta": (lambda c: (c.D,), "spat_q_linear_v/LayerNorm/beta"),
    "att_w": (lambda c: (c.D, 1), "spat_att/compute/score/fc/weights"), "att_b": (lambda c: (1,), "spat_att/compute/score/fc/biases"),
    "gru_gates_w": (lambda c: (c.W + c.L, 2 * c.L), "encode_L_blank/rnn/gru_cell/gates/kernel"),
    "gru_gates_b": (lambda c: (2 * c.L,), "encode_L_blank/rnn/gru_cell/gates/bias"),
    "gru_cand_w": (lambda c: (c.W + c.L, c.L), "encode_L_blank/rnn/gru_cell/candidate/kernel"),
    "gru_cand_b": (lambda c: (c.L,), "encode_L_blank/rnn/gru_cell/candidate/bias"),
    "pl_w": (lambda c: (c.Dv, c.L), "pooled_linear_l/fc/weights"), "pl_b": (lambda c: (c.L,), "pooled_linear_l/fc/biases"),
    "pl_gamma": (lambda c: (c.L,), "pooled_linear_l/LayerNorm/gamma"), "pl_beta": (lambda c: (c.L,), "pooled_linear_l/LayerNorm/beta"),
    "ql_w": (lambda c: (c.L, c.L), "q_linear_l/fc/weights"), "ql_b": (lambda c: (c.L,), "q_linear_l/fc/biases"),
    "ql_gamma": (lambda c: (c.L,), "q_linear_l/LayerNorm/gamma"), "ql_beta": (lambda c: (c.L,), "q_linear_l/LayerNorm/beta"),
    "joint_w": (lambda c: (c.L, 2 * c.L), "joint_fc/fc/weights"), "joint_b": (lambda c: (2 * c.L,), "joint_fc/fc/biases"),
    "joint_gamma": (lambda c: (2 * c.L,), "joint_fc/LayerNorm/gamma"), "joint_beta": (lambda c: (2 * c.L,), "joint_fc/LayerNorm/beta"),
    "cls_w": (lambda c: (2 * c.L, c.A), "classifier/fc/weights"), "cls_b": (lambda c: (c.A,), "classifier/fc/biases"),
    "ws_w": (lambda c: (c.W, c.L), "wordset_ft/fc/weights"), "ws_b": (lambda c: (c.L,), "wordset_ft/fc/biases"),
    "ws_gamma": (lambda c: (c.L,), "wordset_ft/LayerNorm/gamma"), "ws_beta": (lambda c: (c.L,), "wordset_ft/LayerNorm/beta"),
}
TF_NAMES = {k: v[1] for k, v in FIELDS.items()}
GEMM_WEIGHTS = ("pl_w", "ql_w", "joint_w", "cls_w", "ws_w")   # matrices kept as bf16 operand planes beside the fp32 master


def make_config(dims, precision="bf16", keep_att=0.8, keep_joint=0.5, learning_rate=1e-3):
    """dims: B images, K proposals, n entries per kind, Dv feature width, D = V_DIM, L = L_DIM, W = W_DIM, A answers,
    T blank length, Vq vocabulary, Nws wordsets (model_vlmap_bf_or_wordset_withatt_sp.py:8-10, datasets/dataset_vlmap.py:11-14)."""
    c = SimpleNamespace(**dims)
    c.precision, c.keep_att, c.keep_joint, c.learning_rate = precision, keep_att, keep_joint, learning_rate
    return c


class _Planes:
    """A GEMM operand: bf16 plane (+ residual plane in fp32 mode), row-major [rows, ld]."""

    def __init__(self, rows, ld, fp32, dev):
        self.hi = torch.zeros(rows, ld, dtype=torch.bfloat16, device=dev)
        self.lo = torch.zeros(rows, ld, dtype=torch.bfloat16, device=dev) if fp32 else None
        self.ld = ld

    def ptr(self, row=0):
        off = 2 * row * self.ld
        return self.hi.data_ptr() + off, (self.lo.data_ptr() + off) if self.lo is not None else None


def _p(t, off_elems=0):
    return None if t is None else C.c_void_p(t.data_ptr() + off_elems * t.element_size())


class Model:
    def __init__(self, batch, config, is_train=True, params=None, seed=777, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("vqa_transfer_externaldata_b200.memft needs a CUDA (sm_100a) device; there is no CPU path")
        self.config = c = config
        self.is_train = is_train
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lib = L.load()
        h = C.c_void_p()
        L.check(self.lib.vqa_ops_create(C.byref(h)))
        self.ops = h
        self.fp32 = c.precision == "fp32"
        # batch-sharded data parallelism (BASELINE config 4: 8 x B200): one process per GPU, every row of the graph is per
        # image, so the ranks only meet in one all-reduce of the flat gradient buffer (NCCL) and in the two valid-entry
        # counts that normalise the loss. Per-rank dropout streams as in the answer model: seed + 7919 * rank.
        import torch.distributed as dist
        self._dist = dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.rank = self._dist.get_rank() if self._dist else 0
        self.world = self._dist.get_world_size() if self._dist else 1
        self.seed = int(seed) + 7919 * self.rank
        self.global_counts = None
        self.global_step = 0
        self.losses, self.report, self.mid_result = {}, {}, {}
        self.loss = None
        for k in ("D", "L", "A", "Dv"):
            if getattr(c, k) % 8:
                raise ValueError(f"{k} must be a multiple of 8")
        if c.n > 8 or c.K > 256 or c.W % 4:
            raise ValueError("n <= 8 entries per kind, K <= 256 proposals, W a multiple of 4")
        # ---- parameters: one flat fp32 buffer (+ gradient, Adam moments), a view per variable ----
        sizes = {k: int(np.prod(f[0](c))) for k, f in FIELDS.items()}
        self.n_param = sum((s + 3) & ~3 for s in sizes.values())
        z = lambda: torch.zeros(self.n_param, dtype=torch.float32, device=self.dev)   # noqa: E731
        self.flat, self.flat_grad, self.adam_m, self.adam_v = z(), z(), z(), z()
        self.p, self.g = {}, {}
        off = 0
        for k, f in FIELDS.items():
            shp = f[0](c)
            self.p[k] = self.flat[off:off + sizes[k]].view(shp)
            self.g[k] = self.flat_grad[off:off + sizes[k]].view(shp)
            off += (sizes[k] + 3) & ~3
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self._alloc()
        if params is not None:
            self.load_state_dict(params, by_field=True)
        self.batch = None
        if batch is not None:
            self.set_batch(batch)

    # ---- buffers -----------------------------------------------------------------------------------------------------
    def _alloc(self):
        c, dev, fp32 = self.config, self.dev, self.fp32
        B, K, n, D, L_, Dv, A, W = c.B, c.K, c.n, c.D, c.L, c.Dv, c.A, c.W
        E, R = 2 * B * n, 4 * B * n
        self.E, self.R, self.Wp = E, R, (W + 7) & ~7
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)   # noqa: E731
        P = lambda rows, ld: _Planes(rows, ld, fp32, dev)                 # noqa: E731
        b = self.buf = SimpleNamespace()
        # weights in operand form
        b.sv_wp, b.sq_wp = P(64, D), P(64, D)
        b.sv_w64, b.sq_w64 = f(64, D), f(64, D)
        b.w = {k: P(*FIELDS[k][0](c)) for k in GEMM_WEIGHTS}
        # spatial attention
        b.spat_p, b.key_p = P(B * K, 64), P(E, 64)
        b.zv, b.hv, b.mean_v, b.rstd_v = f(B * K, D), P(B * K, D), f(B), f(B)
        b.zq, b.hq, b.mean_q, b.rstd_q = f(E, D), f(E, D), f(2 * B), f(2 * B)
        b.att, b.pooled, b.pooled_p = f(E, K), f(E, Dv), P(E, Dv)
        # heads
        b.zp, b.vl, b.mean_p, b.rstd_p = f(E, L_), f(E, L_), f(2 * B), f(2 * B)
        b.ws_y, b.ws_p, b.zws, b.mean_ws, b.rstd_ws = f(E, W), P(E, self.Wp), f(E, L_), f(2 * B), f(2 * B)
        b.q = f(E, L_)
        b.lang = P(R, L_)
        b.zl, b.ll, b.x, b.mean_l, b.rstd_l = f(R, L_), f(R, L_), P(R, L_), f(4 * B), f(4 * B)
        b.zj, b.jd, b.mean_j, b.rstd_j = f(R, 2 * L_), P(R, 2 * L_), f(4 * B), f(4 * B)
        b.logit = f(R, A)
        b.stats, b.rep = f(R, 4), f(16)
        # gradients of activations
        b.dlogit, b.dlogit_p = f(R, A), P(R, A)
        b.djd, b.dzj, b.part_j = f(R, 2 * L_), P(R, 2 * L_), f(4 * B, 3 * 2 * L_)
        b.dx, b.dzl, b.dvl4, b.part_l = f(R, L_), P(R, L_), f(R, L_), f(4 * B, 3 * L_)
        b.dlang = f(R, L_)
        b.dzws, b.part_ws, b.dws_y = P(E, L_), f(2 * B, 3 * L_), f(E, W)
        b.dzp, b.part_p, b.dpooled = P(E, L_), f(2 * B, 3 * L_), f(E, Dv)
        b.d_hv, b.d_hq, b.part_att = f(2 * B * K, D), f(E, D), f(2 * B, D + 8)
        # keep bits of the attention dropout stored by the forward pass for the backward pass: measured SLOWER than drawing
        # them again (backward 2.92 -> 3.06 ms at cfg4: the kernels are latency-bound with idle ALUs, a dependent byte load
        # per 8 elements is not free), so off unless VQA_MEMFT_KEEP_BITS=1
        self.store_keep_bits = os.environ.get("VQA_MEMFT_KEEP_BITS") == "1" and c.keep_att < 1.0
        b.att_bits = torch.zeros(E * K * (D // 8) if self.store_keep_bits else 8, dtype=torch.uint8, device=dev)
        b.dzq, b.part_q = P(E, D), f(2 * B, 3 * D)
        b.dzv, b.part_v = P(B * K, D), f(B, 3 * D)
        b.sum3 = f(3 * max(2 * L_, D) + 64)
        b.fw_part = f(296, 6 * D)
        # GRU sequence operator
        self._gru = L.VqaGruSeq(B=E, T=c.T, L=L_, W=W, Vq=c.Vq, precision=L.PREC_FP32 if fp32 else L.PREC_BF16)
        nbytes = C.c_uint64()
        L.check(self.lib.vqa_ops_gru_workspace_bytes(C.byref(self._gru), C.byref(nbytes)))
        b.gru_ws = torch.zeros(nbytes.value + 256, dtype=torch.uint8, device=dev)
        base = (b.gru_ws.data_ptr() + 255) & ~255
        self._gru.ws, self._gru.ws_bytes = base, nbytes.value
        # batch on the device
        d = self.dbatch = SimpleNamespace()
        d.image_ft, d.spatial_ft = f(B, K, Dv), f(B, K, 6)
        i32 = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)   # noqa: E731
        d.num_boxes = i32(B)
        d.boxes, d.blanks, d.blanks_len = f(2, B, n, 4), i32(E, c.T), i32(E)
        d.fills4, d.num, d.wordsets = i32(R), i32(2, B), i32(E)
        self._own = {"image_ft": d.image_ft, "spatial_ft": d.spatial_ft}

    def close(self):
        if getattr(self, "ops", None):
            self.lib.vqa_ops_destroy(self.ops)
            self.ops = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters --------------------------------------------------------------------------------------------------
    def state_dict(self):
        """Variables under the reference's checkpoint names."""
        return {TF_NAMES[k]: v.detach().cpu().numpy().copy() for k, v in self.p.items()}

    def load_state_dict(self, params, by_field=False):
        for k in FIELDS:
            key = k if by_field else TF_NAMES[k]
            if key in params:
                self.p[k].copy_(torch.as_tensor(np.asarray(params[key], np.float32).reshape(self.p[k].shape)))
        self._refresh_operands()

    def _split(self, src, planes):
        rows, cols = src.shape
        hi, lo = planes.ptr()
        L.check(self.lib.vqa_ops_split_bf16(_p(src), rows, cols, cols, C.c_void_p(hi), C.c_void_p(lo), planes.ld, self._s()))

    def _refresh_operands(self):
        """fp32 master weights -> bf16 operand planes (after loading and after every optimizer step)."""
        b = self.buf
        for k in GEMM_WEIGHTS:
            self._split(self.p[k], b.w[k])
        for w64, planes, k in ((b.sv_w64, b.sv_wp, "sv_w"), (b.sq_w64, b.sq_wp, "sq_w")):
            w64[:6].copy_(self.p[k])            # the 6-d box features ride a K = 64 operand, rows 6..63 zero
            self._split(w64, planes)

    # ---- batch -------------------------------------------------------------------------------------------------------
    def set_batch(self, batch):
        """Host batch dict with the keys of vlmap_memft/datasets/dataset_vlmap.py (image_ft, spatial_ft, num_boxes,
        {obj,attr}_blank_fill/{normal_boxes, blanks, blanks_len, fills, num, wordsets}) -> device buffers."""
        c, d = self.config, self.dbatch
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(np.asarray(a, dt)))   # noqa: E731
        B, n, T = c.B, c.n, c.T
        if tuple(batch["image_ft"].shape) != (B, c.K, c.Dv):
            raise ValueError(f"image_ft shape {tuple(batch['image_ft'].shape)} != {(B, c.K, c.Dv)}")
        for key in ("image_ft", "spatial_ft"):
            v = batch[key]
            if isinstance(v, torch.Tensor) and v.is_cuda:   # device-resident features (a bank gathered on the device): adopted, no copy
                if v.dtype != torch.float32 or not v.is_contiguous():
                    raise ValueError(f"{key}: device tensors must be contiguous float32")
                setattr(d, key, v)
            else:
                if getattr(d, key).data_ptr() != self._own[key].data_ptr():
                    setattr(d, key, self._own[key])
                getattr(d, key).copy_(t(v, np.float32), non_blocking=True)
        d.num_boxes.copy_(t(batch["num_boxes"], np.int32), non_blocking=True)
        fills = []
        for i, kind in enumerate(KINDS):
            pre = f"{kind}_blank_fill/"
            blanks = np.asarray(batch[pre + "blanks"], np.int32).reshape(B * n, -1)
            if blanks.shape[1] > T:
                raise ValueError(f"blanks longer than the configured T = {T}")
            if blanks.min() < 0 or blanks.max() >= c.Vq:
                raise IndexError("blank token id out of range")
            pad = np.zeros((B * n, T), np.int32)
            pad[:, :blanks.shape[1]] = blanks
            ln = np.asarray(batch[pre + "blanks_len"], np.int32).reshape(-1)
            if ln.min() < 0 or ln.max() > blanks.shape[1]:
                raise ValueError("blanks_len out of range")
            ws = np.asarray(batch[pre + "wordsets"], np.int32).reshape(-1)
            f_ = np.asarray(batch[pre + "fills"], np.int32).reshape(-1)
            if ws.min() < 0 or ws.max() >= c.Nws or f_.min() < 0 or f_.max() >= c.A:
                raise IndexError("wordset / fill id out of range")
            d.boxes[i].copy_(t(batch[pre + "normal_boxes"], np.float32).view(B, n, 4), non_blocking=True)
            d.blanks[i * B * n:(i + 1) * B * n].copy_(torch.from_numpy(pad), non_blocking=True)
            d.blanks_len[i * B * n:(i + 1) * B * n].copy_(torch.from_numpy(ln), non_blocking=True)
            d.wordsets[i * B * n:(i + 1) * B * n].copy_(torch.from_numpy(ws), non_blocking=True)
            d.num[i].copy_(t(batch[pre + "num"], np.int32), non_blocking=True)
            fills.append(f_)
        d.fills4.copy_(torch.from_numpy(np.concatenate(fills + fills)), non_blocking=True)
        if self._dist:
            cnt = torch.tensor([float(np.minimum(np.maximum(np.asarray(batch[f"{k}_blank_fill/num"]), 0), n).sum()) for k in KINDS],
                               dtype=torch.float32, device=self.dev)
            self._dist.all_reduce(cnt)
            self.global_counts = cnt.cpu().tolist()
        torch.cuda.current_stream(self.dev).synchronize()   # the staging arrays above are temporaries
        self.batch = batch

    # ---- operator wrappers ---------------------------------------------------------------------------------------------
    def _s(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _gemm(self, M, N, K, a, a_mn, b, b_mn, bias=None, addend=None, out=None, planes=None, a_row=0, b_row=0, out_row=0,
              ld_out=None):
        """D[M, N] = A[M, K] B[N, K]^T (+ bias) (+ addend). a / b: _Planes; *_mn: operand stored [K, rows] (see VqaGemmDesc)."""
        d = L.VqaGemmDesc()
        d.a_hi, d.a_lo = a.ptr(a_row)
        d.b_hi, d.b_lo = b.ptr(b_row)
        d.lda, d.ldb, d.a_mn_major, d.b_mn_major = a.ld, b.ld, int(a_mn), int(b_mn)
        d.M, d.N, d.K = M, N, K
        d.bias = None if bias is None else bias.data_ptr()
        if addend is not None:
            d.addend, d.ld_addend = addend.data_ptr(), addend.shape[-1]
        if out is not None:
            ld = out.shape[-1] if ld_out is None else ld_out
            d.out_f32, d.ld_f32 = out.data_ptr() + 4 * out_row * ld, ld
        if planes is not None:
            d.out_hi, d.out_lo = planes.ptr(out_row)
            d.ld_bf = planes.ld
        L.check(self.lib.vqa_ops_gemm(self.ops, C.byref(d), 0, self._s()))

    def _slab(self, slabs, n, N, z, gamma, beta, mean, rstd, act=0, mul=None, mul_rows=0, keep=1.0, site0=0, rows_per_site=1,
              step=0, y=None, out_f32=None, out=None, out_row=0, dout=None, dout2=None, dz=None, dmul=None, part=None, bwd=False):
        a = L.VqaSlabLn(slabs=slabs, n=n, N=N, act=act, z=z.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                        mul=None if mul is None else mul.data_ptr(), mul_rows=mul_rows, keep=keep, seed=self.seed, step=step,
                        site0=site0, rows_per_site=rows_per_site, mean=mean.data_ptr(), rstd=rstd.data_ptr())
        if not bwd:
            a.y = None if y is None else y.data_ptr()
            a.out_f32 = None if out_f32 is None else out_f32.data_ptr()
            if out is not None:
                a.out_hi, a.out_lo = out.ptr(out_row)
            L.check(self.lib.vqa_ops_slab_ln_fwd(self.ops, C.byref(a), self._s()))
        else:
            a.dout = dout.data_ptr()
            a.dout2 = None if dout2 is None else dout2.data_ptr()
            if dz is not None:
                a.dz_hi, a.dz_lo = dz.ptr()
            a.dmul = None if dmul is None else dmul.data_ptr()
            a.part = None if part is None else part.data_ptr()
            L.check(self.lib.vqa_ops_slab_ln_bwd(self.ops, C.byref(a), self._s()))

    def _colsum(self, x, rows, cols, ld, out):
        L.check(self.lib.vqa_ops_colsum(self.ops, _p(x), rows, cols, ld, _p(out), self._s()))

    def _ln_param_grads(self, part, slabs, N, g_gamma, g_beta, g_bias):
        """per-slab partials [slabs, 3, N] -> d gamma, d beta, d bias"""
        s = self.buf.sum3[:3 * N]
        self._colsum(part, slabs, 3 * N, 3 * N, s)
        g_gamma.copy_(s[:N])
        g_beta.copy_(s[N:2 * N])
        g_bias.copy_(s[2 * N:3 * N])

    def _spat(self, step):
        c, b, d = self.config, self.buf, self.dbatch
        hv_hi, hv_lo = b.hv.ptr()
        p_hi, p_lo = b.pooled_p.ptr()
        return L.VqaSpatAttn(B=c.B, K=c.K, n=c.n, D=c.D, Dv=c.Dv, kinds=2, hv_hi=hv_hi, hv_lo=hv_lo, hq=b.hq.data_ptr(),
                             att_w=self.p["att_w"].data_ptr(), att_b=self.p["att_b"].data_ptr(),
                             num_boxes=d.num_boxes.data_ptr(), v=d.image_ft.data_ptr(), keep=c.keep_att, seed=self.seed,
                             step=step, site0=SITE_ATT0, att=b.att.data_ptr(), pooled=b.pooled.data_ptr(), pooled_hi=p_hi,
                             pooled_lo=p_lo, d_pooled=b.dpooled.data_ptr(), d_hv=b.d_hv.data_ptr(), d_hq=b.d_hq.data_ptr(),
                             part=b.part_att.data_ptr(), keep_bits=b.att_bits.data_ptr() if self.store_keep_bits else None)

    # ---- the graph ---------------------------------------------------------------------------------------------------
    def forward(self, batch=None, dropout_step=None, with_grad_seed=None):
        """build() of the reference as kernel launches (:56-74). with_grad_seed: also emit d loss / d logits."""
        if batch is not None and batch is not self.batch:
            self.set_batch(batch)
        c, b, d, p = self.config, self.buf, self.dbatch, self.p
        B, K, n, D, L_, Dv, A, W, E, R = c.B, c.K, c.n, c.D, c.L, c.Dv, c.A, c.W, self.E, self.R
        step = self.global_step if dropout_step is None else int(dropout_step)
        self._step = step
        train = self.is_train if with_grad_seed is None else with_grad_seed
        s = self._s()
        # spatial attention (:323-365, :413-455): 6-d box features of the proposals against the key box of each entry
        hi, lo = b.spat_p.ptr()
        L.check(self.lib.vqa_ops_pad_planes(_p(d.spatial_ft), B * K, 6, 0, C.c_void_p(hi), C.c_void_p(lo), 64, s))
        hi, lo = b.key_p.ptr()
        L.check(self.lib.vqa_ops_pad_planes(_p(d.boxes), E, 4, 1, C.c_void_p(hi), C.c_void_p(lo), 64, s))
        self._gemm(B * K, D, 64, b.spat_p, False, b.sv_wp, True, bias=p["sv_b"], out=b.zv)
        self._slab(B, K, D, b.zv, p["sv_gamma"], p["sv_beta"], b.mean_v, b.rstd_v, out=b.hv)          # LN over [K, D]
        self._gemm(E, D, 64, b.key_p, False, b.sq_wp, True, bias=p["sq_b"], out=b.zq)
        self._slab(2 * B, n, D, b.zq, p["sq_gamma"], p["sq_beta"], b.mean_q, b.rstd_q, out_f32=b.hq)  # LN over [n, D]
        sa = self._spat(step)
        L.check(self.lib.vqa_memft_spat_attn_fwd(self.ops, C.byref(sa), s))
        # pooled_linear_l, once per kind (:523-526)
        self._gemm(E, L_, Dv, b.pooled_p, False, b.w["pl_w"], True, bias=p["pl_b"], out=b.zp)
        self._slab(2 * B, n, L_, b.zp, p["pl_gamma"], p["pl_beta"], b.mean_p, b.rstd_p, y=b.vl)
        # language side: blank-fill GRU (:511-519) -> rows [0, E) of `lang`; wordset features (:373-380) -> rows [E, 2E)
        g = self._gru
        g.embed, g.gates_w, g.gates_b = p["l_glove"].data_ptr(), p["gru_gates_w"].data_ptr(), p["gru_gates_b"].data_ptr()
        g.cand_w, g.cand_b = p["gru_cand_w"].data_ptr(), p["gru_cand_b"].data_ptr()
        g.tokens, g.len, g.q = d.blanks.data_ptr(), d.blanks_len.data_ptr(), b.q.data_ptr()
        g.q_hi, g.q_lo = b.lang.ptr(0)
        L.check(self.lib.vqa_ops_gru_fwd(self.ops, C.byref(g), s))
        hi, lo = b.ws_p.ptr()
        L.check(self.lib.vqa_memft_wordset_fwd(_p(p["wordset_map"]), _p(d.wordsets), E, W, c.Nws, _p(b.ws_y), C.c_void_p(hi),
                                               C.c_void_p(lo), self.Wp, s))
        self._gemm(E, L_, W, b.ws_p, False, b.w["ws_w"], True, bias=p["ws_b"], out=b.zws)
        self._slab(2 * B, n, L_, b.zws, p["ws_gamma"], p["ws_beta"], b.mean_ws, b.rstd_ws, act=1, out=b.lang, out_row=E)
        # q_linear_l, Hadamard with v_linear_l, joint_fc + dropout 0.5, classifier (:528-547) for the four heads at once
        self._gemm(R, L_, L_, b.lang, False, b.w["ql_w"], True, bias=p["ql_b"], out=b.zl)
        self._slab(4 * B, n, L_, b.zl, p["ql_gamma"], p["ql_beta"], b.mean_l, b.rstd_l, mul=b.vl, mul_rows=E, y=b.ll, out=b.x)
        self._gemm(R, 2 * L_, L_, b.x, False, b.w["joint_w"], True, bias=p["joint_b"], out=b.zj)
        self._slab(4 * B, n, 2 * L_, b.zj, p["joint_gamma"], p["joint_beta"], b.mean_j, b.rstd_j, keep=c.keep_joint,
                   site0=SITE_JOINT0, rows_per_site=B * n, step=step, out=b.jd)
        self._gemm(R, A, 2 * L_, b.jd, False, b.w["cls_w"], True, bias=p["cls_b"], out=b.logit)
        # n_way_classification_loss of the four heads (:675-706)
        ce = L.VqaSoftmaxCe(heads=4, B=B, n=n, A=A, top_k=TOP_K, logit=b.logit.data_ptr(), fills=d.fills4.data_ptr(),
                            loss_scale=1.0, stats=b.stats.data_ptr(), report=b.rep.data_ptr())
        for h in range(4):
            ce.num[h] = d.num[h % 2].data_ptr()
            if self.global_counts is not None:
                ce.count[h] = self.global_counts[h % 2]
        if train:
            ce.d_logit = b.dlogit.data_ptr()
            ce.d_hi, ce.d_lo = b.dlogit_p.ptr()
        L.check(self.lib.vqa_memft_softmax_ce(self.ops, C.byref(ce), s))
        self.mid_result = {"object_pooled_V_ft": b.pooled[:B * n].view(B, n, Dv), "attribute_pooled_V_ft": b.pooled[B * n:].view(B, n, Dv),
                           "object_att_score": b.att[:B * n], "attribute_att_score": b.att[B * n:]}
        for h, name in enumerate(HEADS):
            key = name.replace("_blank_fill", "_blank_fill/logit").replace("_wordset", "_wordset/logit")
            self.mid_result[key] = b.logit[h * B * n:(h + 1) * B * n].view(B, n, A)

    def backward(self):
        """Hand-derived gradients of the graph above, every variable of the reference (filter_train_vars keeps all, :46-52)."""
        c, b, d, p, g = self.config, self.buf, self.dbatch, self.p, self.g
        B, K, n, D, L_, Dv, A, W, E, R = c.B, c.K, c.n, c.D, c.L, c.Dv, c.A, c.W, self.E, self.R
        step = self._step
        s = self._s()
        w = b.w
        # classifier
        self._gemm(R, 2 * L_, A, b.dlogit_p, False, w["cls_w"], False, out=b.djd)
        self._gemm(2 * L_, A, R, b.jd, True, b.dlogit_p, True, out=g["cls_w"])
        self._colsum(b.dlogit, R, A, A, g["cls_b"])
        # joint_fc
        self._slab(4 * B, n, 2 * L_, b.zj, p["joint_gamma"], p["joint_beta"], b.mean_j, b.rstd_j, keep=c.keep_joint,
                   site0=SITE_JOINT0, rows_per_site=B * n, step=step, dout=b.djd, dz=b.dzj, part=b.part_j, bwd=True)
        self._ln_param_grads(b.part_j, 4 * B, 2 * L_, g["joint_gamma"], g["joint_beta"], g["joint_b"])
        self._gemm(R, L_, 2 * L_, b.dzj, False, w["joint_w"], False, out=b.dx)
        self._gemm(L_, 2 * L_, R, b.x, True, b.dzj, True, out=g["joint_w"])
        # q_linear_l (d X (.) vl through LN / ReLU) and the gradient reaching v_linear_l (d X (.) ll)
        self._slab(4 * B, n, L_, b.zl, p["ql_gamma"], p["ql_beta"], b.mean_l, b.rstd_l, mul=b.vl, mul_rows=E, dout=b.dx,
                   dz=b.dzl, dmul=b.dvl4, part=b.part_l, bwd=True)
        self._ln_param_grads(b.part_l, 4 * B, L_, g["ql_gamma"], g["ql_beta"], g["ql_b"])
        self._gemm(R, L_, L_, b.dzl, False, w["ql_w"], False, out=b.dlang)
        self._gemm(L_, L_, R, b.lang, True, b.dzl, True, out=g["ql_w"])
        # wordset branch: rows [E, 2E) of d lang
        self._slab(2 * B, n, L_, b.zws, p["ws_gamma"], p["ws_beta"], b.mean_ws, b.rstd_ws, act=1, dout=b.dlang[E:], dz=b.dzws,
                   part=b.part_ws, bwd=True)
        self._ln_param_grads(b.part_ws, 2 * B, L_, g["ws_gamma"], g["ws_beta"], g["ws_b"])
        self._gemm(E, W, L_, b.dzws, False, w["ws_w"], False, out=b.dws_y)
        self._gemm(W, L_, E, b.ws_p, True, b.dzws, True, out=g["ws_w"])
        L.check(self.lib.vqa_memft_wordset_bwd(_p(b.dws_y), _p(b.ws_y), _p(d.wordsets), E, W, c.Nws, _p(g["wordset_map"]), s))
        # blank-fill GRU: rows [0, E) of d lang
        gr = self._gru
        gr.dq = b.dlang.data_ptr()
        gr.d_embed, gr.d_gates_w, gr.d_gates_b = g["l_glove"].data_ptr(), g["gru_gates_w"].data_ptr(), g["gru_gates_b"].data_ptr()
        gr.d_cand_w, gr.d_cand_b = g["gru_cand_w"].data_ptr(), g["gru_cand_b"].data_ptr()
        L.check(self.lib.vqa_ops_gru_bwd(self.ops, C.byref(gr), s))
        # pooled_linear_l: the two heads of a kind both read it
        self._slab(2 * B, n, L_, b.zp, p["pl_gamma"], p["pl_beta"], b.mean_p, b.rstd_p, dout=b.dvl4[:E], dout2=b.dvl4[E:],
                   dz=b.dzp, part=b.part_p, bwd=True)
        self._ln_param_grads(b.part_p, 2 * B, L_, g["pl_gamma"], g["pl_beta"], g["pl_b"])
        self._gemm(E, Dv, L_, b.dzp, False, w["pl_w"], False, out=b.dpooled)
        self._gemm(Dv, L_, E, b.pooled_p, True, b.dzp, True, out=g["pl_w"])
        # spatial attention
        sa = self._spat(step)
        L.check(self.lib.vqa_memft_spat_attn_bwd(self.ops, C.byref(sa), s))
        sums = b.sum3[:D + 8]
        self._colsum(b.part_att, 2 * B, D + 8, D + 8, sums)
        g["att_w"].view(-1).copy_(sums[:D])
        g["att_b"].copy_(sums[D:D + 1])
        self._slab(2 * B, n, D, b.zq, p["sq_gamma"], p["sq_beta"], b.mean_q, b.rstd_q, dout=b.d_hq, dz=b.dzq, part=b.part_q, bwd=True)
        self._ln_param_grads(b.part_q, 2 * B, D, g["sq_gamma"], g["sq_beta"], g["sq_b"])
        hi, lo = b.dzq.ptr()
        L.check(self.lib.vqa_ops_feat_wgrad(self.ops, _p(d.boxes), 6, 1, C.c_void_p(hi), C.c_void_p(lo), E, D, _p(b.fw_part),
                                            b.fw_part.shape[0], _p(g["sq_w"]), s))
        self._slab(B, K, D, b.zv, p["sv_gamma"], p["sv_beta"], b.mean_v, b.rstd_v, dout=b.d_hv[:B * K], dout2=b.d_hv[B * K:],
                   dz=b.dzv, part=b.part_v, bwd=True)
        self._ln_param_grads(b.part_v, B, D, g["sv_gamma"], g["sv_beta"], g["sv_b"])
        hi, lo = b.dzv.ptr()
        L.check(self.lib.vqa_ops_feat_wgrad(self.ops, _p(d.spatial_ft), 6, 0, C.c_void_p(hi), C.c_void_p(lo), B * K, D, _p(b.fw_part),
                                            b.fw_part.shape[0], _p(g["sv_w"]), s))

    def adam_step(self, lr=None, clip_norm=20.0, beta1=0.9, beta2=0.999, eps=1e-8):
        lr = self.config.learning_rate if lr is None else lr
        L.check(self.lib.vqa_ops_adam(self.ops, _p(self.flat), _p(self.flat_grad), _p(self.adam_m), _p(self.adam_v), self.n_param,
                                      lr, beta1, beta2, eps, clip_norm, self.global_step + 1, _p(self.grad_norm), self._s()))
        self._refresh_operands()

    def train_step(self, batch=None, lr=None, clip_norm=20.0, apply_optimizer=True, sync=True):
        """run_single_step of vlmap_memft/trainer.py: forward + backward + clip + Adam; returns the loss."""
        self.forward(batch, with_grad_seed=True)
        self.backward()
        if self._dist:
            self._dist.all_reduce(self.flat_grad)
        if apply_optimizer:
            self.adam_step(lr=lr, clip_norm=clip_norm)
        self.global_step += 1
        return self.fetch()[0] if sync else None

    def fetch(self):
        """Synchronise and bind loss / losses / report under the reference's keys (:548-556, :70-73)."""
        r = self.buf.rep[:13].cpu().numpy()
        self.report, self.losses = {}, {}
        for h, name in enumerate(HEADS):
            self.losses[name] = float(r[3 * h])
            self.report[f"{name}_loss"] = float(r[3 * h])
            self.report[f"{name}_acc"] = float(r[3 * h + 1])
            self.report[f"{name}_top_{TOP_K}_acc"] = float(r[3 * h + 2])
        self.loss = float(r[12])
        self.report["total_loss"] = self.loss
        return self.loss, self.report

    def dropout_masks(self, step=None):
        """The keep masks the kernels draw at (seed, step), keyed like the oracle's (parity tests)."""
        c = self.config
        step = self._step if step is None else int(step)
        out = {}

        def site(n_elems, keep, s):
            m = torch.zeros(n_elems, dtype=torch.uint8, device=self.dev)
            L.check(self.lib.vqa_ops_dropout_mask(_p(m), n_elems, keep, self.seed, step, s, self._s()))
            return m.cpu().numpy().astype(np.float64)

        for i, kind in enumerate(KINDS):
            out[f"att/{kind}"] = site(c.B * c.n * c.K * c.D, c.keep_att, SITE_ATT0 + i).reshape(c.B * c.n, c.K, c.D)
        for h, name in enumerate(HEADS):
            out[f"joint/{name}"] = site(c.B * c.n * 2 * c.L, c.keep_joint, SITE_JOINT0 + h).reshape(c.B, c.n, 2 * c.L)
        return out

    def gradients(self):
        torch.cuda.synchronize(self.dev)
        return {k: v.detach().cpu().numpy().copy() for k, v in self.g.items()}


# ---- synthetic workload (BASELINE config 4 shapes; bench.py --mode memft, scripts/gpu_bench_memft.py) ----------------------
CFG4 = dict(B=512, K=36, n=5, Dv=2048, D=1024, L=1024, W=300, A=4000, T=10, Vq=8192, Nws=2000)


def synthetic_batch(dims, seed=0):
    """A batch with the keys / dtypes of vlmap_memft/datasets/dataset_vlmap.py:128-266: non-negative features (post-ReLU
    bottom-up features), 10..K valid proposals, 1..n valid entries per kind, blanks of 1..T tokens."""
    rng = np.random.default_rng(seed)
    B, K, n, T = dims["B"], dims["K"], dims["n"], dims["T"]
    batch = {"image_ft": np.abs(rng.standard_normal((B, K, dims["Dv"]), dtype=np.float32)) * 0.5,
             "spatial_ft": rng.uniform(size=(B, K, 6)).astype(np.float32),
             "num_boxes": rng.integers(min(10, K), K + 1, size=B).astype(np.int32)}
    for kind in KINDS:
        x0, y0 = rng.uniform(0, 0.5, size=(B, n)), rng.uniform(0, 0.5, size=(B, n))
        boxes = np.stack([x0, y0, x0 + rng.uniform(0.1, 0.5, size=(B, n)), y0 + rng.uniform(0.1, 0.5, size=(B, n))], axis=-1)
        ln = rng.integers(1, T + 1, size=(B, n)).astype(np.int32)
        blanks = rng.integers(1, dims["Vq"], size=(B, n, T)).astype(np.int32)
        blanks[np.arange(T)[None, None, :] >= ln[:, :, None]] = 0
        batch.update({f"{kind}_blank_fill/normal_boxes": boxes.astype(np.float32), f"{kind}_blank_fill/blanks": blanks,
                      f"{kind}_blank_fill/blanks_len": ln,
                      f"{kind}_blank_fill/fills": rng.integers(0, dims["A"], size=(B, n)).astype(np.int32),
                      f"{kind}_blank_fill/num": rng.integers(1, n + 1, size=B).astype(np.int32),
                      f"{kind}_blank_fill/wordsets": rng.integers(0, dims["Nws"], size=(B, n)).astype(np.int32)})
    return batch


def xavier_params(cfg, seed=1):
    """Random initial variables as TF creates them (Xavier-uniform matrices, LayerNorm gamma 1 / beta 0, GRU gate bias 1)."""
    rng = np.random.default_rng(seed)
    p = {}
    for k, (shp, _) in FIELDS.items():
        s = shp(cfg)
        if k in ("wordset_map", "l_glove"):
            p[k] = rng.standard_normal(s, dtype=np.float32) * 0.4
        elif k.endswith("_gamma") or k == "gru_gates_b":
            p[k] = np.ones(s, np.float32)
        elif len(s) == 2:
            lim = np.sqrt(6.0 / (s[0] + s[1]))
            p[k] = rng.uniform(-lim, lim, size=s).astype(np.float32)
        else:
            p[k] = np.zeros(s, np.float32)
    return p


def gemm_flops_per_step(c):
    """Dense contractions of one train step (forward x 3: data + weight gradients), SURVEY 8d."""
    E, R = 2 * c.B * c.n, 4 * c.B * c.n
    fwd = 2.0 * (E * c.Dv * c.L + E * c.W * c.L + R * c.L * c.L + R * c.L * 2 * c.L + R * 2 * c.L * c.A +
                 E * c.T * (c.W + c.L) * 3 * c.L)
    return 3.0 * fwd
