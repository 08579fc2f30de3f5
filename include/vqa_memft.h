/* C ABI of the vlmap pre-training path (BASELINE config 4): the kernels behind
 *   vlmap_memft/model_vlmap_bf_or_wordset_withatt_sp.py:56-74, 323-609, 675-706
 * as operator-level entry points. The graph itself (which operator follows which, on which buffer) is written in the
 * host language of the reference -- Python: vqa_transfer_externaldata_b200/memft.py mirrors the reference's
 * Model(batch, config, is_train) class -- and calls these through ctypes, the way the reference's Python calls
 * TensorFlow's operator library. Every entry point enqueues on `stream` and returns; nothing allocates or synchronises
 * except vqa_ops_create / vqa_ops_destroy. All pointers are DEVICE pointers unless said otherwise. No CPU fallback.
 *
 * Shapes: B images, K proposals, n entries per image and kind (5 objects + 5 attributes, datasets/dataset_vlmap.py:11-14),
 * rows = B * n entries of one kind; "slab" = the n consecutive rows of one image: modules.fc_layer on a rank-3 input
 * normalises over the whole [n, dim] slab of a sample (vlmap/modules.py:630-650). */
#ifndef VQA_MEMFT_H_
#define VQA_MEMFT_H_

#include "vqa_answer.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct VqaOps_t* VqaOps;

/* operator context: SM count, split-K semaphores and reduction scratch (device memory owned by the context) */
VQA_API VqaStatus vqa_ops_create(VqaOps* out);
VQA_API VqaStatus vqa_ops_destroy(VqaOps ops);

/* D = A B^T (+bias) (+addend): vqa_gemm without an answer-model handle (tcgen05 kernels of csrc/gemm.cu, gemm_pair.cu).
 * narrow = 1: one CTA pair per output tile, no split-K (GEMMs that run side by side). */
VQA_API VqaStatus vqa_ops_gemm(VqaOps ops, const VqaGemmDesc* d, int32_t narrow, void* stream);
/* out[c] = sum over rows of x[r, c] (bias / LayerNorm parameter gradients), fixed summation order */
VQA_API VqaStatus vqa_ops_colsum(VqaOps ops, const float* x, int64_t rows, int64_t cols, int64_t ld, float* out,
                                 void* stream);
/* global-norm clip + Adam over a flat parameter buffer (vlmap_memft/trainer.py:126-150: clip 20, Adam 1e-3) */
VQA_API VqaStatus vqa_ops_adam(VqaOps ops, float* param, const float* grad, float* m, float* v, int64_t n, float lr,
                               float beta1, float beta2, float eps, float clip_norm, int64_t t, float* grad_norm_out,
                               void* stream);
/* fp32 [rows, cols] (pitch ld) -> bf16 operand planes (hi, and the residual lo if non-NULL) with pitch ld_out */
VQA_API VqaStatus vqa_ops_split_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld, void* hi, void* lo,
                                     int64_t ld_out, void* stream);
/* the keep flags (one byte per element) of dropout site `site` at (seed, step): what the kernels below regenerate */
VQA_API VqaStatus vqa_ops_dropout_mask(uint8_t* out, int64_t n, float keep, uint64_t seed, uint64_t step, uint32_t site,
                                       void* stream);

/* modules.fc_layer's tail on a rank-3 tensor: LayerNorm over each [n, N] slab (gamma / beta over the last axis),
 * activation, optional Hadamard partner, optional dropout:  y = act(LN(z)),  out = y * mul * keep_mask / keep.
 * Forward reads z, writes mean / rstd [slabs] and any of y (fp32), out_f32, out planes (bf16 hi [+ lo residual]).
 * Backward reads dout (+ dout2, added to it) = d loss / d out and writes dz (fp32 and / or planes), dmul = d loss / d mul,
 * part [slabs, 3, N] = per-slab partials of d gamma | d beta | d bias (column sums over the slab's rows).
 * mul row of output row r = r % mul_rows. Dropout site of row r = site0 + r / rows_per_site, element index inside the
 * site = (r % rows_per_site) * N + column. n * N <= 49152. */
typedef struct VqaSlabLn {
  int32_t slabs, n, N;
  int32_t act;                 /* 0 relu, 1 tanh, 2 none */
  const float* z;              /* [slabs * n, N] pre-LN (bias added) */
  const float* gamma; const float* beta;
  const float* mul; int64_t mul_rows;
  float keep; uint64_t seed, step; uint32_t site0; int64_t rows_per_site;
  float* mean; float* rstd;
  float* y; float* out_f32; void* out_hi; void* out_lo;
  const float* dout; const float* dout2;
  float* dz_f32; void* dz_hi; void* dz_lo; float* dmul; float* part;
} VqaSlabLn;
VQA_API VqaStatus vqa_ops_slab_ln_fwd(VqaOps ops, const VqaSlabLn* a, void* stream);
VQA_API VqaStatus vqa_ops_slab_ln_bwd(VqaOps ops, const VqaSlabLn* a, void* stream);

/* modules.fc_layer on a rank-2 input as ONE kernel (csrc/linear_ln.cu; vlmap/modules.py:616-650, call sites
 * vqa/model_vlmap_answer.py:142-181): the tcgen05 product with the layer's tail in its epilogue -- the CTAs of a row tile
 * form a thread-block cluster along N and exchange row statistics through distributed shared memory.
 *   backward = 0:  z = a W + bias;  y = act(LN(z));  out = y * mul * keep_mask / keep        (W: [K, N], TF's [in, out])
 *   backward = 1:  raw = a W'^T (the data gradient through the layer ABOVE: a = d loss / d pre-activation of that layer,
 *                  W' = its weights, the same TF buffer seen as [N, K] with K contiguous);
 *                  dz = this layer's dropout / mul / activation / LayerNorm backward applied to raw
 *                  (z / mean / rstd / gamma / beta / mul / dropout site are those of THIS layer's forward pass)
 * bf16 operands (one plane), fp32 everything else. N = 64 or 128 x {1, 2, 4, 8, 16}; forward needs K % 64 == 0.
 * VQA_ERR_BAD_SHAPE when the shape (or the device: 16-CTA clusters) is not eligible: use vqa_ops_gemm + vqa_ops_slab_ln_*
 * (n = 1). Dropout element index = row * N + column at `site` (as vqa_ops_dropout_mask). */
typedef struct VqaLinearLn {
  int32_t M, N, K;
  int32_t backward;
  const void* a; int64_t lda;
  const void* w; int64_t ldw;
  const float* bias;
  const float* gamma; const float* beta;
  const float* mul;
  int32_t act;                 /* 0 relu, 1 tanh */
  float keep; uint64_t seed, step; uint32_t site;
  float* z; float* mean; float* rstd;      /* forward: outputs; backward: inputs */
  float* y; float* out_f32; void* out_hi;  /* forward outputs (any may be NULL) */
  float* raw; float* dz_f32; void* dz_hi;  /* backward outputs (any may be NULL) */
  float* dgamma_part; float* dbeta_part;   /* backward, optional: [M, N] per-row terms of d gamma / d beta (column-sum them) */
} VqaLinearLn;
VQA_API VqaStatus vqa_ops_linear_ln(VqaOps ops, const VqaLinearLn* a, void* stream);

/* fp32 [rows, cols] -> GEMM operand planes [rows, ld_out] with the columns beyond `cols` zero (6-d box features as a
 * K = 64 operand). boxes != 0: src is [rows, 4] normalised boxes and the operand is (x0, y0, x1, y1, x1 - x0, y1 - y0)
 * (model_vlmap_bf_or_wordset_withatt_sp.py:340-345). */
VQA_API VqaStatus vqa_ops_pad_planes(const float* src, int64_t rows, int32_t cols, int32_t boxes, void* hi, void* lo,
                                     int32_t ld_out, void* stream);

/* Weight gradient of a layer with F <= 8 input features (spat_v_linear_v / spat_q_linear_v: the 6-d box features):
 * out[f, c] = sum over rows of feat[r, f] * dz[r, c]; dz as operand planes [rows, N]. boxes != 0: feat is [rows, 4] normalised
 * boxes and the six features are derived as in vqa_ops_pad_planes. part: scratch [max_parts, F, N] (per-CTA partial sums,
 * reduced in fixed order). F * N <= 8192. */
VQA_API VqaStatus vqa_ops_feat_wgrad(VqaOps ops, const float* feat, int32_t F, int32_t boxes, const void* dz_hi, const void* dz_lo,
                                     int64_t rows, int32_t N, float* part, int32_t max_parts, float* out, void* stream);

/* Spatial Hadamard attention + attended pooling for the n entries of each image and kind (:323-365, :413-455;
 * vlmap/modules.py:67-97, 23-39): score[e, k] = sum_d Hv[b, k, d] Hq[e, d] w[d] keep[e, k, d] / keep_att + bias, -inf
 * beyond num_boxes[b], softmax over k, pooled[e] = sum_k att[e, k] V[b, k, :]. V is read once per image and kind, not
 * tiled n times. kinds = 2 (objects, attributes): rows are [kind][b][e]. Dropout site of kind i = site0 + i. */
typedef struct VqaSpatAttn {
  int32_t B, K, n, D, Dv, kinds;
  const void* hv_hi; const void* hv_lo;   /* [B, K, D] relu(LN(spat Wv)) as planes */
  const float* hq;                         /* [kinds * B * n, D] */
  const float* att_w; const float* att_b;  /* [D], [1] */
  const int32_t* num_boxes;                /* [B] */
  const float* v;                          /* [B, K, Dv] raw image features */
  float keep; uint64_t seed, step; uint32_t site0;
  float* att;                              /* [kinds * B * n, K] */
  float* pooled; void* pooled_hi; void* pooled_lo;   /* [kinds * B * n, Dv] */
  /* backward */
  const float* d_pooled;                   /* [kinds * B * n, Dv] */
  float* d_hv;                             /* [kinds, B, K, D]: one plane per kind (summed over the kind's entries) */
  float* d_hq;                             /* [kinds * B * n, D] */
  float* part;                             /* [kinds * B, D + 8] partials per image and kind: d att_w [D] | d att_b (slot D) */
  uint8_t* keep_bits;                      /* optional [kinds * B * n * K * D / 8]: the forward pass leaves the keep bits of
                                            * every group of 8 feature columns here, the backward pass reads them back */
} VqaSpatAttn;
VQA_API VqaStatus vqa_memft_spat_attn_fwd(VqaOps ops, const VqaSpatAttn* a, void* stream);
VQA_API VqaStatus vqa_memft_spat_attn_bwd(VqaOps ops, const VqaSpatAttn* a, void* stream);

/* n_way_classification_loss (:675-706) for `heads` heads of rows_per_head = B * n rows each: masked softmax
 * cross-entropy (mean over the valid entries e < num[b]), top-1 and top-k accuracy (tf.nn.top_k: the lower index wins a
 * tie). stats [rows, 4] = ce | top-1 | top-k | valid per row; report [heads, 3] = loss | acc | top-k acc, report[3 * heads]
 * = sum of the losses. num_of_head[h] points at the [B] valid counts of head h. d_logit (optional, fp32 and / or planes) =
 * loss_scale * (softmax - onehot) * valid / count(head). */
typedef struct VqaSoftmaxCe {
  int32_t heads, B, n, A, top_k;
  const float* logit;            /* [heads * B * n, A] */
  const int32_t* fills;          /* [heads * B * n] target class per row */
  const int32_t* num[8];         /* per head: [B] */
  float loss_scale;
  float count[8];                /* > 0: the head's valid-entry count to normalise d_logit by instead of this batch's own
                                  * (batch-sharded data parallelism: the count of the GLOBAL batch, so that the sum of the
                                  * ranks' gradients is the gradient of the global loss) */
  float* stats; float* report;
  float* d_logit; void* d_hi; void* d_lo;
} VqaSoftmaxCe;
VQA_API VqaStatus vqa_memft_softmax_ce(VqaOps ops, const VqaSoftmaxCe* a, void* stream);

/* tanh(wordset_map[id]) (:373-374) as fp32 [rows, W] + operand planes [rows, ld]; backward: d map[id] += d * (1 - y^2) */
VQA_API VqaStatus vqa_memft_wordset_fwd(const float* map, const int32_t* ids, int64_t rows, int32_t W, int32_t num_ws,
                                        float* y, void* hi, void* lo, int32_t ld, void* stream);
VQA_API VqaStatus vqa_memft_wordset_bwd(const float* d_y, const float* y, const int32_t* ids, int64_t rows, int32_t W,
                                        int32_t num_ws, float* d_map, void* stream);

/* encode_L_blank (:511-519; vlmap/modules.py:124-140): embedding lookup + GRU over `B` sequences of at most T tokens,
 * final state q [B, L]; backward from d q to the embedding map (dense [Vq, W], zeroed here), the two GRU kernels and
 * biases. The recurrent part runs on the CTA-pair kernels of csrc/gru_pair.cu in waves of 512 sequences (bf16 mode) or as
 * per-step GEMMs (fp32 mode). `ws` (vqa_ops_gru_workspace_bytes, 256-byte aligned) holds the saved states between the
 * two calls. */
typedef struct VqaGruSeq {
  int32_t B, T, L, W, Vq, precision;
  const float* embed; const float* gates_w; const float* gates_b; const float* cand_w; const float* cand_b;
  const int32_t* tokens;         /* [B, T] */
  const int32_t* len;            /* [B] */
  void* ws; uint64_t ws_bytes;
  float* q; void* q_hi; void* q_lo;          /* [B, L] final state: fp32 and operand planes */
  const float* dq;                            /* [B, L] */
  float* d_embed; float* d_gates_w; float* d_gates_b; float* d_cand_w; float* d_cand_b;
} VqaGruSeq;
VQA_API VqaStatus vqa_ops_gru_workspace_bytes(const VqaGruSeq* a, uint64_t* bytes);
VQA_API VqaStatus vqa_ops_gru_fwd(VqaOps ops, const VqaGruSeq* a, void* stream);
VQA_API VqaStatus vqa_ops_gru_bwd(VqaOps ops, const VqaGruSeq* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQA_MEMFT_H_ */
