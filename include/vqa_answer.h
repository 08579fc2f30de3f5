/*
 * vqa_answer.h -- C ABI of the B200-native VQA answer-model hot path.
 *
 * This is the drop-in boundary for ONE path of HyeonwooNoh/VQA-Transfer-ExternalData: the batched
 * forward + backward of the vqa/model_vlmap_answer* family (and vqa/model_standard), i.e. what one
 * session.run([loss, report, optimizer]) executes in the reference (vqa/trainer.py:275-287).
 * The reference has no FFI of its own (pure Python on TensorFlow 1.6); these entry points are what a
 * ctypes stub inside vqa/model_vlmap_answer.py would bind.
 *
 * Conventions
 *   - every function returns a VqaStatus (0 = ok, negative = error); vqa_last_error() gives the text
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers unless the name ends in
 *     _host; the caller owns every buffer (the library keeps a caller-provided workspace only)
 *   - all work is enqueued on the cudaStream_t passed as `void* stream`; no hidden synchronisation
 *   - parameters / gradients are fp32 in TensorFlow layout ([in, out] row-major), keyed by the
 *     reference's checkpoint variable names (comment on each field)
 *   - one handle per GPU / host thread; entry points are re-entrant per handle
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 */
#ifndef VQA_ANSWER_H_
#define VQA_ANSWER_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQA_API __attribute__((visibility("default")))

typedef int32_t VqaStatus;
enum {
  VQA_OK = 0,
  VQA_ERR_BAD_ARG = -1,     /* null pointer, bad enum                                     */
  VQA_ERR_BAD_SHAPE = -2,   /* dimension not supported (see vqa_create)                   */
  VQA_ERR_WORKSPACE = -3,   /* workspace missing or too small                             */
  VQA_ERR_CUDA = -4,        /* a CUDA runtime / driver call failed                        */
  VQA_ERR_NO_DEVICE = -5,   /* no sm_100 device visible                                   */
  VQA_ERR_STATE = -6        /* call order violated (e.g. backward before forward)         */
};

/* model family members served by the same kernels (vqa/importer.py:22-51) */
enum {
  VQA_VARIANT_VLMAP_ANSWER = 0, /* vqa/model_vlmap_answer.py: frozen transfer head, train mask */
  VQA_VARIANT_STANDARD = 1,     /* vqa/model_standard.py: learned classifier, all trainable   */
  /* vqa/model_vlmap_answer2.py:127-131,164: q_L_ft2 = tanh(LN(FC(q))) feeds q_linear_l and is the `condition` output */
  VQA_VARIANT_VLMAP_ANSWER2 = 2,
  /* vqa/model_vlmap_answer_no_noise.py:122-125,157: q_L_mean = FC(q) (no LayerNorm, no activation) feeds q_linear_l */
  VQA_VARIANT_VLMAP_ANSWER_NO_NOISE = 3,
  /* vqa/model_vlmap_answer_noc.py:177-203 (model_vlmap_answer_nocarch.py is the same graph): no Hadamard fusion;
   * pooled_linear_l -> joint_v -> WordWeightAnswerV and q_linear_l -> joint_l -> WordWeightAnswerL, each with its own
   * dropout(0.5), logits added; both heads frozen */
  VQA_VARIANT_VLMAP_ANSWER_NOC = 4,
  /* vqa/model_vlmap_answer_full.py:124-134,166,217-223,272-276: q_linear_l reads q_L_mean + N(0,1) * sqrt(exp(
   * q_L_log_sigma_sq)) (both linear layers of q, trained); loss += 0.1 * KL(q_L_mean, q_L_log_sigma_sq) */
  VQA_VARIANT_VLMAP_ANSWER_FULL = 5,
  /* vqa/model_vlmap_answer_vqa_all.py:188-244: frozen word-weight logits with the absent answers filled by the row
   * minimum, plus TunedWordWeightAnswer(joint) (trained; :215-216 feeds `joint`, not tuned_joint); loss =
   * (BCE(fixed) + BCE(fixed + tuned)) * train_mask; logit / pred from fixed + tuned */
  VQA_VARIANT_VLMAP_ANSWER_VQA_ALL = 6,
  /* vqa/model_vlmap_answer_vqa_all2.py:188-243: no fill; loss = BCE(fixed) * train_mask + BCE(tuned); pred =
   * argmax(fixed * test_mask + tuned * train_mask); logit = fixed + tuned */
  VQA_VARIANT_VLMAP_ANSWER_VQA_ALL2 = 7,
  /* vqa/model_vlmap_answer_adapt.py:132-142: attention pools v_adapt = relu(LN(FC(V))) [K, D] (trained) instead of
   * the raw features, so the pooled vector is D-wide and pooled_linear_l/fc/weights is [D, L] */
  VQA_VARIANT_VLMAP_ANSWER_ADAPT = 8,
  /* vqa/model_vlmap_answer_ent.py:14-16,193-213,284-294: the base model plus a maximum-entropy regulariser: the joint
   * head is evaluated on num_marginal tiles per sample (pooled_linear_l rows of OTHER samples, stop_gradient, times
   * this sample's q_linear_l; LayerNorm over the whole [num_marginal, J] slab; dropout 0.5), softmax over the train &
   * existing answers, mean over the tiles = marginal; loss += 0.1 * mean_b sum_a marg log(marg + 1e-8) */
  VQA_VARIANT_VLMAP_ANSWER_ENT = 9,
  VQA_NUM_VARIANTS
};

/* arithmetic mode of the dense contractions */
enum {
  VQA_PREC_BF16 = 0, /* bf16 operands, fp32 accumulate in TMEM; activations between GEMMs bf16  */
  VQA_PREC_FP32 = 1  /* error-compensated: each fp32 operand = bf16 hi + bf16 lo, 3 MMAs / tile */
};

typedef struct VqaConfig {
  int32_t B;   /* max samples per step on this GPU                                            */
  int32_t K;   /* max_box_num (boxes per image), vqa/model_vlmap_answer.py:66                 */
  int32_t Dv;  /* vfeat_dim (2048)                                                            */
  int32_t D;   /* V_DIM (1024), vqa/model_vlmap_answer.py:12                                  */
  int32_t L;   /* L_DIM (1024), :11                                                           */
  int32_t J;   /* joint dim = 2*L_DIM (2048), :177                                            */
  int32_t A;   /* number of answers                                                           */
  int32_t T;   /* padded question length of the batch (<= 14 in VQA v2)                       */
  int32_t W;   /* W_DIM word-embedding dim (300), :10                                         */
  int32_t Vq;  /* question vocabulary size                                                    */
  int32_t num_train_answer; /* answers [0, num_train_answer) are train answers, :38-42        */
  int32_t variant;          /* VQA_VARIANT_*                                                  */
  int32_t precision;        /* VQA_PREC_*                                                     */
  float keep_att;           /* attention-feature dropout keep prob (0.8), vlmap/modules.py:82 */
  float keep_joint;         /* joint dropout keep prob (0.5), vqa/model_vlmap_answer.py:180   */
  int32_t num_marginal;     /* NUM_MARGINAL of the ent variant (200, model_vlmap_answer_ent.py:16); 0 = 200 */
} VqaConfig;

/*
 * Parameters (and, with the same struct, their gradients). fp32, TF layout [in, out].
 * A NULL pointer in a gradient struct means "do not compute this gradient" (frozen variable).
 */
typedef struct VqaParams {
  float* embed;       /* LearnGloVe/embed_map                        [Vq, W]      */
  float* v_w;         /* v_linear_v/fc/weights                       [Dv, D]      */
  float* v_b;         /* v_linear_v/fc/biases                        [D]          */
  float* v_gamma;     /* v_linear_v/LayerNorm/gamma                  [D]          */
  float* v_beta;      /* v_linear_v/LayerNorm/beta                   [D]          */
  float* gru_gates_w; /* encode_L/rnn/gru_cell/gates/kernel          [W+L, 2L]    */
  float* gru_gates_b; /* encode_L/rnn/gru_cell/gates/bias            [2L]         */
  float* gru_cand_w;  /* encode_L/rnn/gru_cell/candidate/kernel      [W+L, L]     */
  float* gru_cand_b;  /* encode_L/rnn/gru_cell/candidate/bias        [L]          */
  float* qv_w;        /* q_linear_v/fc/weights                       [L, D]       */
  float* qv_b;        /* q_linear_v/fc/biases                        [D]          */
  float* qv_gamma;    /* q_linear_v/LayerNorm/gamma                  [D]          */
  float* qv_beta;     /* q_linear_v/LayerNorm/beta                   [D]          */
  float* att_w;       /* hadamard_attention/compute/score/fc/weights [D, 1]       */
  float* att_b;       /* hadamard_attention/compute/score/fc/biases  [1]          */
  float* pl_w;        /* (reasoning/)pooled_linear_l/fc/weights      [Dv, L]      */
  float* pl_b;        /* (reasoning/)pooled_linear_l/fc/biases       [L]          */
  float* pl_gamma;    /* (reasoning/)pooled_linear_l/LayerNorm/gamma [L]          */
  float* pl_beta;     /* (reasoning/)pooled_linear_l/LayerNorm/beta  [L]          */
  float* ql_w;        /* (reasoning/)q_linear_l/fc/weights           [L, L]       */
  float* ql_b;        /* (reasoning/)q_linear_l/fc/biases            [L]          */
  float* ql_gamma;    /* (reasoning/)q_linear_l/LayerNorm/gamma      [L]          */
  float* ql_beta;     /* (reasoning/)q_linear_l/LayerNorm/beta       [L]          */
  float* joint_w;     /* (reasoning/)joint_fc/fc/weights             [L, J]       */
  float* joint_b;     /* (reasoning/)joint_fc/fc/biases              [J]          */
  float* joint_gamma; /* (reasoning/)joint_fc/LayerNorm/gamma        [J]          */
  float* joint_beta;  /* (reasoning/)joint_fc/LayerNorm/beta         [J]          */
  float* ans_w;       /* WordWeightAnswer/fc/weights | reasoning/classifier/fc/weights [J, A] */
  float* ans_b;       /* WordWeightAnswer/fc/biases  | reasoning/classifier/fc/biases  [A]    */
  /* the extra question layer of the answer2 / no_noise variants (NULL otherwise); L must equal D there */
  float* qp_w;        /* q_L_ft2/fc/weights | q_L_mean/fc/weights     [L, L]       */
  float* qp_b;        /* q_L_ft2/fc/biases  | q_L_mean/fc/biases      [L]          */
  float* qp_gamma;    /* q_L_ft2/LayerNorm/gamma (answer2 only)       [L]          */
  float* qp_beta;     /* q_L_ft2/LayerNorm/beta  (answer2 only)       [L]          */
  /* second branch of the noc / nocarch variants (NULL otherwise); there joint_* is scope joint_v and ans_* is
   * WordWeightAnswerV */
  float* jl_w;        /* joint_l/fc/weights                           [L, J]       */
  float* jl_b;        /* joint_l/fc/biases                            [J]          */
  float* jl_gamma;    /* joint_l/LayerNorm/gamma                      [J]          */
  float* jl_beta;     /* joint_l/LayerNorm/beta                       [J]          */
  float* al_w;        /* WordWeightAnswerL/fc/weights                 [J, A]       */
  float* al_b;        /* WordWeightAnswerL/fc/biases                  [A]          */
  /* full: the log-variance layer next to q_L_mean (= qp_w / qp_b)                    */
  float* qs_w;        /* q_L_log_sigma_sq/fc/weights                  [L, L]       */
  float* qs_b;        /* q_L_log_sigma_sq/fc/biases                   [L]          */
  /* vqa_all / vqa_all2: the tuned head (trained)                                     */
  float* tw_w;        /* TunedWordWeightAnswer/fc/weights             [J, A]       */
  float* tw_b;        /* TunedWordWeightAnswer/fc/biases              [A]          */
  /* adapt: the pooled projection (trained); pl_w is [D, L] in that variant           */
  float* va_w;        /* v_adapt/fc/weights                           [Dv, D]      */
  float* va_b;        /* v_adapt/fc/biases                            [D]          */
  float* va_gamma;    /* v_adapt/LayerNorm/gamma                      [D]          */
  float* va_beta;     /* v_adapt/LayerNorm/beta                       [D]          */
} VqaParams;
#define VQA_NUM_PARAM_TENSORS 29

/* The feature bank the reference holds in host RAM (vqa/model_vlmap_answer.py:57-77); here in HBM. */
typedef struct VqaFeatureBank {
  const float* features;    /* image_features [N, K, Dv] fp32                                 */
  const int32_t* num_boxes; /* num_boxes      [N]                                             */
  int64_t num_images;       /* N                                                              */
  const void* features_bf16; /* optional one-off bf16 copy of image_features [N, K, Dv] (vqa_split_bf16): in bf16 mode
                              * the per-step gather then copies bf16 rows (half the bytes read); the rounding is the
                              * same fp32 -> bf16 conversion done once instead of every step, so results are identical.
                              * NULL = gather from the fp32 bank. Ignored in fp32 mode.                            */
} VqaFeatureBank;

/* One batch, keys as in vqa/datasets/input_ops_vqa_tf_record_memft.py:47-59 */
typedef struct VqaBatch {
  int32_t batch_size;          /* <= config.B (last eval batch may be smaller)                */
  int32_t q_len_max;           /* T of this batch (<= config.T); q_intseq is [batch, T]       */
  const int64_t* image_idx;    /* [batch] index into the feature bank                         */
  const int32_t* q_intseq;     /* [batch, q_len_max], pad id 0                                */
  const int32_t* q_intseq_len; /* [batch]                                                     */
  const float* answer_target;  /* [batch, A] soft scores                                      */
} VqaBatch;

/* answer masks ([A] fp32 each), vqa/model_vlmap_answer.py:37-52 */
typedef struct VqaAnswerMasks {
  const float* is_object;    /* answer_dict['is_object']                                      */
  const float* is_attribute; /* answer_dict['is_attribute']                                   */
  const float* answer_exist; /* modules.AnswerExistMask (vlmap/modules.py:575-586)            */
} VqaAnswerMasks;

/* report scalars, order = vqa/model_vlmap_answer.py:275-288 */
enum {
  VQA_REPORT_ANSWER_TRAIN_LOSS = 0,
  VQA_REPORT_ANSWER_REPORT_LOSS,
  VQA_REPORT_ANSWER_ACC,
  VQA_REPORT_EXIST_ACC,
  VQA_REPORT_TEST_ACC,
  VQA_REPORT_NORMAL_TEST_ACC,
  VQA_REPORT_NORMAL_TEST_OBJECT_ACC,
  VQA_REPORT_NORMAL_TEST_ATTRIBUTE_ACC,
  VQA_REPORT_NORMAL_EXIST_ACC,
  VQA_REPORT_NORMAL_TRAIN_EXIST_ACC,
  VQA_REPORT_MAX_EXIST_ACC,
  VQA_REPORT_TEST_MAX_ACC,
  VQA_REPORT_TEST_MAX_EXIST_ACC,
  /* vqa/model_vlmap_answer_full.py:221-223 (0 for every other variant) */
  VQA_REPORT_LATENT_LOSS,
  VQA_REPORT_TRAIN_LATENT_LOSS,
  /* vqa/model_vlmap_answer_ent.py:292-294 (0 for every other variant) */
  VQA_REPORT_ENTROPY,
  VQA_REPORT_WEIGHTED_ENTROPY,
  VQA_NUM_REPORT
};

/* per-sample outputs, order = model.output keys read at vqa/evaler.py:139-156 */
enum {
  VQA_PS_ALL_SCORE = 0,
  VQA_PS_MAX_TRAIN_SCORE,
  VQA_PS_TEST_OBJ_SCORE,
  VQA_PS_TEST_OBJ_MAX_SCORE,
  VQA_PS_TEST_ATTR_SCORE,
  VQA_PS_TEST_ATTR_MAX_SCORE,
  VQA_NUM_PER_SAMPLE
};

/* Outputs of a forward pass; any pointer may be NULL (that output is then not written). */
typedef struct VqaOutputs {
  float* loss;        /* [1]  model.loss (= answer_train_loss)                                */
  float* report;      /* [VQA_NUM_REPORT]                                                     */
  float* att_score;   /* [batch, K]   model.output['att_score']                               */
  float* logit;       /* [batch, A]   model.output['logit']                                   */
  int32_t* pred;      /* [batch]      model.output['pred'] (argmax, first index on ties)      */
  float* per_sample;  /* [VQA_NUM_PER_SAMPLE, batch]                                          */
  float* condition;   /* [batch, L]   model.heavy_output['condition'] (final GRU state)       */
  float* pooled;      /* [batch, Dv]  model.mid_result['pooled_V_ft'] ([batch, D] in the adapt variant) */
} VqaOutputs;

typedef struct VqaHandle_t* VqaHandle;

/* ---- lifecycle -------------------------------------------------------------------------------- */
/* Supported shapes: Dv, D, L, J, A multiples of 8; D, L, J <= 4096; K <= 256; T <= 64. */
VQA_API VqaStatus vqa_create(const VqaConfig* config, VqaHandle* out);
VQA_API VqaStatus vqa_destroy(VqaHandle h);
VQA_API const char* vqa_last_error(void);
VQA_API int32_t vqa_abi_version(void);
/* which recurrent (GRU) kernels the most recent forward used: bit 0 = the CTA-pair kernels (csrc/gru_pair.cu), bit 1 =
 * the single-CTA kernels (csrc/gru.cu: ~2x slower per phase; taken when the batch needs more row tiles than one wave of
 * CTA pairs or when the cluster + cooperative launch was refused), bit 8 = such a refusal happened in this process
 * (also reported once on stderr and in vqa_last_error()). Replaces nothing in the reference (tf.nn.dynamic_rnn,
 * vlmap/modules.py:131-135, has no kernel choice); it exists so that a silent 2x slowdown cannot hide. */
VQA_API int32_t vqa_gru_kernel_path(void);
/* number of kernels this library has enqueued so far in this process */
VQA_API uint64_t vqa_launch_count(void);

/* CRC-32C (Castagnoli) of a HOST buffer: the checksum of the TFRecord framing the reference's input pipeline reads
 * (vqa/datasets/input_ops_vqa_tf_record_memft.py:17-22); used by the Python mirror's record reader. No GPU involved. */
VQA_API uint32_t vqa_crc32c(const uint8_t* data_host, uint64_t n);

/* ---- input side in native code (vqa/datasets/input_ops_vqa_tf_record_memft.py:17-82; csrc/input_host.cu) -----------------
 * vqa_tfrecord_index_host: payload offsets / lengths of every record of one TFRecord file held in HOST memory (framing
 *   u64 length | u32 masked crc32c | payload | u32 masked crc32c), both checksums verified when verify_crc; *count is
 *   the number of records found (fill at most `capacity`; call again with larger arrays if *count > capacity).
 * vqa_parse_examples_host: parse_fn + padded_batch for n serialized tf.train.Example records in HOST memory: qid ->
 *   id (default -1), image_idx (default -1), q_intseq/list -> row i of q_intseq [n, t_cap] padded with 0, q_intseq/len
 *   (required), *t_longest = longest question of the batch; the soft-score target is emitted SPARSE: triples
 *   (ans_row, ans_id, ans_score), *ans_count of them (a repeated id keeps its last score, as target[ids] = scores
 *   does); image_id is returned as offset / length into each record (optional). Errors (malformed message, missing
 *   q_intseq/len, ids / scores of different length, answer id outside [0, num_answers), a question longer than
 *   t_cap) return VQA_ERR_BAD_SHAPE with the record number in vqa_last_error().
 * vqa_densify_targets: tf.sparse_to_dense on the DEVICE: target[batch, num_answers] = 0, then target[row, id] = score
 *   for the n triples (device pointers), on `stream`. A step then uploads the triples (KBs), not 6 MB of zeros. */
VQA_API VqaStatus vqa_tfrecord_index_host(const uint8_t* file_host, uint64_t size, int32_t verify_crc, uint64_t* offsets,
                                          uint64_t* lengths, int64_t capacity, int64_t* count);
VQA_API VqaStatus vqa_parse_examples_host(const uint8_t* const* records, const uint64_t* lengths, int32_t n,
                                          int32_t num_answers, int32_t t_cap, int64_t* id, int64_t* image_idx,
                                          int32_t* q_intseq, int32_t* q_intseq_len, int32_t* t_longest,
                                          int32_t* ans_row, int32_t* ans_id, float* ans_score, int32_t ans_cap,
                                          int32_t* ans_count, uint32_t* image_id_off, uint32_t* image_id_len);
VQA_API VqaStatus vqa_densify_targets(const int32_t* rows, const int32_t* ids, const float* scores, int32_t n, int32_t batch,
                                      int32_t num_answers, float* target, void* stream);

/* bytes of device workspace this handle needs; attach a buffer of at least that size (256-B aligned) */
VQA_API VqaStatus vqa_workspace_bytes(VqaHandle h, uint64_t* bytes);
VQA_API VqaStatus vqa_set_workspace(VqaHandle h, void* dev_ptr, uint64_t bytes);

/* ---- the path ---------------------------------------------------------------------------------- */
/* Refresh the GEMM-operand shadows of the weights (bf16 [hi, lo] planes). Call before the first forward and
 * after every parameter update; weight matrices whose pointer is NULL are left as they are (after an
 * optimizer step only the trainable ones changed). Replaces nothing in the reference (TF reads fp32 variables). */
VQA_API VqaStatus vqa_prepare_params(VqaHandle h, const VqaParams* params, void* stream);

/* Forward of Model.build() (vqa/model_vlmap_answer.py:102-288): gather -> v-proj -> GRU -> attention ->
 * pooling -> joint head -> logits -> soft-score BCE + report. Dropout always fires (tf.nn.dropout has no
 * train switch); masks come from Philox keyed by (seed, step). Keeps what backward needs. */
VQA_API VqaStatus vqa_forward(VqaHandle h, const VqaParams* params, const VqaFeatureBank* bank,
                              const VqaBatch* batch, const VqaAnswerMasks* masks, uint64_t seed,
                              uint64_t step, const VqaOutputs* out, void* stream);

/* Software pipelining across steps: register the NEXT batch (`batch->image_idx`, `batch->batch_size`; `stream` = the
 * stream its image_idx upload was enqueued on). The feature gather depends on no parameter, so the next vqa_backward
 * launches it into a second set of operand planes next to its weight-gradient GEMMs (an HBM-bound copy beside
 * tensor-bound kernels), and the following vqa_forward with the same image_idx pointer and batch size adopts the
 * planes instead of gathering. Optional: without it (or without a backward pass in between) vqa_forward gathers. */
VQA_API VqaStatus vqa_prefetch_features(VqaHandle h, const VqaFeatureBank* bank, const VqaBatch* batch, void* stream);

/* Backward of the same graph (tf.gradients inside optimize_loss, vqa/trainer.py:106-114) w.r.t. every
 * non-NULL field of `grads`, for the batch of the preceding vqa_forward on this handle.
 * loss_scale multiplies d(loss) (1/world_size under data parallelism, so that summed grads = mean). */
VQA_API VqaStatus vqa_backward(VqaHandle h, const VqaParams* params, const VqaBatch* batch,
                               const VqaParams* grads, float loss_scale, void* stream);

/* Materialise the dropout masks vqa_forward(seed, step) uses, as 0/1 bytes: att [batch, K, D],
 * joint [batch, J]. Test / parity helper (the kernels regenerate the same bits on the fly). */
/* One dropout site's keep mask (0 / 1 bytes) for (seed, step): site 1 = attention features [batch, K, D],
 * 2 = joint (joint_v in noc) [batch, J], 3 = joint_l of the noc variants [batch, J], 5 = the tiled joint of the ent
 * variant [batch, num_marginal, J]. */
VQA_API VqaStatus vqa_dropout_mask_site(VqaHandle h, int32_t site, int32_t batch, uint64_t seed, uint64_t step,
                                        uint8_t* mask, void* stream);

/* Deferred outputs (training loops): nothing downstream of the logits needs the loss / metrics kernels -- the backward
 * pass recomputes d(logits) from the logits -- so with this enabled vqa_forward enqueues them (and the copies into
 * VqaOutputs) on an auxiliary stream, and the following vqa_backward joins that stream before it returns: loss, report,
 * pred, ... are valid on `stream` AFTER vqa_backward instead of after vqa_forward. A forward that is not followed by a
 * backward is joined by the next vqa_forward, or explicitly by vqa_sync_outputs(h, stream). Off by default. */
VQA_API VqaStatus vqa_set_deferred_outputs(VqaHandle h, int32_t enable);
VQA_API VqaStatus vqa_sync_outputs(VqaHandle h, void* stream);

/* Index inputs out of range. The reference fails loudly: tf.nn.embedding_lookup raises on a token id outside [0, Vq)
 * (vqa/model_vlmap_answer.py:134) and np.take raises on an image_idx outside [-N, N) while WRAPPING negative ones
 * (parse_fn's default for a missing feature is -1 = the last image; vqa/model_vlmap_answer.py:110-117,
 * vqa/datasets/input_ops_vqa_tf_record_memft.py:28-46). Kernels cannot raise: the gathers wrap a negative image_idx
 * like np.take, replace anything still out of range by index 0 (no out-of-bounds read, no out-of-bounds atomicAdd in
 * the embedding scatter-add) and count it in a sticky per-process device counter. This call copies the counter to the
 * host (it SYNCHRONISES with the device) and optionally clears it; the Python host raises when it is non-zero. */
VQA_API VqaStatus vqa_input_error_count(uint32_t* count, int32_t reset);

/* Data-parallel overlap (vqa/trainer.py has no distributed code; this is the B200 side of SURVEY 8e).
 * With early gradients enabled, vqa_backward produces the gradients of everything EXCEPT the embedding and the GRU
 * (v_linear_v, q_linear_v, the attention score layer, and the trainable heads of model_standard) BEFORE the GRU's
 * back-propagation through time and records an event there; vqa_stream_wait_early_gradients makes `stream` wait for
 * that event of the most recent vqa_backward, so an all-reduce of that slice can run under the BPTT kernels. */
VQA_API VqaStatus vqa_set_early_gradients(VqaHandle h, int32_t enable);
/* The gradient exchange of the data-parallel step INSIDE vqa_backward. The caller puts the flat gradient buffer (the one
 * the VqaParams gradient struct points into) in symmetric memory that is also mapped as one NVSwitch multicast object
 * (torch.distributed._symmetric_memory does rendezvous and mapping), with 64 spare bytes at float offset `flags_offset`
 * (>= n_total) zeroed on every rank, and registers both addresses here. vqa_backward then
 *   - takes dWv before the BPTT (as with vqa_set_early_gradients) and all-reduces [0, n_early) -- everything but the
 *     GRU and the embedding -- on an auxiliary stream UNDER the cooperative recurrent kernel, as ten 2-CTA clusters on
 *     the TPCs that grid leaves idle (csrc/collective.cu: multimem.ld_reduce / multimem.st, sums formed in the switch);
 *   - all-reduces [n_early, n_total) after the weight-gradient section, on `stream`.
 * Both launches carry their own cross-rank entry / exit barriers (multimem.red on two counters at flags_offset), so the
 * host issues nothing between vqa_backward and the optimizer step, and the sums are valid on `stream` when vqa_backward's
 * work completes. Sizes in floats, multiples of 4. multicast_base = NULL unregisters. (The reference has no distributed
 * code at all: one process per GPU via CUDA_VISIBLE_DEVICES, run.py:25-46.) */
VQA_API VqaStatus vqa_set_gradient_allreduce(VqaHandle h, void* multicast_base, void* local_base, int64_t n_early,
                                             int64_t n_total, int64_t flags_offset, int32_t rank, int32_t world);
VQA_API VqaStatus vqa_stream_wait_early_gradients(VqaHandle h, void* stream);

/* In-switch all-reduce(sum) of n floats that every rank holds at the same offset of a symmetric allocation mapped
 * as one NVSwitch multicast object (`multicast_ptr` = the multicast address of element 0; torch's
 * _symmetric_memory.rendezvous provides it). Rank r reduces and re-broadcasts elements [r, r + 1) * n / world with
 * multimem.ld_reduce / multimem.st. The caller must put a cross-rank barrier before (all gradients written) and
 * after (all slices broadcast) on the same stream. num_ctas = 0 picks a default. */
VQA_API VqaStatus vqa_multimem_all_reduce(void* multicast_ptr, int64_t n, int32_t rank, int32_t world,
                                          int32_t num_ctas, void* stream);

/* The N(0, 1) draw of the full variant's reparameterisation for (seed, step): noise [batch, L] fp32 (Philox4x32-10
 * + Box-Muller; tf.random_normal(seed=123) of vqa/model_vlmap_answer_full.py:133 is not bit-reproducible outside TF,
 * so parity tests feed the oracle this very draw, as they do with the dropout masks). */
VQA_API VqaStatus vqa_reparam_noise(VqaHandle h, int32_t batch, uint64_t seed, uint64_t step, float* noise,
                                    void* stream);

VQA_API VqaStatus vqa_dropout_masks(VqaHandle h, int32_t batch, uint64_t seed, uint64_t step,
                                    uint8_t* att_mask, uint8_t* joint_mask, void* stream);

/* Device pointers of activations the last vqa_forward saved in the workspace (read-only; parity tests use
 * them to account for ReLU gates decided differently at working precision). `bytes` = extent for the last
 * batch. */
enum {
  VQA_ACT_HQ = 0,   /* relu(LN(q Wqv + b))            [batch, D]  fp32                            */
  VQA_ACT_HL,       /* relu(LN(q Wl + b))             [batch, L]  fp32                            */
  VQA_ACT_HP,       /* relu(LN(P Wp + b))             [batch, L]  fp32                            */
  VQA_ACT_JD,       /* dropout(relu(LN(X Wj + b)))    [batch, J]  bf16 (hi plane)                 */
  VQA_ACT_Z,        /* pre-LN v-projection            [batch*K, D] bf16 (PREC_BF16) / fp32        */
  VQA_ACT_JDL,      /* noc: dropout(relu(LN(Hl Wjl + b))) [batch, J]  bf16 (hi plane)                 */
  VQA_ACT_VA,       /* adapt: v_adapt = relu(LN(V Wa + b)) [batch*K, D] bf16 (hi plane)           */
  VQA_NUM_ACT
};
VQA_API VqaStatus vqa_peek_activation(VqaHandle h, int32_t which, const void** dev_ptr, uint64_t* bytes);

/* ---- optimizer step (vqa/trainer.py:87-114: clip_by_global_norm(20) + Adam) ---------------------- */
/* flat fp32 buffers of n elements (the trainable set laid out contiguously by the caller).
 * t = 1-based step count. grad_norm_out [1] receives the pre-clip global norm. */
VQA_API VqaStatus vqa_adam_step(VqaHandle h, float* param, const float* grad, float* m, float* v,
                                int64_t n, float lr, float beta1, float beta2, float eps,
                                float clip_norm, int64_t t, float* grad_norm_out, void* stream);

/* The same step with the parameter refresh folded in: `params` names the tensors of the model; every weight matrix of
 * it that lies inside [param, param + n) has its GEMM-operand shadow rewritten by the Adam pass itself (and the GRU
 * weights repacked), i.e. vqa_adam_step followed by vqa_prepare_params of the trainable matrices, in one pass over
 * the parameters instead of two. */
VQA_API VqaStatus vqa_adam_step_shadowed(VqaHandle h, const VqaParams* params, float* param, const float* grad,
                                         float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                                         float eps, float clip_norm, int64_t t, float* grad_norm_out, void* stream);

/* Pipelined optimizer tail: with tail_begin > 0, vqa_adam_step_shadowed updates parameters [tail_begin, n) of the flat
 * buffer (the embedding and the GRU: ParamStore lays them out last) on an auxiliary stream, so the next vqa_forward's
 * first kernels -- which read only the head of the buffer -- start while the tail is still being updated; that forward's
 * embedding / x-projection branch and its recurrent kernel wait for the tail. 0 switches it off (default). Any OTHER
 * reader of the parameters between the optimizer step and the next vqa_forward must first call vqa_sync_params(h, stream),
 * which makes `stream` wait for the tail (vqa/trainer.py:106-114 applies the update inside the same session.run). */
VQA_API VqaStatus vqa_set_optimizer_tail(VqaHandle h, int64_t tail_begin);
VQA_API VqaStatus vqa_sync_params(VqaHandle h, void* stream);

/* How the reference's clip_by_global_norm(20.0) sees the embedding gradient: tf.nn.embedding_lookup on a variable yields
 * an IndexedSlices (one row per token occurrence), and clip_ops.global_norm takes its `.values` as they are, so
 * LearnGloVe/embed_map contributes sum_{b,t} |dE[b,t]|^2 -- not the norm of the scattered dense gradient (they differ
 * whenever a token occurs more than once in the batch). With a slot set (one device float, e.g. just past the flat
 * gradient buffer so that a data-parallel all-reduce sums it with the gradients), vqa_backward writes that sum into it
 * and vqa_adam_step_shadowed uses it in place of the dense tensor's share of the norm. The Adam update itself is the
 * dense one in both (the reference sums duplicate indices before applying and decays m, v of every row). NULL = off:
 * plain dense norm. */
VQA_API VqaStatus vqa_set_embedding_slice_norm(VqaHandle h, float* slot);

/* ---- per-phase device timing (bench.py roofline): CUDA events recorded on the caller's stream around
 * the phases of vqa_forward / vqa_backward while enabled. Not for use under CUDA-graph capture. ------- */
enum {
  VQA_PH_GATHER = 0,   /* feature gather -> operand planes                                  */
  VQA_PH_VPROJ_FWD,    /* Z = V Wv + bv  (tcgen05 GEMM, [B*K, Dv] x [Dv, D])                 */
  VQA_PH_GRU_FWD,      /* embedding gather + hoisted input GEMMs + T recurrent steps        */
  VQA_PH_QHEADS_FWD,   /* q_linear_v, q_linear_l                                            */
  VQA_PH_ATTN_FWD,     /* attention block forward kernel                                    */
  VQA_PH_HEAD_FWD,     /* pooled_linear_l, joint_fc, answer logits                          */
  VQA_PH_LOSS,         /* BCE + metrics                                                     */
  VQA_PH_HEAD_BWD,     /* d logits ... dP, dq (and the head's wgrads when trainable)        */
  VQA_PH_ATTN_BWD,     /* attention block backward kernel (+ partial reduce)                */
  VQA_PH_QV_BWD,       /* q_linear_v backward                                               */
  VQA_PH_VPROJ_WGRAD,  /* dWv = V^T dZ   (tcgen05 GEMM, [Dv, B*K] x [B*K, D])                */
  VQA_PH_GRU_BWD,      /* BPTT recurrent part                                               */
  VQA_PH_GRU_WGRAD,    /* GRU weight / bias gradients                                       */
  VQA_PH_EMBED_BWD,    /* dE GEMMs + scatter-add                                            */
  VQA_NUM_PHASES
};
/* enable: 0 = off; 1 = per-phase events with every branch SERIALISED on the caller's stream (isolated phase times);
 * 2 = events with the auxiliary-stream forks kept (sections of the real critical path; the concurrent weight-gradient
 *     section is reported as GRU_WGRAD, VPROJ_WGRAD / EMBED_BWD read 0) */
VQA_API VqaStatus vqa_profile_enable(VqaHandle h, int32_t enable);
/* milliseconds per phase of the most recent forward+backward (synchronises on the events) */
VQA_API VqaStatus vqa_profile_read(VqaHandle h, float* ms /* [VQA_NUM_PHASES] */);
VQA_API const char* vqa_phase_name(int32_t phase);

/* ---- per-kernel entry points (unit parity tests, ncu) --------------------------------------------- */
typedef struct VqaGemmDesc {
  /* D[M,N] = A[M,K] * B[N,K]^T (+bias[N]) (+addend[M,N]); operands bf16 planes (lo may be NULL).
   * a_mn_major = 0: A stored [M, K] (K contiguous), pitch lda elements
   * a_mn_major = 1: A stored [K, M] (M contiguous), pitch lda
   * b_mn_major = 0: B stored [N, K] (K contiguous), pitch ldb
   * b_mn_major = 1: B stored [K, N] (N contiguous), pitch ldb                               */
  const void* a_hi; const void* a_lo;
  const void* b_hi; const void* b_lo;
  int64_t lda, ldb;
  int32_t a_mn_major, b_mn_major;
  int32_t M, N, K;
  const float* bias;
  const float* addend; int64_t ld_addend;
  float* out_f32; int64_t ld_f32;           /* optional fp32 output                           */
  void* out_hi; void* out_lo; int64_t ld_bf; /* optional bf16 output planes (lo = residual)    */
  int32_t block_n;                          /* 0 = auto; 64 / 128 / 256 = single-CTA tile width;
                                             * -128 / -256 = force the CTA-pair kernel (256 x |block_n| tiles, bf16 only) */
} VqaGemmDesc;
VQA_API VqaStatus vqa_gemm(VqaHandle h, const VqaGemmDesc* d, void* stream);

/* fp32 [rows, cols] (pitch ld) -> bf16 hi (and lo if non-NULL) planes with pitch ld_out */
VQA_API VqaStatus vqa_split_bf16(VqaHandle h, const float* src, int64_t rows, int64_t cols, int64_t ld,
                                 void* hi, void* lo, int64_t ld_out, void* stream);

/* The attention-feature dropout (tf.nn.dropout(feature, 0.8), vlmap/modules.py:82) of (seed, step) as a bit plane: one
 * byte per group of 8 consecutive elements of the [batch, K, D] tensor, bit j = keep flag of element j -- the very bits
 * vqa_dropout_masks materialises one byte per element. vqa_forward fills the workspace's plane once per step on an
 * auxiliary stream (under the v-projection GEMM) and both attention kernels read it: 1 byte per 8 elements instead of
 * ten Philox rounds per 8 elements in the forward AND in the backward kernel. n = batch*K*D must be a multiple of 8. */
VQA_API VqaStatus vqa_keep_bits(VqaHandle h, int32_t batch, uint64_t seed, uint64_t step, uint8_t* bits, void* stream);

/* attention block forward: per-sample LayerNorm(K*D)+ReLU of z, Hadamard with hq, dropout, score,
 * masked softmax, attended pooling of the features (vlmap/modules.py:67-97, 23-39) */
typedef struct VqaAttnFwd {
  int32_t batch;
  const void* z;            /* [batch*K, D] pre-LN projection; bf16 (PREC_BF16) or fp32 (PREC_FP32) */
  const float* gamma; const float* beta; /* [D]                                               */
  const float* hq;          /* [batch, D]                                                     */
  const float* att_w; const float* att_b;
  const int32_t* nbox;      /* [batch]                                                        */
  const void* v_hi; const void* v_lo; /* gathered features [batch*K, Dv] bf16 planes          */
  uint64_t seed, step;
  float* att;               /* [batch, K]                                                     */
  float* pooled;            /* [batch, Dv]                                                    */
  void* pooled_hi; void* pooled_lo; /* bf16 planes of pooled for the next GEMM (may be NULL)   */
  float* ln_mean; float* ln_rstd;   /* [batch] saved statistics                               */
  const uint8_t* keep_bits; /* [batch*K*D/8] dropout keep bits of (seed, step), one byte per 8 consecutive elements (bit j =
                             * element j; vqa_keep_bits fills it), or NULL: the kernel then draws the same bits itself */
} VqaAttnFwd;
VQA_API VqaStatus vqa_attn_fwd(VqaHandle h, const VqaAttnFwd* a, void* stream);

typedef struct VqaAttnBwd {
  int32_t batch;
  const void* z; const float* gamma; const float* beta; const float* hq;
  const float* att_w; const int32_t* nbox; const void* v_hi; const void* v_lo;
  uint64_t seed, step;
  const float* att; const float* ln_mean; const float* ln_rstd;
  const float* d_pooled;    /* [batch, Dv]                                                    */
  void* dz_hi; void* dz_lo; /* [batch*K, D] bf16 planes of d(pre-LN projection)               */
  float* d_hq;              /* [batch, D]                                                     */
  float* d_att_w; float* d_att_b; float* d_gamma; float* d_beta; float* d_bias; /* [D],[1],[D],[D],[D] */
  const uint8_t* keep_bits; /* as in VqaAttnFwd (the SAME plane: the backward pass regenerates nothing)  */
} VqaAttnBwd;
VQA_API VqaStatus vqa_attn_bwd(VqaHandle h, const VqaAttnBwd* a, void* stream);

/* soft-score BCE + argmax + report (vqa/model_vlmap_answer.py:192-288); d_logit may be NULL */
VQA_API VqaStatus vqa_bce_metrics(VqaHandle h, int32_t batch, const float* logit, const float* target,
                                  const VqaAnswerMasks* masks, float grad_scale, float* loss,
                                  float* report, int32_t* pred, float* per_sample, float* d_logit,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQA_ANSWER_H_ */
