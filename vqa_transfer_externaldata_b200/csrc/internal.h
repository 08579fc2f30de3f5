// Internal (non-ABI) declarations shared by the translation units of libvqa_answer_b200.so
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

struct CUtensorMap_st;

#include "../../include/vqa_answer.h"

namespace vqa {

typedef __nv_bfloat16 bf16;

// ---- error plumbing (thread-local message, returned by vqa_last_error) ----
VqaStatus set_error(VqaStatus code, const char* fmt, ...);
VqaStatus set_cuda_error(cudaError_t e, const char* what);
void count_launch();            // one of OUR kernels was enqueued
unsigned long long launch_count();

#define VQA_CUDA_CHECK(expr)                                         \
  do {                                                               \
    cudaError_t _e = (expr);                                         \
    if (_e != cudaSuccess) return ::vqa::set_cuda_error(_e, #expr);  \
  } while (0)

#define VQA_TRY(expr)                     \
  do {                                    \
    VqaStatus _s = (expr);                \
    if (_s != VQA_OK) return _s;          \
  } while (0)

#define VQA_LAUNCH_CHECK(what)                                       \
  do {                                                               \
    cudaError_t _e = cudaGetLastError();                             \
    if (_e != cudaSuccess) return ::vqa::set_cuda_error(_e, what);   \
    ::vqa::count_launch();                                           \
  } while (0)

// A GEMM operand / activation in "operand form": one bf16 plane (PREC_BF16) or hi+lo (PREC_FP32)
struct Planes {
  bf16* hi = nullptr;
  bf16* lo = nullptr;
};

// ---- gemm.cu / gemm_pair.cu ----
// split-K hand-over semaphores of the pair kernel: a ring of regions inside the workspace (zeroed when the
// workspace is attached; every launch leaves its region zero again), one region per in-flight launch
struct GemmCtx {
  unsigned int* sem = nullptr;
  int regions = 0;
  int region_elems = 0;
  int next_region = 0;
};
// narrow = 1: one pair per output tile and no split-K, for GEMMs that run NEXT TO others on forked streams (the
// GRU weight gradients): together they fill the SMs, and no reduction chain is needed
// narrow: bit 0 = one CTA pair per output tile, no split-K; bit 1 = the B operand is stable (weights, not written by the
// in-stream predecessor): the single-CTA kernels may fetch its first stages before the programmatic dependency resolves
VqaStatus gemm_launch(const VqaGemmDesc& d, int num_sms, cudaStream_t stream, GemmCtx* ctx = nullptr, int narrow = 0);
bool gemm_pair_plan(const VqaGemmDesc& d, int num_sms, const GemmCtx* ctx, int narrow, int* bn_out, int* splits_out);
VqaStatus gemm_pair_launch(const VqaGemmDesc& d, int num_sms, int bn, int splits, GemmCtx* ctx, cudaStream_t stream);
// cached cuTensorMapEncodeTiled: 2-D bf16, 128-byte swizzle, inner extent `inner` (contiguous), row pitch in elements
bool cached_tmap(CUtensorMap_st* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch,
                 uint32_t box_inner, uint32_t box_outer);
// kind: 0 = bf16 / 128-byte swizzle (operands), 1 = fp32 / 128-byte swizzle, 2 = bf16 / no swizzle (TMA stores)
// K-major bf16 [rows, K] as {64 k, rows, k-blocks}: box {64, box_rows, box_kb} = box_kb swizzled k-block tiles per TMA
bool cached_tmap_kblocks(CUtensorMap_st* out, const void* base, uint64_t K, uint64_t rows, uint64_t pitch,
                         uint32_t box_rows, uint32_t box_kb);
// MN-major bf16 [K, MN] as {64 mn, 64 k, mn-blocks, k-blocks}: box {64, 64, box_mnb, box_kb} (K % 64 == 0)
bool cached_tmap_mnblocks(CUtensorMap_st* out, const void* base, uint64_t MN, uint64_t K, uint64_t pitch,
                          uint32_t box_mnb, uint32_t box_kb);
bool cached_tmap_kind(CUtensorMap_st* out, const void* base, int kind, uint64_t inner, uint64_t outer,
                      uint64_t pitch, uint32_t box_inner, uint32_t box_outer);

// ---- elementwise.cu ----
VqaStatus split_bf16_launch(const float* src, long long rows, long long cols, long long ld, bf16* hi,
                            bf16* lo, long long ld_out, cudaStream_t s);
VqaStatus gather_features_launch(const float* bank, const int* num_boxes, const long long* image_idx,
                                 int batch, int K, int Dv, bf16* v_hi, bf16* v_lo, int* nbox, long long num_images,
                                 cudaStream_t s, int max_ctas = 0);
VqaStatus gather_features_bf16_launch(const bf16* bank, const int* num_boxes, const long long* image_idx, int batch,
                                      int K, int Dv, bf16* v_hi, int* nbox, long long num_images, cudaStream_t s,
                                      int max_ctas = 0);
VqaStatus embed_gather_launch(const float* embed, const int* q_intseq, int batch, int T, int Tstride,
                              int W, int Wpad, int vocab, bf16* e_hi, bf16* e_lo, cudaStream_t s);
VqaStatus embed_scatter_add_launch(const float* dE, long long ld_dE, const int* q_intseq,
                                   const int* q_len, int batch, int T, int Tstride, int W, int vocab,
                                   float* d_embed, cudaStream_t s);
VqaStatus input_error_count(unsigned int* out, bool reset);
// ---- collective.cu ----
VqaStatus multimem_allreduce_sync_launch(float* mc, long long n, int rank, int world, unsigned int* mc_flags,
                                         const unsigned int* my_flags, unsigned int* grid_ctr, unsigned int flag_total,
                                         unsigned int* grid_total_io, bool exclusive, int ctas, cudaStream_t s);   // sticky count of out-of-range image_idx / token ids
VqaStatus colsum_launch(const float* x, long long rows, long long cols, long long ld, float* out,
                        float* scratch, cudaStream_t s);
VqaStatus fill_zero_launch(void* p, size_t bytes, cudaStream_t s);

// GRU cell pieces (vlmap/modules.py:124-140; TF-1.6 GRUCell: gate order (r,u), c = tanh([x, r*h] Wc))
VqaStatus gru_gates_launch(const float* G, const float* h, int batch, int L, float* r, float* u,
                           bf16* rh_hi, bf16* rh_lo, cudaStream_t s);
VqaStatus gru_update_launch(const float* C, const float* h, const float* u, const int* q_len, int t,
                            int batch, int L, float* c_out, float* h_next, bf16* hn_hi, bf16* hn_lo,
                            cudaStream_t s);
VqaStatus gru_bwd_update_launch(const float* dh, const float* h, const float* u, const float* c,
                                const int* q_len, int t, int batch, int L, float* du, float* dh_part,
                                float* dC_f32, bf16* dC_hi, bf16* dC_lo, cudaStream_t s);
VqaStatus gru_bwd_gates_launch(const float* dRH, const float* du, const float* h, const float* r,
                               const float* u, const int* q_len, int t, int batch, int L,
                               float* dh_part, float* dG_f32, bf16* dG_hi, bf16* dG_lo,
                               cudaStream_t s);

// ---- gru.cu: persistent recurrent kernels (one cooperative launch per <= num_sms/(L/32) row tiles) ----
struct GruFwdPersistent {
  int B, L, T;
  const int* q_len;
  unsigned int* counter;
  const float* xg; const float* xc;      // hoisted x-parts (+bias): [T*B, 2L], [T*B, L]
  int x_bf16;                            // != 0: xg / xc hold bf16 values (the x-projection GEMMs stored bf16: half the bytes)
  float* h_f32; bf16* h_bf;              // [(T+1)*B, L], block 0 = zero initial state
  bf16* rh_bf;                           // [T*B, L]
  float* r; float* u; float* c;          // [T*B, L]
  const bf16* w_pack;                    // [L/32][96][L] packed K-major weight slices (gru_pack_weights_launch)
};
struct GruBwdPersistent {
  int B, L, T;
  const int* q_len;
  unsigned int* counter;
  const float* h_f32; const float* r; const float* u; const float* c;
  const float* dq;                       // [B, L] gradient of the final state
  const float* dq2;                      // optional second addend of that gradient (the two heads reading q)
  bf16* dG_bf;                           // [T*B, 2L]
  bf16* dC_bf;                           // [T*B, L]
  float* bias_part;                      // [2 * ceil(B/128), 3L] partial bias gradients (gates r | gates u | candidate)
  const bf16* wg_h; const bf16* wc_h;    // bf16 shadows of the h-rows of the TF kernels: [L, 2L], [L, L]
};
bool gru_persistent_supported(int B, int L, int precision, int num_sms);
size_t gru_pack_elems(int L);
size_t gru_bias_part_floats(int B, int L);
int gru_bias_part_rows(int B);   // partial rows of [3L] the BPTT kernels write (two per 128-row tile)
extern unsigned long long* g_gru_trace;
bool gru_pair_supported(int B, int L, int num_sms);
cudaError_t gru_pair_fwd(const GruFwdPersistent& a, int num_sms, cudaStream_t s);
cudaError_t gru_pair_bwd(const GruBwdPersistent& a, int num_sms, cudaStream_t s);
VqaStatus gru_pack_weights_launch(const bf16* wg_h, const bf16* wc_h, int L, bf16* out, cudaStream_t s);
VqaStatus gru_fwd_persistent_launch(const GruFwdPersistent& a, int num_sms, cudaStream_t s);
VqaStatus gru_bwd_persistent_launch(const GruBwdPersistent& a, int num_sms, cudaStream_t s);

// ---- rows.cu: row LayerNorm + ReLU heads (modules.fc_layer with use_ln, vlmap/modules.py:630-650) ----
struct RowLnFwd {
  int rows, N;
  const float* z;        // [rows, N] pre-LN (bias already added)
  const float* gamma;    // [N]
  const float* beta;     // [N]
  const float* mul;      // optional [rows, N]: output *= mul (the q (.) v Hadamard fusion)
  float keep;            // dropout keep prob on the output (1 = none)
  unsigned long long seed, step;
  unsigned int stream_id;
  float* y;              // optional fp32 [rows, N]: relu(LN(z)) BEFORE mul / dropout
  float* out_f32;        // optional fp32 final output
  bf16* out_hi;          // optional operand planes of the final output
  bf16* out_lo;
  float* mean;           // [rows]
  float* rstd;           // [rows]
  int act;               // 0 = ReLU (fc_layer's default here), 1 = tanh (q_L_ft2 of model_vlmap_answer2)
};
VqaStatus row_ln_relu_fwd_launch(const RowLnFwd& a, cudaStream_t s);

struct RowLnBwd {
  int rows, N;
  const float* dout;     // [rows, N] gradient w.r.t. the final output
  const float* z;        // pre-LN
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* rstd;
  const float* mul;      // optional: forward multiplied the activation by this
  float keep;
  unsigned long long seed, step;
  unsigned int stream_id;
  float* dz_f32;         // optional
  bf16* dz_hi;
  bf16* dz_lo;
  float* dgamma_part;    // optional [parts, N] per-CTA partials (parts = rows)
  float* dbeta_part;
  int act;               // as in RowLnFwd
};
VqaStatus row_ln_relu_bwd_launch(const RowLnBwd& a, cudaStream_t s);

// ---- linear_ln.cu ----
// One fc_layer (rank-2 input) as one kernel: tcgen05 product + bias + row LayerNorm + activation + Hadamard partner +
// dropout in the epilogue (forward), or the data-gradient product + dropout / Hadamard / activation / LayerNorm backward
// (backward); the CTAs of a row tile form a cluster along N and exchange row statistics through distributed shared
// memory. bf16 mode (one operand plane). *launched = false (and VQA_OK): shape / device not eligible -- the caller
// runs the GEMM + rows.cu kernels instead. VQA_LINEAR_LN=0 turns it off.
struct LinearLn {
  int M, N, K;
  const bf16* a; long long lda;      // [M, K], K contiguous
  const bf16* b; long long ldb;      // forward: weights [K, N] (b_mn_major = 1); backward: weights [N, K] (K contiguous)
  int b_mn_major;
  int backward;
  const float* bias;                 // forward
  const float* gamma; const float* beta;
  const float* mul;                  // optional [M, N]
  int act;                           // 0 relu, 1 tanh
  float keep;
  unsigned long long seed, step;
  unsigned int stream_id;
  float* z;                          // [M, N] pre-LN: forward output, backward input
  float* mean; float* rstd;          // [M]: forward output, backward input
  float* y; float* out_f32; bf16* out_hi;            // forward outputs (as RowLnFwd)
  float* raw;                        // backward, optional: the product itself (d loss / d layer output), fp32
  float* dz_f32; bf16* dz_hi; float* dgamma_part; float* dbeta_part;   // backward outputs (as RowLnBwd)
};
VqaStatus linear_ln_launch(const LinearLn& d, cudaStream_t s, bool* launched);
bool linear_ln_enabled();

// ---- attn.cu ----
VqaStatus attn_fwd_launch(const VqaAttnFwd& a, int K, int D, int Dv, int precision, float keep,
                          cudaStream_t s);
// reduce_stream: where the reduction of the per-sample partials (d att_w / gamma / beta / bias / att_b) runs; the
// caller has made it wait for nothing -- attn_bwd_launch orders it after the kernel itself (nullptr = s)
// qv (optional): the attention kernel also applies q_linear_v's ReLU / LayerNorm backward to each sample's d_hq row
// (what row_ln_relu_bwd would do next): z / gamma / mean / rstd of that layer's forward pass in, dz (+ the per-row
// d gamma / d beta terms) out
struct AttnQvBwd {
  const float* z; const float* gamma; const float* mean; const float* rstd;
  bf16* dz_hi; bf16* dz_lo; float* dz_f32; float* dgamma_part; float* dbeta_part;
};
VqaStatus attn_bwd_launch(const VqaAttnBwd& a, int K, int D, int Dv, int precision, float keep,
                          float* partials, cudaStream_t s, cudaStream_t reduce_stream = nullptr,
                          cudaEvent_t kernel_done = nullptr, const AttnQvBwd* qv = nullptr);
size_t attn_bwd_partial_floats(int batch, int D);
// attn_pipe.cu: persistent, software-pipelined forward (bf16 mode, buffers must fit one SM)
bool attn_fwd_pipe_supported(int K, int D, int Dv, int precision, bool has_v_lo, bool mask, size_t* smem_out, int* rv_out);
VqaStatus attn_fwd_pipe_launch(const VqaAttnFwd& a, int K, int D, int Dv, float keep, size_t smem, int rv,
                               int num_sms, cudaStream_t s);

// ---- loss.cu ----
VqaStatus bce_metrics_launch(int batch, int A, int num_train_answer, int use_train_mask,
                             const float* logit, const float* target, const VqaAnswerMasks& masks,
                             float grad_scale, float* loss, float* report, int* pred,
                             float* per_sample, float* d_logit_f32, bf16* d_hi, bf16* d_lo,
                             float* scratch, cudaStream_t s);
// two-term form (vqa_all / vqa_all2): loss_b = logits of a second BCE term (mask_b: train-masked or not),
// pred_logit = what the argmax runs on (NULL = logit)
VqaStatus bce_metrics2_launch(int batch, int A, int num_train_answer, int use_train_mask, const float* logit,
                              const float* loss_b, int mask_b, const float* pred_logit, const float* target,
                              const VqaAnswerMasks& masks, float* loss, float* report, int* pred, float* per_sample,
                              float* scratch, cudaStream_t s);
VqaStatus bce_grad_launch(int batch, int A, int num_train_answer, int use_train_mask,
                          const float* logit, const float* target, float grad_scale,
                          float* d_logit_f32, bf16* d_hi, bf16* d_lo, cudaStream_t s);
VqaStatus dropout_mask_launch(unsigned char* out, long long n, float keep, unsigned long long seed,
                              unsigned long long step, unsigned int stream_id, cudaStream_t s);
// one byte per group of 8 elements (bit j = keep of element j): the plane the attention kernels read
VqaStatus keep_bits_launch(unsigned char* out, long long n, float keep, unsigned long long seed,
                           unsigned long long step, unsigned int stream_id, cudaStream_t s, int max_ctas = 0);

// ---- variants.cu: kernels of the later family members (vqa_all / vqa_all2, full, adapt) ----
struct TunedHeadFwd {
  int batch, A, num_train_answer;
  int fill_min;            // 1 = vqa_all (absent answers take the row minimum), 0 = vqa_all2
  const float* logit0;     // [batch, A] word-weight logits
  const float* tuned;      // [batch, A] TunedWordWeightAnswer logits
  const float* exist;      // [A]
  float* l1;               // [batch, A] word-weight logits after the fill
  float* total;            // [batch, A] l1 + tuned  (model.output['logit'])
  float* pred_logit;       // optional [batch, A]: l1 * test_mask + tuned * train_mask (vqa_all2's argmax input)
};
VqaStatus tuned_combine_launch(const TunedHeadFwd& a, cudaStream_t s);
struct TunedHeadBwd {
  int batch, A, num_train_answer, fill_min;
  const float* logit0; const float* l1; const float* total; const float* tuned;
  const float* target; const float* exist;
  float grad_scale;        // loss_scale / batch
  float* d_logit0_f32; bf16* d_logit0_hi; bf16* d_logit0_lo;
  float* d_tuned_f32; bf16* d_tuned_hi; bf16* d_tuned_lo;
};
VqaStatus tuned_grad_launch(const TunedHeadBwd& a, cudaStream_t s);

struct ReparamFwd {
  int batch, L;
  const float* mean; const float* lss;   // [batch, L] q_L_mean, q_L_log_sigma_sq
  unsigned long long seed, step;
  float* out_f32; bf16* out_hi; bf16* out_lo;   // mean + noise * sqrt(exp(lss))
  float* kl_rows;                         // [batch] sum_l (1 + lss - mean^2 - exp(lss))
};
VqaStatus reparam_fwd_launch(const ReparamFwd& a, cudaStream_t s);
// loss[0] += weight * (scale * mean_b rows); report[slot] = scale * mean_b rows, report[slot + 1] = weight * that
VqaStatus latent_finalize_launch(const float* rows, int batch, float scale, float weight, int slot, float* loss,
                                 float* report, cudaStream_t s);

// ent variant (vqa/model_vlmap_answer_ent.py:193-213, 284-294)
struct EntTile {
  int batch, M, L;
  const float* hp; const float* hl;   // [batch, L] pooled_linear_l (no gradient through the tile), q_linear_l
  bf16* out_hi; bf16* out_lo;         // [batch * M, L]
};
VqaStatus ent_tile_launch(const EntTile& a, cudaStream_t s);
struct EntMarginal {
  int batch, M, A, num_train_answer;
  const float* logit2;                // [batch * M, A]
  const float* exist;                 // [A]
  float* marg;                        // [batch, A] marginal probabilities (0 at unselected answers)
  float* row_max; float* row_inv;     // [batch * M] softmax statistics kept for the backward
  float* ent_rows;                    // [batch] sum_a marg log(marg + 1e-8)
};
VqaStatus ent_marginal_launch(const EntMarginal& a, cudaStream_t s);
struct EntMarginalBwd {
  int batch, M, A, num_train_answer;
  const float* logit2; const float* exist; const float* marg; const float* row_max; const float* row_inv;
  float scale;                        // W_ENTROPY * loss_scale / batch
  bf16* d_hi; bf16* d_lo;             // [batch * M, A] d logit2 as GEMM operand planes
};
VqaStatus ent_marginal_bwd_launch(const EntMarginalBwd& a, cudaStream_t s);
struct EntDhl {
  int batch, M, L;
  const float* dX; const float* dX2; const float* hp;
  float* out;                         // [batch, L] gradient w.r.t. q_linear_l's output
};
VqaStatus ent_dhl_launch(const EntDhl& a, cudaStream_t s);
struct ReparamBwd {
  int batch, L;
  const float* d_out; const float* mean; const float* lss;
  unsigned long long seed, step;
  float kl_scale;                         // latent weight * loss_scale / batch
  float* d_mean_f32; bf16* d_mean_hi; bf16* d_mean_lo;
  float* d_lss_f32; bf16* d_lss_hi; bf16* d_lss_lo;
};
VqaStatus reparam_bwd_launch(const ReparamBwd& a, cudaStream_t s);
VqaStatus reparam_noise_launch(float* out, long long n, unsigned long long seed, unsigned long long step,
                               cudaStream_t s);

struct SlabLnFwd {
  int batch, K, D;
  const void* z;                          // [batch, K, D] pre-LN: bf16 (PREC_BF16) / fp32 (PREC_FP32)
  const float* gamma; const float* beta;  // [D]
  bf16* out_hi; bf16* out_lo;             // relu(LN_{K,D}(z)) as operand planes
  float* mean; float* rstd;               // [batch]
  // optional dropout on the output (set thr = 65536 for none): the tiled joint of the ent variant
  unsigned int thr; float inv_keep; unsigned long long seed, step; unsigned int site;
};
VqaStatus slab_ln_relu_fwd_launch(const SlabLnFwd& a, int precision, cudaStream_t s);
struct SlabLnBwd {
  int batch, K, D;
  const void* z; const float* gamma; const float* beta; const float* mean; const float* rstd;
  const float* att;                       // [batch, K]
  const float* d_pooled;                  // [batch, D]
  bf16* dz_hi; bf16* dz_lo;               // [batch, K, D]
  float* part;                            // [batch, 3, D] per-sample partials: d gamma | d beta | d bias (NULL: none)
  // ent: the upstream gradient is a tensor [batch, K, D] that passed through dropout (att / d_pooled unused then)
  const float* dout;
  unsigned int thr; float inv_keep; unsigned long long seed, step; unsigned int site;
};
VqaStatus slab_ln_relu_bwd_launch(const SlabLnBwd& a, int precision, cudaStream_t s);

// ---- optim.cu ----
// weight matrices inside the flat parameter buffer whose GEMM-operand shadows the Adam pass rewrites as it goes
struct AdamShadows {
  int n;
  long long begin4[8], end4[8];   // float4 index range of the tensor inside the flat buffer
  bf16* hi[8];
  bf16* lo[8];
};
VqaStatus adam_step_launch(float* param, const float* grad, float* m, float* v, long long n, float lr,
                           float beta1, float beta2, float eps, float clip_norm, long long t,
                           float* grad_norm_out, float* scratch, int num_sms, cudaStream_t s,
                           const AdamShadows* shadows = nullptr, const float* slice_grad = nullptr, long long slice_n = 0,
                           const float* slice_sumsq = nullptr, long long tail_begin = 0, cudaStream_t tail_stream = nullptr,
                           cudaEvent_t fork_ev = nullptr);
// out[0] = sum of squares of x[rows, cols] (pitch ld); scratch: >= 148 floats
VqaStatus rows_sumsq_launch(const float* x, long long rows, int cols, long long ld, float* out, float* scratch,
                            cudaStream_t s);

// weight of the KL latent loss of the full variant (vqa/model_vlmap_answer_full.py:33)
#define VQA_LATENT_LOSS_WEIGHT 0.1f

// RNG stream ids (which dropout site a Philox draw belongs to)
enum { RNG_STREAM_ATT = 1, RNG_STREAM_JOINT = 2, RNG_STREAM_JOINT_L = 3, RNG_STREAM_NOISE = 4, RNG_STREAM_ENT = 5 };
#define VQA_W_ENTROPY 0.1f        /* vqa/model_vlmap_answer_ent.py:14 */

}  // namespace vqa
