// The per-GPU handle: configuration + the carve-up of the caller-provided workspace.
#pragma once
#include <cstddef>
#include <cstdint>

#include "internal.h"

namespace vqa {

// Weight shadows in GEMM-operand form, same [in, out] layout as the fp32 TF variables.
struct WeightShadows {
  Planes v_w, gru_gates_w, gru_cand_w, qv_w, pl_w, ql_w, joint_w, ans_w, qp_w, jl_w, al_w, qs_w, tw_w, va_w;
};

// Everything forward keeps for backward + scratch, all inside the workspace.
struct Buffers {
  WeightShadows w;
  // inputs in operand form
  Planes v;        // [B*K, Dv] gathered features
  int* nbox;       // [B]
  Planes v_alt;    // second set: vqa_prefetch_features gathers the NEXT batch here while this step's backward runs
  int* nbox_alt;
  Planes e;        // [T*B, Wpad] embedded question tokens, time-major
  // v-projection
  void* z;         // [B*K, D] pre-LN projection: bf16 (PREC_BF16) / fp32 (PREC_FP32)
  Planes z_planes; // PREC_FP32 only: unused (z is consumed by the attention kernels, not by GEMMs)
  float* lnv_mean; float* lnv_rstd;  // [B]
  // GRU
  float* xg;       // [T*B, 2L]  x-part of the gate pre-activations (+ bias)
  float* xc;       // [T*B, L]   x-part of the candidate pre-activation (+ bias)
  float* g_pre;    // [B, 2L]
  float* c_pre;    // [B, L]
  float* h_f32;    // [(T+1)*B, L]   h_0 .. h_T
  Planes h;        // [(T+1)*B, L]
  Planes rh;       // [T*B, L]
  float* r; float* u; float* c;  // [T*B, L] each
  // heads
  float* zq; float* hq; float* lnq_mean; float* lnq_rstd;       // q_linear_v
  float* zl; float* hl; float* lnl_mean; float* lnl_rstd;       // q_linear_l
  // extra question layer of the answer2 / no_noise variants: pre-activation, output (fp32 + operand planes), LN stats
  float* zqp; float* qp_f32; Planes qp; float* lnqp_mean; float* lnqp_rstd;
  float* dqp;                        // [B, L] gradient w.r.t. that layer's output
  // second branch of the noc variants: Hl as a GEMM operand, joint_l pre-LN / stats / output, and its gradients
  Planes hl_op; float* zjl; float* lnjl_mean; float* lnjl_rstd; Planes jdl;
  float* dJl; float* dzjl_f32; Planes dzjl; float* dXl;
  float* dzqp_f32; Planes dzqp;      // [B, L] gradient w.r.t. its pre-activation
  // full: log-variance layer output, per-sample KL sums, gradient w.r.t. the log-variance
  float* lss; float* kl_rows; float* dlss_f32; Planes dlss;
  // vqa_all / vqa_all2: tuned logits, word-weight logits after the fill, their sum, vqa_all2's argmax input, d tuned
  float* tuned; float* logit1; float* logit_total; float* pred_logit; float* dtuned_f32; Planes dtuned;
  // adapt: pre-LN v_adapt projection (bf16 / fp32 like z), v_adapt as the pooling operand, LN stats, gradient planes,
  // per-sample partials [B, 3, D]
  void* za; Planes va; float* lnva_mean; float* lnva_rstd; Planes dza; float* va_part;
  // ent (M = num_marginal): tiled joint input, its pre-LN projection (bf16 / fp32 like z), LN stats, dropped-out
  // output, tile logits, softmax row statistics, marginal, per-sample negative entropies, and the backward tensors
  Planes x2; void* z2; float* ln2_mean; float* ln2_rstd; Planes jd2; float* logit2; float* row_max; float* row_inv;
  float* marg; float* ent_rows; Planes dl2; float* dJ2; Planes dz2; float* dX2; float* dhl_ent;
  unsigned char* att_bits;   // [B*K*D/8] keep bits of the attention dropout for this step (one byte per 8 elements)
  float* att;      // [B, K]
  float* pooled;   // [B, Dv]  ([B, D] in the adapt variant)
  Planes pooled_op;
  float* zp; float* hp; float* lnp_mean; float* lnp_rstd;       // pooled_linear_l
  Planes x;        // [B, L]  hp (.) hl
  float* zj; float* lnj_mean; float* lnj_rstd;                  // joint_fc
  Planes jd;       // [B, J] after dropout
  float* logit;    // [B, A]
  int* pred;       // [B]
  float* per_sample;  // [6, B]
  float* report;   // [13]
  float* loss;     // [1]
  // backward
  float* dlogit_f32; Planes dlogit;  // [B, A]
  float* dJ;       // [B, J]
  float* dzj_f32; Planes dzj;        // [B, J]
  float* dX;       // [B, L]
  float* dzp_f32; Planes dzp;        // [B, L]
  float* dzl_f32; Planes dzl;        // [B, L]
  float* dP;       // [B, Dv]
  float* dq;       // [B, L]  from q_linear_l
  float* dq2;      // [B, L]  from q_linear_v (the BPTT kernels add the two)
  float* dhq;      // [B, D]
  float* dzq_f32; Planes dzq;        // [B, D]
  Planes dzv;      // [B*K, D]
  float* attn_part;  // per-CTA partial sums of the attention backward
  float* dh[2];    // [B, L] ping-pong
  float* du; float* dh_part; float* dRH;  // [B, L]
  float* dC_f32; Planes dC;  // [T*B, L]
  float* dG_f32; Planes dG;  // [T*B, 2L]
  float* dE;       // [T*B, Wpad]
  float* ln_part_g; float* ln_part_b;  // [B, max(D,L,J)] LayerNorm gamma/beta partials (q_linear_v)
  float* ln_parts[5][2];               // the same for joint_fc, joint_l, pooled_linear_l, q_linear_l, the extra question layer
  unsigned int* gemm_sem;     // split-K hand-over semaphores of the pair GEMM (kGemmSemRegions x kGemmSemElems)
  unsigned int* ar_grid_ctr;  // block counter of the in-kernel exit barrier of the gradient all-reduce
  unsigned int* gru_counter;  // per-row-tile phase counters of the persistent GRU kernels
  bf16* gru_pack;             // [L/32][96][L] packed weight slices of the forward recurrent kernel
  float* gru_bias_part;       // [ceil(B/128)+1, 3L] partial bias gradients of the BPTT kernel
  float* scratch;  // column-sum / loss scratch: (1 + kAux) regions of scratch_floats, one per stream
  size_t scratch_floats;
};

}  // namespace vqa

// global scope: this is the struct the C header forward-declares
struct VqaHandle_t {
  VqaConfig cfg;
  int device;
  int num_sms;
  int Wpad;             // W rounded up to 8 (TMA pitch must be a multiple of 16 bytes)
  int planes;           // 1 (bf16) or 2 (fp32 = hi + lo)
  int M;                // num_marginal of the ent variant (0 otherwise)
  void* ws;
  uint64_t ws_bytes;
  uint64_t ws_needed;
  vqa::Buffers buf;
  vqa::GemmCtx gemm_ctx;
  static constexpr int kGemmSemRegions = 64, kGemmSemElems = 1024;
  bool params_ready;
  // state of the last forward (what backward differentiates)
  bool fwd_valid;
  int last_batch, last_T;
  uint64_t last_seed, last_step;
  VqaAnswerMasks last_masks;
  // auxiliary streams: independent branches of the graph (x-projections vs v-projection, the weight-gradient
  // GEMMs) are forked off the caller's stream and joined back with events -- capturable in a CUDA graph
  static constexpr int kAux = 6;
  cudaStream_t aux[kAux];
  cudaEvent_t ev_fork[kAux], ev_join[kAux];
  bool aux_created;
  // vqa_prefetch_features: the alternate feature planes hold the gather of (image_idx pointer, batch size)
  bool prefetched;
  const void* prefetched_idx;
  int prefetched_batch;
  cudaEvent_t ev_prefetch;
  // a registered request that the next vqa_backward launches next to its weight-gradient GEMMs
  bool pf_pending;
  bool pf_joined;          // auxiliary stream 4 has been joined back into the caller's stream
  cudaStream_t pf_joined_into = nullptr;   // ... namely this one (a forward pass on the same stream need not wait for the gather again)
  VqaFeatureBank pf_bank;
  const void* pf_idx;
  int pf_batch;
  cudaEvent_t ev_upload;   // image_idx of the next batch is on the device
  // vqa_set_deferred_outputs: loss / metrics kernels and output copies of vqa_forward run on auxiliary stream 1 and are
  // joined by the following vqa_backward (or the next entry point that needs them)
  bool defer_outputs;
  bool outputs_pending;
  // the GRU weight repack after an optimizer step runs on auxiliary stream 2; the next forward's GRU waits for it
  bool pack_pending;
  cudaEvent_t ev_pack;
  float* slice_slot;       // vqa_set_embedding_slice_norm: where vqa_backward leaves sum |dE rows|^2 (NULL = off)
  // vqa_set_gradient_allreduce: the data-parallel gradient exchange runs INSIDE vqa_backward (csrc/collective.cu)
  struct {
    float* mc;                  // multicast address of the symmetric gradient buffer (NULL = not registered)
    float* local;               // this rank's own mapping of it (= where the VqaParams gradient struct points)
    long long n_early, n_total; // floats: [0, n_early) is complete before the BPTT, [n_early, n_total) after the weight-gradient section
    unsigned int* mc_flags;     // two barrier counters behind the gradients (multicast / own mapping)
    unsigned int* my_flags;
    unsigned int* grid_ctr;     // local block counter of the exit barrier
    unsigned int flag_total, grid_total;   // channel 0
    // channels 1..3: exchanges that may be in flight at the same time (one per branch of the weight-gradient section)
    // each own a pair of flag words (flags[2 ch], flags[2 ch + 1]), a block counter and their running totals
    unsigned int ch_flag_total[4], ch_grid_total[4];
    int rank, world;
  } ar;
  // vqa_set_optimizer_tail: parameters [tail_begin, n) are updated on an auxiliary stream; the next forward's embedding /
  // x-projection branch and recurrent kernel wait for ev_tail, everything else starts right after the head is updated
  long long adam_tail_begin;
  bool tail_pending;
  cudaEvent_t ev_tail;
  bool early_grads;        // vqa_set_early_gradients
  cudaEvent_t ev_early;    // recorded by vqa_backward once the non-GRU gradients are complete
  // optional per-phase timing
  bool profile;
  bool profile_overlapped;   // events are recorded but the branches still fork: sections of the main stream's critical path
  cudaEvent_t ev[VQA_NUM_PHASES][2];
  bool ev_created;
  bool ev_used[VQA_NUM_PHASES];
};

namespace vqa {

// lays out `buf` inside [base, base + bytes); with base == nullptr only measures. Returns bytes used.
uint64_t plan_workspace(VqaHandle_t* h, uint8_t* base);

}  // namespace vqa
