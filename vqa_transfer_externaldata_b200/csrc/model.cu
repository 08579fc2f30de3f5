// Forward / backward of the answer-model graph as a fixed sequence of kernel launches on the caller's
// stream (no allocation, no synchronisation: capturable in a CUDA graph).
//
// Graph followed: vqa/model_vlmap_answer.py:102-203 (and model_standard.py:193-285): gather -> v_linear_v
// -> embedding + GRU -> q_linear_v -> hadamard_attention -> attention_pooling -> pooled_linear_l,
// q_linear_l, joint_fc (+dropout 0.5) -> WordWeightAnswer / classifier -> soft-score BCE.
// Backward is the hand-derived chain of SURVEY Appendix A; gradients are produced only for the non-NULL
// fields of `grads` (vlmap_answer freezes q_linear_l, pooled_linear_l, joint_fc, WordWeightAnswer --
// vqa/model_vlmap_answer.py:81-89 -- so those layers run dgrad only).
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>
#include <vector>

#include "handle.h"
#include "internal.h"
#include "philox.cuh"

using namespace vqa;

namespace {

struct GemmB {  // builder with the conventions of VqaGemmDesc
  VqaGemmDesc d{};
  GemmB(int M, int N, int K) {
    d.M = M; d.N = N; d.K = K;
  }
  GemmB& a(const Planes& p, long long off, long long ld, bool mn_major) {
    d.a_hi = p.hi + off;
    d.a_lo = p.lo ? p.lo + off : nullptr;
    d.lda = ld;
    d.a_mn_major = mn_major;
    return *this;
  }
  GemmB& b(const Planes& p, long long off, long long ld, bool mn_major) {
    d.b_hi = p.hi + off;
    d.b_lo = p.lo ? p.lo + off : nullptr;
    d.ldb = ld;
    d.b_mn_major = mn_major;
    return *this;
  }
  GemmB& bias(const float* p) { d.bias = p; return *this; }
  GemmB& addend(const float* p, long long ld) { d.addend = p; d.ld_addend = ld; return *this; }
  GemmB& f32(float* p, long long ld) { d.out_f32 = p; d.ld_f32 = ld; return *this; }
  GemmB& planes(const Planes& p, long long off, long long ld) {
    d.out_hi = p.hi + off;
    d.out_lo = p.lo ? p.lo + off : nullptr;
    d.ld_bf = ld;
    return *this;
  }
  int narrow_ = 0;
  GemmB& narrow() { narrow_ |= 1; return *this; }
  GemmB& stable_b() { narrow_ |= 2; return *this; }   // B = a weight shadow nobody upstream in this stream writes
  int sms_ = 0;
  GemmB& sms(int n) { sms_ = n; return *this; }   // plan (and size the persistent grid) for n SMs instead of the whole device
  GemmB& bn(int block_n) { d.block_n = block_n; return *this; }   // VqaGemmDesc.block_n: 0 auto, 64 / 128 / 256, -128 / -256 pair
  VqaStatus run(VqaHandle h, cudaStream_t s) { return gemm_launch(d, sms_ > 0 ? sms_ : h->num_sms, s, &h->gemm_ctx, narrow_); }
};

// modules.fc_layer forward on a rank-2 input: z = a W + bias, then the LayerNorm / activation / Hadamard / dropout tail
// described by `r` (r.z = the pre-LN buffer). bf16 mode: one fused kernel when the shape is eligible (linear_ln.cu),
// else -- and in fp32 mode -- the GEMM followed by the row kernel.
// Which call sites take the fused kernel (VQA_LINEAR_LN_SITES, bit per site: 1 q_linear_l, 2 q_linear_v, 4 pooled_linear_l,
// 8 joint_fc, 16 joint_fc backward, 32 pooled_linear_l backward, 64 the variants' extra question layers).
// Measured inside the cfg1 step (B200, profiles/r02_linear_ln.md): the two question layers, which follow the recurrent
// kernel on an otherwise idle machine, gain (19.6 -> 17.5 us for the pair, step 0.997 -> 0.986 ms); pooled_linear_l and
// joint_fc lose ~3 us each and the two backward sites ~7 us each although the kernel alone beats GEMM + row kernel by
// 3 - 4.5 us: a 16-CTA cluster starts only when a whole GPC has drained, which the CTA-by-CTA hand-over from the
// attention kernel / the concurrent weight-gradient GEMMs delays. Default: the question layers only.
enum { LL_QL = 1, LL_QV = 2, LL_PL = 4, LL_JOINT = 8, LL_JOINT_BWD = 16, LL_PL_BWD = 32, LL_EXTRA = 64 };
bool fused_site(int site) {
  static const int mask = getenv("VQA_LINEAR_LN_SITES") ? atoi(getenv("VQA_LINEAR_LN_SITES")) : (LL_QL | LL_QV | LL_EXTRA);
  return (mask & site) != 0;
}

VqaStatus fc_ln_fwd(VqaHandle h, int site, const Planes& a, long long a_off, long long lda, int K, const Planes& w, const float* bias,
                    float* zbuf, RowLnFwd r, cudaStream_t s) {
  r.z = zbuf;
  if (fused_site(site) && !a.lo && !w.lo && !r.out_lo) {
    LinearLn d{};
    d.M = r.rows; d.N = r.N; d.K = K;
    d.a = a.hi + a_off; d.lda = lda; d.b = w.hi; d.ldb = r.N; d.b_mn_major = 1;
    d.bias = bias; d.gamma = r.gamma; d.beta = r.beta; d.mul = r.mul; d.act = r.act; d.keep = r.keep;
    d.seed = r.seed; d.step = r.step; d.stream_id = r.stream_id;
    d.z = zbuf; d.mean = r.mean; d.rstd = r.rstd; d.y = r.y; d.out_f32 = r.out_f32; d.out_hi = r.out_hi;
    bool launched = false;
    VQA_TRY(linear_ln_launch(d, s, &launched));
    if (launched) return VQA_OK;
  }
  VQA_TRY(GemmB(r.rows, r.N, K).a(a, a_off, lda, false).b(w, 0, r.N, true).bias(bias).f32(zbuf, r.N).stable_b().run(h, s));
  return row_ln_relu_fwd_launch(r, s);
}

// The data gradient d = dy W^T of the layer ABOVE followed by this layer's dropout / Hadamard / activation / LayerNorm
// backward (`r`; r.dout = the buffer that receives d when somebody else needs it too, else scratch for the unfused
// path). W: [N, K] as stored (K contiguous). Fused under the same conditions as fc_ln_fwd.
VqaStatus fc_ln_bwd(VqaHandle h, int site, const Planes& dy, long long lddy, int K, const Planes& w, float* dbuf, bool keep_d,
                    RowLnBwd r, cudaStream_t s) {
  r.dout = dbuf;
  if (fused_site(site) && !dy.lo && !w.lo && !r.dz_lo) {
    LinearLn d{};
    d.M = r.rows; d.N = r.N; d.K = K; d.backward = 1;
    d.a = dy.hi; d.lda = lddy; d.b = w.hi; d.ldb = K; d.b_mn_major = 0;
    d.gamma = r.gamma; d.beta = r.beta; d.mul = r.mul; d.act = r.act; d.keep = r.keep;
    d.seed = r.seed; d.step = r.step; d.stream_id = r.stream_id;
    d.z = const_cast<float*>(r.z); d.mean = const_cast<float*>(r.mean); d.rstd = const_cast<float*>(r.rstd);
    d.raw = keep_d ? dbuf : nullptr; d.dz_f32 = r.dz_f32; d.dz_hi = r.dz_hi;
    d.dgamma_part = r.dgamma_part; d.dbeta_part = r.dbeta_part;
    bool launched = false;
    VQA_TRY(linear_ln_launch(d, s, &launched));
    if (launched) return VQA_OK;
  }
  VQA_TRY(GemmB(r.rows, r.N, K).a(dy, 0, lddy, false).b(w, 0, K, false).f32(dbuf, r.N).stable_b().run(h, s));
  return row_ln_relu_bwd_launch(r, s);
}

// How many leading rows of the v-projection run under the recurrent forward kernel (0: none), and on how many SMs.
// The persistent recurrent grid takes (L / 32) x ceil(B / 128) CTAs (gru_pair.cu); what is left holds idle CTA pairs for
// the whole ~T x 12 us of the chain, and the question heads that follow it (~10 us) do not need Z either. As many
// 256-row tiles move there as finish inside that window (a 256 x 256 tile: ~18 us per 2048 of K on one pair, measured:
// 88 tiles on 10 pairs fit under cfg1's 176 us, 112 do not). Measured on cfg1 (profiles/r02_vproj_split.md): main launch
// 73 -> 57 us, step -7 us. VQA_VPROJ_SPLIT=0 turns it off, =n forces n row tiles.
int vproj_split_plan(VqaHandle h, int Bn, int K, int D, int Dv, int L, int T, bool fp32, bool off, bool keep_bits, long long* rows) {
  static const int env = (getenv("VQA_VPROJ_SPLIT") && *getenv("VQA_VPROJ_SPLIT")) ? atoi(getenv("VQA_VPROJ_SPLIT")) : -1;
  *rows = 0;
  if (off || fp32 || env == 0 || D % 256 != 0) return 0;
  if (!gru_persistent_supported(Bn, L, VQA_PREC_BF16, h->num_sms) || !gru_pair_supported(Bn, L, h->num_sms)) return 0;
  const int gru_ctas = (L / 32) * ((Bn + 127) / 128);
  const int idle = (h->num_sms - gru_ctas) & ~1;
  if (gru_ctas > h->num_sms || idle < 8) return 0;
  const long long M = static_cast<long long>(Bn) * K;
  const int rt = static_cast<int>((M + 255) / 256), ct = D / 256, pairs = h->num_sms / 2, ipairs = idle / 2;
  // the attention-dropout keep bits (ten Philox rounds per 8 elements, ALU only) run on the same idle SMs FIRST, on the same
  // stream: side by side they starve the GEMM's issuer warp (inference at K = 100: the moved rows took 270 us instead of 160)
  const double kb_us = keep_bits ? (static_cast<double>(M) * D / 8.0) * 64.0 / (idle * 4.0 * 32.0 * 1900.0) : 0.0;
  const double tile_us = 18.0 * Dv / 2048.0, window_us = 12.0 * T + 10.0 - kb_us;
  const int rounds = static_cast<int>(window_us / tile_us);
  int rta = rounds * ipairs / ct;
  if (rta > rt - (pairs + ct - 1) / ct) rta = rt - (pairs + ct - 1) / ct;   // leave the main launch a full wave
  // tiles are dealt round-robin over the CTA pairs: the main launch only gets shorter by whole waves. Without one saved
  // the moved rows can only lose (inference at K = 100, B = 512: 800 -> 752 tiles are 11 waves either way; 0.717 -> 0.757 ms)
  if (rta > 0 && ((rt - rta) * ct + pairs - 1) / pairs >= (rt * ct + pairs - 1) / pairs) rta = 0;
  if (env > 0) rta = env;
  if (rta <= 0 || rta >= rt) return 0;
  *rows = static_cast<long long>(rta) * 256;
  return idle;
}

// The attention-dropout keep bits as a precomputed plane (vqa_keep_bits) instead of Philox inside the two attention
// kernels. Measured on B200 (profiles/r02_attn_keep_bits.md): the kernels are latency-bound with idle ALUs, so the ten
// Philox rounds were free and a byte fetched from global memory per 8 elements is NOT: forward 68 -> 77 us, backward
// 102 -> 108 us. Off unless VQA_ATTN_KEEP_BITS=1; the entry point and the kernels' plane path stay (tests cover them).
bool use_keep_bits(const VqaConfig& c, int K, int D) {
  static const bool on = getenv("VQA_ATTN_KEEP_BITS") == nullptr || atoi(getenv("VQA_ATTN_KEEP_BITS")) != 0;
  return on && c.keep_att < 1.0f && (static_cast<long long>(K) * D) % 8 == 0;
}

VqaStatus check_ready(VqaHandle h, const char* who) {
  if (!h) return set_error(VQA_ERR_BAD_ARG, "%s: null handle", who);
  if (!h->ws) return set_error(VQA_ERR_WORKSPACE, "%s: no workspace attached (vqa_set_workspace)", who);
  return VQA_OK;
}

// per-phase event markers (no-ops unless vqa_profile_enable(h, 1))
#define PH_BEGIN(tag)                                                               \
  do {                                                                              \
    if (h->profile) VQA_CUDA_CHECK(cudaEventRecord(h->ev[tag][0], s));              \
  } while (0)
#define PH_END(tag)                                                                 \
  do {                                                                              \
    if (h->profile) {                                                               \
      VQA_CUDA_CHECK(cudaEventRecord(h->ev[tag][1], s));                            \
      h->ev_used[tag] = true;                                                       \
    }                                                                               \
  } while (0)

// fork an independent branch onto auxiliary stream i (returns the caller's stream when serialised for
// per-phase profiling); join makes the caller's stream wait for it
VqaStatus fork_stream(VqaHandle h, int i, cudaStream_t s, cudaStream_t* out) {
  if (h->profile && !h->profile_overlapped) { *out = s; return VQA_OK; }
  VQA_CUDA_CHECK(cudaEventRecord(h->ev_fork[i], s));
  VQA_CUDA_CHECK(cudaStreamWaitEvent(h->aux[i], h->ev_fork[i], 0));
  *out = h->aux[i];
  return VQA_OK;
}
VqaStatus join_stream(VqaHandle h, int i, cudaStream_t s) {
  if (h->profile && !h->profile_overlapped) return VQA_OK;
  VQA_CUDA_CHECK(cudaEventRecord(h->ev_join[i], h->aux[i]));
  VQA_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_join[i], 0));
  return VQA_OK;
}

VqaStatus copy_out(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (!dst || bytes == 0) return VQA_OK;
  VQA_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
  return VQA_OK;
}

// launch the registered feature prefetch on auxiliary stream 4 (already forked from the main stream by the caller): an
// HBM-bound copy on the SMs the cooperative BPTT grid (128 CTAs, one per SM) does not use. Joined by the
// weight-gradient section.
VqaStatus launch_pending_prefetch(VqaHandle h, cudaStream_t a4, int ctas = 0) {
  h->pf_pending = false;
  const VqaConfig& c = h->cfg;
  Buffers& b = h->buf;
  int free_sms = ctas != 0 ? ctas : h->num_sms - 128;   // default: the CTA-pair recurrent kernels occupy 128 SMs (gru_pair.cu)
  if (free_sms >= 0 && free_sms < 4) free_sms = 4;
  if (free_sms < 0) free_sms = 0;   // (the plain full-width kernel)
  VQA_CUDA_CHECK(cudaStreamWaitEvent(a4, h->ev_upload, 0));
  if (h->pf_bank.features_bf16 && c.precision == VQA_PREC_BF16)
    VQA_TRY(gather_features_bf16_launch(static_cast<const bf16*>(h->pf_bank.features_bf16), h->pf_bank.num_boxes,
                                        static_cast<const long long*>(h->pf_idx), h->pf_batch, c.K, c.Dv, b.v_alt.hi,
                                        b.nbox_alt, h->pf_bank.num_images, a4, free_sms));
  else
    VQA_TRY(gather_features_launch(h->pf_bank.features, h->pf_bank.num_boxes, static_cast<const long long*>(h->pf_idx),
                                   h->pf_batch, c.K, c.Dv, b.v_alt.hi, b.v_alt.lo, b.nbox_alt, h->pf_bank.num_images, a4, free_sms));
  VQA_CUDA_CHECK(cudaEventRecord(h->ev_prefetch, a4));
  h->prefetched = true;
  h->prefetched_idx = h->pf_idx;
  h->prefetched_batch = h->pf_batch;
  h->pf_joined = false;
  return VQA_OK;
}

}  // namespace

extern "C" {

VQA_API VqaStatus vqa_prepare_params(VqaHandle h, const VqaParams* p, void* stream) {
  VQA_TRY(check_ready(h, "vqa_prepare_params"));
  if (!p) return set_error(VQA_ERR_BAD_ARG, "vqa_prepare_params: null params");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->tail_pending) VQA_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_tail, 0));   // (a split optimizer step may still be updating the tail)
  const VqaConfig& c = h->cfg;
  WeightShadows& w = h->buf.w;
  struct Item { const float* src; Planes* dst; long long rows, cols; } items[] = {
      {p->v_w, &w.v_w, c.Dv, c.D},
      {p->gru_gates_w, &w.gru_gates_w, c.W + c.L, 2LL * c.L},
      {p->gru_cand_w, &w.gru_cand_w, c.W + c.L, c.L},
      {p->qv_w, &w.qv_w, c.L, c.D},
      {p->pl_w, &w.pl_w, c.variant == VQA_VARIANT_VLMAP_ANSWER_ADAPT ? c.D : c.Dv, c.L},   // adapt pools the D-wide v_adapt
      {p->ql_w, &w.ql_w, c.L, c.L},
      {p->joint_w, &w.joint_w, c.L, c.J},
      {p->ans_w, &w.ans_w, c.J, c.A},
      {p->qp_w, &w.qp_w, c.L, c.L},   // answer2 / no_noise only (NULL otherwise)
      {p->jl_w, &w.jl_w, c.L, c.J},   // noc only
      {p->al_w, &w.al_w, c.J, c.A},   // noc only
      {p->qs_w, &w.qs_w, c.L, c.L},   // full only
      {p->tw_w, &w.tw_w, c.J, c.A},   // vqa_all / vqa_all2 only
      {p->va_w, &w.va_w, c.Dv, c.D},  // adapt only
  };
  // the refreshes are independent small memory-bound kernels: spread them over the auxiliary streams instead of
  // queueing them behind one another (they sit between the optimizer and the next step's first GEMM)
  int lane_i = 0;
  bool forked[VqaHandle_t::kAux] = {};
  cudaStream_t gru_stream = s;
  for (auto& it : items) {
    const bool has_qp = c.variant == VQA_VARIANT_VLMAP_ANSWER2 || c.variant == VQA_VARIANT_VLMAP_ANSWER_NO_NOISE ||
                        c.variant == VQA_VARIANT_VLMAP_ANSWER_FULL;
    if (it.dst == &w.qp_w && !has_qp) continue;   // the base variants have no such layer
    if ((it.dst == &w.jl_w || it.dst == &w.al_w) && c.variant != VQA_VARIANT_VLMAP_ANSWER_NOC) continue;
    if (it.dst == &w.qs_w && c.variant != VQA_VARIANT_VLMAP_ANSWER_FULL) continue;
    if (it.dst == &w.tw_w && c.variant != VQA_VARIANT_VLMAP_ANSWER_VQA_ALL && c.variant != VQA_VARIANT_VLMAP_ANSWER_VQA_ALL2) continue;
    if (it.dst == &w.va_w && c.variant != VQA_VARIANT_VLMAP_ANSWER_ADAPT) continue;
    if (!it.src) {
      if (!h->params_ready) return set_error(VQA_ERR_BAD_ARG, "vqa_prepare_params: the first call needs every weight");
      continue;  // unchanged since the last call
    }
    const bool is_gru = it.src == p->gru_gates_w || it.src == p->gru_cand_w;
    int ai = is_gru ? 0 : 1 + (lane_i++ % (VqaHandle_t::kAux - 1));   // both GRU matrices on aux 0: the pack follows them
    cudaStream_t st;
    if (!forked[ai]) {
      VQA_TRY(fork_stream(h, ai, s, &st));
      forked[ai] = true;
    } else {
      st = (h->profile && !h->profile_overlapped) ? s : h->aux[ai];
    }
    if (is_gru) gru_stream = st;
    VQA_TRY(split_bf16_launch(it.src, it.rows, it.cols, it.cols, it.dst->hi, it.dst->lo, it.cols, st));
  }
  if ((p->gru_gates_w || p->gru_cand_w) && gru_persistent_supported(c.B, c.L, c.precision, h->num_sms))
    VQA_TRY(gru_pack_weights_launch(w.gru_gates_w.hi + static_cast<long long>(c.W) * 2 * c.L,
                                    w.gru_cand_w.hi + static_cast<long long>(c.W) * c.L, c.L, h->buf.gru_pack, gru_stream));
  for (int i = 0; i < VqaHandle_t::kAux; ++i)
    if (forked[i]) VQA_TRY(join_stream(h, i, s));
  h->params_ready = true;
  return VQA_OK;
}

VQA_API VqaStatus vqa_forward(VqaHandle h, const VqaParams* p, const VqaFeatureBank* bank,
                              const VqaBatch* batch, const VqaAnswerMasks* masks, uint64_t seed,
                              uint64_t step, const VqaOutputs* out, void* stream) {
  VQA_TRY(check_ready(h, "vqa_forward"));
  if (!p || !bank || !batch || !masks) return set_error(VQA_ERR_BAD_ARG, "vqa_forward: null argument");
  if (!h->params_ready) return set_error(VQA_ERR_STATE, "vqa_forward: call vqa_prepare_params first");
  const VqaConfig& c = h->cfg;
  const int Bn = batch->batch_size, T = batch->q_len_max;
  if (Bn < 0 || Bn > c.B || T <= 0 || T > c.T)
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_forward: batch_size %d (max %d) / q_len_max %d (max %d)", Bn,
                     c.B, T, c.T);
  if (!bank->features || !bank->num_boxes || !batch->image_idx || !batch->q_intseq ||
      !batch->q_intseq_len || !batch->answer_target)
    return set_error(VQA_ERR_BAD_ARG, "vqa_forward: null batch / bank pointer");
  if (bank->num_images <= 0) return set_error(VQA_ERR_BAD_ARG, "vqa_forward: feature bank with num_images %lld", static_cast<long long>(bank->num_images));
  h->fwd_valid = false;
  if (Bn == 0) return VQA_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->outputs_pending) {   // deferred outputs of a forward that no backward pass followed
    VQA_TRY(join_stream(h, 1, s));
    h->outputs_pending = false;
  }
  Buffers& b = h->buf;
  const int K = c.K, Dv = c.Dv, D = c.D, L = c.L, J = c.J, A = c.A, W = c.W, Wp = h->Wpad;
  const bool fp32 = c.precision == VQA_PREC_FP32;
  const long long BL = static_cast<long long>(Bn) * L;
  const bool v_full = c.variant == VQA_VARIANT_VLMAP_ANSWER_FULL;
  const bool v_tuned = c.variant == VQA_VARIANT_VLMAP_ANSWER_VQA_ALL || c.variant == VQA_VARIANT_VLMAP_ANSWER_VQA_ALL2;
  const bool v_adapt = c.variant == VQA_VARIANT_VLMAP_ANSWER_ADAPT;
  const int Pd = v_adapt ? D : Dv;   // width of the pooled vector

  // branch 0 (auxiliary stream): embedding lookup + the hoisted x-parts of the GRU pre-activations; it is
  // epilogue/store-bound and overlaps the MMA-bound v-projection below      (:134-137, modules.py:124-140)
  cudaStream_t s1 = s;
  static const bool x_bf16_env = getenv("VQA_GRU_X_BF16") == nullptr || atoi(getenv("VQA_GRU_X_BF16")) != 0;
  const bool x_bf16 = x_bf16_env && !fp32 && gru_persistent_supported(Bn, L, c.precision, h->num_sms);
  auto gru_inputs = [&](cudaStream_t st) -> VqaStatus {
    if (h->tail_pending) {   // the embedding / GRU parameters of the last optimizer step (vqa_set_optimizer_tail)
      VQA_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_tail, 0));
      h->tail_pending = false;
    }
    VQA_TRY(embed_gather_launch(p->embed, batch->q_intseq, Bn, T, T, W, Wp, c.Vq, b.e.hi, b.e.lo, st));
    // bf16 mode with the persistent recurrent kernels: the hoisted x-parts are stored as bf16 (in the same buffers) --
    // these two products are store-bound (K = 300: 88 MB of fp32 per step at cfg1), and the recurrent kernel reads them back
    GemmB gx(T * Bn, 2 * L, W), cx(T * Bn, L, W);
    gx.a(b.e, 0, Wp, false).b(b.w.gru_gates_w, 0, 2 * L, true).bias(p->gru_gates_b);
    cx.a(b.e, 0, Wp, false).b(b.w.gru_cand_w, 0, L, true).bias(p->gru_cand_b);
    if (x_bf16) {
      Planes pg, pc;
      pg.hi = reinterpret_cast<bf16*>(b.xg); pc.hi = reinterpret_cast<bf16*>(b.xc);
      gx.planes(pg, 0, 2 * L); cx.planes(pc, 0, L);
    } else {
      gx.f32(b.xg, 2 * L); cx.f32(b.xc, L);
    }
    VQA_TRY(gx.run(h, st));
    VQA_TRY(cx.run(h, st));
    VQA_TRY(fill_zero_launch(b.h_f32, sizeof(float) * BL, st));
    VQA_TRY(fill_zero_launch(b.h.hi, sizeof(bf16) * BL, st));
    if (b.h.lo) VQA_TRY(fill_zero_launch(b.h.lo, sizeof(bf16) * BL, st));
    return VQA_OK;
  };
  const bool serial = h->profile && !h->profile_overlapped;
  if (!serial) {
    VQA_TRY(fork_stream(h, 0, s, &s1));
    VQA_TRY(gru_inputs(s1));
  }

  PH_BEGIN(VQA_PH_GATHER);
  // a0: V = features[image_idx], nbox = num_boxes[image_idx]      (model_vlmap_answer.py:110-123)
  if (h->prefetched && h->prefetched_idx == batch->image_idx && h->prefetched_batch == Bn) {
    // vqa_prefetch_features already gathered this batch into the alternate planes (under the previous step's
    // backward): adopt them
    std::swap(b.v, b.v_alt);
    std::swap(b.nbox, b.nbox_alt);
    // (the backward pass that launched the gather has joined its stream into this one already: nothing to wait for then)
    static const bool skip_joined = getenv("VQA_PF_SKIP_WAIT") == nullptr || atoi(getenv("VQA_PF_SKIP_WAIT")) != 0;
    if (!(skip_joined && h->pf_joined && h->pf_joined_into == s)) VQA_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_prefetch, 0));
  } else if (bank->features_bf16 && !fp32) {
    VQA_TRY(gather_features_bf16_launch(static_cast<const bf16*>(bank->features_bf16), bank->num_boxes,
                                        reinterpret_cast<const long long*>(batch->image_idx), Bn, K, Dv, b.v.hi, b.nbox,
                                        bank->num_images, s));
  } else {
    VQA_TRY(gather_features_launch(bank->features, bank->num_boxes,
                                   reinterpret_cast<const long long*>(batch->image_idx), Bn, K, Dv, b.v.hi,
                                   b.v.lo, b.nbox, bank->num_images, s));
  }
  h->prefetched = false;
  h->pf_pending = false;   // a request no backward pass picked up (inference loops): this forward gathered itself
  PH_END(VQA_PH_GATHER);
  PH_BEGIN(VQA_PH_VPROJ_FWD);
  // a1: Z = V Wv + bv (LayerNorm over (K, D) + ReLU are applied inside the attention kernels)
  // The recurrent kernel below is a latency chain on 128 of the 148 SMs; the v-projection is 288 tiles of 256 x 256 on 74
  // CTA pairs. Its first rows therefore run UNDER the recurrent kernel, as a persistent launch on the CTA pairs that grid
  // leaves idle, and only the rest runs here (rows of the product are independent).
  long long vsplit_rows = 0;
  const int v_idle_sms = vproj_split_plan(h, Bn, K, D, Dv, L, T, fp32, serial || v_adapt, use_keep_bits(c, K, D), &vsplit_rows);
  {
    GemmB g(static_cast<int>(Bn * K - vsplit_rows), D, Dv);
    g.a(b.v, vsplit_rows * Dv, Dv, false).b(b.w.v_w, 0, D, true).bias(p->v_b);
    if (fp32) g.f32(static_cast<float*>(b.z) + vsplit_rows * D, D);
    else { Planes zp; zp.hi = static_cast<bf16*>(b.z); g.planes(zp, vsplit_rows * D, D); }
    VQA_TRY(g.run(h, s));
  }
  if (v_adapt) {
    // v_adapt = relu(LN_{K,D}(V Wa + ba)): the tensor attention pools in this variant (model_vlmap_answer_adapt.py:132-142)
    if (!p->va_w || !p->va_b || !p->va_gamma || !p->va_beta)
      return set_error(VQA_ERR_BAD_ARG, "vqa_forward: the adapt variant needs va_*");
    GemmB g(Bn * K, D, Dv);
    g.a(b.v, 0, Dv, false).b(b.w.va_w, 0, D, true).bias(p->va_b);
    if (fp32) g.f32(static_cast<float*>(b.za), D);
    else { Planes zp; zp.hi = static_cast<bf16*>(b.za); g.planes(zp, 0, D); }
    VQA_TRY(g.run(h, s));
    SlabLnFwd f{};
    f.batch = Bn; f.K = K; f.D = D; f.z = b.za; f.gamma = p->va_gamma; f.beta = p->va_beta;
    f.out_hi = b.va.hi; f.out_lo = b.va.lo; f.mean = b.lnva_mean; f.rstd = b.lnva_rstd;
    f.thr = 65536u; f.inv_keep = 1.f;   // no dropout on v_adapt
    VQA_TRY(slab_ln_relu_fwd_launch(f, c.precision, s));
  }
  PH_END(VQA_PH_VPROJ_FWD);
  VQA_TRY(join_stream(h, 0, s));
  PH_BEGIN(VQA_PH_GRU_FWD);
  if (h->pack_pending) {   // the repack of the GRU weights after the last optimizer step ran on an auxiliary stream
    VQA_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_pack, 0));
    h->pack_pending = false;
  }
  if (serial) VQA_TRY(gru_inputs(s));  // serialised for per-phase timing
  // a2: the recurrent part of the GRU
  const bool persistent = gru_persistent_supported(Bn, L, c.precision, h->num_sms);
  const bool want_bits = use_keep_bits(c, K, D);
  bool kb_forked = false;
  if (persistent) {
    GruFwdPersistent a{};
    a.B = Bn; a.L = L; a.T = T; a.q_len = batch->q_intseq_len; a.counter = b.gru_counter;
    a.xg = b.xg; a.xc = b.xc; a.x_bf16 = x_bf16 ? 1 : 0; a.h_f32 = b.h_f32; a.h_bf = b.h.hi; a.rh_bf = b.rh.hi;
    a.r = b.r; a.u = b.u; a.c = b.c;
    a.w_pack = b.gru_pack;
    // The step's attention-dropout keep bits, once, for both attention kernels: an ALU-only kernel (ten Philox rounds
    // per 8 elements) on the SMs the cooperative recurrent grid leaves idle -- forked BEFORE that launch, enqueued AFTER
    // it (the cooperative grid must be first in line), joined before the attention kernel. Beside the v-projection GEMM
    // it cost that GEMM 10 us (measured).
    cudaStream_t a5 = s, a2 = s;
    kb_forked = want_bits && !serial;
    if (kb_forked) VQA_TRY(fork_stream(h, 5, s, &a5));
    if (vsplit_rows > 0 && !kb_forked) VQA_TRY(fork_stream(h, 2, s, &a2));   // (forked BEFORE the cooperative launch, enqueued AFTER it)
    VQA_TRY(gru_fwd_persistent_launch(a, h->num_sms, s));
    if (want_bits)
      VQA_TRY(keep_bits_launch(b.att_bits, static_cast<long long>(Bn) * K * D, c.keep_att, seed, step, RNG_STREAM_ATT, a5,
                               kb_forked ? 8 * (h->num_sms > 128 ? h->num_sms - 128 : 8) : 0));
    if (vsplit_rows > 0) {   // rows [0, vsplit_rows) of the v-projection on the idle CTA pairs (after the keep bits, same stream)
      Planes zp; zp.hi = static_cast<bf16*>(b.z);
      VQA_TRY(GemmB(static_cast<int>(vsplit_rows), D, Dv).a(b.v, 0, Dv, false).b(b.w.v_w, 0, D, true).bias(p->v_b).planes(zp, 0, D)
                  .sms(v_idle_sms).run(h, kb_forked ? a5 : a2));
    }
  } else {
    if (want_bits)
      VQA_TRY(keep_bits_launch(b.att_bits, static_cast<long long>(Bn) * K * D, c.keep_att, seed, step, RNG_STREAM_ATT, s));
    for (int t = 0; t < T; ++t) {
      VQA_TRY(GemmB(Bn, 2 * L, L).a(b.h, t * BL, L, false)
                  .b(b.w.gru_gates_w, static_cast<long long>(W) * 2 * L, 2 * L, true)
                  .addend(b.xg + static_cast<long long>(t) * Bn * 2 * L, 2 * L).f32(b.g_pre, 2 * L).run(h, s));
      VQA_TRY(gru_gates_launch(b.g_pre, b.h_f32 + t * BL, Bn, L, b.r + t * BL, b.u + t * BL,
                               b.rh.hi + t * BL, b.rh.lo ? b.rh.lo + t * BL : nullptr, s));
      VQA_TRY(GemmB(Bn, L, L).a(b.rh, t * BL, L, false)
                  .b(b.w.gru_cand_w, static_cast<long long>(W) * L, L, true)
                  .addend(b.xc + static_cast<long long>(t) * Bn * L, L).f32(b.c_pre, L).run(h, s));
      VQA_TRY(gru_update_launch(b.c_pre, b.h_f32 + t * BL, b.u + t * BL, batch->q_intseq_len, t, Bn, L,
                                b.c + t * BL, b.h_f32 + (t + 1) * BL, b.h.hi + (t + 1) * BL,
                                b.h.lo ? b.h.lo + (t + 1) * BL : nullptr, s));
    }
  }
  const float* q = b.h_f32 + T * BL;
  const long long q_off = T * BL;

  PH_END(VQA_PH_GRU_FWD);
  PH_BEGIN(VQA_PH_QHEADS_FWD);
  // a6 (question half) on the auxiliary stream: Hl = relu(LN(q Wl + b))   (:170-174); needed only by the joint head
  VQA_TRY(fork_stream(h, 0, s, &s1));
  // q_linear_v is enqueued FIRST: the attention block waits for it, and the device co-schedules 7 of the 8 sixteen-CTA
  // clusters the two question layers ask for -- the one left waiting must be q_linear_l's, which nobody needs before the head
  // a3: Hq = relu(LN(q Wqv + b))                                    (:142-145)
  auto run_qv = [&]() -> VqaStatus {
    RowLnFwd r{};
    r.rows = Bn; r.N = D; r.gamma = p->qv_gamma; r.beta = p->qv_beta; r.keep = 1.f;
    r.y = b.hq; r.mean = b.lnq_mean; r.rstd = b.lnq_rstd;
    return fc_ln_fwd(h, LL_QV, b.h, q_off, L, L, b.w.qv_w, p->qv_b, b.zq, r, s);
  };
  static const bool qv_first = getenv("VQA_QV_FIRST") == nullptr || atoi(getenv("VQA_QV_FIRST")) != 0;
  if (qv_first) VQA_TRY(run_qv());
  const bool has_qp = c.variant == VQA_VARIANT_VLMAP_ANSWER2 || c.variant == VQA_VARIANT_VLMAP_ANSWER_NO_NOISE || v_full;
  if (has_qp) {
    if (!p->qp_w || !p->qp_b) return set_error(VQA_ERR_BAD_ARG, "vqa_forward: this variant needs qp_w / qp_b");
    if (v_full) {
      // q_L_mean, q_L_log_sigma_sq (both linear), q_L_mean_noise = mean + N(0,1) * sqrt(exp(lss))   (_full.py:124-134)
      if (!p->qs_w || !p->qs_b) return set_error(VQA_ERR_BAD_ARG, "vqa_forward: the full variant needs qs_w / qs_b");
      VQA_TRY(GemmB(Bn, L, L).a(b.h, q_off, L, false).b(b.w.qp_w, 0, L, true).bias(p->qp_b).f32(b.qp_f32, L).run(h, s1));
      VQA_TRY(GemmB(Bn, L, L).a(b.h, q_off, L, false).b(b.w.qs_w, 0, L, true).bias(p->qs_b).f32(b.lss, L).run(h, s1));
      ReparamFwd r{};
      r.batch = Bn; r.L = L; r.mean = b.qp_f32; r.lss = b.lss; r.seed = seed; r.step = step;
      r.out_hi = b.qp.hi; r.out_lo = b.qp.lo; r.kl_rows = b.kl_rows;
      VQA_TRY(reparam_fwd_launch(r, s1));
    } else if (c.variant == VQA_VARIANT_VLMAP_ANSWER2) {
      // q_L_ft2 = tanh(LN(q W2 + b2))           (vqa/model_vlmap_answer2.py:127-130)
      if (!p->qp_gamma || !p->qp_beta) return set_error(VQA_ERR_BAD_ARG, "vqa_forward: answer2 needs qp_gamma / qp_beta");
      RowLnFwd r{};
      r.rows = Bn; r.N = L; r.gamma = p->qp_gamma; r.beta = p->qp_beta; r.keep = 1.f; r.act = 1;
      r.y = b.qp_f32; r.out_hi = b.qp.hi; r.out_lo = b.qp.lo; r.mean = b.lnqp_mean; r.rstd = b.lnqp_rstd;
      VQA_TRY(fc_ln_fwd(h, LL_EXTRA, b.h, q_off, L, L, b.w.qp_w, p->qp_b, b.zqp, r, s1));
    } else {
      // q_L_mean = q Wm + bm: no LayerNorm, no activation   (vqa/model_vlmap_answer_no_noise.py:122-125)
      VQA_TRY(GemmB(Bn, L, L).a(b.h, q_off, L, false).b(b.w.qp_w, 0, L, true).bias(p->qp_b).f32(b.qp_f32, L)
                  .planes(b.qp, 0, L).run(h, s1));
    }
  }
  const Planes& ql_in = has_qp ? b.qp : b.h;
  const long long ql_in_off = has_qp ? 0 : T * BL;
  {
    RowLnFwd r{};
    r.rows = Bn; r.N = L; r.gamma = p->ql_gamma; r.beta = p->ql_beta; r.keep = 1.f;
    r.y = b.hl; r.mean = b.lnl_mean; r.rstd = b.lnl_rstd;
    if (c.variant == VQA_VARIANT_VLMAP_ANSWER_NOC) { r.out_hi = b.hl_op.hi; r.out_lo = b.hl_op.lo; }   // joint_l reads Hl
    VQA_TRY(fc_ln_fwd(h, LL_QL, ql_in, ql_in_off, L, L, b.w.ql_w, p->ql_b, b.zl, r, s1));
  }
  if (!qv_first) VQA_TRY(run_qv());
  PH_END(VQA_PH_QHEADS_FWD);
  if (kb_forked) VQA_TRY(join_stream(h, 5, s));   // the keep bits
  if (vsplit_rows > 0 && !kb_forked) VQA_TRY(join_stream(h, 2, s));   // the first rows of Z (else: on the keep bits' stream, joined above)
  PH_BEGIN(VQA_PH_ATTN_FWD);
  // a4 + a5: attention + pooling                                     (:151-156)
  {
    VqaAttnFwd a{};
    a.batch = Bn; a.z = b.z; a.gamma = p->v_gamma; a.beta = p->v_beta; a.hq = b.hq;
    a.att_w = p->att_w; a.att_b = p->att_b; a.nbox = b.nbox;
    a.v_hi = v_adapt ? b.va.hi : b.v.hi; a.v_lo = v_adapt ? b.va.lo : b.v.lo;   // adapt pools v_adapt [K, D]
    a.seed = seed; a.step = step; a.att = b.att; a.pooled = b.pooled; a.pooled_hi = b.pooled_op.hi;
    a.pooled_lo = b.pooled_op.lo; a.ln_mean = b.lnv_mean; a.ln_rstd = b.lnv_rstd;
    a.keep_bits = use_keep_bits(c, K, D) ? b.att_bits : nullptr;
    VQA_TRY(attn_fwd_launch(a, K, D, Pd, c.precision, c.keep_att, s));
  }
  PH_END(VQA_PH_ATTN_FWD);
  PH_BEGIN(VQA_PH_HEAD_FWD);
  // a6: Hp = relu(LN(P Wp + b)); X = Hp (.) Hl; Jd = dropout(relu(LN(X Wj + b)), 0.5)   (:163-181)
  VQA_TRY(join_stream(h, 0, s));  // Hl (the Hadamard partner in the layer's epilogue)
  {
    RowLnFwd r{};
    r.rows = Bn; r.N = L; r.gamma = p->pl_gamma; r.beta = p->pl_beta; r.keep = 1.f;
    const bool noc_f = c.variant == VQA_VARIANT_VLMAP_ANSWER_NOC;   // noc: no Hadamard, joint_v reads Hp itself
    r.mul = noc_f ? nullptr : b.hl; r.y = b.hp; r.out_hi = b.x.hi; r.out_lo = b.x.lo; r.mean = b.lnp_mean; r.rstd = b.lnp_rstd;
    VQA_TRY(fc_ln_fwd(h, LL_PL, b.pooled_op, 0, Pd, Pd, b.w.pl_w, p->pl_b, b.zp, r, s));
  }
  {
    RowLnFwd r{};
    r.rows = Bn; r.N = J; r.gamma = p->joint_gamma; r.beta = p->joint_beta;
    r.keep = c.keep_joint; r.seed = seed; r.step = step; r.stream_id = RNG_STREAM_JOINT;
    r.out_hi = b.jd.hi; r.out_lo = b.jd.lo; r.mean = b.lnj_mean; r.rstd = b.lnj_rstd;
    VQA_TRY(fc_ln_fwd(h, LL_JOINT, b.x, 0, L, L, b.w.joint_w, p->joint_b, b.zj, r, s));
  }
  // a7: logits against the (exported vlmap word) weights              (:183-185)
  VQA_TRY(GemmB(Bn, A, J).a(b.jd, 0, J, false).b(b.w.ans_w, 0, A, true).bias(p->ans_b).f32(b.logit, A).stable_b().run(h, s));
  if (c.variant == VQA_VARIANT_VLMAP_ANSWER_NOC) {
    // second branch (model_vlmap_answer_noc.py:184-203): Jl = dropout(relu(LN(Hl Wjl + b))), logit += Jl Wal + bal
    if (!p->jl_w || !p->jl_b || !p->jl_gamma || !p->jl_beta || !p->al_w || !p->al_b)
      return set_error(VQA_ERR_BAD_ARG, "vqa_forward: the noc variant needs jl_* and al_*");
    RowLnFwd r{};
    r.rows = Bn; r.N = J; r.gamma = p->jl_gamma; r.beta = p->jl_beta;
    r.keep = c.keep_joint; r.seed = seed; r.step = step; r.stream_id = RNG_STREAM_JOINT_L;
    r.out_hi = b.jdl.hi; r.out_lo = b.jdl.lo; r.mean = b.lnjl_mean; r.rstd = b.lnjl_rstd;
    VQA_TRY(fc_ln_fwd(h, LL_EXTRA, b.hl_op, 0, L, L, b.w.jl_w, p->jl_b, b.zjl, r, s));
    VQA_TRY(GemmB(Bn, A, J).a(b.jdl, 0, J, false).b(b.w.al_w, 0, A, true).bias(p->al_b).addend(b.logit, A)
                .f32(b.logit, A).run(h, s));
  }
  const float* out_logit = b.logit;
  if (v_tuned) {
    // TunedWordWeightAnswer reads `joint` itself (model_vlmap_answer_vqa_all.py:215-216), then the logits are combined
    if (!p->tw_w || !p->tw_b) return set_error(VQA_ERR_BAD_ARG, "vqa_forward: the vqa_all variants need tw_w / tw_b");
    VQA_TRY(GemmB(Bn, A, J).a(b.jd, 0, J, false).b(b.w.tw_w, 0, A, true).bias(p->tw_b).f32(b.tuned, A).run(h, s));
    TunedHeadFwd t{};
    t.batch = Bn; t.A = A; t.num_train_answer = c.num_train_answer;
    t.fill_min = c.variant == VQA_VARIANT_VLMAP_ANSWER_VQA_ALL;
    t.logit0 = b.logit; t.tuned = b.tuned; t.exist = masks->answer_exist; t.l1 = b.logit1; t.total = b.logit_total;
    t.pred_logit = t.fill_min ? nullptr : b.pred_logit;
    VQA_TRY(tuned_combine_launch(t, s));
    out_logit = b.logit_total;
  }
  const bool v_ent = c.variant == VQA_VARIANT_VLMAP_ANSWER_ENT;
  if (v_ent) {
    // maximum-entropy regulariser (model_vlmap_answer_ent.py:193-213): the joint head on M tiles per sample
    const int M = h->M, BM = Bn * M;
    EntTile t{};
    t.batch = Bn; t.M = M; t.L = L; t.hp = b.hp; t.hl = b.hl; t.out_hi = b.x2.hi; t.out_lo = b.x2.lo;
    VQA_TRY(ent_tile_launch(t, s));
    {
      GemmB g(BM, J, L);
      g.a(b.x2, 0, L, false).b(b.w.joint_w, 0, J, true).bias(p->joint_b);
      if (fp32) g.f32(static_cast<float*>(b.z2), J);
      else { Planes zp; zp.hi = static_cast<bf16*>(b.z2); g.planes(zp, 0, J); }
      VQA_TRY(g.run(h, s));
    }
    SlabLnFwd f{};   // LayerNorm over the whole [M, J] slab, ReLU, dropout 0.5
    f.batch = Bn; f.K = M; f.D = J; f.z = b.z2; f.gamma = p->joint_gamma; f.beta = p->joint_beta;
    f.out_hi = b.jd2.hi; f.out_lo = b.jd2.lo; f.mean = b.ln2_mean; f.rstd = b.ln2_rstd;
    f.thr = keep_threshold(c.keep_joint); f.inv_keep = 1.f / c.keep_joint; f.seed = seed; f.step = step; f.site = RNG_STREAM_ENT;
    VQA_TRY(slab_ln_relu_fwd_launch(f, c.precision, s));
    VQA_TRY(GemmB(BM, A, J).a(b.jd2, 0, J, false).b(b.w.ans_w, 0, A, true).bias(p->ans_b).f32(b.logit2, A).run(h, s));
    EntMarginal e{};
    e.batch = Bn; e.M = M; e.A = A; e.num_train_answer = c.num_train_answer; e.logit2 = b.logit2;
    e.exist = masks->answer_exist; e.marg = b.marg; e.row_max = b.row_max; e.row_inv = b.row_inv; e.ent_rows = b.ent_rows;
    VQA_TRY(ent_marginal_launch(e, s));
  }
  PH_END(VQA_PH_HEAD_FWD);
  PH_BEGIN(VQA_PH_LOSS);
  // a8 + a9: loss, pred, report                                       (:192-288)
  // Nothing downstream of the logits needs the loss / metrics kernels (the backward pass recomputes d logits from the
  // logits), so a training step may defer them to auxiliary stream 1 (vqa_set_deferred_outputs): they then run beside
  // the first kernels of vqa_backward, which joins them.
  const bool defer = h->defer_outputs && !(h->profile && !h->profile_overlapped);
  cudaStream_t s_main = s;
  float* loss_scratch = b.scratch;
  if (defer) {
    VQA_TRY(fork_stream(h, 1, s_main, &s));          // `s` is the loss / output stream until the end of this function
    loss_scratch = b.scratch + b.scratch_floats;      // auxiliary stream 1's scratch region (as in the weight-gradient section)
  }
  const int use_tm = c.variant != VQA_VARIANT_STANDARD;  // every vlmap_answer* variant masks the loss to the train answers
  if (c.variant == VQA_VARIANT_VLMAP_ANSWER_VQA_ALL) {
    // (BCE(fixed) + BCE(fixed + tuned)) * train_mask; pred from fixed + tuned              (_vqa_all.py:234-244)
    VQA_TRY(bce_metrics2_launch(Bn, A, c.num_train_answer, 1, b.logit_total, b.logit1, 1, nullptr, batch->answer_target,
                                *masks, b.loss, b.report, b.pred, b.per_sample, loss_scratch, s));
  } else if (c.variant == VQA_VARIANT_VLMAP_ANSWER_VQA_ALL2) {
    // BCE(fixed) * train_mask + BCE(tuned); pred from fixed * test_mask + tuned * train_mask (_vqa_all2.py:231-242)
    VQA_TRY(bce_metrics2_launch(Bn, A, c.num_train_answer, 1, b.logit1, b.tuned, 0, b.pred_logit, batch->answer_target,
                                *masks, b.loss, b.report, b.pred, b.per_sample, loss_scratch, s));
  } else {
    VQA_TRY(bce_metrics_launch(Bn, A, c.num_train_answer, use_tm, b.logit, batch->answer_target, *masks, 0.f,
                               b.loss, b.report, b.pred, b.per_sample, nullptr, nullptr, nullptr, loss_scratch, s));
  }
  if (v_full)   // loss += 0.1 * KL(q_L_mean, q_L_log_sigma_sq); report latent_loss / train_latent_loss  (_full.py:217-223)
    VQA_TRY(latent_finalize_launch(b.kl_rows, Bn, -0.5f, VQA_LATENT_LOSS_WEIGHT, VQA_REPORT_LATENT_LOSS, b.loss, b.report, s));
  if (v_ent)    // loss += 0.1 * negative entropy of the marginal; report entropy / weighted_entropy     (_ent.py:284-294)
    VQA_TRY(latent_finalize_launch(b.ent_rows, Bn, 1.0f, VQA_W_ENTROPY, VQA_REPORT_ENTROPY, b.loss, b.report, s));
  PH_END(VQA_PH_LOSS);
  if (out) {
    VQA_TRY(copy_out(out->loss, b.loss, sizeof(float), s));
    VQA_TRY(copy_out(out->report, b.report, sizeof(float) * VQA_NUM_REPORT, s));
    VQA_TRY(copy_out(out->att_score, b.att, sizeof(float) * Bn * K, s));
    VQA_TRY(copy_out(out->logit, out_logit, sizeof(float) * Bn * A, s));
    VQA_TRY(copy_out(out->pred, b.pred, sizeof(int) * Bn, s));
    VQA_TRY(copy_out(out->per_sample, b.per_sample, sizeof(float) * VQA_NUM_PER_SAMPLE * Bn, s));
    // heavy_output['condition']: q (base), q_L_ft2 (vqa/model_vlmap_answer2.py:131)
    VQA_TRY(copy_out(out->condition, c.variant == VQA_VARIANT_VLMAP_ANSWER2 ? b.qp_f32 : q, sizeof(float) * BL, s));
    VQA_TRY(copy_out(out->pooled, b.pooled, sizeof(float) * Bn * Pd, s));
  }
  if (defer) h->outputs_pending = true;   // joined by vqa_backward (or by the next vqa_forward / vqa_sync_outputs)
  h->fwd_valid = true;
  h->last_batch = Bn;
  h->last_T = T;
  h->last_seed = seed;
  h->last_step = step;
  h->last_masks = *masks;
  return VQA_OK;
}

VQA_API VqaStatus vqa_backward(VqaHandle h, const VqaParams* p, const VqaBatch* batch,
                               const VqaParams* g, float loss_scale, void* stream) {
  VQA_TRY(check_ready(h, "vqa_backward"));
  if (!p || !batch || !g) return set_error(VQA_ERR_BAD_ARG, "vqa_backward: null argument");
  if (!h->fwd_valid) return set_error(VQA_ERR_STATE, "vqa_backward: no forward pass to differentiate");
  const VqaConfig& c = h->cfg;
  const int Bn = h->last_batch, T = h->last_T;
  if (batch->batch_size != Bn || batch->q_len_max != T)
    return set_error(VQA_ERR_STATE, "vqa_backward: batch differs from the one given to vqa_forward");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Buffers& b = h->buf;
  const int K = c.K, Dv = c.Dv, D = c.D, L = c.L, J = c.J, A = c.A, W = c.W, Wp = h->Wpad;
  const long long BL = static_cast<long long>(Bn) * L;
  const uint64_t seed = h->last_seed, step = h->last_step;
  const int use_tm = c.variant != VQA_VARIANT_STANDARD;  // every vlmap_answer* variant masks the loss to the train answers
  const long long q_off = T * BL;
  const bool v_full = c.variant == VQA_VARIANT_VLMAP_ANSWER_FULL;
  const bool v_tuned = c.variant == VQA_VARIANT_VLMAP_ANSWER_VQA_ALL || c.variant == VQA_VARIANT_VLMAP_ANSWER_VQA_ALL2;
  const bool v_adapt = c.variant == VQA_VARIANT_VLMAP_ANSWER_ADAPT;
  const int Pd = v_adapt ? D : Dv;

  // Parameter gradients of the head layers (weight-gradient GEMMs, bias / LayerNorm column sums) are needed by nobody
  // before the optimizer: each is enqueued on auxiliary stream 3, ordered after the main stream at that point, and the
  // weight-gradient section (or the end of this function) joins it. Each layer has its own LayerNorm partial buffers.
  cudaStream_t pg = s;
  float* pg_scr = b.scratch;
  bool pg_used = false;
  auto pg_fork = [&]() -> VqaStatus {
    if (h->profile && !h->profile_overlapped) return VQA_OK;   // isolated phase timing: everything stays on s
    VQA_CUDA_CHECK(cudaEventRecord(h->ev_fork[3], s));
    VQA_CUDA_CHECK(cudaStreamWaitEvent(h->aux[3], h->ev_fork[3], 0));
    pg = h->aux[3];
    pg_scr = b.scratch + 4 * b.scratch_floats;
    pg_used = true;
    return VQA_OK;
  };
  // The next batch's feature gather (vqa_prefetch_features) goes out FIRST, beside the M = 512 head kernels below, which
  // leave most SMs idle (CTAs that claim an SM each), instead of beside the BPTT on the 20 SMs its grid leaves free
  // (116 us there, and its 0.15 GB of traffic slows the recurrence: BPTT 188 -> 183 us without it).
  // VQA_PREFETCH_EARLY=0: beside the BPTT as before; =n: n CTAs; =-1: the plain full-width gather kernel.
  static const int pf_early = (getenv("VQA_PREFETCH_EARLY") && *getenv("VQA_PREFETCH_EARLY")) ? atoi(getenv("VQA_PREFETCH_EARLY")) : 64;
  if (pf_early != 0 && h->pf_pending && !(h->profile && !h->profile_overlapped) && gru_pair_supported(Bn, L, h->num_sms) &&
      gru_persistent_supported(Bn, L, c.precision, h->num_sms)) {
    cudaStream_t a4;
    VQA_TRY(fork_stream(h, 4, s, &a4));
    VQA_TRY(launch_pending_prefetch(h, a4, pf_early == 1 ? 64 : pf_early));
  }
  PH_BEGIN(VQA_PH_HEAD_BWD);
  if (v_tuned) {
    // gradients of the two-term loss w.r.t. the word-weight logits (through the min fill in vqa_all) and the tuned ones
    TunedHeadBwd t{};
    t.batch = Bn; t.A = A; t.num_train_answer = c.num_train_answer;
    t.fill_min = c.variant == VQA_VARIANT_VLMAP_ANSWER_VQA_ALL;
    t.logit0 = b.logit; t.l1 = b.logit1; t.total = b.logit_total; t.tuned = b.tuned;
    t.target = batch->answer_target; t.exist = h->last_masks.answer_exist;
    t.grad_scale = loss_scale / static_cast<float>(Bn);
    t.d_logit0_f32 = b.dlogit_f32; t.d_logit0_hi = b.dlogit.hi; t.d_logit0_lo = b.dlogit.lo;
    t.d_tuned_f32 = b.dtuned_f32; t.d_tuned_hi = b.dtuned.hi; t.d_tuned_lo = b.dtuned.lo;
    VQA_TRY(tuned_grad_launch(t, s));
    if (g->tw_w) {
      VQA_TRY(pg_fork());
      VQA_TRY(GemmB(J, A, Bn).a(b.jd, 0, J, true).b(b.dtuned, 0, A, true).f32(g->tw_w, A).run(h, pg));
    }
    if (g->tw_b) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.dtuned_f32, Bn, A, A, g->tw_b, pg_scr, pg));
    }
  } else {
    // d logit = (sigmoid(x) - z) * train_mask / B
    VQA_TRY(bce_grad_launch(Bn, A, c.num_train_answer, use_tm, b.logit, batch->answer_target,
                            loss_scale / static_cast<float>(Bn), b.dlogit_f32, b.dlogit.hi, b.dlogit.lo, s));
  }
  if (g->ans_w) {
    VQA_TRY(pg_fork());
    VQA_TRY(GemmB(J, A, Bn).a(b.jd, 0, J, true).b(b.dlogit, 0, A, true).f32(g->ans_w, A).run(h, pg));
  }
  if (g->ans_b) {
    VQA_TRY(pg_fork());
    VQA_TRY(colsum_launch(b.dlogit_f32, Bn, A, A, g->ans_b, pg_scr, pg));
  }
  // dJd = dlogit Wa^T, then joint_fc's dropout / ReLU / LayerNorm backward (one kernel in bf16 mode: linear_ln.cu)
  {
    RowLnBwd r{};
    r.rows = Bn; r.N = J; r.dout = b.dJ; r.z = b.zj; r.gamma = p->joint_gamma; r.beta = p->joint_beta;
    r.mean = b.lnj_mean; r.rstd = b.lnj_rstd; r.keep = c.keep_joint; r.seed = seed; r.step = step;
    r.stream_id = RNG_STREAM_JOINT; r.dz_f32 = b.dzj_f32; r.dz_hi = b.dzj.hi; r.dz_lo = b.dzj.lo;
    if (g->joint_gamma || g->joint_beta) { r.dgamma_part = b.ln_parts[0][0]; r.dbeta_part = b.ln_parts[0][1]; }
    if (v_tuned) {   // both heads read the same joint: dJd += d tuned Wt^T
      VQA_TRY(GemmB(Bn, J, A).a(b.dlogit, 0, A, false).b(b.w.ans_w, 0, A, false).f32(b.dJ, J).run(h, s));
      VQA_TRY(GemmB(Bn, J, A).a(b.dtuned, 0, A, false).b(b.w.tw_w, 0, A, false).addend(b.dJ, J).f32(b.dJ, J).run(h, s));
      VQA_TRY(row_ln_relu_bwd_launch(r, s));
    } else {
      if (!g->joint_b) r.dz_f32 = nullptr;   // (only the bias gradient reads the fp32 copy)
      VQA_TRY(fc_ln_bwd(h, LL_JOINT_BWD, b.dlogit, A, A, b.w.ans_w, b.dJ, false, r, s));
    }
    if (g->joint_gamma) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.ln_parts[0][0], Bn, J, J, g->joint_gamma, pg_scr, pg));
    }
    if (g->joint_beta) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.ln_parts[0][1], Bn, J, J, g->joint_beta, pg_scr, pg));
    }
  }
  if (g->joint_w) {
    VQA_TRY(pg_fork());
    VQA_TRY(GemmB(L, J, Bn).a(b.x, 0, L, true).b(b.dzj, 0, J, true).f32(g->joint_w, J).run(h, pg));
  }
  if (g->joint_b) {
    VQA_TRY(pg_fork());
    VQA_TRY(colsum_launch(b.dzj_f32, Bn, J, J, g->joint_b, pg_scr, pg));
  }
  // dX = dZj Wj^T ; dHp = dX (.) Hl ; dHl = dX (.) Hp   (noc: dHp = dX, dHl comes from the joint_l branch)
  // (pooled_linear_l's Hadamard / ReLU / LayerNorm backward rides in the epilogue of the dX product; dX itself is
  //  stored for the q_linear_l branch)
  const bool noc = c.variant == VQA_VARIANT_VLMAP_ANSWER_NOC;
  {
    RowLnBwd r{};
    r.rows = Bn; r.N = L; r.dout = b.dX; r.mul = noc ? nullptr : b.hl; r.z = b.zp; r.gamma = p->pl_gamma; r.beta = p->pl_beta;
    r.mean = b.lnp_mean; r.rstd = b.lnp_rstd; r.keep = 1.f; r.dz_f32 = g->pl_b ? b.dzp_f32 : nullptr; r.dz_hi = b.dzp.hi;
    r.dz_lo = b.dzp.lo;
    if (g->pl_gamma || g->pl_beta) { r.dgamma_part = b.ln_parts[2][0]; r.dbeta_part = b.ln_parts[2][1]; }
    VQA_TRY(fc_ln_bwd(h, LL_PL_BWD, b.dzj, J, J, b.w.joint_w, b.dX, true, r, s));
    if (g->pl_gamma) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.ln_parts[2][0], Bn, L, L, g->pl_gamma, pg_scr, pg));
    }
    if (g->pl_beta) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.ln_parts[2][1], Bn, L, L, g->pl_beta, pg_scr, pg));
    }
  }
  const float* d_hl_src = b.dX;
  if (noc) {
    // joint_l branch: dJl = dlogit Wal^T -> dropout / ReLU / LN backward -> dHl = dZjl Wjl^T
    if (g->al_w) {
      VQA_TRY(pg_fork());
      VQA_TRY(GemmB(J, A, Bn).a(b.jdl, 0, J, true).b(b.dlogit, 0, A, true).f32(g->al_w, A).run(h, pg));
    }
    if (g->al_b) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.dlogit_f32, Bn, A, A, g->al_b, pg_scr, pg));
    }
    VQA_TRY(GemmB(Bn, J, A).a(b.dlogit, 0, A, false).b(b.w.al_w, 0, A, false).f32(b.dJl, J).run(h, s));
    RowLnBwd r{};
    r.rows = Bn; r.N = J; r.dout = b.dJl; r.z = b.zjl; r.gamma = p->jl_gamma; r.beta = p->jl_beta;
    r.mean = b.lnjl_mean; r.rstd = b.lnjl_rstd; r.keep = c.keep_joint; r.seed = seed; r.step = step;
    r.stream_id = RNG_STREAM_JOINT_L; r.dz_f32 = b.dzjl_f32; r.dz_hi = b.dzjl.hi; r.dz_lo = b.dzjl.lo;
    if (g->jl_gamma || g->jl_beta) { r.dgamma_part = b.ln_parts[1][0]; r.dbeta_part = b.ln_parts[1][1]; }
    VQA_TRY(row_ln_relu_bwd_launch(r, s));
    if (g->jl_gamma) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.ln_parts[1][0], Bn, J, J, g->jl_gamma, pg_scr, pg));
    }
    if (g->jl_beta) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.ln_parts[1][1], Bn, J, J, g->jl_beta, pg_scr, pg));
    }
    if (g->jl_w) {
      VQA_TRY(pg_fork());
      VQA_TRY(GemmB(L, J, Bn).a(b.hl_op, 0, L, true).b(b.dzjl, 0, J, true).f32(g->jl_w, J).run(h, pg));
    }
    if (g->jl_b) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.dzjl_f32, Bn, J, J, g->jl_b, pg_scr, pg));
    }
    VQA_TRY(GemmB(Bn, L, J).a(b.dzjl, 0, J, false).b(b.w.jl_w, 0, J, false).f32(b.dXl, L).run(h, s));
    d_hl_src = b.dXl;
  }
  // The q_linear_l branch (LN / ReLU backward + dq = dZl Wl^T) is independent of the pooled_linear_l branch: when it is
  // the plain frozen layer (no extra question layer, no parameter gradients, no regulariser) it runs on auxiliary
  // stream 2 beside the pooled branch and the attention backward; the BPTT, which consumes dq, joins it.
  const bool has_qp_early = c.variant == VQA_VARIANT_VLMAP_ANSWER2 || c.variant == VQA_VARIANT_VLMAP_ANSWER_NO_NOISE || v_full;
  const bool fork_ql = !has_qp_early && !noc && c.variant != VQA_VARIANT_VLMAP_ANSWER_ENT && !g->ql_w && !g->ql_b &&
                       !g->ql_gamma && !g->ql_beta && !(h->profile && !h->profile_overlapped);
  if (fork_ql) {
    cudaStream_t sq;
    VQA_TRY(fork_stream(h, 2, s, &sq));
    RowLnBwd r{};
    r.rows = Bn; r.N = L; r.dout = b.dX; r.mul = b.hp; r.z = b.zl; r.gamma = p->ql_gamma; r.beta = p->ql_beta;
    r.mean = b.lnl_mean; r.rstd = b.lnl_rstd; r.keep = 1.f; r.dz_f32 = b.dzl_f32; r.dz_hi = b.dzl.hi; r.dz_lo = b.dzl.lo;
    VQA_TRY(row_ln_relu_bwd_launch(r, sq));
    VQA_TRY(GemmB(Bn, L, L).a(b.dzl, 0, L, false).b(b.w.ql_w, 0, L, false).f32(b.dq, L).stable_b().run(h, sq));
  }
  const bool v_ent = c.variant == VQA_VARIANT_VLMAP_ANSWER_ENT;
  if (v_ent) {
    // regulariser backward: d marginal -> d tile logits -> tiled joint head (frozen: data gradients only) -> dHl
    const int M = h->M, BM = Bn * M;
    EntMarginalBwd e{};
    e.batch = Bn; e.M = M; e.A = A; e.num_train_answer = c.num_train_answer; e.logit2 = b.logit2;
    e.exist = h->last_masks.answer_exist; e.marg = b.marg; e.row_max = b.row_max; e.row_inv = b.row_inv;
    e.scale = VQA_W_ENTROPY * loss_scale / static_cast<float>(Bn); e.d_hi = b.dl2.hi; e.d_lo = b.dl2.lo;
    VQA_TRY(ent_marginal_bwd_launch(e, s));
    VQA_TRY(GemmB(BM, J, A).a(b.dl2, 0, A, false).b(b.w.ans_w, 0, A, false).f32(b.dJ2, J).run(h, s));
    SlabLnBwd r{};
    r.batch = Bn; r.K = M; r.D = J; r.z = b.z2; r.gamma = p->joint_gamma; r.beta = p->joint_beta;
    r.mean = b.ln2_mean; r.rstd = b.ln2_rstd; r.dz_hi = b.dz2.hi; r.dz_lo = b.dz2.lo; r.part = nullptr;
    r.dout = b.dJ2; r.thr = keep_threshold(c.keep_joint); r.inv_keep = 1.f / c.keep_joint; r.seed = seed; r.step = step;
    r.site = RNG_STREAM_ENT;
    VQA_TRY(slab_ln_relu_bwd_launch(r, c.precision, s));
    VQA_TRY(GemmB(BM, L, J).a(b.dz2, 0, J, false).b(b.w.joint_w, 0, J, false).f32(b.dX2, L).run(h, s));
    EntDhl d{};
    d.batch = Bn; d.M = M; d.L = L; d.dX = b.dX; d.dX2 = b.dX2; d.hp = b.hp; d.out = b.dhl_ent;
    VQA_TRY(ent_dhl_launch(d, s));
    d_hl_src = b.dhl_ent;
  }
  if (!fork_ql) {
    RowLnBwd r{};
    r.rows = Bn; r.N = L; r.dout = d_hl_src; r.mul = (noc || v_ent) ? nullptr : b.hp; r.z = b.zl; r.gamma = p->ql_gamma; r.beta = p->ql_beta;
    r.mean = b.lnl_mean; r.rstd = b.lnl_rstd; r.keep = 1.f; r.dz_f32 = b.dzl_f32; r.dz_hi = b.dzl.hi;
    r.dz_lo = b.dzl.lo;
    if (g->ql_gamma || g->ql_beta) { r.dgamma_part = b.ln_parts[3][0]; r.dbeta_part = b.ln_parts[3][1]; }
    VQA_TRY(row_ln_relu_bwd_launch(r, s));
    if (g->ql_gamma) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.ln_parts[3][0], Bn, L, L, g->ql_gamma, pg_scr, pg));
    }
    if (g->ql_beta) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.ln_parts[3][1], Bn, L, L, g->ql_beta, pg_scr, pg));
    }
  }
  if (g->pl_w) {
    VQA_TRY(pg_fork());
    VQA_TRY(GemmB(Pd, L, Bn).a(b.pooled_op, 0, Pd, true).b(b.dzp, 0, L, true).f32(g->pl_w, L).run(h, pg));
  }
  if (g->pl_b) {
    VQA_TRY(pg_fork());
    VQA_TRY(colsum_launch(b.dzp_f32, Bn, L, L, g->pl_b, pg_scr, pg));
  }
  const bool has_qp = c.variant == VQA_VARIANT_VLMAP_ANSWER2 || c.variant == VQA_VARIANT_VLMAP_ANSWER_NO_NOISE || v_full;
  {
    const Planes& ql_in = has_qp ? b.qp : b.h;   // q_linear_l reads the extra layer's output in those variants
    if (g->ql_w) {
      VQA_TRY(pg_fork());
      VQA_TRY(GemmB(L, L, Bn).a(ql_in, has_qp ? 0 : q_off, L, true).b(b.dzl, 0, L, true).f32(g->ql_w, L).run(h, pg));
    }
  }
  if (g->ql_b) {
    VQA_TRY(pg_fork());
    VQA_TRY(colsum_launch(b.dzl_f32, Bn, L, L, g->ql_b, pg_scr, pg));
  }
  // dP = dZp Wp^T ; dq = dZl Wl^T
  VQA_TRY(GemmB(Bn, Pd, L).a(b.dzp, 0, L, false).b(b.w.pl_w, 0, L, false).f32(b.dP, Pd).stable_b().run(h, s));
  if (fork_ql) {
    // dq is being produced on auxiliary stream 2
  } else if (!has_qp) {
    VQA_TRY(GemmB(Bn, L, L).a(b.dzl, 0, L, false).b(b.w.ql_w, 0, L, false).f32(b.dq, L).run(h, s));
  } else if (c.variant == VQA_VARIANT_VLMAP_ANSWER2) {
    // d(q_L_ft2) -> tanh / LayerNorm backward -> parameter gradients of q_L_ft2 -> dq
    VQA_TRY(GemmB(Bn, L, L).a(b.dzl, 0, L, false).b(b.w.ql_w, 0, L, false).f32(b.dqp, L).run(h, s));
    RowLnBwd r{};
    r.rows = Bn; r.N = L; r.dout = b.dqp; r.z = b.zqp; r.gamma = p->qp_gamma; r.beta = p->qp_beta;
    r.mean = b.lnqp_mean; r.rstd = b.lnqp_rstd; r.keep = 1.f; r.act = 1; r.dz_f32 = b.dzqp_f32;
    r.dz_hi = b.dzqp.hi; r.dz_lo = b.dzqp.lo;
    if (g->qp_gamma || g->qp_beta) { r.dgamma_part = b.ln_parts[4][0]; r.dbeta_part = b.ln_parts[4][1]; }
    VQA_TRY(row_ln_relu_bwd_launch(r, s));
    if (g->qp_gamma) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.ln_parts[4][0], Bn, L, L, g->qp_gamma, pg_scr, pg));
    }
    if (g->qp_beta) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.ln_parts[4][1], Bn, L, L, g->qp_beta, pg_scr, pg));
    }
  } else if (v_full) {
    // d(q_L_mean_noise) -> d mean (+ KL), d log_sigma_sq (+ KL)                        (_full.py:132-134, 272-276)
    VQA_TRY(GemmB(Bn, L, L).a(b.dzl, 0, L, false).b(b.w.ql_w, 0, L, false).f32(b.dqp, L).run(h, s));
    ReparamBwd r{};
    r.batch = Bn; r.L = L; r.d_out = b.dqp; r.mean = b.qp_f32; r.lss = b.lss; r.seed = seed; r.step = step;
    r.kl_scale = VQA_LATENT_LOSS_WEIGHT * loss_scale / static_cast<float>(Bn);
    r.d_mean_f32 = b.dzqp_f32; r.d_mean_hi = b.dzqp.hi; r.d_mean_lo = b.dzqp.lo;
    r.d_lss_f32 = b.dlss_f32; r.d_lss_hi = b.dlss.hi; r.d_lss_lo = b.dlss.lo;
    VQA_TRY(reparam_bwd_launch(r, s));
  } else {
    // q_L_mean is linear: the gradient of its output IS the gradient of its pre-activation
    VQA_TRY(GemmB(Bn, L, L).a(b.dzl, 0, L, false).b(b.w.ql_w, 0, L, false).f32(b.dzqp_f32, L).planes(b.dzqp, 0, L)
                .run(h, s));
  }
  if (has_qp) {
    if (g->qp_w) {
      VQA_TRY(pg_fork());
      VQA_TRY(GemmB(L, L, Bn).a(b.h, q_off, L, true).b(b.dzqp, 0, L, true).f32(g->qp_w, L).run(h, pg));
    }
    if (g->qp_b) {
      VQA_TRY(pg_fork());
      VQA_TRY(colsum_launch(b.dzqp_f32, Bn, L, L, g->qp_b, pg_scr, pg));
    }
    VQA_TRY(GemmB(Bn, L, L).a(b.dzqp, 0, L, false).b(b.w.qp_w, 0, L, false).f32(b.dq, L).run(h, s));
    if (v_full) {
      if (g->qs_w) {
        VQA_TRY(pg_fork());
        VQA_TRY(GemmB(L, L, Bn).a(b.h, q_off, L, true).b(b.dlss, 0, L, true).f32(g->qs_w, L).run(h, pg));
      }
      if (g->qs_b) {
        VQA_TRY(pg_fork());
        VQA_TRY(colsum_launch(b.dlss_f32, Bn, L, L, g->qs_b, pg_scr, pg));
      }
      VQA_TRY(GemmB(Bn, L, L).a(b.dlss, 0, L, false).b(b.w.qs_w, 0, L, false).addend(b.dq, L).f32(b.dq, L).run(h, s));
    }
  }
  PH_END(VQA_PH_HEAD_BWD);
  PH_BEGIN(VQA_PH_ATTN_BWD);
  // attention block backward
  static const bool qv_in_attn = getenv("VQA_QV_BWD_IN_ATTN") == nullptr || atoi(getenv("VQA_QV_BWD_IN_ATTN")) != 0;
  {
    VqaAttnBwd a{};
    a.batch = Bn; a.z = b.z; a.gamma = p->v_gamma; a.beta = p->v_beta; a.hq = b.hq; a.att_w = p->att_w;
    a.nbox = b.nbox; a.v_hi = v_adapt ? b.va.hi : b.v.hi; a.v_lo = v_adapt ? b.va.lo : b.v.lo;
    a.seed = seed; a.step = step; a.att = b.att;
    a.ln_mean = b.lnv_mean; a.ln_rstd = b.lnv_rstd; a.d_pooled = b.dP; a.dz_hi = b.dzv.hi;
    a.dz_lo = b.dzv.lo; a.d_hq = b.dhq; a.d_att_w = g->att_w; a.d_att_b = g->att_b;
    a.d_gamma = g->v_gamma; a.d_beta = g->v_beta; a.d_bias = g->v_b;
    a.keep_bits = use_keep_bits(c, K, D) ? b.att_bits : nullptr;
    const bool side = !(h->profile && !h->profile_overlapped);
    // q_linear_v's ReLU / LayerNorm backward rides in the same kernel (one CTA holds the whole d_hq row of its sample)
    AttnQvBwd qv{};
    if (qv_in_attn) {
      qv.z = b.zq; qv.gamma = p->qv_gamma; qv.mean = b.lnq_mean; qv.rstd = b.lnq_rstd;
      qv.dz_hi = b.dzq.hi; qv.dz_lo = b.dzq.lo; qv.dz_f32 = b.dzq_f32;
      if (g->qv_gamma || g->qv_beta) { qv.dgamma_part = b.ln_part_g; qv.dbeta_part = b.ln_part_b; }
    }
    VQA_TRY(attn_bwd_launch(a, K, D, Pd, c.precision, c.keep_att, b.attn_part, s, side ? h->aux[3] : nullptr,
                            side ? h->ev_fork[3] : nullptr, qv_in_attn ? &qv : nullptr));
  }
  if (v_adapt) {
    // d v_adapt[k, :] = a_k dP -> ReLU / LayerNorm(K*D) backward -> dZa; parameter gradients of v_adapt
    SlabLnBwd r{};
    r.batch = Bn; r.K = K; r.D = D; r.z = b.za; r.gamma = p->va_gamma; r.beta = p->va_beta;
    r.mean = b.lnva_mean; r.rstd = b.lnva_rstd; r.att = b.att; r.d_pooled = b.dP;
    r.dz_hi = b.dza.hi; r.dz_lo = b.dza.lo; r.part = b.va_part;
    r.dout = nullptr; r.thr = 65536u; r.inv_keep = 1.f;
    VQA_TRY(slab_ln_relu_bwd_launch(r, c.precision, s));
    if (g->va_gamma) VQA_TRY(colsum_launch(b.va_part, Bn, D, 3 * D, g->va_gamma, b.scratch, s));
    if (g->va_beta) VQA_TRY(colsum_launch(b.va_part + D, Bn, D, 3 * D, g->va_beta, b.scratch, s));
    if (g->va_b) VQA_TRY(colsum_launch(b.va_part + 2 * D, Bn, D, 3 * D, g->va_b, b.scratch, s));
    if (g->va_w) VQA_TRY(GemmB(Dv, D, Bn * K).a(b.v, 0, Dv, true).b(b.dza, 0, D, true).f32(g->va_w, D).run(h, s));
  }
  PH_END(VQA_PH_ATTN_BWD);
  if (fork_ql) VQA_TRY(join_stream(h, 2, s));   // dq (q_linear_l branch) is needed by the BPTT below
  // data-parallel runs take dWv here, ahead of the BPTT, so that its all-reduce can overlap the recurrent kernels
  // in-library gradient exchange (vqa_set_gradient_allreduce): the slice that is final before the BPTT is reduced under it
  const bool ar_on = h->ar.mc != nullptr && h->ar.world > 1 && g->v_w != nullptr &&
                     g->v_w >= h->ar.local && g->v_w < h->ar.local + h->ar.n_total;
  const bool early = h->early_grads && !(h->profile && !h->profile_overlapped);
  // one exchange of gradient floats [off, off + cnt) on barrier channel ch (0 = the sequential channel)
  auto ar_range = [&](long long off, long long cnt, int ch, bool exclusive, int ctas, cudaStream_t st) -> VqaStatus {
    if (cnt <= 0) return VQA_OK;
    unsigned int* ft = ch == 0 ? &h->ar.flag_total : &h->ar.ch_flag_total[ch];
    unsigned int* gt = ch == 0 ? &h->ar.grid_total : &h->ar.ch_grid_total[ch];
    *ft += static_cast<unsigned int>(h->ar.world);
    return multimem_allreduce_sync_launch(h->ar.mc + off, cnt, h->ar.rank, h->ar.world, h->ar.mc_flags + 2 * ch,
                                          h->ar.my_flags + 2 * ch, h->ar.grid_ctr + ch, *ft, gt, exclusive, ctas, st);
  };
  // Branch-wise exchange (default for the in-library collective): the small non-GRU slice goes out under the BPTT on the
  // TPCs the recurrent grid leaves idle; in the weight-gradient section every branch's gradients are exchanged as soon as
  // that branch is done, on its own barrier channel, beside the GEMMs still running -- only the exchange of the branch
  // that finishes last is exposed. Needs dWv at the head of the buffer (ParamStore puts it there).
  static const bool ar_branch_env = getenv("VQA_DP_BRANCH") != nullptr && atoi(getenv("VQA_DP_BRANCH")) != 0;   // (opt-in until measured on >= 2 GPUs)
  const long long vw_floats = ((static_cast<long long>(Dv) * D + 63) / 64) * 64;
  const bool ar_branch = ar_on && !early && ar_branch_env && g->v_w == h->ar.local && g->embed && g->gru_gates_w &&
                         g->embed == h->ar.local + h->ar.n_early && vw_floats <= h->ar.n_early &&
                         !(h->profile && !h->profile_overlapped);
  if (early && g->v_w)
    VQA_TRY(GemmB(Dv, D, Bn * K).a(b.v, 0, Dv, true).b(b.dzv, 0, D, true).f32(g->v_w, D).run(h, s));
  PH_BEGIN(VQA_PH_QV_BWD);
  // q_linear_v backward. Only the data gradient dq2 = dZqv Wqv^T is on the critical path (the BPTT kernel adds it to
  // dq itself); the parameter gradients go to auxiliary stream 3, which the weight-gradient section joins later.
  {
    RowLnBwd r{};
    r.rows = Bn; r.N = D; r.dout = b.dhq; r.z = b.zq; r.gamma = p->qv_gamma; r.beta = p->qv_beta;
    r.mean = b.lnq_mean; r.rstd = b.lnq_rstd; r.keep = 1.f; r.dz_f32 = b.dzq_f32; r.dz_hi = b.dzq.hi;
    r.dz_lo = b.dzq.lo;
    if (g->qv_gamma || g->qv_beta) { r.dgamma_part = b.ln_part_g; r.dbeta_part = b.ln_part_b; }
    if (!qv_in_attn) VQA_TRY(row_ln_relu_bwd_launch(r, s));
  }
  {
    cudaStream_t sp = s;
    float* scr = b.scratch;
    const bool side = !(h->profile && !h->profile_overlapped);
    if (side) {
      // aux[3] may still be reducing the attention partials: stream order keeps both correct
      VQA_CUDA_CHECK(cudaEventRecord(h->ev_fork[3], s));
      VQA_CUDA_CHECK(cudaStreamWaitEvent(h->aux[3], h->ev_fork[3], 0));
      sp = h->aux[3];
      scr = b.scratch + 4 * b.scratch_floats;
    }
    if (g->qv_gamma) VQA_TRY(colsum_launch(b.ln_part_g, Bn, D, D, g->qv_gamma, scr, sp));
    if (g->qv_beta) VQA_TRY(colsum_launch(b.ln_part_b, Bn, D, D, g->qv_beta, scr, sp));
    if (g->qv_w) VQA_TRY(GemmB(L, D, Bn).a(b.h, q_off, L, true).b(b.dzq, 0, D, true).f32(g->qv_w, D).run(h, sp));
    if (g->qv_b) VQA_TRY(colsum_launch(b.dzq_f32, Bn, D, D, g->qv_b, scr, sp));
  }
  if (gru_persistent_supported(Bn, L, c.precision, h->num_sms)) {
    VQA_TRY(GemmB(Bn, L, D).a(b.dzq, 0, D, false).b(b.w.qv_w, 0, D, false).f32(b.dq2, L).stable_b().run(h, s));
  } else {  // the per-step fallback kernels take one gradient tensor: accumulate into dq
    VQA_TRY(GemmB(Bn, L, D).a(b.dzq, 0, D, false).b(b.w.qv_w, 0, D, false).addend(b.dq, L).f32(b.dq, L).run(h, s));
  }
  PH_END(VQA_PH_QV_BWD);
  if (early) {
    // every gradient but the embedding's and the GRU's is complete once auxiliary stream 3 has caught up
    VQA_TRY(join_stream(h, 3, s));
    VQA_CUDA_CHECK(cudaEventRecord(h->ev_early, s));
  }
  // dWv = V^T dZv (the largest weight gradient: [Dv, D] over B*K rows). Independent of everything below: it
  // runs on an auxiliary stream next to the GRU weight-gradient GEMMs (after BPTT, whose cooperative grid
  // needs the SMs to itself)
  // (Tried: the first 512 rows of dWv beside the BPTT on the CTA pairs its grid leaves idle, once the gather had moved out
  //  from there -- 8 tiles with K = 18432 each. They outlast the BPTT under its L2 traffic and the weight-gradient section
  //  waits for them: step +20 us. Removed.)
  auto vproj_wgrad = [&](cudaStream_t st) -> VqaStatus {
    if (g->v_w && !early) VQA_TRY(GemmB(Dv, D, Bn * K).a(b.v, 0, Dv, true).b(b.dzv, 0, D, true).f32(g->v_w, D).run(h, st));
    return VQA_OK;
  };
  bool ar_early_done = false, ar_late_done = false, ar_all_done = false;
  // GRU: back-propagation through time from dq
  const bool need_gru = g->gru_gates_w || g->gru_gates_b || g->gru_cand_w || g->gru_cand_b || g->embed;
  if (need_gru) {
    PH_BEGIN(VQA_PH_GRU_BWD);
    float* dh_cur = b.dq;
    int pp = 0;
    const bool persistent = gru_persistent_supported(Bn, L, c.precision, h->num_sms);
    if (persistent) {
      // one cooperative launch: head of step T-1 from dq, then all 2T-1 dependent matmuls
      GruBwdPersistent a{};
      a.B = Bn; a.L = L; a.T = T; a.q_len = batch->q_intseq_len; a.counter = b.gru_counter;
      a.h_f32 = b.h_f32; a.r = b.r; a.u = b.u; a.c = b.c; a.dq = dh_cur; a.dq2 = b.dq2;   // dq (q_linear_l) + dq2 (q_linear_v)
      a.dG_bf = b.dG.hi; a.dC_bf = b.dC.hi; a.bias_part = b.gru_bias_part;
      a.wg_h = b.w.gru_gates_w.hi + static_cast<long long>(W) * 2 * L;
      a.wc_h = b.w.gru_cand_w.hi + static_cast<long long>(W) * L;
      // the next batch's feature gather goes to the SMs the recurrent grid leaves idle: fork BEFORE the launch (so the
      // gather is ordered after the same prefix of this step), enqueue it AFTER (so the cooperative grid is first in line)
      const bool pf = h->pf_pending && !(h->profile && !h->profile_overlapped) && gru_pair_supported(Bn, L, h->num_sms);
      cudaStream_t a4 = s, a5 = s;
      if (pf) VQA_TRY(fork_stream(h, 4, s, &a4));
      const bool ar_early = ar_on && early && h->ar.n_early > 0;
      const bool ar_small = ar_branch && h->ar.n_early > vw_floats;
      if (ar_early || ar_small) VQA_TRY(fork_stream(h, 5, s, &a5));   // forked BEFORE the cooperative launch, enqueued AFTER it
      VQA_TRY(gru_bwd_persistent_launch(a, h->num_sms, s));
      if (ar_small) {   // everything final before the BPTT except dWv (which is taken in the weight-gradient section)
        // (those parameter gradients were enqueued on auxiliary stream 3: the exchange waits for it, the BPTT does not)
        VQA_CUDA_CHECK(cudaEventRecord(h->ev_join[3], h->aux[3]));
        VQA_CUDA_CHECK(cudaStreamWaitEvent(a5, h->ev_join[3], 0));
      }
      if (ar_small)
        VQA_TRY(ar_range(vw_floats, h->ar.n_early - vw_floats, 0, gru_pair_supported(Bn, L, h->num_sms), 0, a5));
      if (ar_early) {
        // ten 2-CTA clusters that each take a whole SM: the TPCs the 64 CTA pairs of the recurrent grid leave free
        h->ar.flag_total += static_cast<unsigned int>(h->ar.world);
        VQA_TRY(multimem_allreduce_sync_launch(h->ar.mc, h->ar.n_early, h->ar.rank, h->ar.world, h->ar.mc_flags, h->ar.my_flags,
                                               h->ar.grid_ctr, h->ar.flag_total, &h->ar.grid_total,
                                               gru_pair_supported(Bn, L, h->num_sms), 0, a5));
        ar_early_done = true;
      }
      if (pf) VQA_TRY(launch_pending_prefetch(h, a4));
      (void)pp;
    } else {
      if (ar_branch && h->ar.n_early > vw_floats) {   // the small slice, beside the per-step kernels
        cudaStream_t a5;
        VQA_TRY(fork_stream(h, 5, s, &a5));
        VQA_CUDA_CHECK(cudaEventRecord(h->ev_join[3], h->aux[3]));
        VQA_CUDA_CHECK(cudaStreamWaitEvent(a5, h->ev_join[3], 0));
        VQA_TRY(ar_range(vw_floats, h->ar.n_early - vw_floats, 0, false, 16, a5));
      }
      for (int t = T - 1; t >= 0; --t) {
        VQA_TRY(gru_bwd_update_launch(dh_cur, b.h_f32 + t * BL, b.u + t * BL, b.c + t * BL,
                                      batch->q_intseq_len, t, Bn, L, b.du, b.dh_part, b.dC_f32 + t * BL,
                                      b.dC.hi + t * BL, b.dC.lo ? b.dC.lo + t * BL : nullptr, s));
        VQA_TRY(GemmB(Bn, L, L).a(b.dC, t * BL, L, false)
                    .b(b.w.gru_cand_w, static_cast<long long>(W) * L, L, false).f32(b.dRH, L).run(h, s));
        VQA_TRY(gru_bwd_gates_launch(b.dRH, b.du, b.h_f32 + t * BL, b.r + t * BL, b.u + t * BL,
                                     batch->q_intseq_len, t, Bn, L, b.dh_part, b.dG_f32 + 2 * t * BL,
                                     b.dG.hi + 2 * t * BL, b.dG.lo ? b.dG.lo + 2 * t * BL : nullptr, s));
        float* dh_next = b.dh[pp];
        pp ^= 1;
        VQA_TRY(GemmB(Bn, L, 2 * L).a(b.dG, 2 * t * BL, 2 * L, false)
                    .b(b.w.gru_gates_w, static_cast<long long>(W) * 2 * L, 2 * L, false)
                    .addend(b.dh_part, L).f32(dh_next, L).run(h, s));
        dh_cur = dh_next;
      }
    }
    PH_END(VQA_PH_GRU_BWD);
    const int TB = T * Bn;
    float* scratch1 = b.scratch + b.scratch_floats;       // per-stream column-sum scratch
    float* scratch2 = b.scratch + 2 * b.scratch_floats;
    // the four GRU weight gradients: x-rows (K = T*B, M = W) and h-rows (M = L) of the gate and candidate kernels.
    // Each runs "narrow" (one CTA pair per output tile, no split-K): 32 + 16 + 16 + 8 pairs fit the machine side by
    // side, so on forked streams they overlap instead of queueing four split-K reduction chains
    auto gates_h_wgrad = [&](cudaStream_t st, float* scr) -> VqaStatus {
      if (g->gru_gates_w)
        VQA_TRY(GemmB(L, 2 * L, TB).a(b.h, 0, L, true).b(b.dG, 0, 2 * L, true)
                    .f32(g->gru_gates_w + static_cast<long long>(W) * 2 * L, 2 * L).narrow().run(h, st));
      if (g->gru_gates_b) {
        if (persistent) VQA_TRY(colsum_launch(b.gru_bias_part, gru_bias_part_rows(Bn), 2 * L, 3 * L, g->gru_gates_b, scr, st));
        else VQA_TRY(colsum_launch(b.dG_f32, TB, 2 * L, 2 * L, g->gru_gates_b, scr, st));
      }
      return VQA_OK;
    };
    auto cand_h_wgrad = [&](cudaStream_t st, float* scr) -> VqaStatus {
      if (g->gru_cand_w)
        VQA_TRY(GemmB(L, L, TB).a(b.rh, 0, L, true).b(b.dC, 0, L, true)
                    .f32(g->gru_cand_w + static_cast<long long>(W) * L, L).narrow().run(h, st));
      if (g->gru_cand_b) {
        if (persistent) VQA_TRY(colsum_launch(b.gru_bias_part + 2 * L, gru_bias_part_rows(Bn), L, 3 * L, g->gru_cand_b, scr, st));
        else VQA_TRY(colsum_launch(b.dC_f32, TB, L, L, g->gru_cand_b, scr, st));
      }
      return VQA_OK;
    };
    auto x_wgrad = [&](cudaStream_t st) -> VqaStatus {
      if (g->gru_gates_w)
        VQA_TRY(GemmB(W, 2 * L, TB).a(b.e, 0, Wp, true).b(b.dG, 0, 2 * L, true).f32(g->gru_gates_w, 2 * L).narrow().run(h, st));
      if (g->gru_cand_w)
        VQA_TRY(GemmB(W, L, TB).a(b.e, 0, Wp, true).b(b.dC, 0, L, true).f32(g->gru_cand_w, L).narrow().run(h, st));
      return VQA_OK;
    };
    auto embed_bwd = [&](cudaStream_t st) -> VqaStatus {
      if (g->embed) {
        // dE = dG Wg[:W]^T + dC Wc[:W]^T ; d embed = scatter_add(q_intseq, dE)
        static const int de_bn = getenv("VQA_DE_BN") ? atoi(getenv("VQA_DE_BN")) : 0;
        VQA_TRY(GemmB(TB, W, 2 * L).a(b.dG, 0, 2 * L, false).b(b.w.gru_gates_w, 0, 2 * L, false)
                    .f32(b.dE, Wp).bn(de_bn).run(h, st));
        VQA_TRY(GemmB(TB, W, L).a(b.dC, 0, L, false).b(b.w.gru_cand_w, 0, L, false).addend(b.dE, Wp)
                    .f32(b.dE, Wp).bn(de_bn).run(h, st));
        // what clip_ops.global_norm sees for an IndexedSlices gradient: its rows as they are (vqa_set_embedding_slice_norm)
        if (h->slice_slot)
          VQA_TRY(rows_sumsq_launch(b.dE, TB, W, Wp, h->slice_slot, b.scratch + 5 * b.scratch_floats, st));   // auxiliary stream 4's scratch region: its gather uses none
        VQA_TRY(fill_zero_launch(g->embed, sizeof(float) * c.Vq * W, st));
        VQA_TRY(embed_scatter_add_launch(b.dE, Wp, batch->q_intseq, batch->q_intseq_len, Bn, T, T, W, c.Vq,
                                         g->embed, st));
      }
      return VQA_OK;
    };
    if (h->profile && !h->profile_overlapped) {  // serialised, one phase after the other
      PH_BEGIN(VQA_PH_VPROJ_WGRAD);
      VQA_TRY(vproj_wgrad(s));
      PH_END(VQA_PH_VPROJ_WGRAD);
      PH_BEGIN(VQA_PH_GRU_WGRAD);
      VQA_TRY(gates_h_wgrad(s, b.scratch));
      VQA_TRY(cand_h_wgrad(s, b.scratch));
      VQA_TRY(x_wgrad(s));
      PH_END(VQA_PH_GRU_WGRAD);
      PH_BEGIN(VQA_PH_EMBED_BWD);
      VQA_TRY(embed_bwd(s));
      PH_END(VQA_PH_EMBED_BWD);
    } else {
      // five independent branches (profile mode 2 times the whole section as GRU_WGRAD on the main stream)
      PH_BEGIN(VQA_PH_GRU_WGRAD);
      cudaStream_t a0, a1, a2, a3;
      static const int order_env = getenv("VQA_WGRAD_ORDER") ? atoi(getenv("VQA_WGRAD_ORDER")) : -1;
      // data-parallel runs (in-library exchange, dWv not taken early): the GRU / embedding gradients FIRST, side by side;
      // their all-reduce then runs under the v-projection weight gradient, the largest GEMM of the section, which
      // leaves only the non-GRU slice for the exposed tail. (Single-GPU order: all five side by side.)
      const bool staged = order_env >= 0 ? order_env != 0 : (ar_on && !early && !ar_branch);
      if (staged) {
        VQA_TRY(fork_stream(h, 1, s, &a1));
        VQA_TRY(fork_stream(h, 2, s, &a2));
        VQA_TRY(fork_stream(h, 3, s, &a3));
        VQA_TRY(gates_h_wgrad(a1, scratch1));
        VQA_TRY(cand_h_wgrad(a2, scratch2));
        VQA_TRY(x_wgrad(a3));
        VQA_TRY(embed_bwd(s));
        VQA_TRY(join_stream(h, 1, s));
        VQA_TRY(join_stream(h, 2, s));
        VQA_TRY(join_stream(h, 3, s));   // (also: every head / attention parameter gradient of auxiliary stream 3)
        if (ar_on && !early) {
          cudaStream_t a5;
          VQA_TRY(fork_stream(h, 5, s, &a5));
          h->ar.flag_total += static_cast<unsigned int>(h->ar.world);
          VQA_TRY(multimem_allreduce_sync_launch(h->ar.mc + h->ar.n_early, h->ar.n_total - h->ar.n_early, h->ar.rank,
                                                 h->ar.world, h->ar.mc_flags, h->ar.my_flags, h->ar.grid_ctr,
                                                 h->ar.flag_total, &h->ar.grid_total, false, 32, a5));
          ar_late_done = true;
        }
        VQA_TRY(vproj_wgrad(s));
        if (ar_late_done) VQA_TRY(join_stream(h, 5, s));
      } else {
        if (ar_branch) VQA_TRY(join_stream(h, 5, s));   // the small slice's exchange under the BPTT (its stream is reused below)
        VQA_TRY(fork_stream(h, 0, s, &a0));
        VQA_TRY(fork_stream(h, 1, s, &a1));
        VQA_TRY(fork_stream(h, 2, s, &a2));
        VQA_TRY(fork_stream(h, 3, s, &a3));
        static const int dwv_first = getenv("VQA_WGRAD_DWV_FIRST") ? atoi(getenv("VQA_WGRAD_DWV_FIRST")) : 0;
        if (dwv_first) VQA_TRY(vproj_wgrad(a0));
        VQA_TRY(gates_h_wgrad(a1, scratch1));
        VQA_TRY(cand_h_wgrad(a2, scratch2));
        VQA_TRY(x_wgrad(a3));
        if (!dwv_first) VQA_TRY(vproj_wgrad(a0));
        if (ar_branch) VQA_TRY(ar_range(0, vw_floats, 1, false, 24, a0));            // dWv, as soon as its GEMM is done
        VQA_TRY(embed_bwd(s));
        if (ar_branch) {
          const long long emb_floats = (g->gru_gates_w - g->embed);
          cudaStream_t a5;
          VQA_TRY(fork_stream(h, 5, s, &a5));
          VQA_TRY(ar_range(h->ar.n_early, emb_floats, 2, false, 24, a5));             // the embedding gradient
          VQA_TRY(join_stream(h, 1, s));
          VQA_TRY(join_stream(h, 2, s));
          VQA_TRY(join_stream(h, 3, s));
          // the GRU kernels / biases + the tail slot (the embedding slice norm, written by embed_bwd on this stream)
          VQA_TRY(ar_range(h->ar.n_early + emb_floats, h->ar.n_total - h->ar.n_early - emb_floats, 3, false, 48, s));
          VQA_TRY(join_stream(h, 5, s));
          VQA_TRY(join_stream(h, 0, s));
          ar_all_done = true;
        } else {
        VQA_TRY(join_stream(h, 0, s));
        VQA_TRY(join_stream(h, 1, s));
        VQA_TRY(join_stream(h, 2, s));
        VQA_TRY(join_stream(h, 3, s));
        }
      }
      if (h->prefetched && !h->pf_joined) {   // the background gather
        VQA_TRY(join_stream(h, 4, s));
        h->pf_joined = true;
        h->pf_joined_into = s;
      }
      PH_END(VQA_PH_GRU_WGRAD);
    }
  } else {
    PH_BEGIN(VQA_PH_VPROJ_WGRAD);
    VQA_TRY(vproj_wgrad(s));
    PH_END(VQA_PH_VPROJ_WGRAD);
  }
  if (pg_used && !need_gru) VQA_TRY(join_stream(h, 3, s));   // (the weight-gradient section joins auxiliary stream 3 otherwise)
  if (h->outputs_pending) {   // the forward's deferred loss / metrics / output copies (auxiliary stream 1)
    VQA_TRY(join_stream(h, 1, s));
    h->outputs_pending = false;
  }
  if (ar_on && !ar_all_done) {
    // what has not been exchanged yet, one wide launch on `s` whose own entry / exit barriers make the sums valid when it
    // completes: the non-GRU slice [0, n_early) when the GRU / embedding slice went out under the dWv GEMM (default),
    // the GRU / embedding slice when the early slice went out under the BPTT (vqa_set_early_gradients), else everything
    long long off = 0, cnt = h->ar.n_total;
    if (ar_early_done) {
      VQA_TRY(join_stream(h, 5, s));
      off = h->ar.n_early;
      cnt = h->ar.n_total - off;
    } else if (ar_late_done) {
      cnt = h->ar.n_early;
    }
    if (cnt > 0) {
      h->ar.flag_total += static_cast<unsigned int>(h->ar.world);
      VQA_TRY(multimem_allreduce_sync_launch(h->ar.mc + off, cnt, h->ar.rank, h->ar.world, h->ar.mc_flags, h->ar.my_flags,
                                             h->ar.grid_ctr, h->ar.flag_total, &h->ar.grid_total, false, 0, s));
    }
  }
  return VQA_OK;
}

VQA_API VqaStatus vqa_set_gradient_allreduce(VqaHandle h, void* multicast_base, void* local_base, int64_t n_early,
                                             int64_t n_total, int64_t flags_offset, int32_t rank, int32_t world) {
  VQA_TRY(check_ready(h, "vqa_set_gradient_allreduce"));
  if (!multicast_base) {   // unregister
    h->ar = {};
    return VQA_OK;
  }
  if (!local_base || world <= 0 || rank < 0 || rank >= world || n_total <= 0 || n_early < 0 || n_early > n_total ||
      (n_early & 3) || (n_total & 3) || flags_offset < n_total || (flags_offset & 3) ||
      (reinterpret_cast<uintptr_t>(multicast_base) & 15) || (reinterpret_cast<uintptr_t>(local_base) & 15))
    return set_error(VQA_ERR_BAD_ARG, "vqa_set_gradient_allreduce: bad argument (sizes are floats, multiples of 4; 16-byte aligned buffers)");
  h->ar.mc = static_cast<float*>(multicast_base);
  h->ar.local = static_cast<float*>(local_base);
  h->ar.n_early = n_early;
  h->ar.n_total = n_total;
  h->ar.mc_flags = reinterpret_cast<unsigned int*>(h->ar.mc + flags_offset);
  h->ar.my_flags = reinterpret_cast<unsigned int*>(h->ar.local + flags_offset);
  h->ar.grid_ctr = h->buf.ar_grid_ctr;
  h->ar.flag_total = 0;
  h->ar.grid_total = 0;
  for (int i = 0; i < 4; ++i) h->ar.ch_flag_total[i] = h->ar.ch_grid_total[i] = 0;
  h->ar.rank = rank;
  h->ar.world = world;
  // (the caller zeroes the flag words on every rank and synchronises the ranks before the first vqa_backward)
  VQA_CUDA_CHECK(cudaMemset(h->buf.ar_grid_ctr, 0, 4 * sizeof(unsigned int)));
  return VQA_OK;
}

VQA_API VqaStatus vqa_dropout_mask_site(VqaHandle h, int32_t site, int32_t batch, uint64_t seed, uint64_t step,
                                        uint8_t* mask, void* stream) {
  if (!h || !mask) return set_error(VQA_ERR_BAD_ARG, "vqa_dropout_mask_site: null argument");
  const VqaConfig& c = h->cfg;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (site) {
    case RNG_STREAM_ATT:
      return dropout_mask_launch(mask, static_cast<long long>(batch) * c.K * c.D, c.keep_att, seed, step, RNG_STREAM_ATT, s);
    case RNG_STREAM_ENT:
      return dropout_mask_launch(mask, static_cast<long long>(batch) * h->M * c.J, c.keep_joint, seed, step, RNG_STREAM_ENT, s);
    case RNG_STREAM_JOINT:
    case RNG_STREAM_JOINT_L:
      return dropout_mask_launch(mask, static_cast<long long>(batch) * c.J, c.keep_joint, seed, step,
                                 static_cast<unsigned int>(site), s);
    default:
      return set_error(VQA_ERR_BAD_ARG, "vqa_dropout_mask_site: unknown site %d", site);
  }
}

VQA_API VqaStatus vqa_prefetch_features(VqaHandle h, const VqaFeatureBank* bank, const VqaBatch* batch, void* stream) {
  VQA_TRY(check_ready(h, "vqa_prefetch_features"));
  if (!bank || !batch || !bank->features || !bank->num_boxes || !batch->image_idx)
    return set_error(VQA_ERR_BAD_ARG, "vqa_prefetch_features: null argument");
  const VqaConfig& c = h->cfg;
  if (batch->batch_size < 0 || batch->batch_size > c.B)
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_prefetch_features: batch_size %d (max %d)", batch->batch_size, c.B);
  // only registered here: the next vqa_backward launches the gather beside its cooperative BPTT kernel, on the SMs that
  // grid leaves idle. Launched right away it takes SMs from the latency-bound head kernels and delays the cooperative
  // grid; launched beside the weight-gradient GEMMs it slows those by as much as it saves (both measured).
  VQA_CUDA_CHECK(cudaEventRecord(h->ev_upload, static_cast<cudaStream_t>(stream)));
  h->pf_bank = *bank;
  h->pf_idx = batch->image_idx;
  h->pf_batch = batch->batch_size;
  h->pf_pending = true;
  return VQA_OK;
}

VQA_API VqaStatus vqa_reparam_noise(VqaHandle h, int32_t batch, uint64_t seed, uint64_t step, float* noise,
                                    void* stream) {
  if (!h || !noise) return set_error(VQA_ERR_BAD_ARG, "vqa_reparam_noise: null argument");
  return reparam_noise_launch(noise, static_cast<long long>(batch) * h->cfg.L, seed, step,
                              static_cast<cudaStream_t>(stream));
}

VQA_API VqaStatus vqa_input_error_count(uint32_t* count, int32_t reset) {
  if (!count) return set_error(VQA_ERR_BAD_ARG, "vqa_input_error_count: null argument");
  unsigned int v = 0;
  VQA_TRY(input_error_count(&v, reset != 0));
  *count = v;
  return VQA_OK;
}

VQA_API VqaStatus vqa_set_early_gradients(VqaHandle h, int32_t enable) {
  if (!h) return set_error(VQA_ERR_BAD_ARG, "vqa_set_early_gradients: null handle");
  h->early_grads = enable != 0;
  return VQA_OK;
}

VQA_API VqaStatus vqa_stream_wait_early_gradients(VqaHandle h, void* stream) {
  if (!h) return set_error(VQA_ERR_BAD_ARG, "vqa_stream_wait_early_gradients: null handle");
  if (!h->early_grads) return set_error(VQA_ERR_STATE, "vqa_stream_wait_early_gradients: enable early gradients first");
  VQA_CUDA_CHECK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->ev_early, 0));
  return VQA_OK;
}

VQA_API VqaStatus vqa_dropout_masks(VqaHandle h, int32_t batch, uint64_t seed, uint64_t step,
                                    uint8_t* att_mask, uint8_t* joint_mask, void* stream) {
  if (!h) return set_error(VQA_ERR_BAD_ARG, "vqa_dropout_masks: null handle");
  const VqaConfig& c = h->cfg;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (att_mask)
    VQA_TRY(dropout_mask_launch(att_mask, static_cast<long long>(batch) * c.K * c.D, c.keep_att, seed, step,
                                RNG_STREAM_ATT, s));
  if (joint_mask)
    VQA_TRY(dropout_mask_launch(joint_mask, static_cast<long long>(batch) * c.J, c.keep_joint, seed, step,
                                RNG_STREAM_JOINT, s));
  return VQA_OK;
}

VQA_API VqaStatus vqa_keep_bits(VqaHandle h, int32_t batch, uint64_t seed, uint64_t step, uint8_t* bits, void* stream) {
  if (!h || !bits) return set_error(VQA_ERR_BAD_ARG, "vqa_keep_bits: null argument");
  const VqaConfig& c = h->cfg;
  return keep_bits_launch(bits, static_cast<long long>(batch) * c.K * c.D, c.keep_att, seed, step, RNG_STREAM_ATT,
                          static_cast<cudaStream_t>(stream));
}

VQA_API VqaStatus vqa_peek_activation(VqaHandle h, int32_t which, const void** dev_ptr, uint64_t* bytes) {
  VQA_TRY(check_ready(h, "vqa_peek_activation"));
  if (!dev_ptr || !bytes) return set_error(VQA_ERR_BAD_ARG, "vqa_peek_activation: null argument");
  if (!h->fwd_valid) return set_error(VQA_ERR_STATE, "vqa_peek_activation: no forward pass yet");
  const VqaConfig& c = h->cfg;
  const uint64_t Bn = static_cast<uint64_t>(h->last_batch);
  const Buffers& b = h->buf;
  switch (which) {
    case VQA_ACT_HQ: *dev_ptr = b.hq; *bytes = Bn * c.D * 4; break;
    case VQA_ACT_HL: *dev_ptr = b.hl; *bytes = Bn * c.L * 4; break;
    case VQA_ACT_HP: *dev_ptr = b.hp; *bytes = Bn * c.L * 4; break;
    case VQA_ACT_JD: *dev_ptr = b.jd.hi; *bytes = Bn * c.J * 2; break;
    case VQA_ACT_JDL: *dev_ptr = b.jdl.hi; *bytes = Bn * c.J * 2; break;
    case VQA_ACT_VA: *dev_ptr = b.va.hi; *bytes = Bn * c.K * c.D * 2; break;
    case VQA_ACT_Z: *dev_ptr = b.z; *bytes = Bn * c.K * c.D * (c.precision == VQA_PREC_FP32 ? 4 : 2); break;
    default: return set_error(VQA_ERR_BAD_ARG, "vqa_peek_activation: unknown activation %d", which);
  }
  return VQA_OK;
}

VQA_API VqaStatus vqa_adam_step(VqaHandle h, float* param, const float* grad, float* m, float* v, int64_t n,
                                float lr, float beta1, float beta2, float eps, float clip_norm, int64_t t,
                                float* grad_norm_out, void* stream) {
  VQA_TRY(check_ready(h, "vqa_adam_step"));
  return adam_step_launch(param, grad, m, v, n, lr, beta1, beta2, eps, clip_norm, t, grad_norm_out,
                          h->buf.scratch, h->num_sms, static_cast<cudaStream_t>(stream));
}

VQA_API VqaStatus vqa_adam_step_shadowed(VqaHandle h, const VqaParams* p, float* param, const float* grad, float* m,
                                         float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                                         float clip_norm, int64_t t, float* grad_norm_out, void* stream) {
  VQA_TRY(check_ready(h, "vqa_adam_step_shadowed"));
  if (!p || !param) return set_error(VQA_ERR_BAD_ARG, "vqa_adam_step_shadowed: null argument");
  if (!h->params_ready) return set_error(VQA_ERR_STATE, "vqa_adam_step_shadowed: call vqa_prepare_params first");
  const VqaConfig& c = h->cfg;
  WeightShadows& w = h->buf.w;
  const bool adapt = c.variant == VQA_VARIANT_VLMAP_ANSWER_ADAPT;
  struct Item { const float* src; Planes* dst; long long elems; } items[] = {
      {p->v_w, &w.v_w, 1LL * c.Dv * c.D},
      {p->gru_gates_w, &w.gru_gates_w, 1LL * (c.W + c.L) * 2 * c.L},
      {p->gru_cand_w, &w.gru_cand_w, 1LL * (c.W + c.L) * c.L},
      {p->qv_w, &w.qv_w, 1LL * c.L * c.D},
      {p->pl_w, &w.pl_w, 1LL * (adapt ? c.D : c.Dv) * c.L},
      {p->ql_w, &w.ql_w, 1LL * c.L * c.L},
      {p->joint_w, &w.joint_w, 1LL * c.L * c.J},
      {p->ans_w, &w.ans_w, 1LL * c.J * c.A},
      {p->qp_w, &w.qp_w, 1LL * c.L * c.L},
      {p->jl_w, &w.jl_w, 1LL * c.L * c.J},
      {p->al_w, &w.al_w, 1LL * c.J * c.A},
      {p->qs_w, &w.qs_w, 1LL * c.L * c.L},
      {p->tw_w, &w.tw_w, 1LL * c.J * c.A},
      {p->va_w, &w.va_w, 1LL * c.Dv * c.D},
  };
  AdamShadows tab{};
  std::vector<const Item*> rest;   // more than 8 trainable matrices (model_standard): the others are refreshed after
  bool gru = false;
  for (const Item& it : items) {
    if (!it.src || it.src < param || it.src + it.elems > param + n) continue;   // frozen / absent: unchanged
    const long long off = it.src - param;
    if ((off & 3) || (it.elems & 3)) return set_error(VQA_ERR_BAD_ARG, "vqa_adam_step_shadowed: tensors must start on 16-byte boundaries");
    if (it.src == p->gru_gates_w || it.src == p->gru_cand_w) gru = true;
    if (tab.n < 8) {
      tab.begin4[tab.n] = off >> 2;
      tab.end4[tab.n] = (off + it.elems) >> 2;
      tab.hi[tab.n] = it.dst->hi;
      tab.lo[tab.n] = it.dst->lo;
      ++tab.n;
    } else {
      rest.push_back(&it);
    }
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // the embedding's gradient is an IndexedSlices in the reference: its share of the clip norm is the sum of squares of
  // the slice rows (left by vqa_backward in the slot), not that of the scattered dense gradient
  const float* slice_grad = nullptr;
  long long slice_n = 0;
  if (h->slice_slot && p->embed && p->embed >= param && p->embed + static_cast<long long>(c.Vq) * c.W <= param + n) {
    slice_grad = grad + (p->embed - param);
    slice_n = static_cast<long long>(c.Vq) * c.W;
  }
  const bool tail = h->adam_tail_begin > 0 && h->adam_tail_begin < n && rest.empty() && h->aux_created &&
                    !(h->profile && !h->profile_overlapped);
  if (h->tail_pending) VQA_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_tail, 0));   // (two optimizer steps without a forward between them)
  VQA_TRY(adam_step_launch(param, grad, m, v, n, lr, beta1, beta2, eps, clip_norm, t, grad_norm_out, h->buf.scratch,
                           h->num_sms, s, &tab, slice_grad, slice_n, slice_grad ? h->slice_slot : nullptr,
                           tail ? h->adam_tail_begin : 0, tail ? h->aux[2] : nullptr, tail ? h->ev_fork[2] : nullptr));
  if (tail) {
    VQA_CUDA_CHECK(cudaEventRecord(h->ev_tail, h->aux[2]));
    h->tail_pending = true;
  }
  for (const Item* it : rest)
    VQA_TRY(split_bf16_launch(it->src, 1, it->elems, it->elems, it->dst->hi, it->dst->lo, it->elems, s));
  if (gru && gru_persistent_supported(c.B, c.L, c.precision, h->num_sms)) {
    // only the next forward's recurrent kernel reads the packed weights: repack on auxiliary stream 2, under the
    // first kernels of that forward, which waits for ev_pack right before its GRU launch
    cudaStream_t sp;
    VQA_TRY(fork_stream(h, 2, s, &sp));
    VQA_TRY(gru_pack_weights_launch(w.gru_gates_w.hi + static_cast<long long>(c.W) * 2 * c.L,
                                    w.gru_cand_w.hi + static_cast<long long>(c.W) * c.L, c.L, h->buf.gru_pack, sp));
    if (sp != s) {
      VQA_CUDA_CHECK(cudaEventRecord(h->ev_pack, sp));
      h->pack_pending = true;
    }
  }
  return VQA_OK;
}

VQA_API VqaStatus vqa_set_optimizer_tail(VqaHandle h, int64_t tail_begin) {
  if (!h) return set_error(VQA_ERR_BAD_ARG, "vqa_set_optimizer_tail: null handle");
  if (tail_begin < 0 || (tail_begin & 3)) return set_error(VQA_ERR_BAD_ARG, "vqa_set_optimizer_tail: a non-negative multiple of 4 floats");
  h->adam_tail_begin = tail_begin;
  return VQA_OK;
}

VQA_API VqaStatus vqa_sync_params(VqaHandle h, void* stream) {
  if (!h) return set_error(VQA_ERR_BAD_ARG, "vqa_sync_params: null handle");
  if (h->tail_pending) VQA_CUDA_CHECK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->ev_tail, 0));
  return VQA_OK;
}

VQA_API VqaStatus vqa_set_embedding_slice_norm(VqaHandle h, float* slot) {
  if (!h) return set_error(VQA_ERR_BAD_ARG, "vqa_set_embedding_slice_norm: null handle");
  h->slice_slot = slot;
  return VQA_OK;
}

VQA_API VqaStatus vqa_set_deferred_outputs(VqaHandle h, int32_t enable) {
  if (!h) return set_error(VQA_ERR_BAD_ARG, "vqa_set_deferred_outputs: null handle");
  h->defer_outputs = enable != 0;
  return VQA_OK;
}

VQA_API VqaStatus vqa_sync_outputs(VqaHandle h, void* stream) {
  if (!h) return set_error(VQA_ERR_BAD_ARG, "vqa_sync_outputs: null handle");
  if (h->outputs_pending) {
    VQA_TRY(join_stream(h, 1, static_cast<cudaStream_t>(stream)));
    h->outputs_pending = false;
  }
  return VQA_OK;
}

VQA_API VqaStatus vqa_attn_fwd(VqaHandle h, const VqaAttnFwd* a, void* stream) {
  if (!h || !a) return set_error(VQA_ERR_BAD_ARG, "vqa_attn_fwd: null argument");
  const VqaConfig& c = h->cfg;
  return attn_fwd_launch(*a, c.K, c.D, c.Dv, c.precision, c.keep_att, static_cast<cudaStream_t>(stream));
}

VQA_API VqaStatus vqa_attn_bwd(VqaHandle h, const VqaAttnBwd* a, void* stream) {
  VQA_TRY(check_ready(h, "vqa_attn_bwd"));
  if (!a) return set_error(VQA_ERR_BAD_ARG, "vqa_attn_bwd: null argument");
  const VqaConfig& c = h->cfg;
  if (a->batch > c.B) return set_error(VQA_ERR_BAD_SHAPE, "vqa_attn_bwd: batch exceeds config.B");
  return attn_bwd_launch(*a, c.K, c.D, c.Dv, c.precision, c.keep_att, h->buf.attn_part,
                         static_cast<cudaStream_t>(stream));
}

VQA_API VqaStatus vqa_bce_metrics(VqaHandle h, int32_t batch, const float* logit, const float* target,
                                  const VqaAnswerMasks* masks, float grad_scale, float* loss, float* report,
                                  int32_t* pred, float* per_sample, float* d_logit, void* stream) {
  VQA_TRY(check_ready(h, "vqa_bce_metrics"));
  if (!masks) return set_error(VQA_ERR_BAD_ARG, "vqa_bce_metrics: null masks");
  const VqaConfig& c = h->cfg;
  if (batch > c.B) return set_error(VQA_ERR_BAD_SHAPE, "vqa_bce_metrics: batch exceeds config.B");
  return bce_metrics_launch(batch, c.A, c.num_train_answer, c.variant != VQA_VARIANT_STANDARD, logit, target,
                            *masks, grad_scale, loss, report, pred, per_sample, d_logit, nullptr, nullptr,
                            h->buf.scratch, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
