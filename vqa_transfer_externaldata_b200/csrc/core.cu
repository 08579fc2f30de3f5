// Handle lifecycle, error plumbing, workspace plan and the per-kernel C-ABI entry points.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "handle.h"
#include "internal.h"

namespace vqa {

namespace {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
}  // namespace

VqaStatus set_error(VqaStatus code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

VqaStatus set_cuda_error(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s", static_cast<int>(e),
           cudaGetErrorString(e), what);
  return VQA_ERR_CUDA;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

namespace {

struct Bump {
  uint8_t* base;
  uint64_t off = 0;
  template <typename T>
  T* take(uint64_t count) {
    off = (off + 255) & ~static_cast<uint64_t>(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

}  // namespace

uint64_t plan_workspace(VqaHandle_t* h, uint8_t* base) {
  const VqaConfig& c = h->cfg;
  const uint64_t B = c.B, K = c.K, Dv = c.Dv, D = c.D, L = c.L, J = c.J, A = c.A, T = c.T, W = c.W;
  const uint64_t Wp = h->Wpad;
  const bool two = h->planes == 2;
  Bump a{base};
  Buffers& b = h->buf;
  auto planes = [&](Planes& p, uint64_t n) {
    p.hi = a.take<bf16>(n);
    p.lo = two ? a.take<bf16>(n) : nullptr;
  };
  const bool v_full = c.variant == VQA_VARIANT_VLMAP_ANSWER_FULL;
  const bool v_tuned = c.variant == VQA_VARIANT_VLMAP_ANSWER_VQA_ALL || c.variant == VQA_VARIANT_VLMAP_ANSWER_VQA_ALL2;
  const bool v_adapt = c.variant == VQA_VARIANT_VLMAP_ANSWER_ADAPT;
  const uint64_t Pd = v_adapt ? D : Dv;           // width of the pooled vector
  const uint64_t Pmax = D > Dv ? D : Dv;
  planes(b.w.v_w, Dv * D);
  planes(b.w.gru_gates_w, (W + L) * 2 * L);
  planes(b.w.gru_cand_w, (W + L) * L);
  planes(b.w.qv_w, L * D);
  planes(b.w.pl_w, Pd * L);
  planes(b.w.ql_w, L * L);
  planes(b.w.joint_w, L * J);
  planes(b.w.ans_w, J * A);
  planes(b.w.qp_w, L * L);
  planes(b.w.jl_w, L * J);
  planes(b.w.al_w, J * A);
  planes(b.w.qs_w, v_full ? L * L : 0);
  planes(b.w.tw_w, v_tuned ? J * A : 0);
  planes(b.w.va_w, v_adapt ? Dv * D : 0);

  planes(b.v, B * K * Dv);
  b.nbox = a.take<int>(B);
  planes(b.v_alt, B * K * Dv);
  b.nbox_alt = a.take<int>(B);
  planes(b.e, T * B * Wp);
  b.z = two ? static_cast<void*>(a.take<float>(B * K * D)) : static_cast<void*>(a.take<bf16>(B * K * D));
  b.z_planes = Planes();
  b.lnv_mean = a.take<float>(B);
  b.lnv_rstd = a.take<float>(B);

  b.xg = a.take<float>(T * B * 2 * L);
  b.xc = a.take<float>(T * B * L);
  b.g_pre = a.take<float>(B * 2 * L);
  b.c_pre = a.take<float>(B * L);
  b.h_f32 = a.take<float>((T + 1) * B * L);
  planes(b.h, (T + 1) * B * L);
  planes(b.rh, T * B * L);
  b.r = a.take<float>(T * B * L);
  b.u = a.take<float>(T * B * L);
  b.c = a.take<float>(T * B * L);

  b.zq = a.take<float>(B * D); b.hq = a.take<float>(B * D);
  b.lnq_mean = a.take<float>(B); b.lnq_rstd = a.take<float>(B);
  planes(b.hl_op, B * L);
  b.zjl = a.take<float>(B * J);
  b.lnjl_mean = a.take<float>(B); b.lnjl_rstd = a.take<float>(B);
  planes(b.jdl, B * J);
  b.dJl = a.take<float>(B * J);
  b.dzjl_f32 = a.take<float>(B * J);
  planes(b.dzjl, B * J);
  b.dXl = a.take<float>(B * L);
  b.zqp = a.take<float>(B * L); b.qp_f32 = a.take<float>(B * L);
  planes(b.qp, B * L);
  b.lnqp_mean = a.take<float>(B); b.lnqp_rstd = a.take<float>(B);
  b.dqp = a.take<float>(B * L);
  b.dzqp_f32 = a.take<float>(B * L);
  planes(b.dzqp, B * L);
  b.zl = a.take<float>(B * L); b.hl = a.take<float>(B * L);
  b.lnl_mean = a.take<float>(B); b.lnl_rstd = a.take<float>(B);
  b.lss = a.take<float>(v_full ? B * L : 0);
  b.kl_rows = a.take<float>(v_full ? B : 0);
  b.dlss_f32 = a.take<float>(v_full ? B * L : 0);
  planes(b.dlss, v_full ? B * L : 0);
  b.tuned = a.take<float>(v_tuned ? B * A : 0);
  b.logit1 = a.take<float>(v_tuned ? B * A : 0);
  b.logit_total = a.take<float>(v_tuned ? B * A : 0);
  b.pred_logit = a.take<float>(v_tuned ? B * A : 0);
  b.dtuned_f32 = a.take<float>(v_tuned ? B * A : 0);
  planes(b.dtuned, v_tuned ? B * A : 0);
  const bool v_ent = c.variant == VQA_VARIANT_VLMAP_ANSWER_ENT;
  const uint64_t BM = v_ent ? B * static_cast<uint64_t>(h->M) : 0;
  planes(b.x2, BM * L);
  b.z2 = two ? static_cast<void*>(a.take<float>(BM * J)) : static_cast<void*>(a.take<bf16>(BM * J));
  b.ln2_mean = a.take<float>(v_ent ? B : 0);
  b.ln2_rstd = a.take<float>(v_ent ? B : 0);
  planes(b.jd2, BM * J);
  b.logit2 = a.take<float>(BM * A);
  b.row_max = a.take<float>(BM);
  b.row_inv = a.take<float>(BM);
  b.marg = a.take<float>(v_ent ? B * A : 0);
  b.ent_rows = a.take<float>(v_ent ? B : 0);
  planes(b.dl2, BM * A);
  b.dJ2 = a.take<float>(BM * J);
  planes(b.dz2, BM * J);
  b.dX2 = a.take<float>(BM * L);
  b.dhl_ent = a.take<float>(v_ent ? B * L : 0);
  const uint64_t nza = v_adapt ? B * K * D : 0;
  b.za = two ? static_cast<void*>(a.take<float>(nza)) : static_cast<void*>(a.take<bf16>(nza));
  planes(b.va, nza);
  b.lnva_mean = a.take<float>(v_adapt ? B : 0);
  b.lnva_rstd = a.take<float>(v_adapt ? B : 0);
  planes(b.dza, nza);
  b.va_part = a.take<float>(v_adapt ? B * 3 * D : 0);
  b.att_bits = a.take<unsigned char>((B * K * D + 7) / 8 + 16);
  b.att = a.take<float>(B * K);
  b.pooled = a.take<float>(B * Pmax);
  planes(b.pooled_op, B * Pmax);
  b.zp = a.take<float>(B * L); b.hp = a.take<float>(B * L);
  b.lnp_mean = a.take<float>(B); b.lnp_rstd = a.take<float>(B);
  planes(b.x, B * L);
  b.zj = a.take<float>(B * J);
  b.lnj_mean = a.take<float>(B); b.lnj_rstd = a.take<float>(B);
  planes(b.jd, B * J);
  b.logit = a.take<float>(B * A);
  b.pred = a.take<int>(B);
  b.per_sample = a.take<float>(VQA_NUM_PER_SAMPLE * B);
  b.report = a.take<float>(VQA_NUM_REPORT);
  b.loss = a.take<float>(1);

  b.dlogit_f32 = a.take<float>(B * A);
  planes(b.dlogit, B * A);
  b.dJ = a.take<float>(B * J);
  b.dzj_f32 = a.take<float>(B * J);
  planes(b.dzj, B * J);
  b.dX = a.take<float>(B * L);
  b.dzp_f32 = a.take<float>(B * L);
  planes(b.dzp, B * L);
  b.dzl_f32 = a.take<float>(B * L);
  planes(b.dzl, B * L);
  b.dP = a.take<float>(B * Pmax);
  b.dq = a.take<float>(B * L);
  b.dq2 = a.take<float>(B * L);
  b.dhq = a.take<float>(B * D);
  b.dzq_f32 = a.take<float>(B * D);
  planes(b.dzq, B * D);
  planes(b.dzv, B * K * D);
  b.attn_part = a.take<float>(attn_bwd_partial_floats(static_cast<int>(B), static_cast<int>(D)));
  b.dh[0] = a.take<float>(B * L);
  b.dh[1] = a.take<float>(B * L);
  b.du = a.take<float>(B * L);
  b.dh_part = a.take<float>(B * L);
  b.dRH = a.take<float>(B * L);
  b.dC_f32 = a.take<float>(T * B * L);
  planes(b.dC, T * B * L);
  b.dG_f32 = a.take<float>(T * B * 2 * L);
  planes(b.dG, T * B * 2 * L);
  b.dE = a.take<float>(T * B * Wp);
  uint64_t maxn = D > L ? D : L;
  if (J > maxn) maxn = J;
  b.ln_part_g = a.take<float>(B * maxn);
  b.ln_part_b = a.take<float>(B * maxn);
  for (int i = 0; i < 5; ++i) {   // per-layer partials: their column sums run on an auxiliary stream while the next layer works
    const uint64_t n = (i < 2 ? J : L) * B;
    b.ln_parts[i][0] = a.take<float>(n);
    b.ln_parts[i][1] = a.take<float>(n);
  }
  uint64_t maxc = 3 * L;
  if (A > maxc) maxc = A;
  if (J > maxc) maxc = J;
  if (Dv > maxc) maxc = Dv;
  b.gemm_sem = a.take<unsigned int>(VqaHandle_t::kGemmSemRegions * VqaHandle_t::kGemmSemElems);
  b.ar_grid_ctr = a.take<unsigned int>(4);
  b.gru_counter = a.take<unsigned int>(64);
  b.gru_pack = a.take<bf16>(gru_pack_elems(static_cast<int>(L)));
  b.gru_bias_part = a.take<float>(gru_bias_part_floats(static_cast<int>(B), static_cast<int>(L)));
  b.scratch_floats = 32 * maxc + 16 * B + 4096;
  b.scratch = a.take<float>(b.scratch_floats * (1 + VqaHandle_t::kAux));
  return (a.off + 255) & ~static_cast<uint64_t>(255);
}

}  // namespace vqa

using namespace vqa;

extern "C" {

VQA_API int32_t vqa_abi_version(void) { return 1; }

VQA_API const char* vqa_last_error(void) { return g_err; }

VQA_API VqaStatus vqa_create(const VqaConfig* config, VqaHandle* out) {
  if (!config || !out) return set_error(VQA_ERR_BAD_ARG, "vqa_create: null argument");
  const VqaConfig& c = *config;
  if (c.B <= 0 || c.K <= 0 || c.T <= 0 || c.A <= 0 || c.W <= 0 || c.Vq <= 0)
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_create: non-positive dimension");
  if ((c.Dv & 7) || (c.D & 7) || (c.L & 7) || (c.J & 7) || (c.A & 7))
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_create: Dv, D, L, J, A must be multiples of 8");
  if (c.D > 4096 || c.L > 4096 || c.J > 4096 || c.K > 256 || c.T > 64)
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_create: D, L, J <= 4096; K <= 256; T <= 64");
  if (c.variant < VQA_VARIANT_VLMAP_ANSWER || c.variant >= VQA_NUM_VARIANTS)
    return set_error(VQA_ERR_BAD_ARG, "vqa_create: unknown variant %d", c.variant);
  if ((c.variant == VQA_VARIANT_VLMAP_ANSWER2 || c.variant == VQA_VARIANT_VLMAP_ANSWER_NO_NOISE ||
       c.variant == VQA_VARIANT_VLMAP_ANSWER_FULL) && c.D != c.L)
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_create: the answer2 / no_noise / full variants need D == L (V_DIM == L_DIM as in the reference)");
  if (c.precision != VQA_PREC_BF16 && c.precision != VQA_PREC_FP32)
    return set_error(VQA_ERR_BAD_ARG, "vqa_create: unknown precision %d", c.precision);
  if (!(c.keep_att > 0.f && c.keep_att <= 1.f) || !(c.keep_joint > 0.f && c.keep_joint <= 1.f))
    return set_error(VQA_ERR_BAD_ARG, "vqa_create: keep probabilities must be in (0, 1]");
  if (c.num_marginal < 0 || c.num_marginal > 4096)
    return set_error(VQA_ERR_BAD_ARG, "vqa_create: num_marginal out of range");
  if (c.variant == VQA_VARIANT_VLMAP_ANSWER_ENT && c.A > 4096)
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_create: the ent variant supports A <= 4096");
  if (c.num_train_answer < 0 || c.num_train_answer > c.A)
    return set_error(VQA_ERR_BAD_ARG, "vqa_create: num_train_answer out of range");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(VQA_ERR_NO_DEVICE, "vqa_create: no CUDA device (this library has no CPU path)");
  }
  int dev = 0;
  VQA_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  VQA_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return set_error(VQA_ERR_NO_DEVICE, "vqa_create: device %d is sm_%d%d; this library is sm_100a only",
                     dev, prop.major, prop.minor);

  VqaHandle_t* h = new (std::nothrow) VqaHandle_t();
  if (!h) return set_error(VQA_ERR_BAD_ARG, "vqa_create: out of host memory");
  h->cfg = c;
  h->device = dev;
  h->num_sms = prop.multiProcessorCount;
  h->Wpad = (c.W + 7) & ~7;
  h->planes = c.precision == VQA_PREC_FP32 ? 2 : 1;
  h->M = c.variant == VQA_VARIANT_VLMAP_ANSWER_ENT ? (c.num_marginal > 0 ? c.num_marginal : 200) : 0;
  h->ws = nullptr;
  h->ws_bytes = 0;
  h->params_ready = false;
  h->fwd_valid = false;
  h->profile = false;
  h->profile_overlapped = false;
  h->ev_created = false;
  for (int i = 0; i < VQA_NUM_PHASES; ++i) h->ev_used[i] = false;
  h->aux_created = false;
  for (int i = 0; i < VqaHandle_t::kAux; ++i) {
    if (cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming) != cudaSuccess) {
      delete h;
      return set_error(VQA_ERR_CUDA, "vqa_create: could not create the auxiliary streams");
    }
  }
  h->aux_created = true;
  h->early_grads = false;
  h->ar = {};
  h->prefetched = false;
  h->prefetched_idx = nullptr;
  h->prefetched_batch = 0;
  h->pf_pending = false;
  h->pf_joined = true;
  h->slice_slot = nullptr;
  h->defer_outputs = false;
  h->outputs_pending = false;
  h->pack_pending = false;
  h->adam_tail_begin = 0;
  h->tail_pending = false;
  if (cudaEventCreateWithFlags(&h->ev_tail, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_early, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_prefetch, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_upload, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_pack, cudaEventDisableTiming) != cudaSuccess) {
    delete h;
    return set_error(VQA_ERR_CUDA, "vqa_create: could not create the early-gradient event");
  }
  h->ws_needed = plan_workspace(h, nullptr);
  *out = h;
  return VQA_OK;
}

VQA_API VqaStatus vqa_destroy(VqaHandle h) {
  if (h && h->ev_created)
    for (int i = 0; i < VQA_NUM_PHASES; ++i) {
      cudaEventDestroy(h->ev[i][0]);
      cudaEventDestroy(h->ev[i][1]);
    }
  if (h && h->aux_created) {
    cudaEventDestroy(h->ev_tail);
    cudaEventDestroy(h->ev_early);
    cudaEventDestroy(h->ev_prefetch);
    cudaEventDestroy(h->ev_upload);
    cudaEventDestroy(h->ev_pack);
  }
  if (h && h->aux_created)
    for (int i = 0; i < VqaHandle_t::kAux; ++i) {
      cudaStreamDestroy(h->aux[i]);
      cudaEventDestroy(h->ev_fork[i]);
      cudaEventDestroy(h->ev_join[i]);
    }
  delete h;
  return VQA_OK;
}

VQA_API VqaStatus vqa_workspace_bytes(VqaHandle h, uint64_t* bytes) {
  if (!h || !bytes) return set_error(VQA_ERR_BAD_ARG, "vqa_workspace_bytes: null argument");
  *bytes = h->ws_needed;
  return VQA_OK;
}

VQA_API VqaStatus vqa_set_workspace(VqaHandle h, void* dev_ptr, uint64_t bytes) {
  if (!h || !dev_ptr) return set_error(VQA_ERR_BAD_ARG, "vqa_set_workspace: null argument");
  if (reinterpret_cast<uintptr_t>(dev_ptr) & 255)
    return set_error(VQA_ERR_WORKSPACE, "vqa_set_workspace: pointer must be 256-byte aligned");
  if (bytes < h->ws_needed)
    return set_error(VQA_ERR_WORKSPACE, "vqa_set_workspace: %llu bytes given, %llu needed",
                     static_cast<unsigned long long>(bytes),
                     static_cast<unsigned long long>(h->ws_needed));
  h->ws = dev_ptr;
  h->ws_bytes = bytes;
  plan_workspace(h, static_cast<uint8_t*>(dev_ptr));
  VQA_CUDA_CHECK(cudaMemset(h->buf.gemm_sem, 0,
                            sizeof(unsigned int) * VqaHandle_t::kGemmSemRegions * VqaHandle_t::kGemmSemElems));
  h->gemm_ctx.sem = h->buf.gemm_sem;
  h->gemm_ctx.regions = VqaHandle_t::kGemmSemRegions;
  h->gemm_ctx.region_elems = VqaHandle_t::kGemmSemElems;
  h->gemm_ctx.next_region = 0;
  h->params_ready = false;
  h->fwd_valid = false;
  h->prefetched = false;
  h->pf_pending = false;
  return VQA_OK;
}

VQA_API VqaStatus vqa_gemm(VqaHandle h, const VqaGemmDesc* d, void* stream) {
  if (!h || !d) return set_error(VQA_ERR_BAD_ARG, "vqa_gemm: null argument");
  return gemm_launch(*d, h->num_sms, static_cast<cudaStream_t>(stream), h->ws ? &h->gemm_ctx : nullptr);
}

VQA_API VqaStatus vqa_split_bf16(VqaHandle h, const float* src, int64_t rows, int64_t cols, int64_t ld,
                                 void* hi, void* lo, int64_t ld_out, void* stream) {
  if (!h || !src || !hi) return set_error(VQA_ERR_BAD_ARG, "vqa_split_bf16: null argument");
  return split_bf16_launch(src, rows, cols, ld, static_cast<bf16*>(hi), static_cast<bf16*>(lo), ld_out,
                           static_cast<cudaStream_t>(stream));
}

VQA_API VqaStatus vqa_profile_enable(VqaHandle h, int32_t enable) {
  if (!h) return set_error(VQA_ERR_BAD_ARG, "vqa_profile_enable: null handle");
  if (enable && !h->ev_created) {
    for (int i = 0; i < VQA_NUM_PHASES; ++i) {
      VQA_CUDA_CHECK(cudaEventCreate(&h->ev[i][0]));
      VQA_CUDA_CHECK(cudaEventCreate(&h->ev[i][1]));
    }
    h->ev_created = true;
  }
  h->profile = enable != 0;
  h->profile_overlapped = enable == 2;   // 2: keep the auxiliary-stream forks (see vqa_answer.h)
  for (int i = 0; i < VQA_NUM_PHASES; ++i) h->ev_used[i] = false;
  return VQA_OK;
}

VQA_API VqaStatus vqa_profile_read(VqaHandle h, float* ms) {
  if (!h || !ms) return set_error(VQA_ERR_BAD_ARG, "vqa_profile_read: null argument");
  for (int i = 0; i < VQA_NUM_PHASES; ++i) {
    ms[i] = 0.f;
    if (h->ev_created && h->ev_used[i]) {
      VQA_CUDA_CHECK(cudaEventSynchronize(h->ev[i][1]));
      VQA_CUDA_CHECK(cudaEventElapsedTime(&ms[i], h->ev[i][0], h->ev[i][1]));
    }
  }
  return VQA_OK;
}

VQA_API const char* vqa_phase_name(int32_t phase) {
  static const char* names[VQA_NUM_PHASES] = {
      "gather", "vproj_fwd", "gru_fwd", "qheads_fwd", "attn_fwd", "head_fwd", "loss",
      "head_bwd", "attn_bwd", "qv_bwd", "vproj_wgrad", "gru_bwd", "gru_wgrad", "embed_bwd"};
  return (phase >= 0 && phase < VQA_NUM_PHASES) ? names[phase] : "?";
}

/* CRC-32C (Castagnoli) of a host buffer: the checksum of TFRecord framing (input_ops.py reads the reference's shards) */
VQA_API uint32_t vqa_crc32c(const uint8_t* data_host, uint64_t n) {
  static uint32_t table[8][256];
  static bool ready = false;
  if (!ready) {   // slicing-by-8 tables of the reflected polynomial 0x82F63B78
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 0xFFu];
    ready = true;
  }
  uint32_t c = 0xFFFFFFFFu;
  uint64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint32_t lo, hi;
    memcpy(&lo, data_host + i, 4);
    memcpy(&hi, data_host + i + 4, 4);
    lo ^= c;
    c = table[7][lo & 0xFFu] ^ table[6][(lo >> 8) & 0xFFu] ^ table[5][(lo >> 16) & 0xFFu] ^ table[4][lo >> 24] ^
        table[3][hi & 0xFFu] ^ table[2][(hi >> 8) & 0xFFu] ^ table[1][(hi >> 16) & 0xFFu] ^ table[0][hi >> 24];
  }
  for (; i < n; ++i) c = table[0][(c ^ data_host[i]) & 0xFFu] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

/* number of kernels this library has enqueued in this process (bench.py's gpu_launches) */
VQA_API uint64_t vqa_launch_count(void) { return launch_count(); }

}  // extern "C"
