// Soft-score BCE, its gradient, argmax and the 13 report scalars in one pass over [B, A]
// (vqa/model_vlmap_answer.py:192-288; model_standard.py:283-376 drops the train mask on the loss).
// One CTA per sample: 128-bit loads of logit and target, warp-shuffle reductions; a one-block second
// kernel averages the per-sample values over the batch. Also the dropout-mask materialiser used by
// the parity tests.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "internal.h"
#include "launch.cuh"
#include "philox.cuh"

namespace vqa {

namespace {

constexpr int LOSS_THREADS = 256;
constexpr int NRED = 10;  // per-sample reduced values kept for the batch mean

// slots of the per-sample scratch row
enum {
  S_LOSS_TRAIN = 0, S_LOSS_ALL, S_ALL_SCORE, S_EXIST_SCORE, S_TEST_SCORE, S_TEST_OBJ_SCORE,
  S_TEST_ATTR_SCORE, S_TRAIN_EXIST_SCORE, S_MAX_TRAIN, S_TEST_OBJ_MAX, S_TEST_ATTR_MAX, S_MAX_EXIST,
  S_MAX_TRAIN_EXIST, S_TEST_MAX, S_TEST_MAX_EXIST, S_COUNT
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct LossArgs {
  int A, num_train_answer, use_train_mask;
  const float* logit; const float* target;
  const float* loss_b; int mask_b;   // optional second BCE term (vqa_all / vqa_all2)
  const float* pred_logit;           // optional: argmax input when it is not `logit`
  const float* is_object; const float* is_attribute; const float* answer_exist;
  float grad_scale;
  int* pred; float* per_sample; int batch;
  float* d_f32; bf16* d_hi; bf16* d_lo;
  float* rows;  // [batch, S_COUNT]
};

__global__ void __launch_bounds__(LOSS_THREADS) bce_metrics_kernel(LossArgs a) {
  __shared__ float red[LOSS_THREADS / 32][NRED];
  __shared__ int red_idx[LOSS_THREADS / 32];
  __shared__ float red_val[LOSS_THREADS / 32];
  const int b = blockIdx.x, A = a.A, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_sync();
  const float* x = a.logit + static_cast<long long>(b) * A;
  const float* z = a.target + static_cast<long long>(b) * A;
  const float* xb = a.loss_b ? a.loss_b + static_cast<long long>(b) * A : nullptr;
  const float* xp = a.pred_logit ? a.pred_logit + static_cast<long long>(b) * A : nullptr;
  float l_train = 0.f, l_all = 0.f;
  float best = -CUDART_INF_F;
  int best_i = 0x7fffffff;
  // maxima of target * mask products; products are >= 0 whenever target >= 0, but start from -inf to
  // follow tf.reduce_max literally
  float mx[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) mx[i] = -CUDART_INF_F;

  for (int c = tid * 4; c < A; c += LOSS_THREADS * 4) {
    const float4 xv = *reinterpret_cast<const float4*>(x + c);
    const float4 zv = *reinterpret_cast<const float4*>(z + c);
    const float4 ob = *reinterpret_cast<const float4*>(a.is_object + c);
    const float4 at = *reinterpret_cast<const float4*>(a.is_attribute + c);
    const float4 ex = *reinterpret_cast<const float4*>(a.answer_exist + c);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, zs[4] = {zv.x, zv.y, zv.z, zv.w};
    const float obs[4] = {ob.x, ob.y, ob.z, ob.w}, ats[4] = {at.x, at.y, at.z, at.w};
    const float exs[4] = {ex.x, ex.y, ex.z, ex.w};
    float bs[4] = {0.f, 0.f, 0.f, 0.f}, ps[4] = {xv.x, xv.y, xv.z, xv.w};
    if (xb) {
      const float4 v = *reinterpret_cast<const float4*>(xb + c);
      bs[0] = v.x; bs[1] = v.y; bs[2] = v.z; bs[3] = v.w;
    }
    if (xp) {
      const float4 v = *reinterpret_cast<const float4*>(xp + c);
      ps[0] = v.x; ps[1] = v.y; ps[2] = v.z; ps[3] = v.w;
    }
    float dl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = c + j;
      const float tm = i < a.num_train_answer ? 1.f : 0.f, te = 1.f - tm;
      const float xx = xs[j], zz = zs[j];
      const float e = expf(-fabsf(xx));
      const float l = fmaxf(xx, 0.f) - xx * zz + log1pf(e);
      l_all += l;
      const float lm = a.use_train_mask ? tm : 1.f;
      l_train += l * lm;
      if (xb) {
        const float lb = fmaxf(bs[j], 0.f) - bs[j] * zz + log1pf(expf(-fabsf(bs[j])));
        l_all += lb;
        l_train += lb * (a.mask_b ? tm : 1.f);
      }
      // sigmoid(x) from the same exponential: x >= 0: 1/(1+e), x < 0: e/(1+e)
      const float sg = xx >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
      dl[j] = (sg - zz) * lm * a.grad_scale;
      if (ps[j] > best) {  // ascending i within a thread: strict > keeps the first maximal index
        best = ps[j];
        best_i = i;
      }
      mx[0] = fmaxf(mx[0], zz * tm);
      mx[1] = fmaxf(mx[1], zz * te * obs[j]);
      mx[2] = fmaxf(mx[2], zz * te * ats[j]);
      mx[3] = fmaxf(mx[3], zz * exs[j]);
      mx[4] = fmaxf(mx[4], zz * exs[j] * tm);
      mx[5] = fmaxf(mx[5], zz * te);
      mx[6] = fmaxf(mx[6], zz * exs[j] * te);
    }
    const long long o = static_cast<long long>(b) * A + c;
    if (a.d_f32) *reinterpret_cast<float4*>(a.d_f32 + o) = make_float4(dl[0], dl[1], dl[2], dl[3]);
    if (a.d_hi) {
      const bf16 h0 = __float2bfloat16_rn(dl[0]), h1 = __float2bfloat16_rn(dl[1]),
                 h2 = __float2bfloat16_rn(dl[2]), h3 = __float2bfloat16_rn(dl[3]);
      __nv_bfloat162 p0(h0, h1), p1(h2, h3);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0);
      pk.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(a.d_hi + o) = pk;
      if (a.d_lo) {
        __nv_bfloat162 q0(__float2bfloat16_rn(dl[0] - __bfloat162float(h0)),
                          __float2bfloat16_rn(dl[1] - __bfloat162float(h1)));
        __nv_bfloat162 q1(__float2bfloat16_rn(dl[2] - __bfloat162float(h2)),
                          __float2bfloat16_rn(dl[3] - __bfloat162float(h3)));
        pk.x = *reinterpret_cast<uint32_t*>(&q0);
        pk.y = *reinterpret_cast<uint32_t*>(&q1);
        *reinterpret_cast<uint2*>(a.d_lo + o) = pk;
      }
    }
  }
  // argmax: larger value wins, ties -> lower index (tf.argmax returns the first maximal index)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) {
      best = ov;
      best_i = oi;
    }
  }
  l_train = wsum(l_train);
  l_all = wsum(l_all);
#pragma unroll
  for (int i = 0; i < 7; ++i) mx[i] = wmax(mx[i]);
  if (lane == 0) {
    red[warp][0] = l_train;
    red[warp][1] = l_all;
#pragma unroll
    for (int i = 0; i < 7; ++i) red[warp][2 + i] = mx[i];
    red_val[warp] = best;
    red_idx[warp] = best_i;
  }
  __syncthreads();
  if (tid == 0) {
    float lt = 0.f, la = 0.f, m[7];
    for (int i = 0; i < 7; ++i) m[i] = -CUDART_INF_F;
    float bv = -CUDART_INF_F;
    int bi = 0x7fffffff;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) {
      lt += red[w][0];
      la += red[w][1];
      for (int i = 0; i < 7; ++i) m[i] = fmaxf(m[i], red[w][2 + i]);
      if (red_val[w] > bv || (red_val[w] == bv && red_idx[w] < bi)) {
        bv = red_val[w];
        bi = red_idx[w];
      }
    }
    if (bi >= A) bi = 0;  // all-NaN row: tf.argmax returns 0
    const float zp = z[bi];
    const float tm = bi < a.num_train_answer ? 1.f : 0.f, te = 1.f - tm;
    const float ob = a.is_object[bi], at = a.is_attribute[bi], ex = a.answer_exist[bi];
    float* r = a.rows + static_cast<long long>(b) * S_COUNT;
    r[S_LOSS_TRAIN] = lt;
    r[S_LOSS_ALL] = la;
    r[S_ALL_SCORE] = zp;
    r[S_EXIST_SCORE] = zp * ex;
    r[S_TEST_SCORE] = zp * te;
    r[S_TEST_OBJ_SCORE] = zp * te * ob;
    r[S_TEST_ATTR_SCORE] = zp * te * at;
    r[S_TRAIN_EXIST_SCORE] = zp * ex * tm;
    r[S_MAX_TRAIN] = m[0];
    r[S_TEST_OBJ_MAX] = m[1];
    r[S_TEST_ATTR_MAX] = m[2];
    r[S_MAX_EXIST] = m[3];
    r[S_MAX_TRAIN_EXIST] = m[4];
    r[S_TEST_MAX] = m[5];
    r[S_TEST_MAX_EXIST] = m[6];
    if (a.pred) a.pred[b] = bi;
    if (a.per_sample) {
      a.per_sample[VQA_PS_ALL_SCORE * a.batch + b] = zp;
      a.per_sample[VQA_PS_MAX_TRAIN_SCORE * a.batch + b] = m[0];
      a.per_sample[VQA_PS_TEST_OBJ_SCORE * a.batch + b] = zp * te * ob;
      a.per_sample[VQA_PS_TEST_OBJ_MAX_SCORE * a.batch + b] = m[1];
      a.per_sample[VQA_PS_TEST_ATTR_SCORE * a.batch + b] = zp * te * at;
      a.per_sample[VQA_PS_TEST_ATTR_MAX_SCORE * a.batch + b] = m[2];
    }
  }
}

__device__ __forceinline__ float normal_ratio(float num, float den) {
  return den == 0.f ? den : num / den;  // tf.where(tf.equal(den, 0), den, num / den)
}

// batch means of the per-sample rows -> report scalars (fixed summation order: deterministic)
__global__ void __launch_bounds__(256) report_finalize_kernel(const float* __restrict__ rows, int batch,
                                                              float* __restrict__ loss,
                                                              float* __restrict__ report) {
  __shared__ float sm[8][S_COUNT];
  pdl_sync();
  float acc[S_COUNT];
#pragma unroll
  for (int i = 0; i < S_COUNT; ++i) acc[i] = 0.f;
  for (int b = threadIdx.x; b < batch; b += 256)
#pragma unroll
    for (int i = 0; i < S_COUNT; ++i) acc[i] += rows[static_cast<long long>(b) * S_COUNT + i];
#pragma unroll
  for (int i = 0; i < S_COUNT; ++i) acc[i] = wsum(acc[i]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < S_COUNT; ++i) sm[warp][i] = acc[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    float m[S_COUNT];
    for (int i = 0; i < S_COUNT; ++i) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += sm[w][i];
      m[i] = s / batch;
    }
    if (loss) loss[0] = m[S_LOSS_TRAIN];
    if (report) {
      report[VQA_REPORT_ANSWER_TRAIN_LOSS] = m[S_LOSS_TRAIN];
      report[VQA_REPORT_ANSWER_REPORT_LOSS] = m[S_LOSS_ALL];
      report[VQA_REPORT_ANSWER_ACC] = m[S_ALL_SCORE];
      report[VQA_REPORT_EXIST_ACC] = m[S_EXIST_SCORE];
      report[VQA_REPORT_TEST_ACC] = m[S_TEST_SCORE];
      report[VQA_REPORT_NORMAL_TEST_ACC] = normal_ratio(m[S_TEST_SCORE], m[S_TEST_MAX]);
      report[VQA_REPORT_NORMAL_TEST_OBJECT_ACC] = normal_ratio(m[S_TEST_OBJ_SCORE], m[S_TEST_OBJ_MAX]);
      report[VQA_REPORT_NORMAL_TEST_ATTRIBUTE_ACC] = normal_ratio(m[S_TEST_ATTR_SCORE], m[S_TEST_ATTR_MAX]);
      report[VQA_REPORT_NORMAL_EXIST_ACC] = normal_ratio(m[S_EXIST_SCORE], m[S_MAX_EXIST]);
      report[VQA_REPORT_NORMAL_TRAIN_EXIST_ACC] = normal_ratio(m[S_TRAIN_EXIST_SCORE], m[S_MAX_TRAIN_EXIST]);
      report[VQA_REPORT_MAX_EXIST_ACC] = m[S_MAX_EXIST];
      report[VQA_REPORT_TEST_MAX_ACC] = m[S_TEST_MAX];
      report[VQA_REPORT_TEST_MAX_EXIST_ACC] = m[S_TEST_MAX_EXIST];
      report[VQA_REPORT_LATENT_LOSS] = 0.f;          // the 'full' variant overwrites these (latent_finalize_kernel)
      report[VQA_REPORT_TRAIN_LATENT_LOSS] = 0.f;
      report[VQA_REPORT_ENTROPY] = 0.f;              // the 'ent' variant overwrites these
      report[VQA_REPORT_WEIGHTED_ENTROPY] = 0.f;
    }
  }
}

// d(loss)/d(logit) = (sigmoid(x) - z) * train_mask * grad_scale   
__global__ void bce_grad_kernel(const float* __restrict__ logit, const float* __restrict__ target,
                                long long total4, int A4, int num_train_answer, int use_train_mask,
                                float grad_scale, float* __restrict__ d_f32, bf16* __restrict__ d_hi,
                                bf16* __restrict__ d_lo) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % A4) * 4;
    const float4 xv = reinterpret_cast<const float4*>(logit)[i];
    const float4 zv = reinterpret_cast<const float4*>(target)[i];
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, zs[4] = {zv.x, zv.y, zv.z, zv.w};
    float dl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lm = (!use_train_mask || (c + j) < num_train_answer) ? 1.f : 0.f;
      const float e = expf(-fabsf(xs[j]));
      const float sg = xs[j] >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
      dl[j] = (sg - zs[j]) * lm * grad_scale;
    }
    if (d_f32) reinterpret_cast<float4*>(d_f32)[i] = make_float4(dl[0], dl[1], dl[2], dl[3]);
    if (d_hi) {
      const bf16 h0 = __float2bfloat16_rn(dl[0]), h1 = __float2bfloat16_rn(dl[1]),
                 h2 = __float2bfloat16_rn(dl[2]), h3 = __float2bfloat16_rn(dl[3]);
      __nv_bfloat162 p0(h0, h1), p1(h2, h3);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0);
      pk.y = *reinterpret_cast<uint32_t*>(&p1);
      reinterpret_cast<uint2*>(d_hi)[i] = pk;
      if (d_lo) {
        __nv_bfloat162 q0(__float2bfloat16_rn(dl[0] - __bfloat162float(h0)),
                          __float2bfloat16_rn(dl[1] - __bfloat162float(h1)));
        __nv_bfloat162 q1(__float2bfloat16_rn(dl[2] - __bfloat162float(h2)),
                          __float2bfloat16_rn(dl[3] - __bfloat162float(h3)));
        pk.x = *reinterpret_cast<uint32_t*>(&q0);
        pk.y = *reinterpret_cast<uint32_t*>(&q1);
        reinterpret_cast<uint2*>(d_lo)[i] = pk;
      }
    }
  }
}

__global__ void dropout_mask_kernel(unsigned char* __restrict__ out, long long groups, uint32_t thr,
                                    unsigned long long seed, unsigned long long step, uint32_t site) {
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < groups;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t bits = thr < 65536u ? philox_keep_bits(philox4x32_10(g, site, seed, step), thr) : 0xFFu;
    uint2 v;
    v.x = (bits & 1u) | ((bits >> 1 & 1u) << 8) | ((bits >> 2 & 1u) << 16) | ((bits >> 3 & 1u) << 24);
    v.y = (bits >> 4 & 1u) | ((bits >> 5 & 1u) << 8) | ((bits >> 6 & 1u) << 16) | ((bits >> 7 & 1u) << 24);
    *reinterpret_cast<uint2*>(out + g * 8) = v;
  }
}

// the same bits as ONE byte per group of 8 elements (bit j = element j of the group): what the attention kernels read
// instead of running the ten Philox rounds themselves, twice per step (forward and backward)
__global__ void keep_bits_kernel(unsigned char* __restrict__ out, long long groups, uint32_t thr,
                                 unsigned long long seed, unsigned long long step, uint32_t site) {
  pdl_sync();
  // four groups per thread -> one 32-bit store
  const long long quads = (groups + 3) >> 2;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < quads;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long g = q * 4 + i;
      if (g < groups) w |= philox_keep_bits(philox4x32_10(g, site, seed, step), thr) << (8 * i);
    }
    if (q * 4 + 3 < groups) {
      *reinterpret_cast<uint32_t*>(out + q * 4) = w;
    } else {
      for (int i = 0; q * 4 + i < groups; ++i) out[q * 4 + i] = static_cast<unsigned char>(w >> (8 * i));
    }
  }
}

}  // namespace

VqaStatus keep_bits_launch(unsigned char* out, long long n, float keep, unsigned long long seed,
                           unsigned long long step, unsigned int stream_id, cudaStream_t s, int max_ctas) {
  if (n == 0) return VQA_OK;
  if (n & 7) return set_error(VQA_ERR_BAD_SHAPE, "keep_bits: n must be a multiple of 8");
  const long long groups = n / 8, quads = (groups + 3) / 4;
  long long g = (quads + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (max_ctas > 0 && g > max_ctas) g = max_ctas;
  launch_pdl(keep_bits_kernel, dim3(static_cast<int>(g)), dim3(256), 0, s, out, groups, keep_threshold(keep), seed, step,
             stream_id);
  VQA_LAUNCH_CHECK("keep_bits");
  return VQA_OK;
}

VqaStatus bce_metrics_launch(int batch, int A, int num_train_answer, int use_train_mask,
                             const float* logit, const float* target, const VqaAnswerMasks& masks,
                             float grad_scale, float* loss, float* report, int* pred,
                             float* per_sample, float* d_logit_f32, bf16* d_hi, bf16* d_lo,
                             float* scratch, cudaStream_t s) {
  if (batch == 0) return VQA_OK;
  if (!logit || !target || !masks.is_object || !masks.is_attribute || !masks.answer_exist || !scratch)
    return set_error(VQA_ERR_BAD_ARG, "vqa_bce_metrics: null argument");
  if (A & 3) return set_error(VQA_ERR_BAD_SHAPE, "vqa_bce_metrics: A must be a multiple of 4");
  LossArgs a;
  a.A = A; a.num_train_answer = num_train_answer; a.use_train_mask = use_train_mask;
  a.logit = logit; a.target = target;
  a.loss_b = nullptr; a.mask_b = 0; a.pred_logit = nullptr;
  a.is_object = masks.is_object; a.is_attribute = masks.is_attribute; a.answer_exist = masks.answer_exist;
  a.grad_scale = grad_scale; a.pred = pred; a.per_sample = per_sample; a.batch = batch;
  a.d_f32 = d_logit_f32; a.d_hi = d_hi; a.d_lo = d_lo; a.rows = scratch;
  launch_pdl(bce_metrics_kernel, dim3(batch), dim3(LOSS_THREADS), 0, s, a);
  VQA_LAUNCH_CHECK("bce_metrics");
  launch_pdl(report_finalize_kernel, dim3(1), dim3(256), 0, s, scratch, batch, loss, report);
  VQA_LAUNCH_CHECK("report_finalize");
  return VQA_OK;
}

VqaStatus bce_metrics2_launch(int batch, int A, int num_train_answer, int use_train_mask, const float* logit,
                              const float* loss_b, int mask_b, const float* pred_logit, const float* target,
                              const VqaAnswerMasks& masks, float* loss, float* report, int* pred, float* per_sample,
                              float* scratch, cudaStream_t s) {
  if (batch == 0) return VQA_OK;
  if (!logit || !target || !masks.is_object || !masks.is_attribute || !masks.answer_exist || !scratch)
    return set_error(VQA_ERR_BAD_ARG, "bce_metrics2: null argument");
  if (A & 3) return set_error(VQA_ERR_BAD_SHAPE, "bce_metrics2: A must be a multiple of 4");
  LossArgs a;
  a.A = A; a.num_train_answer = num_train_answer; a.use_train_mask = use_train_mask;
  a.logit = logit; a.target = target;
  a.loss_b = loss_b; a.mask_b = mask_b; a.pred_logit = pred_logit;
  a.is_object = masks.is_object; a.is_attribute = masks.is_attribute; a.answer_exist = masks.answer_exist;
  a.grad_scale = 0.f; a.pred = pred; a.per_sample = per_sample; a.batch = batch;
  a.d_f32 = nullptr; a.d_hi = nullptr; a.d_lo = nullptr; a.rows = scratch;
  launch_pdl(bce_metrics_kernel, dim3(batch), dim3(LOSS_THREADS), 0, s, a);
  VQA_LAUNCH_CHECK("bce_metrics");
  launch_pdl(report_finalize_kernel, dim3(1), dim3(256), 0, s, scratch, batch, loss, report);
  VQA_LAUNCH_CHECK("report_finalize");
  return VQA_OK;
}

VqaStatus bce_grad_launch(int batch, int A, int num_train_answer, int use_train_mask,
                          const float* logit, const float* target, float grad_scale,
                          float* d_logit_f32, bf16* d_hi, bf16* d_lo, cudaStream_t s) {
  if (batch == 0) return VQA_OK;
  if (A & 3) return set_error(VQA_ERR_BAD_SHAPE, "bce_grad: A must be a multiple of 4");
  const long long total4 = static_cast<long long>(batch) * A / 4;
  long long g = (total4 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  launch_pdl(bce_grad_kernel, dim3(static_cast<int>(g)), dim3(256), 0, s, logit, target, total4, A / 4, num_train_answer,
             use_train_mask, grad_scale, d_logit_f32, d_hi, d_lo);
  VQA_LAUNCH_CHECK("bce_grad");
  return VQA_OK;
}

VqaStatus dropout_mask_launch(unsigned char* out, long long n, float keep, unsigned long long seed,
                              unsigned long long step, unsigned int stream_id, cudaStream_t s) {
  if (n == 0) return VQA_OK;
  if (n & 7) return set_error(VQA_ERR_BAD_SHAPE, "dropout mask: element count must be a multiple of 8");
  const long long groups = n / 8;
  long long g = (groups + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  dropout_mask_kernel<<<static_cast<int>(g), 256, 0, s>>>(out, groups, keep_threshold(keep), seed, step,
                                                          stream_id);
  VQA_LAUNCH_CHECK("dropout_mask");
  return VQA_OK;
}

}  // namespace vqa
