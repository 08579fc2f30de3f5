// Kernels only the later members of the answer-model family need:
//   * vqa_all / vqa_all2 (vqa/model_vlmap_answer_vqa_all.py:188-244, _vqa_all2.py:188-243): the frozen word-weight
//     logits (absent answers filled with the row minimum in vqa_all) plus TunedWordWeightAnswer logits, the two-term
//     BCE and its gradients;
//   * full (vqa/model_vlmap_answer_full.py:124-134, 217-223, 272-276): q_L_mean + N(0,1) * sqrt(exp(q_L_log_sigma_sq)),
//     the KL latent loss (weight 0.1) and their gradients;
//   * adapt (vqa/model_vlmap_answer_adapt.py:132-142): v_adapt = relu(LN_{K,D}(FC(V))) materialised as the pooling
//     operand, and its backward from d v_adapt[k, :] = a_k * dP.
// All are memory-bound row / slab kernels: one CTA per row (sample), 128-bit accesses, warp-shuffle reductions.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "internal.h"
#include "philox.cuh"

namespace vqa {

namespace {

constexpr int VT = 256;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmin(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide reductions over VT threads; `red` holds VT / 32 floats
__device__ __forceinline__ float bsum(float v, float* red) {
  v = wsum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < VT / 32; ++w) t += red[w];
  return t;
}
__device__ __forceinline__ float bmin(float v, float* red) {
  v = wmin(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = CUDART_INF_F;
#pragma unroll
  for (int w = 0; w < VT / 32; ++w) t = fminf(t, red[w]);
  return t;
}

__device__ __forceinline__ float sigmoid_stable(float x) {
  const float e = expf(-fabsf(x));
  return x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
}

__device__ __forceinline__ void store_planes4(bf16* hi, bf16* lo, long long o, const float (&d)[4]) {
  const bf16 h0 = __float2bfloat16_rn(d[0]), h1 = __float2bfloat16_rn(d[1]), h2 = __float2bfloat16_rn(d[2]),
             h3 = __float2bfloat16_rn(d[3]);
  __nv_bfloat162 p0(h0, h1), p1(h2, h3);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&p0);
  pk.y = *reinterpret_cast<uint32_t*>(&p1);
  *reinterpret_cast<uint2*>(hi + o) = pk;
  if (lo) {
    __nv_bfloat162 q0(__float2bfloat16_rn(d[0] - __bfloat162float(h0)), __float2bfloat16_rn(d[1] - __bfloat162float(h1)));
    __nv_bfloat162 q1(__float2bfloat16_rn(d[2] - __bfloat162float(h2)), __float2bfloat16_rn(d[3] - __bfloat162float(h3)));
    pk.x = *reinterpret_cast<uint32_t*>(&q0);
    pk.y = *reinterpret_cast<uint32_t*>(&q1);
    *reinterpret_cast<uint2*>(lo + o) = pk;
  }
}

// 8-element loads (fp32 or bf16 source) and operand-plane stores shared by the slab / tile kernels below
template <typename ZT>
__device__ __forceinline__ void ldz8(const ZT* p, float (&x)[8]);
template <>
__device__ __forceinline__ void ldz8<float>(const float* p, float (&x)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
template <>
__device__ __forceinline__ void ldz8<bf16>(const bf16* p, float (&x)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    x[2 * j] = f.x;
    x[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void ldf8(const float* p, float (&x)[8]) { ldz8<float>(p, x); }
__device__ __forceinline__ void st_planes8(bf16* hi, bf16* lo, long long off, const float (&x)[8]) {
  __nv_bfloat162 h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bf16 a = __float2bfloat16_rn(x[2 * j]), b = __float2bfloat16_rn(x[2 * j + 1]);
    h[j] = __nv_bfloat162(a, b);
    l[j] = __nv_bfloat162(__float2bfloat16_rn(x[2 * j] - __bfloat162float(a)),
                          __float2bfloat16_rn(x[2 * j + 1] - __bfloat162float(b)));
  }
  *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<uint4*>(h);
  if (lo) *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<uint4*>(l);
}

// ------------------------------------------------------------------------------------------------------------------
// vqa_all / vqa_all2: combine the two heads' logits
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VT) tuned_combine_kernel(TunedHeadFwd a) {
  __shared__ float red[VT / 32];
  const int b = blockIdx.x, A = a.A;
  const long long base = static_cast<long long>(b) * A;
  float mn = 0.f;
  if (a.fill_min) {  // tf.reduce_min(logit, axis=1): row minimum of the word-weight logits
    float m = CUDART_INF_F;
    for (int c = threadIdx.x * 4; c < A; c += VT * 4) {
      const float4 x = *reinterpret_cast<const float4*>(a.logit0 + base + c);
      m = fminf(fminf(m, x.x), fminf(fminf(x.y, x.z), x.w));
    }
    mn = bmin(m, red);
  }
  for (int c = threadIdx.x * 4; c < A; c += VT * 4) {
    const float4 xv = *reinterpret_cast<const float4*>(a.logit0 + base + c);
    const float4 tv = *reinterpret_cast<const float4*>(a.tuned + base + c);
    const float4 ev = *reinterpret_cast<const float4*>(a.exist + c);
    const float x[4] = {xv.x, xv.y, xv.z, xv.w}, t[4] = {tv.x, tv.y, tv.z, tv.w}, ex[4] = {ev.x, ev.y, ev.z, ev.w};
    float l1[4], tot[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // logit * exist + min * (1 - exist)  (_vqa_all.py:192-194); vqa_all2 keeps the logits as they are
      l1[j] = a.fill_min ? x[j] * ex[j] + mn * (1.f - ex[j]) : x[j];
      tot[j] = l1[j] + t[j];
      const float tm = (c + j) < a.num_train_answer ? 1.f : 0.f;
      pl[j] = l1[j] * (1.f - tm) + t[j] * tm;   // _vqa_all2.py:241-242
    }
    *reinterpret_cast<float4*>(a.l1 + base + c) = make_float4(l1[0], l1[1], l1[2], l1[3]);
    *reinterpret_cast<float4*>(a.total + base + c) = make_float4(tot[0], tot[1], tot[2], tot[3]);
    if (a.pred_logit) *reinterpret_cast<float4*>(a.pred_logit + base + c) = make_float4(pl[0], pl[1], pl[2], pl[3]);
  }
}

__global__ void __launch_bounds__(VT) tuned_grad_kernel(TunedHeadBwd a) {
  __shared__ float red[VT / 32];
  const int b = blockIdx.x, A = a.A;
  const long long base = static_cast<long long>(b) * A;
  float mn = 0.f, share = 0.f;
  if (a.fill_min) {
    // gradient reaching the row minimum: sum over the absent answers of dL1, shared equally by the minimal entries
    // (tf.reduce_min's gradient, math_grad._MinOrMaxGrad)
    float m = CUDART_INF_F;
    for (int c = threadIdx.x * 4; c < A; c += VT * 4) {
      const float4 x = *reinterpret_cast<const float4*>(a.logit0 + base + c);
      m = fminf(fminf(m, x.x), fminf(fminf(x.y, x.z), x.w));
    }
    mn = bmin(m, red);
    float dmin = 0.f, cnt = 0.f;
    for (int c = threadIdx.x * 4; c < A; c += VT * 4) {
      const float4 xv = *reinterpret_cast<const float4*>(a.logit0 + base + c);
      const float4 lv = *reinterpret_cast<const float4*>(a.l1 + base + c);
      const float4 tv = *reinterpret_cast<const float4*>(a.total + base + c);
      const float4 zv = *reinterpret_cast<const float4*>(a.target + base + c);
      const float4 ev = *reinterpret_cast<const float4*>(a.exist + c);
      const float x[4] = {xv.x, xv.y, xv.z, xv.w}, l1[4] = {lv.x, lv.y, lv.z, lv.w}, tt[4] = {tv.x, tv.y, tv.z, tv.w};
      const float z[4] = {zv.x, zv.y, zv.z, zv.w}, ex[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float tm = (c + j) < a.num_train_answer ? 1.f : 0.f;
        const float dl1 = ((sigmoid_stable(tt[j]) - z[j]) + (sigmoid_stable(l1[j]) - z[j])) * tm * a.grad_scale;
        dmin += dl1 * (1.f - ex[j]);
        cnt += x[j] == mn ? 1.f : 0.f;
      }
    }
    dmin = bsum(dmin, red);
    cnt = bsum(cnt, red);
    share = dmin / cnt;
  }
  for (int c = threadIdx.x * 4; c < A; c += VT * 4) {
    const float4 xv = *reinterpret_cast<const float4*>(a.logit0 + base + c);
    const float4 lv = *reinterpret_cast<const float4*>(a.l1 + base + c);
    const float4 tv = *reinterpret_cast<const float4*>(a.total + base + c);
    const float4 uv = *reinterpret_cast<const float4*>(a.tuned + base + c);
    const float4 zv = *reinterpret_cast<const float4*>(a.target + base + c);
    const float4 ev = *reinterpret_cast<const float4*>(a.exist + c);
    const float x[4] = {xv.x, xv.y, xv.z, xv.w}, l1[4] = {lv.x, lv.y, lv.z, lv.w}, tt[4] = {tv.x, tv.y, tv.z, tv.w};
    const float u[4] = {uv.x, uv.y, uv.z, uv.w}, z[4] = {zv.x, zv.y, zv.z, zv.w}, ex[4] = {ev.x, ev.y, ev.z, ev.w};
    float d0[4], dt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float tm = (c + j) < a.num_train_answer ? 1.f : 0.f;
      if (a.fill_min) {
        // loss = (BCE(L1) + BCE(L1 + tuned)) * train_mask                      (_vqa_all.py:234-243)
        dt[j] = (sigmoid_stable(tt[j]) - z[j]) * tm * a.grad_scale;
        const float dl1 = dt[j] + (sigmoid_stable(l1[j]) - z[j]) * tm * a.grad_scale;
        d0[j] = dl1 * ex[j] + (x[j] == mn ? share : 0.f);
      } else {
        // loss = BCE(logit) * train_mask + BCE(tuned)                           (_vqa_all2.py:231-239)
        d0[j] = (sigmoid_stable(x[j]) - z[j]) * tm * a.grad_scale;
        dt[j] = (sigmoid_stable(u[j]) - z[j]) * a.grad_scale;
      }
    }
    *reinterpret_cast<float4*>(a.d_logit0_f32 + base + c) = make_float4(d0[0], d0[1], d0[2], d0[3]);
    *reinterpret_cast<float4*>(a.d_tuned_f32 + base + c) = make_float4(dt[0], dt[1], dt[2], dt[3]);
    store_planes4(a.d_logit0_hi, a.d_logit0_lo, base + c, d0);
    store_planes4(a.d_tuned_hi, a.d_tuned_lo, base + c, dt);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// full: reparameterisation noise (Philox4x32-10 + Box-Muller: one call = 4 normals), forward, KL, backward
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void normal4(unsigned long long group, unsigned long long seed, unsigned long long step,
                                        float (&n)[4]) {
  const Philox8 p = philox4x32_10(group, RNG_STREAM_NOISE, seed, step);
  // uniforms in (0, 1): 24 high bits + half an ulp
  const float u0 = (static_cast<float>(p.w[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u1 = (static_cast<float>(p.w[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = (static_cast<float>(p.w[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u3 = (static_cast<float>(p.w[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  n[0] = r0 * c0; n[1] = r0 * s0; n[2] = r1 * c1; n[3] = r1 * s1;
}

__global__ void reparam_noise_kernel(float* __restrict__ out, long long groups, unsigned long long seed,
                                     unsigned long long step) {
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < groups;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    float n[4];
    normal4(g, seed, step, n);
    *reinterpret_cast<float4*>(out + g * 4) = make_float4(n[0], n[1], n[2], n[3]);
  }
}

// one CTA per sample: q_noise = mean + noise * sqrt(exp(lss)) -> operand planes; kl_rows[b] = sum_l(1 + lss - mean^2 - exp(lss))
__global__ void __launch_bounds__(VT) reparam_fwd_kernel(ReparamFwd a) {
  __shared__ float red[VT / 32];
  const int b = blockIdx.x, L = a.L;
  const long long base = static_cast<long long>(b) * L;
  float kl = 0.f;
  for (int c = threadIdx.x * 4; c < L; c += VT * 4) {
    const float4 mv = *reinterpret_cast<const float4*>(a.mean + base + c);
    const float4 sv = *reinterpret_cast<const float4*>(a.lss + base + c);
    const float m[4] = {mv.x, mv.y, mv.z, mv.w}, s[4] = {sv.x, sv.y, sv.z, sv.w};
    float n[4], q[4];
    normal4(static_cast<unsigned long long>(base + c) >> 2, a.seed, a.step, n);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float ex = expf(s[j]);
      q[j] = fmaf(n[j], sqrtf(ex), m[j]);            // tf.sqrt(tf.exp(log_sigma_sq))  (_full.py:132-134)
      kl += 1.f + s[j] - m[j] * m[j] - ex;           // latent_loss (:272-276)
    }
    if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + base + c) = make_float4(q[0], q[1], q[2], q[3]);
    store_planes4(a.out_hi, a.out_lo, base + c, q);
  }
  kl = bsum(kl, red);
  if (threadIdx.x == 0) a.kl_rows[b] = kl;
}

// latent = -0.5 * mean_b kl_rows; loss += weight * latent; report slots (fixed summation order)
__global__ void __launch_bounds__(VT) latent_finalize_kernel(const float* __restrict__ kl_rows, int batch, float scale,
                                                             float weight, int slot, float* __restrict__ loss,
                                                             float* __restrict__ report) {
  __shared__ float red[VT / 32];
  float s = 0.f;
  for (int b = threadIdx.x; b < batch; b += VT) s += kl_rows[b];
  s = bsum(s, red);
  if (threadIdx.x == 0) {
    const float extra = scale * s / batch;   // full: -0.5 * mean_b KL sums; ent: mean_b negative entropies
    if (loss) loss[0] += weight * extra;
    if (report) {
      report[slot] = extra;
      report[slot + 1] = weight * extra;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// ent: tiled joint input, marginal softmax + negative entropy, and their backward (model_vlmap_answer_ent.py:193-213)
// ------------------------------------------------------------------------------------------------------------------
// x2[b, m, :] = Hp[(b*M + m) mod B, :] * Hl[b, :]   (tf.reshape(tf.tile(stop_gradient(Hp), [M, 1]), [-1, M, L]) * Hl[:, None])
__global__ void ent_tile_kernel(EntTile a, long long total8) {
  const int CH = a.L >> 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / CH;            // b * M + m
    const int c = static_cast<int>(i - row * CH);
    const int b = static_cast<int>(row / a.M);
    const int src = static_cast<int>(row % a.batch);
    float hp[8], hl[8];
    ldf8(a.hp + static_cast<long long>(src) * a.L + c * 8, hp);
    ldf8(a.hl + static_cast<long long>(b) * a.L + c * 8, hl);
#pragma unroll
    for (int j = 0; j < 8; ++j) hp[j] *= hl[j];
    st_planes8(a.out_hi, a.out_lo, i * 8, hp);
  }
}

// one CTA per sample: for every marginal row m, softmax over the selected (train & existing) answers; the mean over m is
// the marginal; ent_rows[b] = sum_a marg log(marg + 1e-8). Row maxima and reciprocal sums are kept for the backward.
__global__ void __launch_bounds__(VT) ent_marginal_kernel(EntMarginal a) {
  __shared__ float red[VT / 32];
  const int b = blockIdx.x, A = a.A, M = a.M;
  constexpr int MAXC = 4;   // float4 column chunks per thread: A <= 4 * 4 * VT = 4096
  float4 acc[MAXC], selm[MAXC];
#pragma unroll
  for (int t = 0; t < MAXC; ++t) {
    acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = (threadIdx.x + t * VT) * 4;
    selm[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < A) {
      const float4 ex = *reinterpret_cast<const float4*>(a.exist + c);
      selm[t].x = (ex.x > 0.5f && c + 0 < a.num_train_answer) ? 1.f : 0.f;
      selm[t].y = (ex.y > 0.5f && c + 1 < a.num_train_answer) ? 1.f : 0.f;
      selm[t].z = (ex.z > 0.5f && c + 2 < a.num_train_answer) ? 1.f : 0.f;
      selm[t].w = (ex.w > 0.5f && c + 3 < a.num_train_answer) ? 1.f : 0.f;
    }
  }
  const float inv_m = 1.0f / M;
  for (int m = 0; m < M; ++m) {
    const float* x = a.logit2 + (static_cast<long long>(b) * M + m) * A;
    float4 xv[MAXC];
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int t = 0; t < MAXC; ++t) {
      const int c = (threadIdx.x + t * VT) * 4;
      xv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < A) {
        xv[t] = *reinterpret_cast<const float4*>(x + c);
        if (selm[t].x > 0.f) mx = fmaxf(mx, xv[t].x);
        if (selm[t].y > 0.f) mx = fmaxf(mx, xv[t].y);
        if (selm[t].z > 0.f) mx = fmaxf(mx, xv[t].z);
        if (selm[t].w > 0.f) mx = fmaxf(mx, xv[t].w);
      }
    }
    mx = -bmin(-mx, red);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < MAXC; ++t) {
      xv[t].x = selm[t].x > 0.f ? expf(xv[t].x - mx) : 0.f;
      xv[t].y = selm[t].y > 0.f ? expf(xv[t].y - mx) : 0.f;
      xv[t].z = selm[t].z > 0.f ? expf(xv[t].z - mx) : 0.f;
      xv[t].w = selm[t].w > 0.f ? expf(xv[t].w - mx) : 0.f;
      sum += xv[t].x + xv[t].y + xv[t].z + xv[t].w;
    }
    sum = bsum(sum, red);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < MAXC; ++t) {
      acc[t].x += xv[t].x * inv; acc[t].y += xv[t].y * inv; acc[t].z += xv[t].z * inv; acc[t].w += xv[t].w * inv;
    }
    if (threadIdx.x == 0) {
      a.row_max[static_cast<long long>(b) * M + m] = mx;
      a.row_inv[static_cast<long long>(b) * M + m] = inv;
    }
  }
  float ent = 0.f;
#pragma unroll
  for (int t = 0; t < MAXC; ++t) {
    const int c = (threadIdx.x + t * VT) * 4;
    if (c < A) {
      const float4 mg = make_float4(acc[t].x * inv_m, acc[t].y * inv_m, acc[t].z * inv_m, acc[t].w * inv_m);
      *reinterpret_cast<float4*>(a.marg + static_cast<long long>(b) * A + c) = mg;
      // unselected answers are not part of the boolean_mask'ed tensor: they contribute nothing
      ent += selm[t].x * mg.x * logf(mg.x + 1e-8f) + selm[t].y * mg.y * logf(mg.y + 1e-8f) +
             selm[t].z * mg.z * logf(mg.z + 1e-8f) + selm[t].w * mg.w * logf(mg.w + 1e-8f);
    }
  }
  ent = bsum(ent, red);
  if (threadIdx.x == 0) a.ent_rows[b] = ent;
}

// d logit2[b, m, a] = prob * (dprob - sum_a' prob dprob), dprob = c (log(marg + eps) + marg / (marg + eps)) / M
__global__ void __launch_bounds__(VT) ent_marginal_bwd_kernel(EntMarginalBwd a) {
  __shared__ float red[VT / 32];
  const int b = blockIdx.x, A = a.A, M = a.M;
  constexpr int MAXC = 4;
  float4 dpr[MAXC];
#pragma unroll
  for (int t = 0; t < MAXC; ++t) {
    const int c = (threadIdx.x + t * VT) * 4;
    dpr[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < A) {
      const float4 ex = *reinterpret_cast<const float4*>(a.exist + c);
      const float4 mg = *reinterpret_cast<const float4*>(a.marg + static_cast<long long>(b) * A + c);
      const float k = a.scale / M;
      if (ex.x > 0.5f && c + 0 < a.num_train_answer) dpr[t].x = k * (logf(mg.x + 1e-8f) + mg.x / (mg.x + 1e-8f));
      if (ex.y > 0.5f && c + 1 < a.num_train_answer) dpr[t].y = k * (logf(mg.y + 1e-8f) + mg.y / (mg.y + 1e-8f));
      if (ex.z > 0.5f && c + 2 < a.num_train_answer) dpr[t].z = k * (logf(mg.z + 1e-8f) + mg.z / (mg.z + 1e-8f));
      if (ex.w > 0.5f && c + 3 < a.num_train_answer) dpr[t].w = k * (logf(mg.w + 1e-8f) + mg.w / (mg.w + 1e-8f));
    }
  }
  for (int m = 0; m < M; ++m) {
    const long long row = static_cast<long long>(b) * M + m;
    const float* x = a.logit2 + row * A;
    const float mx = a.row_max[row], inv = a.row_inv[row];
    float4 pr[MAXC];
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < MAXC; ++t) {
      const int c = (threadIdx.x + t * VT) * 4;
      pr[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < A) {
        const float4 xv = *reinterpret_cast<const float4*>(x + c);
        const float4 ex = *reinterpret_cast<const float4*>(a.exist + c);
        if (ex.x > 0.5f && c + 0 < a.num_train_answer) pr[t].x = expf(xv.x - mx) * inv;
        if (ex.y > 0.5f && c + 1 < a.num_train_answer) pr[t].y = expf(xv.y - mx) * inv;
        if (ex.z > 0.5f && c + 2 < a.num_train_answer) pr[t].z = expf(xv.z - mx) * inv;
        if (ex.w > 0.5f && c + 3 < a.num_train_answer) pr[t].w = expf(xv.w - mx) * inv;
        dot += pr[t].x * dpr[t].x + pr[t].y * dpr[t].y + pr[t].z * dpr[t].z + pr[t].w * dpr[t].w;
      }
    }
    dot = bsum(dot, red);
#pragma unroll
    for (int t = 0; t < MAXC; ++t) {
      const int c = (threadIdx.x + t * VT) * 4;
      if (c < A) {
        const float d[4] = {pr[t].x * (dpr[t].x - dot), pr[t].y * (dpr[t].y - dot), pr[t].z * (dpr[t].z - dot),
                            pr[t].w * (dpr[t].w - dot)};
        store_planes4(a.d_hi, a.d_lo, row * A + c, d);
      }
    }
  }
}

// dHl[b, :] = dX[b, :] * Hp[b, :] + sum_m dX2[b, m, :] * Hp[(b*M + m) mod B, :]
__global__ void ent_dhl_kernel(EntDhl a, long long total4) {
  const int L4 = a.L >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / L4), c = static_cast<int>(i - static_cast<long long>(b) * L4) * 4;
    const float4 dx = *reinterpret_cast<const float4*>(a.dX + static_cast<long long>(b) * a.L + c);
    const float4 hp = *reinterpret_cast<const float4*>(a.hp + static_cast<long long>(b) * a.L + c);
    float4 acc = make_float4(dx.x * hp.x, dx.y * hp.y, dx.z * hp.z, dx.w * hp.w);
    for (int m = 0; m < a.M; ++m) {
      const long long row = static_cast<long long>(b) * a.M + m;
      const int src = static_cast<int>(row % a.batch);
      const float4 d2 = *reinterpret_cast<const float4*>(a.dX2 + row * a.L + c);
      const float4 h2 = *reinterpret_cast<const float4*>(a.hp + static_cast<long long>(src) * a.L + c);
      acc.x = fmaf(d2.x, h2.x, acc.x); acc.y = fmaf(d2.y, h2.y, acc.y);
      acc.z = fmaf(d2.z, h2.z, acc.z); acc.w = fmaf(d2.w, h2.w, acc.w);
    }
    *reinterpret_cast<float4*>(a.out + static_cast<long long>(b) * a.L + c) = acc;
  }
}

__global__ void reparam_bwd_kernel(ReparamBwd a, long long total4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 dv = reinterpret_cast<const float4*>(a.d_out)[i];
    const float4 mv = reinterpret_cast<const float4*>(a.mean)[i];
    const float4 sv = reinterpret_cast<const float4*>(a.lss)[i];
    const float d[4] = {dv.x, dv.y, dv.z, dv.w}, m[4] = {mv.x, mv.y, mv.z, mv.w}, s[4] = {sv.x, sv.y, sv.z, sv.w};
    float n[4], dm[4], ds[4];
    normal4(static_cast<unsigned long long>(i), a.seed, a.step, n);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float ex = expf(s[j]);
      dm[j] = d[j] + a.kl_scale * m[j];
      ds[j] = d[j] * n[j] * sqrtf(ex) * 0.5f - 0.5f * a.kl_scale * (1.f - ex);
    }
    reinterpret_cast<float4*>(a.d_mean_f32)[i] = make_float4(dm[0], dm[1], dm[2], dm[3]);
    reinterpret_cast<float4*>(a.d_lss_f32)[i] = make_float4(ds[0], ds[1], ds[2], ds[3]);
    store_planes4(a.d_mean_hi, a.d_mean_lo, i * 4, dm);
    store_planes4(a.d_lss_hi, a.d_lss_lo, i * 4, ds);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// adapt: LayerNorm over the whole [K, D] slab of a sample + ReLU, materialised; and its backward
// ------------------------------------------------------------------------------------------------------------------
template <typename ZT>
__global__ void __launch_bounds__(VT) slab_ln_relu_fwd_kernel(SlabLnFwd a) {
  __shared__ float red[VT / 32];
  const int b = blockIdx.x, CH = a.D >> 3, n8 = a.K * CH;
  const long long base = static_cast<long long>(b) * a.K * a.D;
  const ZT* z = static_cast<const ZT*>(a.z) + base;
  const float N = static_cast<float>(a.K) * a.D;
  float s = 0.f;
  for (int i = threadIdx.x; i < n8; i += VT) {
    float x[8];
    ldz8<ZT>(z + static_cast<long long>(i) * 8, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j];
  }
  const float mean = bsum(s, red) / N;
  float q = 0.f;
  for (int i = threadIdx.x; i < n8; i += VT) {   // tf.nn.moments: two-pass variance (the slab is L1 / L2 resident)
    float x[8];
    ldz8<ZT>(z + static_cast<long long>(i) * 8, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) q += (x[j] - mean) * (x[j] - mean);
  }
  const float rstd = 1.0f / sqrtf(bsum(q, red) / N + 1e-12f);
  for (int i = threadIdx.x; i < n8; i += VT) {
    const int c = i % CH;
    float x[8], g[8], bt[8], y[8];
    ldz8<ZT>(z + static_cast<long long>(i) * 8, x);
    ldf8(a.gamma + c * 8, g);
    ldf8(a.beta + c * 8, bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = fmaxf(fmaf((x[j] - mean) * rstd, g[j], bt[j]), 0.f);
    if (a.thr < 65536u) {   // tf.nn.dropout on the layer's output (the tiled joint of the ent variant)
      const uint32_t bits = philox_keep_bits(
          philox4x32_10(static_cast<unsigned long long>(base >> 3) + i, a.site, a.seed, a.step), a.thr);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = ((bits >> j) & 1u) ? y[j] * a.inv_keep : 0.f;
    }
    st_planes8(a.out_hi, a.out_lo, base + static_cast<long long>(i) * 8, y);
  }
  if (threadIdx.x == 0) {
    a.mean[b] = mean;
    a.rstd[b] = rstd;
  }
}

// d v_adapt[k, d] = att[k] * dP[d]; ReLU gate from the recomputed LN output; LN backward over the slab; per-sample
// partials of d gamma / d beta / d bias (reduced over the batch by colsum afterwards: deterministic)
template <typename ZT>
__global__ void __launch_bounds__(VT) slab_ln_relu_bwd_kernel(SlabLnBwd a) {
  __shared__ float red[VT / 32];
  const int b = blockIdx.x, D = a.D, CH = D >> 3, n8 = a.K * CH;
  const long long base = static_cast<long long>(b) * a.K * D;
  const ZT* z = static_cast<const ZT*>(a.z) + base;
  const float N = static_cast<float>(a.K) * D;
  const float mean = a.mean[b], rstd = a.rstd[b];
  const float* att = a.dout ? nullptr : a.att + static_cast<long long>(b) * a.K;
  const float* dP = a.dout ? nullptr : a.d_pooled + static_cast<long long>(b) * D;
  const float* dout = a.dout ? a.dout + base : nullptr;
  // upstream gradient of 8 outputs at flat chunk i (row k, column chunk c): a_k * dP (adapt) or the given tensor through
  // the regenerated dropout mask (ent)
  auto upstream = [&](int i, int k, int c, float (&up)[8]) {
    if (dout) {
      ldf8(dout + static_cast<long long>(i) * 8, up);
      if (a.thr < 65536u) {
        const uint32_t bits = philox_keep_bits(
            philox4x32_10(static_cast<unsigned long long>(base >> 3) + i, a.site, a.seed, a.step), a.thr);
#pragma unroll
        for (int j = 0; j < 8; ++j) up[j] = ((bits >> j) & 1u) ? up[j] * a.inv_keep : 0.f;
      }
    } else {
      ldf8(dP + c * 8, up);
      const float ak = att[k];
#pragma unroll
      for (int j = 0; j < 8; ++j) up[j] *= ak;
    }
  };
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < n8; i += VT) {
    const int k = i / CH, c = i % CH;
    float x[8], g[8], bt[8], dp[8];
    ldz8<ZT>(z + static_cast<long long>(i) * 8, x);
    ldf8(a.gamma + c * 8, g);
    ldf8(a.beta + c * 8, bt);
    upstream(i, k, c, dp);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (x[j] - mean) * rstd;
      const float dy = fmaf(xh, g[j], bt[j]) > 0.f ? dp[j] : 0.f;
      const float dxh = dy * g[j];
      s1 += dxh;
      s2 = fmaf(dxh, xh, s2);
    }
  }
  const float m1 = bsum(s1, red) / N;
  const float m2 = bsum(s2, red) / N;
  // second pass: a thread owns column chunks and walks the K rows, so the per-column partial sums stay in registers
  for (int c = threadIdx.x; c < CH; c += VT) {
    float g[8], bt[8], dg[8], db[8], dbias[8];
    ldf8(a.gamma + c * 8, g);
    ldf8(a.beta + c * 8, bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[j] = db[j] = dbias[j] = 0.f;
    for (int k = 0; k < a.K; ++k) {
      float x[8], dz[8], dp[8];
      const long long o = static_cast<long long>(k) * D + c * 8;
      ldz8<ZT>(z + o, x);
      upstream(k * CH + c, k, c, dp);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[j] - mean) * rstd;
        const float dy = fmaf(xh, g[j], bt[j]) > 0.f ? dp[j] : 0.f;
        dg[j] = fmaf(dy, xh, dg[j]);
        db[j] += dy;
        dz[j] = rstd * (dy * g[j] - m1 - xh * m2);
        dbias[j] += dz[j];
      }
      st_planes8(a.dz_hi, a.dz_lo, base + o, dz);
    }
    if (!a.part) continue;   // frozen layer: no parameter gradients wanted
    float* pg = a.part + (static_cast<long long>(b) * 3 + 0) * D + c * 8;
    float* pb = a.part + (static_cast<long long>(b) * 3 + 1) * D + c * 8;
    float* ps = a.part + (static_cast<long long>(b) * 3 + 2) * D + c * 8;
    *reinterpret_cast<float4*>(pg) = make_float4(dg[0], dg[1], dg[2], dg[3]);
    *reinterpret_cast<float4*>(pg + 4) = make_float4(dg[4], dg[5], dg[6], dg[7]);
    *reinterpret_cast<float4*>(pb) = make_float4(db[0], db[1], db[2], db[3]);
    *reinterpret_cast<float4*>(pb + 4) = make_float4(db[4], db[5], db[6], db[7]);
    *reinterpret_cast<float4*>(ps) = make_float4(dbias[0], dbias[1], dbias[2], dbias[3]);
    *reinterpret_cast<float4*>(ps + 4) = make_float4(dbias[4], dbias[5], dbias[6], dbias[7]);
  }
}

int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

VqaStatus tuned_combine_launch(const TunedHeadFwd& a, cudaStream_t s) {
  if (a.batch == 0) return VQA_OK;
  if (a.A & 3) return set_error(VQA_ERR_BAD_SHAPE, "tuned_combine: A must be a multiple of 4");
  tuned_combine_kernel<<<a.batch, VT, 0, s>>>(a);
  VQA_LAUNCH_CHECK("tuned_combine");
  return VQA_OK;
}

VqaStatus tuned_grad_launch(const TunedHeadBwd& a, cudaStream_t s) {
  if (a.batch == 0) return VQA_OK;
  if (a.A & 3) return set_error(VQA_ERR_BAD_SHAPE, "tuned_grad: A must be a multiple of 4");
  tuned_grad_kernel<<<a.batch, VT, 0, s>>>(a);
  VQA_LAUNCH_CHECK("tuned_grad");
  return VQA_OK;
}

VqaStatus reparam_noise_launch(float* out, long long n, unsigned long long seed, unsigned long long step,
                               cudaStream_t s) {
  if (n == 0) return VQA_OK;
  if (n & 3) return set_error(VQA_ERR_BAD_SHAPE, "reparam noise: element count must be a multiple of 4");
  reparam_noise_kernel<<<grid_for(n / 4, 256), 256, 0, s>>>(out, n / 4, seed, step);
  VQA_LAUNCH_CHECK("reparam_noise");
  return VQA_OK;
}

VqaStatus reparam_fwd_launch(const ReparamFwd& a, cudaStream_t s) {
  if (a.batch == 0) return VQA_OK;
  if (a.L & 3) return set_error(VQA_ERR_BAD_SHAPE, "reparam: L must be a multiple of 4");
  reparam_fwd_kernel<<<a.batch, VT, 0, s>>>(a);
  VQA_LAUNCH_CHECK("reparam_fwd");
  return VQA_OK;
}

VqaStatus latent_finalize_launch(const float* rows, int batch, float scale, float weight, int slot, float* loss,
                                 float* report, cudaStream_t s) {
  if (batch == 0) return VQA_OK;
  latent_finalize_kernel<<<1, VT, 0, s>>>(rows, batch, scale, weight, slot, loss, report);
  VQA_LAUNCH_CHECK("latent_finalize");
  return VQA_OK;
}

VqaStatus ent_tile_launch(const EntTile& a, cudaStream_t s) {
  if (a.batch == 0) return VQA_OK;
  if (a.L & 7) return set_error(VQA_ERR_BAD_SHAPE, "ent_tile: L must be a multiple of 8");
  const long long total8 = static_cast<long long>(a.batch) * a.M * (a.L >> 3);
  ent_tile_kernel<<<grid_for(total8, 256), 256, 0, s>>>(a, total8);
  VQA_LAUNCH_CHECK("ent_tile");
  return VQA_OK;
}

VqaStatus ent_marginal_launch(const EntMarginal& a, cudaStream_t s) {
  if (a.batch == 0) return VQA_OK;
  if ((a.A & 3) || a.A > 16 * VT) return set_error(VQA_ERR_BAD_SHAPE, "ent_marginal: A must be a multiple of 4 and <= 4096");
  ent_marginal_kernel<<<a.batch, VT, 0, s>>>(a);
  VQA_LAUNCH_CHECK("ent_marginal");
  return VQA_OK;
}

VqaStatus ent_marginal_bwd_launch(const EntMarginalBwd& a, cudaStream_t s) {
  if (a.batch == 0) return VQA_OK;
  if ((a.A & 3) || a.A > 16 * VT) return set_error(VQA_ERR_BAD_SHAPE, "ent_marginal: A must be a multiple of 4 and <= 4096");
  ent_marginal_bwd_kernel<<<a.batch, VT, 0, s>>>(a);
  VQA_LAUNCH_CHECK("ent_marginal_bwd");
  return VQA_OK;
}

VqaStatus ent_dhl_launch(const EntDhl& a, cudaStream_t s) {
  if (a.batch == 0) return VQA_OK;
  if (a.L & 3) return set_error(VQA_ERR_BAD_SHAPE, "ent_dhl: L must be a multiple of 4");
  const long long total4 = static_cast<long long>(a.batch) * (a.L >> 2);
  ent_dhl_kernel<<<grid_for(total4, 128), 128, 0, s>>>(a, total4);
  VQA_LAUNCH_CHECK("ent_dhl");
  return VQA_OK;
}

VqaStatus reparam_bwd_launch(const ReparamBwd& a, cudaStream_t s) {
  if (a.batch == 0) return VQA_OK;
  if (a.L & 3) return set_error(VQA_ERR_BAD_SHAPE, "reparam: L must be a multiple of 4");
  const long long total4 = static_cast<long long>(a.batch) * a.L / 4;
  reparam_bwd_kernel<<<grid_for(total4, 256), 256, 0, s>>>(a, total4);
  VQA_LAUNCH_CHECK("reparam_bwd");
  return VQA_OK;
}

VqaStatus slab_ln_relu_fwd_launch(const SlabLnFwd& a, int precision, cudaStream_t s) {
  if (a.batch == 0) return VQA_OK;
  if (a.D & 7) return set_error(VQA_ERR_BAD_SHAPE, "slab_ln_relu: D must be a multiple of 8");
  if (precision == VQA_PREC_FP32) slab_ln_relu_fwd_kernel<float><<<a.batch, VT, 0, s>>>(a);
  else slab_ln_relu_fwd_kernel<bf16><<<a.batch, VT, 0, s>>>(a);
  VQA_LAUNCH_CHECK("slab_ln_relu_fwd");
  return VQA_OK;
}

VqaStatus slab_ln_relu_bwd_launch(const SlabLnBwd& a, int precision, cudaStream_t s) {
  if (a.batch == 0) return VQA_OK;
  if (a.D & 7) return set_error(VQA_ERR_BAD_SHAPE, "slab_ln_relu: D must be a multiple of 8");
  if (precision == VQA_PREC_FP32) slab_ln_relu_bwd_kernel<float><<<a.batch, VT, 0, s>>>(a);
  else slab_ln_relu_bwd_kernel<bf16><<<a.batch, VT, 0, s>>>(a);
  VQA_LAUNCH_CHECK("slab_ln_relu_bwd");
  return VQA_OK;
}

}  // namespace vqa
