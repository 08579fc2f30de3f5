// Optimizer step of the reference trainer (vqa/trainer.py:87-114): tf.contrib.layers.optimize_loss with
// AdamOptimizer and clip_gradients=20.0, i.e. clip_by_global_norm over the train variables followed by
// Adam (beta1 0.9, beta2 0.999, eps 1e-8, lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)).
// The trainable set is one flat fp32 buffer (the caller lays the variables out contiguously), so the
// step is: (1) sum of squares, two-level deterministic reduction; (2) one fused, 128-bit vectorised
// pass over param / grad / m / v. HBM-bound: 4 reads + 3 writes of 4 bytes per parameter.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>

#include "internal.h"
#include "launch.cuh"

namespace vqa {

namespace {

constexpr int OPT_THREADS = 256;

// write 4 updated parameters as GEMM-operand planes (bf16 hi, optional lo residual)
__device__ __forceinline__ void shadow_store4(bf16* hi, bf16* lo, long long o, const float4& x) {
  const bf16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y), h2 = __float2bfloat16_rn(x.z),
             h3 = __float2bfloat16_rn(x.w);
  __nv_bfloat162 p0(h0, h1), p1(h2, h3);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&p0);
  pk.y = *reinterpret_cast<uint32_t*>(&p1);
  *reinterpret_cast<uint2*>(hi + o) = pk;
  if (lo) {
    __nv_bfloat162 q0(__float2bfloat16_rn(x.x - __bfloat162float(h0)), __float2bfloat16_rn(x.y - __bfloat162float(h1)));
    __nv_bfloat162 q1(__float2bfloat16_rn(x.z - __bfloat162float(h2)), __float2bfloat16_rn(x.w - __bfloat162float(h3)));
    pk.x = *reinterpret_cast<uint32_t*>(&q0);
    pk.y = *reinterpret_cast<uint32_t*>(&q1);
    *reinterpret_cast<uint2*>(lo + o) = pk;
  }
}

// float4 chunks [skip_begin4, skip_end4) are left out (the dense gradient of the IndexedSlices variable, see norm_final)
__global__ void __launch_bounds__(OPT_THREADS) sumsq_partial_kernel(const float* __restrict__ g, long long n,
                                                                    long long skip_begin4, long long skip_end4,
                                                                    float* __restrict__ part) {
  __shared__ float red[OPT_THREADS / 32];
  pdl_sync();
  float acc = 0.f;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(OPT_THREADS) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * OPT_THREADS) {
    if (i >= skip_begin4 && i < skip_end4) continue;
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) acc += g[i] * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
    part[blockIdx.x] = s;
  }
}

// norm = sqrt(sum(part) + *plus): `plus` is the sum of squares of the embedding gradient's IndexedSlices rows (what
// clip_ops.global_norm sees in the reference); the partials then leave the dense embedding gradient out
__global__ void __launch_bounds__(OPT_THREADS) norm_final_kernel(const float* __restrict__ part, int parts,
                                                                 const float* __restrict__ plus,
                                                                 float* __restrict__ norm_out,
                                                                 float* __restrict__ user_out) {
  __shared__ double red[OPT_THREADS / 32];
  pdl_sync();
  double acc = 0.0;
  for (int i = threadIdx.x; i < parts; i += OPT_THREADS) acc += static_cast<double>(part[i]);
  if (threadIdx.x == 0 && plus) acc += static_cast<double>(plus[0]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
    const float nrm = static_cast<float>(sqrt(s > 0.0 ? s : 0.0));
    norm_out[0] = nrm;
    if (user_out) user_out[0] = nrm;
  }
}

__global__ void __launch_bounds__(OPT_THREADS) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                           float* __restrict__ m, float* __restrict__ v,
                                                           long long n, float lr_t, float b1, float b2,
                                                           float eps, float clip,
                                                           const float* __restrict__ norm, AdamShadows tab) {
  pdl_sync();
  const float nrm = norm[0];
  const float scale = clip > 0.f ? clip / fmaxf(nrm, clip) : 1.f;  // clip_by_global_norm
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(OPT_THREADS) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * OPT_THREADS) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define VQA_ADAM1(c)                                   \
  {                                                    \
    const float gs = gg.c * scale;                     \
    mm.c = b1 * mm.c + (1.f - b1) * gs;                \
    vv.c = b2 * vv.c + (1.f - b2) * gs * gs;           \
    pp.c -= lr_t * mm.c / (sqrtf(vv.c) + eps);         \
  }
    VQA_ADAM1(x) VQA_ADAM1(y) VQA_ADAM1(z) VQA_ADAM1(w)
#undef VQA_ADAM1
    reinterpret_cast<float4*>(p)[i] = pp;
    // the weight matrices' bf16 operand shadows are refreshed in the same pass (tensor starts are 64-element aligned,
    // so a float4 never straddles two tensors)
    for (int k = 0; k < tab.n; ++k)
      if (i >= tab.begin4[k] && i < tab.end4[k]) shadow_store4(tab.hi[k], tab.lo[k], (i - tab.begin4[k]) * 4, pp);
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) {
      const float gs = g[i] * scale;
      m[i] = b1 * m[i] + (1.f - b1) * gs;
      v[i] = b2 * v[i] + (1.f - b2) * gs * gs;
      p[i] -= lr_t * m[i] / (sqrtf(v[i]) + eps);
    }
}

// sum of squares of a [rows, cols] matrix with pitch ld (the pad columns are not part of it): per-block partials
__global__ void __launch_bounds__(OPT_THREADS) rows_sumsq_partial_kernel(const float* __restrict__ x, long long rows, int cols,
                                                                         long long ld, float* __restrict__ part) {
  __shared__ float red[OPT_THREADS / 32];
  pdl_sync();
  float acc = 0.f;
  // a warp per row, lanes across the columns (no per-element division: the first version spent 35 us on 2 M elements)
  const int lane = threadIdx.x & 31;
  const long long warp0 = blockIdx.x * static_cast<long long>(OPT_THREADS / 32) + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * (OPT_THREADS / 32);
  const bool vec = (cols & 3) == 0 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  for (long long r = warp0; r < rows; r += nwarps) {
    const float* row = x + r * ld;
    if (vec) {
      for (int c = lane * 4; c < cols; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(row + c);
        acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
      }
    } else {
      for (int c = lane; c < cols; c += 32) acc = fmaf(row[c], row[c], acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < OPT_THREADS / 32; ++w) t += red[w];
    part[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(OPT_THREADS) sum_parts_kernel(const float* __restrict__ part, int parts, float* __restrict__ out) {
  __shared__ double red[OPT_THREADS / 32];
  pdl_sync();
  double acc = 0.0;
  for (int i = threadIdx.x; i < parts; i += OPT_THREADS) acc += static_cast<double>(part[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < OPT_THREADS / 32; ++w) t += red[w];
    out[0] = static_cast<float>(t);
  }
}

}  // namespace

VqaStatus rows_sumsq_launch(const float* x, long long rows, int cols, long long ld, float* out, float* scratch,
                            cudaStream_t s) {
  if (!out) return VQA_OK;
  const int blocks = 148;
  launch_pdl(rows_sumsq_partial_kernel, dim3(blocks), dim3(OPT_THREADS), 0, s, x, rows, cols, ld, scratch);
  VQA_LAUNCH_CHECK("rows_sumsq_partial");
  launch_pdl(sum_parts_kernel, dim3(1), dim3(OPT_THREADS), 0, s, static_cast<const float*>(scratch), blocks, out);
  VQA_LAUNCH_CHECK("sum_parts");
  return VQA_OK;
}

VqaStatus adam_step_launch(float* param, const float* grad, float* m, float* v, long long n, float lr,
                           float beta1, float beta2, float eps, float clip_norm, long long t,
                           float* grad_norm_out, float* scratch, int num_sms, cudaStream_t s, const AdamShadows* shadows,
                           const float* slice_grad, long long slice_n, const float* slice_sumsq, long long tail_begin,
                           cudaStream_t tail_stream, cudaEvent_t fork_ev) {
  if (n <= 0) return VQA_OK;
  if (!param || !grad || !m || !v || !scratch) return set_error(VQA_ERR_BAD_ARG, "vqa_adam_step: null argument");
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
       reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15)
    return set_error(VQA_ERR_BAD_ARG, "vqa_adam_step: buffers must be 16-byte aligned");
  if (t < 1) return set_error(VQA_ERR_BAD_ARG, "vqa_adam_step: t is the 1-based step count");
  int blocks = num_sms * 8;
  const long long need = (n / 4 + OPT_THREADS - 1) / OPT_THREADS;
  if (need < blocks) blocks = need < 1 ? 1 : static_cast<int>(need);
  if (blocks > 2048) blocks = 2048;
  // the dense gradient of the IndexedSlices variable is left out of the partial sums; its slice rows' sum comes in as `plus`
  long long skip_b = 0, skip_e = 0;
  const bool slices = slice_grad && slice_sumsq && slice_n > 0 && ((slice_grad - grad) & 3) == 0 && (slice_n & 3) == 0;
  if (slices) {
    skip_b = (slice_grad - grad) >> 2;
    skip_e = skip_b + (slice_n >> 2);
  }
  launch_pdl(sumsq_partial_kernel, dim3(blocks), dim3(OPT_THREADS), 0, s, grad, n, skip_b, skip_e, scratch + 8);
  VQA_LAUNCH_CHECK("sumsq_partial");
  launch_pdl(norm_final_kernel, dim3(1), dim3(OPT_THREADS), 0, s, static_cast<const float*>(scratch + 8), blocks,
             slices ? slice_sumsq : static_cast<const float*>(nullptr), scratch, grad_norm_out);
  VQA_LAUNCH_CHECK("norm_final");
  const double lr_t = static_cast<double>(lr) * std::sqrt(1.0 - std::pow(static_cast<double>(beta2), static_cast<double>(t))) /
                      (1.0 - std::pow(static_cast<double>(beta1), static_cast<double>(t)));
  AdamShadows tab{};
  if (shadows) tab = *shadows;
  if (tail_begin > 0 && tail_begin < n && (tail_begin & 3) == 0 && tail_stream && fork_ev) {
    // two launches: [0, tail_begin) on `s`, the tail on `tail_stream` (ordered after the norm) -- the caller lets the next
    // step's first kernels, which read only the head of the buffer, start while the tail is still being updated
    auto sub = [&](long long b, long long e, cudaStream_t st) -> VqaStatus {
      AdamShadows t2{};
      for (int k = 0; k < tab.n; ++k) {
        const long long lo4 = tab.begin4[k] > (b >> 2) ? tab.begin4[k] : (b >> 2);
        const long long hi4 = tab.end4[k] < (e >> 2) ? tab.end4[k] : (e >> 2);
        if (lo4 >= hi4) continue;
        if (lo4 != tab.begin4[k] || hi4 != tab.end4[k]) return set_error(VQA_ERR_BAD_ARG, "vqa_adam_step: the tail boundary splits a weight matrix");
        t2.begin4[t2.n] = lo4 - (b >> 2);
        t2.end4[t2.n] = hi4 - (b >> 2);
        t2.hi[t2.n] = tab.hi[k];
        t2.lo[t2.n] = tab.lo[k];
        ++t2.n;
      }
      int bl = blocks;
      const long long need2 = ((e - b) / 4 + OPT_THREADS - 1) / OPT_THREADS;
      if (need2 < bl) bl = need2 < 1 ? 1 : static_cast<int>(need2);
      launch_pdl(adam_kernel, dim3(bl), dim3(OPT_THREADS), 0, st, param + b, grad + b, m + b, v + b, e - b,
                 static_cast<float>(lr_t), beta1, beta2, eps, clip_norm, scratch, t2);
      VQA_LAUNCH_CHECK("adam");
      return VQA_OK;
    };
    VQA_CUDA_CHECK(cudaEventRecord(fork_ev, s));
    VQA_CUDA_CHECK(cudaStreamWaitEvent(tail_stream, fork_ev, 0));
    VQA_TRY(sub(0, tail_begin, s));
    VQA_TRY(sub(tail_begin, n, tail_stream));
    return VQA_OK;
  }
  launch_pdl(adam_kernel, dim3(blocks), dim3(OPT_THREADS), 0, s, param, grad, m, v, n, static_cast<float>(lr_t), beta1,
             beta2, eps, clip_norm, scratch, tab);
  VQA_LAUNCH_CHECK("adam");
  return VQA_OK;
}

}  // namespace vqa
