// Operators of the vlmap pre-training path (BASELINE config 4; include/vqa_memft.h):
//   vlmap_memft/model_vlmap_bf_or_wordset_withatt_sp.py:323-365, 413-455  spatial attention + attended pooling
//   :505-609 heads (pooled_linear_l, q_linear_l, joint_fc, classifier), :367-411 wordset branch, :675-706 loss
//   vlmap/modules.py:630-650 fc_layer (LayerNorm over the [n, dim] slab of a rank-3 input), :67-97, :23-39, :124-140
// The dense contractions run on the tcgen05 GEMM kernels (gemm.cu / gemm_pair.cu) and the recurrent kernels
// (gru_pair.cu / gru.cu); what is written here is the memory-bound rest -- slab LayerNorm, the attention block over
// 6-d box features (V read once per image and kind instead of tf.tile x n), softmax cross-entropy with top-k, the
// wordset lookup -- plus the operator context and the GRU sequence operator built from the existing launchers.
// Everything accumulates in fp32; GEMM operands leave as bf16 planes (hi, + lo residual in fp32 mode).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <new>

#include "../../include/vqa_memft.h"
#include "internal.h"
#include "launch.cuh"
#include "philox.cuh"

struct VqaOps_t {
  int num_sms;
  vqa::GemmCtx gemm_ctx;
  unsigned int* sem;
  float* scratch;            // colsum / Adam reduction scratch
  long long scratch_floats;
};

namespace vqa {

namespace {

constexpr int NMAX = 8;            // entries per image and kind
constexpr int kSemRegions = 64, kSemElems = 1024;
constexpr int SL_THREADS = 512;

__device__ __forceinline__ void st_planes(bf16* hi, bf16* lo, long long i, float v) {
  const bf16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}
__device__ __forceinline__ float ld_planes(const bf16* hi, const bf16* lo, long long i) {
  float v = __bfloat162float(hi[i]);
  if (lo) v += __bfloat162float(lo[i]);
  return v;
}
// 8 consecutive bf16 (16-byte aligned) -> 8 floats
__device__ __forceinline__ void ld8_planes(const bf16* hi, const bf16* lo, long long i, float (&v)[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(hi + i);
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = __uint_as_float(w[j] << 16);
    v[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
  if (lo) {
    const uint4 b = *reinterpret_cast<const uint4*>(lo + i);
    const uint32_t x[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] += __uint_as_float(x[j] << 16);
      v[2 * j + 1] += __uint_as_float(x[j] & 0xFFFF0000u);
    }
  }
}
__device__ __forceinline__ void st8_planes(bf16* hi, bf16* lo, long long i, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bf16 a = __float2bfloat16_rn(v[2 * j]), b = __float2bfloat16_rn(v[2 * j + 1]);
    h[j] = static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
    const bf16 c = __float2bfloat16_rn(v[2 * j] - __bfloat162float(a)), d = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(b));
    l[j] = static_cast<uint32_t>(__bfloat16_as_ushort(c)) | (static_cast<uint32_t>(__bfloat16_as_ushort(d)) << 16);
  }
  *reinterpret_cast<uint4*>(hi + i) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + i) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over the block in a fixed order; every thread gets the result. red: >= 33 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();   // red may still be read from a previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? red[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ---------------------------------------------------------------------------------------------------------------------
// slab LayerNorm + activation (+ Hadamard partner, + dropout): one CTA per slab, the slab in shared memory
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(float pre, int act) {
  return act == 0 ? fmaxf(pre, 0.f) : (act == 1 ? tanhf(pre) : pre);
}
__device__ __forceinline__ uint32_t slab_keep_bits(const VqaSlabLn& a, long long row, int c, uint32_t thr) {
  if (thr >= 65536u) return 0xFFu;
  const uint32_t site = a.site0 + static_cast<uint32_t>(row / a.rows_per_site);
  const unsigned long long group = (static_cast<unsigned long long>(row % a.rows_per_site) * a.N + c) >> 3;
  return philox_keep_bits(philox4x32_10(group, site, a.seed, a.step), thr);
}

__global__ void __launch_bounds__(SL_THREADS) slab_ln_fwd_kernel(const VqaSlabLn a) {
  extern __shared__ float x[];
  __shared__ float red[33];
  const int N = a.N, S = a.n * a.N, tid = threadIdx.x, nt = blockDim.x;
  const long long base = static_cast<long long>(blockIdx.x) * S;
  float s = 0.f;
  for (int i = tid * 4; i < S; i += nt * 4) {
    const float4 v = *reinterpret_cast<const float4*>(a.z + base + i);
    *reinterpret_cast<float4*>(x + i) = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = block_sum(s, red) / static_cast<float>(S);
  float q = 0.f;
  for (int i = tid; i < S; i += nt) {
    const float d = x[i] - mean;
    q = fmaf(d, d, q);
  }
  const float var = block_sum(q, red) / static_cast<float>(S);
  const float rstd = 1.0f / sqrtf(var + 1e-12f);
  if (tid == 0) {
    a.mean[blockIdx.x] = mean;
    a.rstd[blockIdx.x] = rstd;
  }
  const uint32_t thr = keep_threshold(a.keep);
  const float inv_keep = 1.0f / a.keep;
  bf16* ohi = static_cast<bf16*>(a.out_hi);
  bf16* olo = static_cast<bf16*>(a.out_lo);
  for (int g = tid; g < S / 8; g += nt) {
    const int i = g * 8, r = i / N, c = i - r * N;
    const long long row = static_cast<long long>(blockIdx.x) * a.n + r;
    const long long o = row * N + c;
    float y[8], out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float pre = fmaf(a.gamma[c + j], (x[i + j] - mean) * rstd, a.beta[c + j]);
      y[j] = act_fwd(pre, a.act);
      out[j] = y[j];
    }
    if (a.mul) {
      const float* m = a.mul + (row % a.mul_rows) * N + c;
#pragma unroll
      for (int j = 0; j < 8; ++j) out[j] *= m[j];
    }
    const uint32_t bits = slab_keep_bits(a, row, c, thr);
    if (thr < 65536u) {
#pragma unroll
      for (int j = 0; j < 8; ++j) out[j] = ((bits >> j) & 1u) ? out[j] * inv_keep : 0.f;
    }
    if (a.y) {
      *reinterpret_cast<float4*>(a.y + o) = make_float4(y[0], y[1], y[2], y[3]);
      *reinterpret_cast<float4*>(a.y + o + 4) = make_float4(y[4], y[5], y[6], y[7]);
    }
    if (a.out_f32) {
      *reinterpret_cast<float4*>(a.out_f32 + o) = make_float4(out[0], out[1], out[2], out[3]);
      *reinterpret_cast<float4*>(a.out_f32 + o + 4) = make_float4(out[4], out[5], out[6], out[7]);
    }
    if (ohi) st8_planes(ohi, olo, o, out);
  }
}

__global__ void __launch_bounds__(SL_THREADS) slab_ln_bwd_kernel(const VqaSlabLn a) {
  extern __shared__ float dp[];   // d loss / d pre-activation of the slab
  __shared__ float red[33];
  const int N = a.N, S = a.n * a.N, tid = threadIdx.x, nt = blockDim.x;
  const long long base = static_cast<long long>(blockIdx.x) * S;
  const float mean = a.mean[blockIdx.x], rstd = a.rstd[blockIdx.x];
  const uint32_t thr = keep_threshold(a.keep);
  const float inv_keep = 1.0f / a.keep;
  float s1 = 0.f, s2 = 0.f;
  for (int g = tid; g < S / 8; g += nt) {
    const int i = g * 8, r = i / N, c = i - r * N;
    const long long row = static_cast<long long>(blockIdx.x) * a.n + r;
    const long long o = row * N + c;
    const uint32_t bits = slab_keep_bits(a, row, c, thr);
    const float* m = a.mul ? a.mul + (row % a.mul_rows) * N + c : nullptr;
    float dm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (a.z[base + i + j] - mean) * rstd;
      const float pre = fmaf(a.gamma[c + j], xh, a.beta[c + j]);
      const float y = act_fwd(pre, a.act);
      float up = a.dout[o + j];
      if (a.dout2) up += a.dout2[o + j];
      if (thr < 65536u) up = ((bits >> j) & 1u) ? up * inv_keep : 0.f;
      dm[j] = up * y;                                  // d loss / d mul
      const float dy = m ? up * m[j] : up;
      const float dpre = a.act == 0 ? (pre > 0.f ? dy : 0.f) : (a.act == 1 ? dy * (1.0f - y * y) : dy);
      dp[i + j] = dpre;
      const float gg = dpre * a.gamma[c + j];
      s1 += gg;
      s2 = fmaf(gg, xh, s2);
    }
    if (a.dmul) {
      *reinterpret_cast<float4*>(a.dmul + o) = make_float4(dm[0], dm[1], dm[2], dm[3]);
      *reinterpret_cast<float4*>(a.dmul + o + 4) = make_float4(dm[4], dm[5], dm[6], dm[7]);
    }
  }
  const float m1 = block_sum(s1, red) / static_cast<float>(S);
  const float m2 = block_sum(s2, red) / static_cast<float>(S);
  bf16* dhi = static_cast<bf16*>(a.dz_hi);
  bf16* dlo = static_cast<bf16*>(a.dz_lo);
  // column pass: dz = rstd (g - mean(g) - xhat mean(g xhat)), g = dpre gamma; per-slab column sums for the parameters
  for (int c = tid; c < N; c += nt) {
    const float gam = a.gamma[c];
    float sg = 0.f, sb = 0.f, sz = 0.f;
    for (int r = 0; r < a.n; ++r) {
      const int i = r * N + c;
      const float xh = (a.z[base + i] - mean) * rstd;
      const float dpre = dp[i];
      const float dz = rstd * (dpre * gam - m1 - xh * m2);
      sg = fmaf(dpre, xh, sg);
      sb += dpre;
      sz += dz;
      if (a.dz_f32) a.dz_f32[base + i] = dz;
      if (dhi) st_planes(dhi, dlo, base + i, dz);
    }
    if (a.part) {
      float* p = a.part + static_cast<long long>(blockIdx.x) * 3 * N;
      p[c] = sg;
      p[N + c] = sb;
      p[2 * N + c] = sz;
    }
  }
}

// Register-resident variants (the default): thread (rg, cg) owns the 8 columns of column group cg in the rows
// rg, rg + RG, ... of its slab -- at most RPT rows -- so the slab never touches shared memory, every global access is a
// coalesced 32-byte piece per thread, and the per-column sums of the backward pass need no pass of their own (a thread
// already holds whole column segments; RG > 1 adds one fixed-order combine through shared memory).
// blockDim = (N / 8) * RG. ncu of the shared-memory kernels above at cfg4 shapes (profiles/r02_memft_ncu.md): 1 CTA of
// 16 warps per SM for the [36, 1024] slab, 25 % occupancy, 1 TB/s.
template <int RPT>
__global__ void __launch_bounds__(RPT <= 5 ? 256 : 512) slab_ln_fwd_reg_kernel(const VqaSlabLn a, int RG) {
  __shared__ float red[33];
  const int N = a.N, CG = N >> 3, tid = threadIdx.x;
  const int rg = tid / CG, c = (tid - rg * CG) << 3;
  const long long row_base = static_cast<long long>(blockIdx.x) * a.n;
  const float S = static_cast<float>(a.n) * static_cast<float>(N);
  float x[RPT][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int r = rg + i * RG;
    if (r < a.n) {
      const float* z = a.z + (row_base + r) * N + c;
      const float4 u = *reinterpret_cast<const float4*>(z), v = *reinterpret_cast<const float4*>(z + 4);
      x[i][0] = u.x; x[i][1] = u.y; x[i][2] = u.z; x[i][3] = u.w; x[i][4] = v.x; x[i][5] = v.y; x[i][6] = v.z; x[i][7] = v.w;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += x[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[i][j] = 0.f;
    }
  }
  const float mean = block_sum(s, red) / S;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    if (rg + i * RG < a.n) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = x[i][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float var = block_sum(q, red) / S;
  const float rstd = 1.0f / sqrtf(var + 1e-12f);
  if (tid == 0) {
    a.mean[blockIdx.x] = mean;
    a.rstd[blockIdx.x] = rstd;
  }
  float gam[8], bet[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gam[j] = a.gamma[c + j];
    bet[j] = a.beta[c + j];
  }
  const uint32_t thr = keep_threshold(a.keep);
  const float inv_keep = 1.0f / a.keep;
  bf16* ohi = static_cast<bf16*>(a.out_hi);
  bf16* olo = static_cast<bf16*>(a.out_lo);
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int r = rg + i * RG;
    if (r < a.n) {
      const long long row = row_base + r;
      const long long o = row * N + c;
      float y[8], out[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y[j] = act_fwd(fmaf(gam[j], (x[i][j] - mean) * rstd, bet[j]), a.act);
        out[j] = y[j];
      }
      if (a.mul) {
        const float* m = a.mul + (row % a.mul_rows) * N + c;
        const float4 u = *reinterpret_cast<const float4*>(m), v = *reinterpret_cast<const float4*>(m + 4);
        out[0] *= u.x; out[1] *= u.y; out[2] *= u.z; out[3] *= u.w; out[4] *= v.x; out[5] *= v.y; out[6] *= v.z; out[7] *= v.w;
      }
      if (thr < 65536u) {
        const uint32_t bits = slab_keep_bits(a, row, c, thr);
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = ((bits >> j) & 1u) ? out[j] * inv_keep : 0.f;
      }
      if (a.y) {
        *reinterpret_cast<float4*>(a.y + o) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(a.y + o + 4) = make_float4(y[4], y[5], y[6], y[7]);
      }
      if (a.out_f32) {
        *reinterpret_cast<float4*>(a.out_f32 + o) = make_float4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<float4*>(a.out_f32 + o + 4) = make_float4(out[4], out[5], out[6], out[7]);
      }
      if (ohi) st8_planes(ohi, olo, o, out);
    }
  }
}

template <int RPT>
__global__ void __launch_bounds__(512) slab_ln_bwd_reg_kernel(const VqaSlabLn a, int RG) {
  extern __shared__ float comb[];   // RG > 1: [RG][3][N] per-row-group column sums
  __shared__ float red[33];
  const int N = a.N, CG = N >> 3, tid = threadIdx.x;
  const int rg = tid / CG, c = (tid - rg * CG) << 3;
  const long long row_base = static_cast<long long>(blockIdx.x) * a.n;
  const float S = static_cast<float>(a.n) * static_cast<float>(N);
  const float mean = a.mean[blockIdx.x], rstd = a.rstd[blockIdx.x];
  const uint32_t thr = keep_threshold(a.keep);
  const float inv_keep = 1.0f / a.keep;
  float gam[8], bet[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gam[j] = a.gamma[c + j];
    bet[j] = a.beta[c + j];
  }
  float dp[RPT][8];   // d loss / d pre-activation
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int r = rg + i * RG;
#pragma unroll
    for (int j = 0; j < 8; ++j) dp[i][j] = 0.f;
    if (r < a.n) {
      const long long row = row_base + r;
      const long long o = row * N + c;
      const float4 z0 = *reinterpret_cast<const float4*>(a.z + o), z1 = *reinterpret_cast<const float4*>(a.z + o + 4);
      const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
      const float4 u0 = *reinterpret_cast<const float4*>(a.dout + o), u1 = *reinterpret_cast<const float4*>(a.dout + o + 4);
      float up[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
      if (a.dout2) {
        const float4 v0 = *reinterpret_cast<const float4*>(a.dout2 + o), v1 = *reinterpret_cast<const float4*>(a.dout2 + o + 4);
        up[0] += v0.x; up[1] += v0.y; up[2] += v0.z; up[3] += v0.w; up[4] += v1.x; up[5] += v1.y; up[6] += v1.z; up[7] += v1.w;
      }
      if (thr < 65536u) {
        const uint32_t bits = slab_keep_bits(a, row, c, thr);
#pragma unroll
        for (int j = 0; j < 8; ++j) up[j] = ((bits >> j) & 1u) ? up[j] * inv_keep : 0.f;
      }
      float mu[8];
      if (a.mul) {
        const float* m = a.mul + (row % a.mul_rows) * N + c;
        const float4 m0 = *reinterpret_cast<const float4*>(m), m1 = *reinterpret_cast<const float4*>(m + 4);
        mu[0] = m0.x; mu[1] = m0.y; mu[2] = m0.z; mu[3] = m0.w; mu[4] = m1.x; mu[5] = m1.y; mu[6] = m1.z; mu[7] = m1.w;
      }
      float dm[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (zz[j] - mean) * rstd;
        const float pre = fmaf(gam[j], xh, bet[j]);
        const float y = act_fwd(pre, a.act);
        dm[j] = up[j] * y;
        const float dy = a.mul ? up[j] * mu[j] : up[j];
        const float dpre = a.act == 0 ? (pre > 0.f ? dy : 0.f) : (a.act == 1 ? dy * (1.0f - y * y) : dy);
        dp[i][j] = dpre;
        const float gg = dpre * gam[j];
        s1 += gg;
        s2 = fmaf(gg, xh, s2);
      }
      if (a.dmul) {
        *reinterpret_cast<float4*>(a.dmul + o) = make_float4(dm[0], dm[1], dm[2], dm[3]);
        *reinterpret_cast<float4*>(a.dmul + o + 4) = make_float4(dm[4], dm[5], dm[6], dm[7]);
      }
    }
  }
  const float m1 = block_sum(s1, red) / S;
  const float m2 = block_sum(s2, red) / S;
  bf16* dhi = static_cast<bf16*>(a.dz_hi);
  bf16* dlo = static_cast<bf16*>(a.dz_lo);
  float sg[8], sb[8], sz[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sg[j] = sb[j] = sz[j] = 0.f;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int r = rg + i * RG;
    if (r < a.n) {
      const long long o = (row_base + r) * N + c;
      const float4 z0 = *reinterpret_cast<const float4*>(a.z + o), z1 = *reinterpret_cast<const float4*>(a.z + o + 4);
      const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
      float dz[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (zz[j] - mean) * rstd;
        dz[j] = rstd * (dp[i][j] * gam[j] - m1 - xh * m2);
        sg[j] = fmaf(dp[i][j], xh, sg[j]);
        sb[j] += dp[i][j];
        sz[j] += dz[j];
      }
      if (a.dz_f32) {
        *reinterpret_cast<float4*>(a.dz_f32 + o) = make_float4(dz[0], dz[1], dz[2], dz[3]);
        *reinterpret_cast<float4*>(a.dz_f32 + o + 4) = make_float4(dz[4], dz[5], dz[6], dz[7]);
      }
      if (dhi) st8_planes(dhi, dlo, o, dz);
    }
  }
  if (a.part) {
    float* p = a.part + static_cast<long long>(blockIdx.x) * 3 * N;
    if (RG > 1) {
      float* mine = comb + static_cast<long long>(rg) * 3 * N;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mine[c + j] = sg[j];
        mine[N + c + j] = sb[j];
        mine[2 * N + c + j] = sz[j];
      }
      __syncthreads();
      if (rg == 0) {
        for (int g2 = 1; g2 < RG; ++g2) {   // fixed order
          const float* o2 = comb + static_cast<long long>(g2) * 3 * N;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            sg[j] += o2[c + j];
            sb[j] += o2[N + c + j];
            sz[j] += o2[2 * N + c + j];
          }
        }
      }
    }
    if (rg == 0) {
      *reinterpret_cast<float4*>(p + c) = make_float4(sg[0], sg[1], sg[2], sg[3]);
      *reinterpret_cast<float4*>(p + c + 4) = make_float4(sg[4], sg[5], sg[6], sg[7]);
      *reinterpret_cast<float4*>(p + N + c) = make_float4(sb[0], sb[1], sb[2], sb[3]);
      *reinterpret_cast<float4*>(p + N + c + 4) = make_float4(sb[4], sb[5], sb[6], sb[7]);
      *reinterpret_cast<float4*>(p + 2 * N + c) = make_float4(sz[0], sz[1], sz[2], sz[3]);
      *reinterpret_cast<float4*>(p + 2 * N + c + 4) = make_float4(sz[4], sz[5], sz[6], sz[7]);
    }
  }
}

// launch plan of the register-resident kernels: rows per thread and row groups, or 0 = use the shared-memory kernels
inline int slab_reg_plan(int n, int N, int max_threads, int* rg_out) {
  const int CG = N >> 3;
  if (getenv("VQA_SLAB_SMEM")) return 0;
  if (n <= 5 && CG <= 256) { *rg_out = 1; return 5; }
  const int RG = (n + 8) / 9;
  if (CG * RG <= max_threads) { *rg_out = RG; return 9; }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// spatial attention + attended pooling: one CTA per image, the kinds one after the other
// ---------------------------------------------------------------------------------------------------------------------
template <int NE>   // NE >= n entries: the unrolled per-entry accumulators (5 in the reference, datasets/dataset_vlmap.py:11-14)
__global__ void __launch_bounds__(256, 3) spat_attn_fwd_kernel(const VqaSpatAttn a) {
  extern __shared__ float sm[];
  const int D = a.D, K = a.K, n = a.n, Dv = a.Dv, B = a.B;
  float* hq_s = sm;             // [n][D]
  float* w_s = hq_s + n * D;    // [D]
  float* sc = w_s + D;          // [n][K] scores, then attention weights
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int nb = a.num_boxes[b];
  nb = nb < 0 ? 0 : (nb > K ? K : nb);
  const float inv_keep = 1.0f / a.keep;
  const uint32_t thr = keep_threshold(a.keep);
  const bf16* hv_hi = static_cast<const bf16*>(a.hv_hi);
  const bf16* hv_lo = static_cast<const bf16*>(a.hv_lo);
  bf16* p_hi = static_cast<bf16*>(a.pooled_hi);
  bf16* p_lo = static_cast<bf16*>(a.pooled_lo);
  for (int i = tid; i < D; i += 256) w_s[i] = a.att_w[i];
  const float bias = a.att_b[0];
  {
    const int kind = blockIdx.y;   // one CTA per (image, kind)
    const long long row0 = (static_cast<long long>(kind) * B + b) * n;
    for (int i = tid; i < n * D; i += 256) hq_s[i] = a.hq[row0 * D + i];
    __syncthreads();
    for (int k = warp; k < K; k += 8) {
      float acc[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) acc[e] = 0.f;
      const long long hv_base = (static_cast<long long>(b) * K + k) * D;
      for (int d8 = lane; d8 < D / 8; d8 += 32) {
        float hw[8];
        ld8_planes(hv_hi, hv_lo, hv_base + d8 * 8, hw);
#pragma unroll
        for (int j = 0; j < 8; ++j) hw[j] *= w_s[d8 * 8 + j];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          if (e < n) {
            uint32_t bits = 0xFFu;
            if (thr < 65536u) {
              const unsigned long long group =
                  ((static_cast<unsigned long long>(b) * n + e) * K + k) * (D / 8) + d8;
              bits = philox_keep_bits(philox4x32_10(group, a.site0 + kind, a.seed, a.step), thr);
              // the backward pass reads the byte back instead of drawing again (a draw is ~150 integer instructions)
              if (a.keep_bits) a.keep_bits[static_cast<unsigned long long>(kind) * B * n * K * (D / 8) + group] = static_cast<uint8_t>(bits);
            }
            const float* hq = hq_s + e * D + d8 * 8;
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) t = ((bits >> j) & 1u) ? fmaf(hw[j], hq[j], t) : t;
            acc[e] += t;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        if (e < n) {
          const float t = warp_sum(acc[e]);
          if (lane == 0) sc[e * K + k] = fmaf(t, inv_keep, bias);
        }
      }
    }
    __syncthreads();
    if (warp < n) {   // masked softmax of entry `warp` over the valid boxes (exact zeros beyond num_boxes)
      float* s = sc + warp * K;
      float m = -INFINITY;
      for (int k = lane; k < nb; k += 32) m = fmaxf(m, s[k]);
      m = warp_max(m);
      float z = 0.f;
      for (int k = lane; k < nb; k += 32) z += __expf(s[k] - m);
      z = warp_sum(z);
      const float inv = nb > 0 ? 1.0f / z : 0.f;
      for (int k = lane; k < K; k += 32) {
        const float p = k < nb ? __expf(s[k] - m) * inv : 0.f;
        s[k] = p;
        a.att[(row0 + warp) * K + k] = p;
      }
    }
    __syncthreads();
    for (int c = tid * 4; c < Dv; c += 1024) {
      float4 acc[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) acc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < nb; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(a.v + (static_cast<long long>(b) * K + k) * Dv + c);
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          if (e < n) {
            const float p = sc[e * K + k];
            acc[e].x = fmaf(p, v.x, acc[e].x);
            acc[e].y = fmaf(p, v.y, acc[e].y);
            acc[e].z = fmaf(p, v.z, acc[e].z);
            acc[e].w = fmaf(p, v.w, acc[e].w);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        if (e < n) {
          const long long o = (row0 + e) * Dv + c;
          if (a.pooled) *reinterpret_cast<float4*>(a.pooled + o) = acc[e];
          if (p_hi) {
            st_planes(p_hi, p_lo, o, acc[e].x);
            st_planes(p_hi, p_lo, o + 1, acc[e].y);
            st_planes(p_hi, p_lo, o + 2, acc[e].z);
            st_planes(p_hi, p_lo, o + 3, acc[e].w);
          }
        }
      }
    }
  }
}

template <int NE>
__global__ void __launch_bounds__(256, 2) spat_attn_bwd_kernel(const VqaSpatAttn a) {
  extern __shared__ float sm[];
  const int D = a.D, K = a.K, n = a.n, Dv = a.Dv, B = a.B;
  float* hq_s = sm;               // [n][D]
  float* w_s = hq_s + n * D;      // [D]
  float* dw_s = w_s + D;          // [D] d att_w of this image
  float* at = dw_s + D;           // [n][K] attention
  float* ds = at + n * K;         // [n][K] d a, then d score
  float* dpool = ds + n * K;      // [n][Dv]
  __shared__ float red[33];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int nb = a.num_boxes[b];
  nb = nb < 0 ? 0 : (nb > K ? K : nb);
  const float inv_keep = 1.0f / a.keep;
  const uint32_t thr = keep_threshold(a.keep);
  const bf16* hv_hi = static_cast<const bf16*>(a.hv_hi);
  const bf16* hv_lo = static_cast<const bf16*>(a.hv_lo);
  for (int i = tid; i < D; i += 256) {
    w_s[i] = a.att_w[i];
    dw_s[i] = 0.f;
  }
  float dbias = 0.f;
  const int kind = blockIdx.y;   // one CTA per (image, kind): d Hv of each kind goes to its own plane [kinds, B, K, D]
  float* d_hv = a.d_hv + static_cast<long long>(kind) * B * K * D;
  {
    const long long row0 = (static_cast<long long>(kind) * B + b) * n;
    for (int i = tid; i < n * D; i += 256) hq_s[i] = a.hq[row0 * D + i];
    for (int i = tid; i < n * K; i += 256) at[i] = a.att[row0 * K + i];
    for (int i = tid; i < n * Dv; i += 256) dpool[i] = a.d_pooled[row0 * Dv + i];
    __syncthreads();
    // d a[e, k] = <d pooled[e], V[b, k]>
    for (int k = warp; k < K; k += 8) {
      float acc[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) acc[e] = 0.f;
      if (k < nb) {
        for (int c = lane * 4; c < Dv; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(a.v + (static_cast<long long>(b) * K + k) * Dv + c);
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            if (e < n) {
              const float4 g = *reinterpret_cast<const float4*>(dpool + e * Dv + c);
              acc[e] = fmaf(g.x, v.x, fmaf(g.y, v.y, fmaf(g.z, v.z, fmaf(g.w, v.w, acc[e]))));
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        if (e < n) {
          const float t = warp_sum(acc[e]);
          if (lane == 0) ds[e * K + k] = t;
        }
      }
    }
    __syncthreads();
    if (warp < n) {   // softmax backward of entry `warp`: d s = a (d a - <a, d a>)
      float dot = 0.f;
      for (int k = lane; k < nb; k += 32) dot = fmaf(at[warp * K + k], ds[warp * K + k], dot);
      dot = warp_sum(dot);
      float sb = 0.f;
      for (int k = lane; k < K; k += 32) {
        const float v = k < nb ? at[warp * K + k] * (ds[warp * K + k] - dot) : 0.f;
        ds[warp * K + k] = v;
        sb += v;
      }
      sb = warp_sum(sb);
      if (lane == 0) red[warp] = sb;   // d att_b share of this entry (zero up to rounding: soft-max shift invariance)
    }
    __syncthreads();
    if (tid == 0)
      for (int e = 0; e < n; ++e) dbias += red[e];
    // thread = 4 consecutive feature columns, all boxes: no reduction across threads for d Hq / d Hv / d w
    for (int d0 = tid * 4; d0 < D; d0 += 1024) {
      float dq[NE][4];
      float dwa[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < NE; ++e)
#pragma unroll
        for (int j = 0; j < 4; ++j) dq[e][j] = 0.f;
      const float w4[4] = {w_s[d0], w_s[d0 + 1], w_s[d0 + 2], w_s[d0 + 3]};
      for (int k = 0; k < K; ++k) {
        const long long o = (static_cast<long long>(b) * K + k) * D + d0;
        float dhv[4] = {0.f, 0.f, 0.f, 0.f};
        if (k < nb) {
          float hv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) hv[j] = ld_planes(hv_hi, hv_lo, o + j);
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            if (e < n) {
              uint32_t bits = 0xFu;
              if (thr < 65536u) {
                const unsigned long long group =
                    ((static_cast<unsigned long long>(b) * n + e) * K + k) * (D / 8) + (d0 >> 3);
                const uint32_t all = a.keep_bits ? a.keep_bits[static_cast<unsigned long long>(kind) * B * n * K * (D / 8) + group]
                                                 : philox_keep_bits(philox4x32_10(group, a.site0 + kind, a.seed, a.step), thr);
                bits = (all >> (d0 & 4)) & 0xFu;
              }
              const float dsv = ds[e * K + k] * inv_keep;
              const float* hq = hq_s + e * D + d0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if ((bits >> j) & 1u) {
                  dhv[j] = fmaf(dsv * hq[j], w4[j], dhv[j]);
                  dq[e][j] = fmaf(dsv * hv[j], w4[j], dq[e][j]);
                  dwa[j] = fmaf(dsv * hv[j], hq[j], dwa[j]);
                }
              }
            }
          }
        }
        *reinterpret_cast<float4*>(d_hv + o) = make_float4(dhv[0], dhv[1], dhv[2], dhv[3]);
      }
#pragma unroll
      for (int e = 0; e < NE; ++e)
        if (e < n)
          *reinterpret_cast<float4*>(a.d_hq + (row0 + e) * D + d0) = make_float4(dq[e][0], dq[e][1], dq[e][2], dq[e][3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) dw_s[d0 + j] += dwa[j];
    }
  }
  __syncthreads();
  float* p = a.part + (static_cast<long long>(kind) * B + b) * (D + 8);
  for (int i = tid; i < D; i += 256) p[i] = dw_s[i];
  if (tid == 0) p[D] = dbias;
}

// ---------------------------------------------------------------------------------------------------------------------
// masked softmax cross-entropy with top-1 / top-k, one CTA per row; finalize: one CTA
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_ce_kernel(const VqaSoftmaxCe a) {
  extern __shared__ float x[];   // [A]
  __shared__ float red[33];
  const int A = a.A, tid = threadIdx.x;
  const long long row = blockIdx.x;
  const int rows_per_head = a.B * a.n;
  const int head = static_cast<int>(row / rows_per_head);
  const int rin = static_cast<int>(row - static_cast<long long>(head) * rows_per_head);
  const int b = rin / a.n, e = rin - b * a.n;
  const int* num = a.num[head];
  // count of valid entries of this head (tf.sequence_mask(num, maxlen = n))
  float cnt = 0.f;
  for (int i = tid; i < a.B; i += 256) {
    const int v = num[i];
    cnt += static_cast<float>(v < 0 ? 0 : (v > a.n ? a.n : v));
  }
  cnt = block_sum(cnt, red);
  const bool valid = e < num[b];
  const float* l = a.logit + row * A;
  float m = -INFINITY;
  for (int j = tid; j < A; j += 256) {
    const float v = l[j];
    x[j] = v;
    m = fmaxf(m, v);
  }
  m = block_max(m, red);
  int tgt = a.fills[row];
  tgt = tgt < 0 ? 0 : (tgt >= A ? A - 1 : tgt);
  const float lt = x[tgt];
  float z = 0.f, rank = 0.f;
  for (int j = tid; j < A; j += 256) {
    const float v = x[j];
    z += __expf(v - m);
    rank += (v > lt || (v == lt && j < tgt)) ? 1.f : 0.f;
  }
  z = block_sum(z, red);
  rank = block_sum(rank, red);
  if (tid == 0) {
    float* st = a.stats + row * 4;
    st[0] = logf(z) + m - lt;
    st[1] = rank < 0.5f ? 1.f : 0.f;
    st[2] = rank < static_cast<float>(a.top_k) - 0.5f ? 1.f : 0.f;
    st[3] = valid ? 1.f : 0.f;
  }
  if (a.d_logit || a.d_hi) {
    const float norm = a.count[head] > 0.f ? a.count[head] : cnt;
    const float scale = (valid && norm > 0.f) ? a.loss_scale / norm : 0.f;
    const float inv = 1.0f / z;
    bf16* hi = static_cast<bf16*>(a.d_hi);
    bf16* lo = static_cast<bf16*>(a.d_lo);
    for (int j = tid; j < A; j += 256) {
      const float g = (__expf(x[j] - m) * inv - (j == tgt ? 1.f : 0.f)) * scale;
      if (a.d_logit) a.d_logit[row * A + j] = g;
      if (hi) st_planes(hi, lo, row * A + j, g);
    }
  }
}

__global__ void __launch_bounds__(256) softmax_ce_finalize_kernel(const float* __restrict__ stats, int heads,
                                                                  int rows_per_head, float* __restrict__ report) {
  __shared__ float red[33];
  float total = 0.f;
  for (int h = 0; h < heads; ++h) {
    float ce = 0.f, t1 = 0.f, tk = 0.f, cnt = 0.f;
    for (int r = threadIdx.x; r < rows_per_head; r += 256) {
      const float4 s = *reinterpret_cast<const float4*>(stats + (static_cast<long long>(h) * rows_per_head + r) * 4);
      ce = fmaf(s.x, s.w, ce);
      t1 = fmaf(s.y, s.w, t1);
      tk = fmaf(s.z, s.w, tk);
      cnt += s.w;
    }
    ce = block_sum(ce, red);
    t1 = block_sum(t1, red);
    tk = block_sum(tk, red);
    cnt = block_sum(cnt, red);
    if (threadIdx.x == 0) {
      report[3 * h] = ce / cnt;
      report[3 * h + 1] = t1 / cnt;
      report[3 * h + 2] = tk / cnt;
    }
    total += ce / cnt;
  }
  if (threadIdx.x == 0) report[3 * heads] = total;
}

// ---------------------------------------------------------------------------------------------------------------------
// small element-wise kernels
// ---------------------------------------------------------------------------------------------------------------------
__global__ void pad_planes_kernel(const float* __restrict__ src, long long rows, int cols, int boxes,
                                  bf16* __restrict__ hi, bf16* __restrict__ lo, int ld) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  const long long r = i / ld;
  const int c = static_cast<int>(i - r * ld);
  float v = 0.f;
  if (boxes) {
    const float* bx = src + r * 4;
    if (c < 4) v = bx[c];
    else if (c == 4) v = bx[2] - bx[0];
    else if (c == 5) v = bx[3] - bx[1];
  } else if (c < cols) {
    v = src[r * cols + c];
  }
  st_planes(hi, lo, i, v);
}

__global__ void wordset_fwd_kernel(const float* __restrict__ map, const int* __restrict__ ids, int W, int num_ws,
                                   float* __restrict__ y, bf16* __restrict__ hi, bf16* __restrict__ lo, int ld) {
  const long long row = blockIdx.x;
  int id = ids[row];
  id = id < 0 ? 0 : (id >= num_ws ? num_ws - 1 : id);
  for (int c = threadIdx.x; c < ld; c += blockDim.x) {
    const float v = c < W ? tanhf(map[static_cast<long long>(id) * W + c]) : 0.f;
    if (c < W) y[row * W + c] = v;
    st_planes(hi, lo, row * ld + c, v);
  }
}

__global__ void wordset_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const int* __restrict__ ids,
                                   int W, int num_ws, float* __restrict__ d_map) {
  const long long row = blockIdx.x;
  int id = ids[row];
  id = id < 0 ? 0 : (id >= num_ws ? num_ws - 1 : id);
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    const float v = y[row * W + c];
    atomicAdd(d_map + static_cast<long long>(id) * W + c, dy[row * W + c] * (1.0f - v * v));
  }
}

// Weight gradient of a layer whose input has F <= 8 features (the 6-d box features): out[f, c] = sum_r feat[r, f] dz[r, c].
// As a tensor-core GEMM this is one 64-row tile with K = rows (18432): 16 CTAs walking 288 k-blocks, 70 us. Here every
// CTA takes a slice of rows, a thread four columns; per-CTA partials [ctas, F, N] are summed in fixed order afterwards.
constexpr int FW_THREADS = 256;
__global__ void __launch_bounds__(FW_THREADS) feat_wgrad_kernel(const float* __restrict__ feat, int F, int boxes,
                                                                const bf16* __restrict__ dz_hi, const bf16* __restrict__ dz_lo,
                                                                long long rows, int N, float* __restrict__ part) {
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
  for (int c = threadIdx.x * 4; c < N; c += FW_THREADS * 4) {
    float acc[8][4];
#pragma unroll
    for (int f = 0; f < 8; ++f)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[f][j] = 0.f;
    for (long long r = r0; r < r1; ++r) {
      float x[8];
      if (boxes) {   // (x0, y0, x1, y1, x1 - x0, y1 - y0) of a normalised box
        const float4 b = *reinterpret_cast<const float4*>(feat + r * 4);
        x[0] = b.x; x[1] = b.y; x[2] = b.z; x[3] = b.w; x[4] = b.z - b.x; x[5] = b.w - b.y; x[6] = x[7] = 0.f;
      } else {
#pragma unroll
        for (int f = 0; f < 8; ++f) x[f] = f < F ? feat[r * F + f] : 0.f;
      }
      float d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = ld_planes(dz_hi, dz_lo, r * N + c + j);
#pragma unroll
      for (int f = 0; f < 8; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[f][j] = fmaf(x[f], d[j], acc[f][j]);
    }
    for (int f = 0; f < F; ++f)
      *reinterpret_cast<float4*>(part + (static_cast<long long>(blockIdx.x) * F + f) * N + c) =
          make_float4(acc[f][0], acc[f][1], acc[f][2], acc[f][3]);
  }
}

template <typename Kern>
VqaStatus ensure_smem(Kern kern, size_t bytes, const char* what) {
  if (bytes > 48 * 1024) {
    if (bytes > 227 * 1024) return set_error(VQA_ERR_BAD_SHAPE, "%s: needs %zu bytes of shared memory (limit 227 KB)", what, bytes);
    VQA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  }
  return VQA_OK;
}

VqaStatus check_slab(const VqaSlabLn* a, const char* who) {
  if (!a || !a->z || !a->gamma || !a->beta || !a->mean || !a->rstd) return set_error(VQA_ERR_BAD_ARG, "%s: null argument", who);
  if (a->slabs < 0 || a->n <= 0 || a->N <= 0 || (a->N & 7) || static_cast<long long>(a->n) * a->N > 49152)
    return set_error(VQA_ERR_BAD_SHAPE, "%s: n * N = %lld (N a multiple of 8, n * N <= 49152)", who, static_cast<long long>(a->n) * a->N);
  if (!(a->keep > 0.f && a->keep <= 1.f)) return set_error(VQA_ERR_BAD_ARG, "%s: keep must be in (0, 1]", who);
  if (a->keep < 1.f && a->rows_per_site <= 0) return set_error(VQA_ERR_BAD_ARG, "%s: rows_per_site", who);
  if (a->mul && a->mul_rows <= 0) return set_error(VQA_ERR_BAD_ARG, "%s: mul_rows", who);
  return VQA_OK;
}

// ---- GRU sequence operator: workspace carve ----
struct GruWs {
  Planes wg, wc;          // weight planes [(W + L), 2L], [(W + L), L]
  bf16* pack;             // packed recurrent weights of the persistent forward kernel
  Planes e;               // [T*B, Wp]
  float* xg; float* xc;   // [T*B, 2L], [T*B, L]
  float* h_f32; Planes h; Planes rh;
  float* r; float* u; float* c;
  float* g_pre; float* c_pre;
  unsigned int* counter;
  Planes dG, dC; float* dG_f32; float* dC_f32;
  float* du; float* dh_part; float* dRH; float* dh[2];
  float* bias_part; float* dE;
  int Wp;
  size_t bytes;
};

GruWs carve_gru(const VqaGruSeq& a, void* base) {
  GruWs w{};
  const size_t B = a.B, T = a.T, L = a.L, W = a.W;
  const bool fp32 = a.precision == VQA_PREC_FP32;
  w.Wp = (a.W + 7) & ~7;
  const size_t Wp = w.Wp;
  size_t off = 0;
  auto take = [&](size_t bytes) -> void* {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += (bytes + 255) & ~static_cast<size_t>(255);
    return p;
  };
  auto planes = [&](size_t elems) {
    Planes p;
    p.hi = static_cast<bf16*>(take(elems * 2));
    p.lo = fp32 ? static_cast<bf16*>(take(elems * 2)) : nullptr;
    return p;
  };
  auto f32 = [&](size_t elems) { return static_cast<float*>(take(elems * 4)); };
  w.wg = planes((W + L) * 2 * L);
  w.wc = planes((W + L) * L);
  w.pack = static_cast<bf16*>(take(gru_pack_elems(a.L) * 2));
  w.e = planes(T * B * Wp);
  w.xg = f32(T * B * 2 * L);
  w.xc = f32(T * B * L);
  w.h_f32 = f32((T + 1) * B * L);
  w.h = planes((T + 1) * B * L);
  w.rh = planes(T * B * L);
  w.r = f32(T * B * L);
  w.u = f32(T * B * L);
  w.c = f32(T * B * L);
  w.g_pre = f32(B * 2 * L);
  w.c_pre = f32(B * L);
  w.counter = static_cast<unsigned int*>(take(64 * 4));
  w.dG = planes(T * B * 2 * L);
  w.dC = planes(T * B * L);
  w.dG_f32 = f32(T * B * 2 * L);   // (read by the per-step path only: fp32 mode, or L not a multiple of 64)
  w.dC_f32 = f32(T * B * L);
  w.du = f32(B * L);
  w.dh_part = f32(B * L);
  w.dRH = f32(B * L);
  w.dh[0] = f32(B * L);
  w.dh[1] = f32(B * L);
  w.bias_part = f32(gru_bias_part_floats(a.B, a.L));
  w.dE = f32(T * B * Wp);
  w.bytes = off;
  return w;
}

VqaStatus check_gru(const VqaGruSeq* a, const char* who, bool need_ws) {
  if (!a) return set_error(VQA_ERR_BAD_ARG, "%s: null argument", who);
  if (a->B <= 0 || a->T <= 0 || a->L <= 0 || a->W <= 0 || a->Vq <= 0 || (a->L & 7))
    return set_error(VQA_ERR_BAD_SHAPE, "%s: bad dimension", who);
  if (a->precision != VQA_PREC_BF16 && a->precision != VQA_PREC_FP32) return set_error(VQA_ERR_BAD_ARG, "%s: precision", who);
  if (need_ws) {
    if (!a->ws || (reinterpret_cast<uintptr_t>(a->ws) & 255)) return set_error(VQA_ERR_WORKSPACE, "%s: workspace missing / not 256-byte aligned", who);
    if (a->ws_bytes < carve_gru(*a, nullptr).bytes) return set_error(VQA_ERR_WORKSPACE, "%s: workspace too small", who);
  }
  return VQA_OK;
}

struct G {   // GEMM builder over Planes (conventions of VqaGemmDesc)
  VqaGemmDesc d{};
  G(long long M, long long N, long long K) { d.M = static_cast<int>(M); d.N = static_cast<int>(N); d.K = static_cast<int>(K); }
  G& a(const Planes& p, long long off, long long ld, bool mn) { d.a_hi = p.hi + off; d.a_lo = p.lo ? p.lo + off : nullptr; d.lda = ld; d.a_mn_major = mn; return *this; }
  G& b(const Planes& p, long long off, long long ld, bool mn) { d.b_hi = p.hi + off; d.b_lo = p.lo ? p.lo + off : nullptr; d.ldb = ld; d.b_mn_major = mn; return *this; }
  G& bias(const float* p) { d.bias = p; return *this; }
  G& addend(const float* p, long long ld) { d.addend = p; d.ld_addend = ld; return *this; }
  G& f32(float* p, long long ld) { d.out_f32 = p; d.ld_f32 = ld; return *this; }
  G& bf(bf16* p, long long ld) { d.out_hi = p; d.ld_bf = ld; return *this; }   // one bf16 output plane
  VqaStatus run(VqaOps ops, cudaStream_t s) { return gemm_launch(d, ops->num_sms, s, &ops->gemm_ctx, 0); }
};

}  // namespace

}  // namespace vqa

using namespace vqa;

extern "C" {

VQA_API VqaStatus vqa_ops_create(VqaOps* out) {
  if (!out) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(VQA_ERR_NO_DEVICE, "vqa_ops_create: no CUDA device (this library has no CPU path)");
  }
  int dev = 0;
  VQA_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  VQA_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return set_error(VQA_ERR_NO_DEVICE, "vqa_ops_create: device %d is sm_%d%d; this library is sm_100a only", dev, prop.major, prop.minor);
  VqaOps_t* o = new (std::nothrow) VqaOps_t();
  if (!o) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_create: out of host memory");
  o->num_sms = prop.multiProcessorCount;
  o->sem = nullptr;
  o->scratch = nullptr;
  o->scratch_floats = 32LL * 8192 + 16384;
  if (cudaMalloc(&o->sem, sizeof(unsigned int) * kSemRegions * kSemElems) != cudaSuccess ||
      cudaMalloc(&o->scratch, sizeof(float) * o->scratch_floats) != cudaSuccess ||
      cudaMemset(o->sem, 0, sizeof(unsigned int) * kSemRegions * kSemElems) != cudaSuccess) {
    cudaFree(o->sem);
    cudaFree(o->scratch);
    delete o;
    return set_cuda_error(cudaGetLastError(), "vqa_ops_create: device allocation");
  }
  o->gemm_ctx.sem = o->sem;
  o->gemm_ctx.regions = kSemRegions;
  o->gemm_ctx.region_elems = kSemElems;
  o->gemm_ctx.next_region = 0;
  *out = o;
  return VQA_OK;
}

VQA_API VqaStatus vqa_ops_destroy(VqaOps ops) {
  if (ops) {
    cudaFree(ops->sem);
    cudaFree(ops->scratch);
    delete ops;
  }
  return VQA_OK;
}

VQA_API VqaStatus vqa_ops_gemm(VqaOps ops, const VqaGemmDesc* d, int32_t narrow, void* stream) {
  if (!ops || !d) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_gemm: null argument");
  return gemm_launch(*d, ops->num_sms, static_cast<cudaStream_t>(stream), &ops->gemm_ctx, narrow);
}

VQA_API VqaStatus vqa_ops_colsum(VqaOps ops, const float* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream) {
  if (!ops || !x || !out) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_colsum: null argument");
  if (cols > 8192) return set_error(VQA_ERR_BAD_SHAPE, "vqa_ops_colsum: at most 8192 columns");
  return colsum_launch(x, rows, cols, ld, out, ops->scratch, static_cast<cudaStream_t>(stream));
}

VQA_API VqaStatus vqa_ops_adam(VqaOps ops, float* param, const float* grad, float* m, float* v, int64_t n, float lr,
                               float beta1, float beta2, float eps, float clip_norm, int64_t t, float* grad_norm_out,
                               void* stream) {
  if (!ops || !param || !grad || !m || !v) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_adam: null argument");
  return adam_step_launch(param, grad, m, v, n, lr, beta1, beta2, eps, clip_norm, t, grad_norm_out, ops->scratch,
                          ops->num_sms, static_cast<cudaStream_t>(stream));
}

VQA_API VqaStatus vqa_ops_split_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld, void* hi, void* lo,
                                     int64_t ld_out, void* stream) {
  if (!src || !hi) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_split_bf16: null argument");
  return split_bf16_launch(src, rows, cols, ld, static_cast<bf16*>(hi), static_cast<bf16*>(lo), ld_out,
                           static_cast<cudaStream_t>(stream));
}

VQA_API VqaStatus vqa_ops_dropout_mask(uint8_t* out, int64_t n, float keep, uint64_t seed, uint64_t step, uint32_t site,
                                       void* stream) {
  if (!out) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_dropout_mask: null argument");
  return dropout_mask_launch(out, n, keep, seed, step, site, static_cast<cudaStream_t>(stream));
}

VQA_API VqaStatus vqa_ops_slab_ln_fwd(VqaOps ops, const VqaSlabLn* a, void* stream) {
  if (!ops) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_slab_ln_fwd: null context");
  VQA_TRY(check_slab(a, "vqa_ops_slab_ln_fwd"));
  if (a->slabs == 0) return VQA_OK;
  int RG = 1;
  const int rpt = slab_reg_plan(a->n, a->N, 512, &RG);
  if (rpt == 5) {
    slab_ln_fwd_reg_kernel<5><<<a->slabs, (a->N >> 3) * RG, 0, static_cast<cudaStream_t>(stream)>>>(*a, RG);
  } else if (rpt == 9) {
    slab_ln_fwd_reg_kernel<9><<<a->slabs, (a->N >> 3) * RG, 0, static_cast<cudaStream_t>(stream)>>>(*a, RG);
  } else {
    const size_t smem = sizeof(float) * a->n * a->N;
    VQA_TRY(ensure_smem(slab_ln_fwd_kernel, smem, "vqa_ops_slab_ln_fwd"));
    slab_ln_fwd_kernel<<<a->slabs, SL_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(*a);
  }
  VQA_LAUNCH_CHECK("slab_ln_fwd");
  return VQA_OK;
}

VQA_API VqaStatus vqa_ops_slab_ln_bwd(VqaOps ops, const VqaSlabLn* a, void* stream) {
  if (!ops) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_slab_ln_bwd: null context");
  VQA_TRY(check_slab(a, "vqa_ops_slab_ln_bwd"));
  if (!a->dout) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_slab_ln_bwd: dout is NULL");
  if (a->slabs == 0) return VQA_OK;
  int RG = 1;
  const int rpt = slab_reg_plan(a->n, a->N, 512, &RG);
  const size_t comb = RG > 1 ? sizeof(float) * RG * 3 * a->N : 0;
  if (rpt == 5) {
    slab_ln_bwd_reg_kernel<5><<<a->slabs, (a->N >> 3) * RG, comb, static_cast<cudaStream_t>(stream)>>>(*a, RG);
  } else if (rpt == 9 && comb <= 96 * 1024) {
    if (comb > 32 * 1024)   // (+ the kernel's static shared memory: past the 48 KB default)
      VQA_CUDA_CHECK(cudaFuncSetAttribute(slab_ln_bwd_reg_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    slab_ln_bwd_reg_kernel<9><<<a->slabs, (a->N >> 3) * RG, comb, static_cast<cudaStream_t>(stream)>>>(*a, RG);
  } else {
    const size_t smem = sizeof(float) * a->n * a->N;
    VQA_TRY(ensure_smem(slab_ln_bwd_kernel, smem, "vqa_ops_slab_ln_bwd"));
    slab_ln_bwd_kernel<<<a->slabs, SL_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(*a);
  }
  VQA_LAUNCH_CHECK("slab_ln_bwd");
  return VQA_OK;
}

VQA_API VqaStatus vqa_ops_linear_ln(VqaOps ops, const VqaLinearLn* a, void* stream) {
  if (!ops || !a) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_linear_ln: null argument");
  if (!a->a || !a->w || !a->gamma || !a->beta || !a->z || !a->mean || !a->rstd || (!a->backward && !a->bias))
    return set_error(VQA_ERR_BAD_ARG, "vqa_ops_linear_ln: a / w / gamma / beta / z / mean / rstd (and bias, forward) must be given");
  if (a->keep <= 0.f || a->keep > 1.f) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_linear_ln: keep %g", a->keep);
  LinearLn d{};
  d.M = a->M; d.N = a->N; d.K = a->K;
  d.a = static_cast<const bf16*>(a->a); d.lda = a->lda;
  d.b = static_cast<const bf16*>(a->w); d.ldb = a->ldw;
  d.b_mn_major = a->backward ? 0 : 1;
  d.backward = a->backward;
  d.bias = a->bias; d.gamma = a->gamma; d.beta = a->beta; d.mul = a->mul;
  d.act = a->act; d.keep = a->keep; d.seed = a->seed; d.step = a->step; d.stream_id = a->site;
  d.z = a->z; d.mean = a->mean; d.rstd = a->rstd;
  d.y = a->y; d.out_f32 = a->out_f32; d.out_hi = static_cast<bf16*>(a->out_hi);
  d.raw = a->raw; d.dz_f32 = a->dz_f32; d.dz_hi = static_cast<bf16*>(a->dz_hi);
  d.dgamma_part = a->dgamma_part; d.dbeta_part = a->dbeta_part;
  bool launched = false;
  VQA_TRY(linear_ln_launch(d, static_cast<cudaStream_t>(stream), &launched));
  if (!launched)
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_ops_linear_ln: M %d N %d K %d is not eligible for the fused kernel on this device",
                     a->M, a->N, a->K);
  return VQA_OK;
}

VQA_API VqaStatus vqa_ops_pad_planes(const float* src, int64_t rows, int32_t cols, int32_t boxes, void* hi, void* lo,
                                     int32_t ld_out, void* stream) {
  if (!src || !hi || rows < 0 || ld_out <= 0 || (ld_out & 7) || cols > ld_out || (boxes && ld_out < 6))
    return set_error(VQA_ERR_BAD_ARG, "vqa_ops_pad_planes: bad argument");
  if (rows == 0) return VQA_OK;
  const long long total = rows * ld_out;
  pad_planes_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, rows, cols, boxes, static_cast<bf16*>(hi), static_cast<bf16*>(lo), ld_out);
  VQA_LAUNCH_CHECK("pad_planes");
  return VQA_OK;
}

VQA_API VqaStatus vqa_ops_feat_wgrad(VqaOps ops, const float* feat, int32_t F, int32_t boxes, const void* dz_hi, const void* dz_lo,
                                     int64_t rows, int32_t N, float* part, int32_t max_parts, float* out, void* stream) {
  if (!ops || !feat || !dz_hi || !part || !out || rows < 0 || N <= 0 || (N & 3) || F <= 0 || F > 8 || max_parts <= 0 ||
      static_cast<long long>(F) * N > 8192 || (boxes && F != 6))
    return set_error(VQA_ERR_BAD_ARG, "vqa_ops_feat_wgrad: bad argument (F <= 8, N a multiple of 4, F * N <= 8192)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int ctas = ops->num_sms * 2;
  if (ctas > max_parts) ctas = max_parts;
  if (ctas > rows) ctas = rows > 0 ? static_cast<int>(rows) : 1;
  feat_wgrad_kernel<<<ctas, FW_THREADS, 0, s>>>(feat, F, boxes, static_cast<const bf16*>(dz_hi), static_cast<const bf16*>(dz_lo),
                                                 rows, N, part);
  VQA_LAUNCH_CHECK("feat_wgrad");
  return colsum_launch(part, ctas, static_cast<long long>(F) * N, static_cast<long long>(F) * N, out, ops->scratch, s);
}

static VqaStatus check_spat(const VqaSpatAttn* a, const char* who) {
  if (!a || !a->hv_hi || !a->hq || !a->att_w || !a->att_b || !a->num_boxes || !a->v || !a->att)
    return set_error(VQA_ERR_BAD_ARG, "%s: null argument", who);
  if (a->B < 0 || a->K <= 0 || a->K > 256 || a->n <= 0 || a->n > NMAX || a->D <= 0 || (a->D & 7) || a->Dv <= 0 || (a->Dv & 3) ||
      a->kinds <= 0 || a->kinds > 4)
    return set_error(VQA_ERR_BAD_SHAPE, "%s: K <= 256, n <= 8, D a multiple of 8, Dv of 4, kinds <= 4", who);
  if (!(a->keep > 0.f && a->keep <= 1.f)) return set_error(VQA_ERR_BAD_ARG, "%s: keep must be in (0, 1]", who);
  return VQA_OK;
}

VQA_API VqaStatus vqa_memft_spat_attn_fwd(VqaOps ops, const VqaSpatAttn* a, void* stream) {
  if (!ops) return set_error(VQA_ERR_BAD_ARG, "vqa_memft_spat_attn_fwd: null context");
  VQA_TRY(check_spat(a, "vqa_memft_spat_attn_fwd"));
  if (!a->pooled && !a->pooled_hi) return set_error(VQA_ERR_BAD_ARG, "vqa_memft_spat_attn_fwd: no pooled output");
  if (a->B == 0) return VQA_OK;
  const size_t smem = sizeof(float) * (static_cast<size_t>(a->n) * a->D + a->D + static_cast<size_t>(a->n) * a->K);
  const dim3 grid(a->B, a->kinds);
  if (a->n <= 5) {
    VQA_TRY(ensure_smem(spat_attn_fwd_kernel<5>, smem, "vqa_memft_spat_attn_fwd"));
    spat_attn_fwd_kernel<5><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(*a);
  } else {
    VQA_TRY(ensure_smem(spat_attn_fwd_kernel<NMAX>, smem, "vqa_memft_spat_attn_fwd"));
    spat_attn_fwd_kernel<NMAX><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(*a);
  }
  VQA_LAUNCH_CHECK("spat_attn_fwd");
  return VQA_OK;
}

VQA_API VqaStatus vqa_memft_spat_attn_bwd(VqaOps ops, const VqaSpatAttn* a, void* stream) {
  if (!ops) return set_error(VQA_ERR_BAD_ARG, "vqa_memft_spat_attn_bwd: null context");
  VQA_TRY(check_spat(a, "vqa_memft_spat_attn_bwd"));
  if (!a->d_pooled || !a->d_hv || !a->d_hq || !a->part) return set_error(VQA_ERR_BAD_ARG, "vqa_memft_spat_attn_bwd: null gradient buffer");
  if (a->B == 0) return VQA_OK;
  const size_t smem = sizeof(float) * (static_cast<size_t>(a->n) * a->D + 2 * static_cast<size_t>(a->D) +
                                       2 * static_cast<size_t>(a->n) * a->K + static_cast<size_t>(a->n) * a->Dv);
  const dim3 grid(a->B, a->kinds);
  if (a->n <= 5) {
    VQA_TRY(ensure_smem(spat_attn_bwd_kernel<5>, smem, "vqa_memft_spat_attn_bwd"));
    spat_attn_bwd_kernel<5><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(*a);
  } else {
    VQA_TRY(ensure_smem(spat_attn_bwd_kernel<NMAX>, smem, "vqa_memft_spat_attn_bwd"));
    spat_attn_bwd_kernel<NMAX><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(*a);
  }
  VQA_LAUNCH_CHECK("spat_attn_bwd");
  return VQA_OK;
}

VQA_API VqaStatus vqa_memft_softmax_ce(VqaOps ops, const VqaSoftmaxCe* a, void* stream) {
  if (!ops || !a || !a->logit || !a->fills || !a->stats || !a->report)
    return set_error(VQA_ERR_BAD_ARG, "vqa_memft_softmax_ce: null argument");
  if (a->heads <= 0 || a->heads > 8 || a->B <= 0 || a->n <= 0 || a->A <= 0 || a->A > 12288 || a->top_k <= 0)
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_memft_softmax_ce: heads <= 8, A <= 12288");
  for (int h = 0; h < a->heads; ++h)
    if (!a->num[h]) return set_error(VQA_ERR_BAD_ARG, "vqa_memft_softmax_ce: num[%d] is NULL", h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows_per_head = a->B * a->n;
  softmax_ce_kernel<<<a->heads * rows_per_head, 256, sizeof(float) * a->A, s>>>(*a);
  VQA_LAUNCH_CHECK("softmax_ce");
  softmax_ce_finalize_kernel<<<1, 256, 0, s>>>(a->stats, a->heads, rows_per_head, a->report);
  VQA_LAUNCH_CHECK("softmax_ce_finalize");
  return VQA_OK;
}

VQA_API VqaStatus vqa_memft_wordset_fwd(const float* map, const int32_t* ids, int64_t rows, int32_t W, int32_t num_ws,
                                        float* y, void* hi, void* lo, int32_t ld, void* stream) {
  if (!map || !ids || !y || !hi || rows < 0 || W <= 0 || ld < W || (ld & 7) || num_ws <= 0)
    return set_error(VQA_ERR_BAD_ARG, "vqa_memft_wordset_fwd: bad argument");
  if (rows == 0) return VQA_OK;
  wordset_fwd_kernel<<<static_cast<unsigned>(rows), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      map, ids, W, num_ws, y, static_cast<bf16*>(hi), static_cast<bf16*>(lo), ld);
  VQA_LAUNCH_CHECK("wordset_fwd");
  return VQA_OK;
}

VQA_API VqaStatus vqa_memft_wordset_bwd(const float* d_y, const float* y, const int32_t* ids, int64_t rows, int32_t W,
                                        int32_t num_ws, float* d_map, void* stream) {
  if (!d_y || !y || !ids || !d_map || rows < 0 || W <= 0 || num_ws <= 0)
    return set_error(VQA_ERR_BAD_ARG, "vqa_memft_wordset_bwd: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VQA_CUDA_CHECK(cudaMemsetAsync(d_map, 0, sizeof(float) * static_cast<size_t>(num_ws) * W, s));
  if (rows == 0) return VQA_OK;
  wordset_bwd_kernel<<<static_cast<unsigned>(rows), 128, 0, s>>>(d_y, y, ids, W, num_ws, d_map);
  VQA_LAUNCH_CHECK("wordset_bwd");
  return VQA_OK;
}

VQA_API VqaStatus vqa_ops_gru_workspace_bytes(const VqaGruSeq* a, uint64_t* bytes) {
  VQA_TRY(check_gru(a, "vqa_ops_gru_workspace_bytes", false));
  if (!bytes) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_gru_workspace_bytes: null argument");
  *bytes = carve_gru(*a, nullptr).bytes;
  return VQA_OK;
}

VQA_API VqaStatus vqa_ops_gru_fwd(VqaOps ops, const VqaGruSeq* ap, void* stream) {
  if (!ops) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_gru_fwd: null context");
  VQA_TRY(check_gru(ap, "vqa_ops_gru_fwd", true));
  const VqaGruSeq& a = *ap;
  if (!a.embed || !a.gates_w || !a.gates_b || !a.cand_w || !a.cand_b || !a.tokens || !a.len || !a.q)
    return set_error(VQA_ERR_BAD_ARG, "vqa_ops_gru_fwd: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GruWs w = carve_gru(a, a.ws);
  const int B = a.B, T = a.T, L = a.L, W = a.W, Wp = w.Wp;
  const long long BL = static_cast<long long>(B) * L;
  // operand planes of the two GRU kernels (TF layout [W + L, out]: x rows first, then h rows)
  VQA_TRY(split_bf16_launch(a.gates_w, W + L, 2 * L, 2 * L, w.wg.hi, w.wg.lo, 2 * L, s));
  VQA_TRY(split_bf16_launch(a.cand_w, W + L, L, L, w.wc.hi, w.wc.lo, L, s));
  const bool persistent = gru_persistent_supported(B, L, a.precision, ops->num_sms);
  if (persistent)
    VQA_TRY(gru_pack_weights_launch(w.wg.hi + static_cast<long long>(W) * 2 * L, w.wc.hi + static_cast<long long>(W) * L, L, w.pack, s));
  VQA_TRY(embed_gather_launch(a.embed, a.tokens, B, T, T, W, Wp, a.Vq, w.e.hi, w.e.lo, s));
  // bf16 mode on the persistent kernels: the hoisted x-parts are stored as bf16 in the same buffers (these products are
  // store-bound: 630 MB of fp32 per step for the 5120 blank sequences of BASELINE config 4); VQA_GRU_X_BF16=0: fp32
  static const bool x_bf16_env = getenv("VQA_GRU_X_BF16") == nullptr || atoi(getenv("VQA_GRU_X_BF16")) != 0;
  const bool x_bf16 = x_bf16_env && persistent && a.precision == VQA_PREC_BF16;
  G gx(static_cast<long long>(T) * B, 2 * L, W), cx(static_cast<long long>(T) * B, L, W);
  gx.a(w.e, 0, Wp, false).b(w.wg, 0, 2 * L, true).bias(a.gates_b);
  cx.a(w.e, 0, Wp, false).b(w.wc, 0, L, true).bias(a.cand_b);
  if (x_bf16) { gx.bf(reinterpret_cast<bf16*>(w.xg), 2 * L); cx.bf(reinterpret_cast<bf16*>(w.xc), L); }
  else { gx.f32(w.xg, 2 * L); cx.f32(w.xc, L); }
  VQA_TRY(gx.run(ops, s));
  VQA_TRY(cx.run(ops, s));
  VQA_TRY(fill_zero_launch(w.h_f32, sizeof(float) * BL, s));
  VQA_TRY(fill_zero_launch(w.h.hi, sizeof(bf16) * BL, s));
  if (w.h.lo) VQA_TRY(fill_zero_launch(w.h.lo, sizeof(bf16) * BL, s));
  if (persistent) {
    GruFwdPersistent g{};
    g.B = B; g.L = L; g.T = T; g.q_len = a.len; g.counter = w.counter;
    g.xg = w.xg; g.xc = w.xc; g.x_bf16 = x_bf16 ? 1 : 0; g.h_f32 = w.h_f32; g.h_bf = w.h.hi; g.rh_bf = w.rh.hi;
    g.r = w.r; g.u = w.u; g.c = w.c; g.w_pack = w.pack;
    VQA_TRY(gru_fwd_persistent_launch(g, ops->num_sms, s));
  } else {
    for (int t = 0; t < T; ++t) {
      VQA_TRY(G(B, 2 * L, L).a(w.h, t * BL, L, false).b(w.wg, static_cast<long long>(W) * 2 * L, 2 * L, true)
                  .addend(w.xg + static_cast<long long>(t) * B * 2 * L, 2 * L).f32(w.g_pre, 2 * L).run(ops, s));
      VQA_TRY(gru_gates_launch(w.g_pre, w.h_f32 + t * BL, B, L, w.r + t * BL, w.u + t * BL, w.rh.hi + t * BL,
                               w.rh.lo ? w.rh.lo + t * BL : nullptr, s));
      VQA_TRY(G(B, L, L).a(w.rh, t * BL, L, false).b(w.wc, static_cast<long long>(W) * L, L, true)
                  .addend(w.xc + static_cast<long long>(t) * B * L, L).f32(w.c_pre, L).run(ops, s));
      VQA_TRY(gru_update_launch(w.c_pre, w.h_f32 + t * BL, w.u + t * BL, a.len, t, B, L, w.c + t * BL,
                                w.h_f32 + (t + 1) * BL, w.h.hi + (t + 1) * BL, w.h.lo ? w.h.lo + (t + 1) * BL : nullptr, s));
    }
  }
  VQA_CUDA_CHECK(cudaMemcpyAsync(a.q, w.h_f32 + T * BL, sizeof(float) * BL, cudaMemcpyDeviceToDevice, s));
  if (a.q_hi) VQA_CUDA_CHECK(cudaMemcpyAsync(a.q_hi, w.h.hi + T * BL, sizeof(bf16) * BL, cudaMemcpyDeviceToDevice, s));
  if (a.q_lo) {
    if (w.h.lo) VQA_CUDA_CHECK(cudaMemcpyAsync(a.q_lo, w.h.lo + T * BL, sizeof(bf16) * BL, cudaMemcpyDeviceToDevice, s));
    else VQA_TRY(fill_zero_launch(a.q_lo, sizeof(bf16) * BL, s));
  }
  return VQA_OK;
}

VQA_API VqaStatus vqa_ops_gru_bwd(VqaOps ops, const VqaGruSeq* ap, void* stream) {
  if (!ops) return set_error(VQA_ERR_BAD_ARG, "vqa_ops_gru_bwd: null context");
  VQA_TRY(check_gru(ap, "vqa_ops_gru_bwd", true));
  const VqaGruSeq& a = *ap;
  if (!a.tokens || !a.len || !a.dq || !a.d_embed || !a.d_gates_w || !a.d_gates_b || !a.d_cand_w || !a.d_cand_b)
    return set_error(VQA_ERR_BAD_ARG, "vqa_ops_gru_bwd: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GruWs w = carve_gru(a, a.ws);
  const int B = a.B, T = a.T, L = a.L, W = a.W, Wp = w.Wp;
  const long long BL = static_cast<long long>(B) * L;
  const long long TB = static_cast<long long>(T) * B;
  const bool persistent = gru_persistent_supported(B, L, a.precision, ops->num_sms);
  if (persistent) {
    GruBwdPersistent g{};
    g.B = B; g.L = L; g.T = T; g.q_len = a.len; g.counter = w.counter;
    g.h_f32 = w.h_f32; g.r = w.r; g.u = w.u; g.c = w.c; g.dq = a.dq; g.dq2 = nullptr;
    g.dG_bf = w.dG.hi; g.dC_bf = w.dC.hi; g.bias_part = w.bias_part;
    g.wg_h = w.wg.hi + static_cast<long long>(W) * 2 * L;
    g.wc_h = w.wc.hi + static_cast<long long>(W) * L;
    VQA_TRY(gru_bwd_persistent_launch(g, ops->num_sms, s));
  } else {
    const float* dh_cur = a.dq;
    int pp = 0;
    for (int t = T - 1; t >= 0; --t) {
      VQA_TRY(gru_bwd_update_launch(dh_cur, w.h_f32 + t * BL, w.u + t * BL, w.c + t * BL, a.len, t, B, L, w.du, w.dh_part,
                                    w.dC_f32 + t * BL, w.dC.hi + t * BL, w.dC.lo ? w.dC.lo + t * BL : nullptr, s));
      VQA_TRY(G(B, L, L).a(w.dC, t * BL, L, false).b(w.wc, static_cast<long long>(W) * L, L, false).f32(w.dRH, L).run(ops, s));
      VQA_TRY(gru_bwd_gates_launch(w.dRH, w.du, w.h_f32 + t * BL, w.r + t * BL, w.u + t * BL, a.len, t, B, L, w.dh_part,
                                   w.dG_f32 + 2 * t * BL, w.dG.hi + 2 * t * BL, w.dG.lo ? w.dG.lo + 2 * t * BL : nullptr, s));
      float* dh_next = w.dh[pp];
      pp ^= 1;
      VQA_TRY(G(B, L, 2 * L).a(w.dG, 2 * t * BL, 2 * L, false).b(w.wg, static_cast<long long>(W) * 2 * L, 2 * L, false)
                  .addend(w.dh_part, L).f32(dh_next, L).run(ops, s));
      dh_cur = dh_next;
    }
  }
  // weight gradients: h rows, x rows, biases
  VQA_TRY(G(L, 2 * L, TB).a(w.h, 0, L, true).b(w.dG, 0, 2 * L, true).f32(a.d_gates_w + static_cast<long long>(W) * 2 * L, 2 * L).run(ops, s));
  VQA_TRY(G(L, L, TB).a(w.rh, 0, L, true).b(w.dC, 0, L, true).f32(a.d_cand_w + static_cast<long long>(W) * L, L).run(ops, s));
  VQA_TRY(G(W, 2 * L, TB).a(w.e, 0, Wp, true).b(w.dG, 0, 2 * L, true).f32(a.d_gates_w, 2 * L).run(ops, s));
  VQA_TRY(G(W, L, TB).a(w.e, 0, Wp, true).b(w.dC, 0, L, true).f32(a.d_cand_w, L).run(ops, s));
  if (persistent) {
    VQA_TRY(colsum_launch(w.bias_part, gru_bias_part_rows(B), 2 * L, 3 * L, a.d_gates_b, ops->scratch, s));
    VQA_TRY(colsum_launch(w.bias_part + 2 * L, gru_bias_part_rows(B), L, 3 * L, a.d_cand_b, ops->scratch, s));
  } else {
    VQA_TRY(colsum_launch(w.dG_f32, TB, 2 * L, 2 * L, a.d_gates_b, ops->scratch, s));
    VQA_TRY(colsum_launch(w.dC_f32, TB, L, L, a.d_cand_b, ops->scratch, s));
  }
  // d E = d G Wg[:W]^T + d C Wc[:W]^T, scattered into the embedding map
  VQA_TRY(G(TB, W, 2 * L).a(w.dG, 0, 2 * L, false).b(w.wg, 0, 2 * L, false).f32(w.dE, Wp).run(ops, s));
  VQA_TRY(G(TB, W, L).a(w.dC, 0, L, false).b(w.wc, 0, L, false).addend(w.dE, Wp).f32(w.dE, Wp).run(ops, s));
  VQA_TRY(fill_zero_launch(a.d_embed, sizeof(float) * static_cast<size_t>(a.Vq) * W, s));
  VQA_TRY(embed_scatter_add_launch(w.dE, Wp, a.tokens, a.len, B, T, T, W, a.Vq, a.d_embed, s));
  return VQA_OK;
}

}  // extern "C"
