// Row-wise LayerNorm + ReLU heads: the LN / activation / Hadamard / dropout tail of modules.fc_layer
// for the 2-D heads q_linear_v, pooled_linear_l, q_linear_l, joint_fc (vlmap/modules.py:630-650,
// vqa/model_vlmap_answer.py:142-181), forward and backward. One CTA per row, each thread owns chunks
// of 8 consecutive columns (two 128-bit accesses), two-pass statistics (tf.nn.moments), eps 1e-12.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "internal.h"
#include "launch.cuh"
#include "philox.cuh"

namespace vqa {

namespace {

constexpr int ROW_THREADS = 256;
constexpr int MAX_CHUNKS = 2;  // N <= 4096

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect red from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < ROW_THREADS / 32; ++w) t += red[w];
  return t;
}

__device__ __forceinline__ void ld8(const float* p, float (&x)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
  x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&x)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
}
__device__ __forceinline__ void st8_planes(bf16* hi, bf16* lo, long long off, const float (&x)[8]) {
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

__global__ void __launch_bounds__(ROW_THREADS) row_ln_relu_fwd_kernel(RowLnFwd a, uint32_t thr) {
  __shared__ float red[ROW_THREADS / 32];
  const int row = blockIdx.x, N = a.N, CH = N >> 3;
  const long long base = static_cast<long long>(row) * N;
  pdl_sync();
  float x[MAX_CHUNKS][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < CH) {
      ld8(a.z + base + c * 8, x[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += x[i][j];
    }
  }
  const float mean = block_sum(s, red) / N;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < CH) {
#pragma unroll
      for (int j = 0; j < 8; ++j) q += (x[i][j] - mean) * (x[i][j] - mean);
    }
  }
  const float var = block_sum(q, red) / N;
  const float rstd = 1.0f / sqrtf(var + 1e-12f);
  const float inv_keep = 1.0f / a.keep;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < CH) {
      float g[8], bt[8], y[8];
      ld8(a.gamma + c * 8, g);
      ld8(a.beta + c * 8, bt);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float ypre = fmaf((x[i][j] - mean) * rstd, g[j], bt[j]);
        y[j] = a.act == 1 ? tanhf(ypre) : fmaxf(ypre, 0.f);
      }
      if (a.y) st8(a.y + base + c * 8, y);
      if (a.mul) {
        float m[8];
        ld8(a.mul + base + c * 8, m);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] *= m[j];
      }
      if (thr < 65536u) {
        const uint32_t bits = philox_keep_bits(
            philox4x32_10(static_cast<unsigned long long>(row) * CH + c, a.stream_id, a.seed, a.step), thr);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = ((bits >> j) & 1u) ? y[j] * inv_keep : 0.f;
      } else if (a.keep < 1.0f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] *= inv_keep;
      }
      if (a.out_f32) st8(a.out_f32 + base + c * 8, y);
      if (a.out_hi) st8_planes(a.out_hi, a.out_lo, base + c * 8, y);
    }
  }
  if (threadIdx.x == 0) {
    a.mean[row] = mean;
    a.rstd[row] = rstd;
  }
}

__global__ void __launch_bounds__(ROW_THREADS) row_ln_relu_bwd_kernel(RowLnBwd a, uint32_t thr) {
  __shared__ float red[ROW_THREADS / 32];
  const int row = blockIdx.x, N = a.N, CH = N >> 3;
  const long long base = static_cast<long long>(row) * N;
  pdl_sync();
  const float mean = a.mean[row], rstd = a.rstd[row];
  const float inv_keep = 1.0f / a.keep;
  float xh[MAX_CHUNKS][8], dxh[MAX_CHUNKS][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < CH) {
      float z[8], g[8], bt[8], d[8];
      ld8(a.z + base + c * 8, z);
      ld8(a.gamma + c * 8, g);
      ld8(a.beta + c * 8, bt);
      ld8(a.dout + base + c * 8, d);
      if (thr < 65536u) {
        const uint32_t bits = philox_keep_bits(
            philox4x32_10(static_cast<unsigned long long>(row) * CH + c, a.stream_id, a.seed, a.step), thr);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = ((bits >> j) & 1u) ? d[j] * inv_keep : 0.f;
      } else if (a.keep < 1.0f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] *= inv_keep;
      }
      if (a.mul) {
        float m[8];
        ld8(a.mul + base + c * 8, m);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] *= m[j];
      }
      float dg[8], db[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (z[j] - mean) * rstd;
        const float ypre = fmaf(xh[i][j], g[j], bt[j]);
        float dy;
        if (a.act == 1) {
          const float t = tanhf(ypre);
          dy = d[j] * (1.0f - t * t);
        } else {
          dy = ypre > 0.f ? d[j] : 0.f;
        }
        dg[j] = dy * xh[i][j];
        db[j] = dy;
        dxh[i][j] = dy * g[j];
        s1 += dxh[i][j];
        s2 = fmaf(dxh[i][j], xh[i][j], s2);
      }
      if (a.dgamma_part) {
        st8(a.dgamma_part + base + c * 8, dg);
        st8(a.dbeta_part + base + c * 8, db);
      }
    }
  }
  const float m1 = block_sum(s1, red) / N;
  const float m2 = block_sum(s2, red) / N;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * ROW_THREADS;
    if (c < CH) {
      float dz[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dz[j] = rstd * (dxh[i][j] - m1 - xh[i][j] * m2);
      if (a.dz_f32) st8(a.dz_f32 + base + c * 8, dz);
      if (a.dz_hi) st8_planes(a.dz_hi, a.dz_lo, base + c * 8, dz);
    }
  }
}

}  // namespace

VqaStatus row_ln_relu_fwd_launch(const RowLnFwd& a, cudaStream_t s) {
  if (a.rows == 0) return VQA_OK;
  if ((a.N & 7) || a.N > 8 * ROW_THREADS * MAX_CHUNKS)
    return set_error(VQA_ERR_BAD_SHAPE, "row_ln_relu: N must be a multiple of 8 and <= 4096");
  launch_pdl(row_ln_relu_fwd_kernel, dim3(a.rows), dim3(ROW_THREADS), 0, s, a, keep_threshold(a.keep));
  VQA_LAUNCH_CHECK("row_ln_relu_fwd");
  return VQA_OK;
}

VqaStatus row_ln_relu_bwd_launch(const RowLnBwd& a, cudaStream_t s) {
  if (a.rows == 0) return VQA_OK;
  if ((a.N & 7) || a.N > 8 * ROW_THREADS * MAX_CHUNKS)
    return set_error(VQA_ERR_BAD_SHAPE, "row_ln_relu: N must be a multiple of 8 and <= 4096");
  launch_pdl(row_ln_relu_bwd_kernel, dim3(a.rows), dim3(ROW_THREADS), 0, s, a, keep_threshold(a.keep));
  VQA_LAUNCH_CHECK("row_ln_relu_bwd");
  return VQA_OK;
}

}  // namespace vqa
