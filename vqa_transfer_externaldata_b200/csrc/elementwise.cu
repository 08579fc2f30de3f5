// Memory-bound helpers around the GEMMs: operand-plane conversion, feature / embedding gather,
// embedding scatter-add, column sums (bias gradients) and the element-wise parts of the GRU cell.
// All are 128-bit vectorised, coalesced, grid-stride kernels.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "internal.h"
#include "launch.cuh"

namespace vqa {

namespace {

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  __nv_bfloat162 p0(__float2bfloat16_rn(a), __float2bfloat16_rn(b));
  __nv_bfloat162 p1(__float2bfloat16_rn(c), __float2bfloat16_rn(d));
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&p0);
  r.y = *reinterpret_cast<uint32_t*>(&p1);
  return r;
}

// write x (4 floats) as hi (+ lo residual) planes at element offset o
__device__ __forceinline__ void store_planes4(bf16* hi, bf16* lo, long long o, float4 x) {
  const bf16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y),
             h2 = __float2bfloat16_rn(x.z), h3 = __float2bfloat16_rn(x.w);
  __nv_bfloat162 p0(h0, h1), p1(h2, h3);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&p0);
  pk.y = *reinterpret_cast<uint32_t*>(&p1);
  *reinterpret_cast<uint2*>(hi + o) = pk;
  if (lo) {
    *reinterpret_cast<uint2*>(lo + o) =
        pack4_bf16(x.x - __bfloat162float(h0), x.y - __bfloat162float(h1),
                   x.z - __bfloat162float(h2), x.w - __bfloat162float(h3));
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------
__global__ void split_bf16_kernel(const float* __restrict__ src, long long rows, long long cols4,
                                  long long ld, bf16* __restrict__ hi, bf16* __restrict__ lo,
                                  long long ld_out) {
  pdl_sync();
  const long long total = rows * cols4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols4, c = (i - r * cols4) * 4;
    const float4 x = *reinterpret_cast<const float4*>(src + r * ld + c);
    store_planes4(hi, lo, r * ld_out + c, x);
  }
}

// Range checks of the two index inputs (the reference: np.take wraps negative image_idx -- parse_fn's default for a
// missing feature is -1 -> the LAST image, vqa/model_vlmap_answer.py:110-117 --, raises on the rest; tf.nn.embedding_lookup
// raises on an id outside [0, Vq)). A kernel cannot raise: it counts the offence in a sticky device counter that the
// host reads with vqa_input_error_count (Engine.read_scalars raises) and uses index 0, so nothing is read or, in the
// scatter-add, WRITTEN out of bounds.
__device__ unsigned int g_input_errors = 0;
__device__ __forceinline__ long long checked_image(long long img, long long num_images, bool count) {
  if (img < 0) img += num_images;
  if (img < 0 || img >= num_images) {
    if (count) atomicAdd(&g_input_errors, 1u);
    img = 0;
  }
  return img;
}
__device__ __forceinline__ int checked_token(int id, int vocab, bool count) {
  if (id < 0 || id >= vocab) {
    if (count) atomicAdd(&g_input_errors, 1u);
    id = 0;
  }
  return id;
}

// the same gather from the one-off bf16 copy of the bank: a 16-byte row copy
__global__ void __launch_bounds__(1024) gather_features_bf16_kernel(const bf16* __restrict__ bank, const int* __restrict__ num_boxes,
                                            const long long* __restrict__ image_idx, int batch, long long per_image8,
                                            bf16* __restrict__ v_hi, int* __restrict__ nbox, long long num_images) {
  pdl_sync();
  for (int b = blockIdx.y; b < batch; b += gridDim.y) {
    const long long img = checked_image(image_idx[b], num_images, blockIdx.x == 0 && threadIdx.x == 0);
    if (blockIdx.x == 0 && threadIdx.x == 0) nbox[b] = num_boxes[img];
    const uint4* src = reinterpret_cast<const uint4*>(bank) + img * per_image8;
    uint4* dst = reinterpret_cast<uint4*>(v_hi) + static_cast<long long>(b) * per_image8;
#pragma unroll 4
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < per_image8;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
      dst[i] = __ldg(src + i);
  }
}

// V[b] = features[image_idx[b]] (vqa/model_vlmap_answer.py:110-123) written as GEMM operand planes
__global__ void __launch_bounds__(1024) gather_features_kernel(const float* __restrict__ bank, const int* __restrict__ num_boxes,
                                       const long long* __restrict__ image_idx, int batch,
                                       long long per_image4, bf16* __restrict__ v_hi,
                                       bf16* __restrict__ v_lo, int* __restrict__ nbox, long long num_images) {
  pdl_sync();
  // gridDim.y == batch for the in-step gather; the background prefetch runs a small grid that walks the samples
  for (int b = blockIdx.y; b < batch; b += gridDim.y) {
    const long long img = checked_image(image_idx[b], num_images, blockIdx.x == 0 && threadIdx.x == 0);
    if (blockIdx.x == 0 && threadIdx.x == 0) nbox[b] = num_boxes[img];
    const float4* src = reinterpret_cast<const float4*>(bank) + img * per_image4;
    const long long dst0 = static_cast<long long>(b) * per_image4 * 4;
#pragma unroll 4
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < per_image4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const float4 x = __ldg(src + i);
      store_planes4(v_hi, v_lo, dst0 + i * 4, x);
    }
  }
}

// E[t*batch + b, :] = embed[q_intseq[b, t], :]  (tf.nn.embedding_lookup, model_vlmap_answer.py:134),
// time-major so that each GRU step reads a contiguous [batch, W] slab. Columns [W, Wpad) are zero.
__global__ void embed_gather_kernel(const float* __restrict__ embed, const int* __restrict__ q_intseq,
                                    int batch, int T, int Tstride, int W, int Wpad,
                                    bf16* __restrict__ e_hi, bf16* __restrict__ e_lo, int vocab) {
  const int row = blockIdx.x;  // t * batch + b
  const int t = row / batch, b = row - t * batch;
  pdl_sync();
  const int id = checked_token(q_intseq[b * Tstride + t], vocab, threadIdx.x == 0);
  const float* src = embed + static_cast<long long>(id) * W;
  for (int c = threadIdx.x; c < Wpad; c += blockDim.x) {
    const float x = c < W ? src[c] : 0.0f;
    const bf16 h = __float2bfloat16_rn(x);
    e_hi[static_cast<long long>(row) * Wpad + c] = h;
    if (e_lo) e_lo[static_cast<long long>(row) * Wpad + c] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

// d_embed[q[b,t]] += dE[t*batch+b] for t < len[b] (gradient of the gather = scatter-add;
// TF produces IndexedSlices and Adam applies them densely). fp32 atomics: order-dependent rounding.
__global__ void embed_scatter_add_kernel(const float* __restrict__ dE, long long ld,
                                         const int* __restrict__ q_intseq, const int* __restrict__ q_len,
                                         int batch, int T, int Tstride, int W,
                                         float* __restrict__ d_embed, int vocab) {
  const int row = blockIdx.x;
  const int t = row / batch, b = row - t * batch;
  pdl_sync();
  if (t >= q_len[b]) return;
  const int id = checked_token(q_intseq[b * Tstride + t], vocab, false);   // (the forward gather counted it)
  float* dst = d_embed + static_cast<long long>(id) * W;
  const float* src = dE + static_cast<long long>(row) * ld;
  for (int c = threadIdx.x; c < W; c += blockDim.x) atomicAdd(dst + c, src[c]);
}

// column sums, deterministic two-pass: grid (ceil(cols/32), RS); block (32, 32)
__global__ void colsum_partial_kernel(const float* __restrict__ x, long long rows, long long cols,
                                      long long ld, float* __restrict__ part) {
  __shared__ float sm[32][33];
  const long long c = blockIdx.x * 32LL + threadIdx.x;
  pdl_sync();
  float acc = 0.0f;
  if (c < cols)
    for (long long r = blockIdx.y * 32LL + threadIdx.y; r < rows; r += 32LL * gridDim.y)
      acc += x[r * ld + c];
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += sm[j][threadIdx.x];
    part[blockIdx.y * cols + c] = s;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int parts, long long cols,
                                    float* __restrict__ out) {
  const long long c = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  pdl_sync();
  if (c >= cols) return;
  float s = 0.0f;
  for (int p = 0; p < parts; ++p) s += part[p * cols + c];
  out[c] = s;
}

// ---- GRU cell element-wise parts -------------------------------------------------------------
// G = [x, h] Wg + bg (pre-activation); r,u = sigmoid; rh = r * h
__global__ void gru_gates_kernel(const float* __restrict__ G, const float* __restrict__ h, int batch,
                                 int L4, float* __restrict__ r_out, float* __restrict__ u_out,
                                 bf16* __restrict__ rh_hi, bf16* __restrict__ rh_lo) {
  const long long total = static_cast<long long>(batch) * L4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / L4, j = i - b * L4;
    const float4 gr = reinterpret_cast<const float4*>(G)[b * 2 * L4 + j];
    const float4 gu = reinterpret_cast<const float4*>(G)[b * 2 * L4 + L4 + j];
    const float4 hh = reinterpret_cast<const float4*>(h)[i];
    float4 r, u, rh;
    r.x = sigmoidf_(gr.x); r.y = sigmoidf_(gr.y); r.z = sigmoidf_(gr.z); r.w = sigmoidf_(gr.w);
    u.x = sigmoidf_(gu.x); u.y = sigmoidf_(gu.y); u.z = sigmoidf_(gu.z); u.w = sigmoidf_(gu.w);
    rh.x = r.x * hh.x; rh.y = r.y * hh.y; rh.z = r.z * hh.z; rh.w = r.w * hh.w;
    reinterpret_cast<float4*>(r_out)[i] = r;
    reinterpret_cast<float4*>(u_out)[i] = u;
    store_planes4(rh_hi, rh_lo, i * 4, rh);
  }
}

// c = tanh(C); h' = u*h + (1-u)*c; state is copied through for t >= len (dynamic_rnn sequence_length)
__global__ void gru_update_kernel(const float* __restrict__ C, const float* __restrict__ h,
                                  const float* __restrict__ u, const int* __restrict__ q_len, int t,
                                  int batch, int L4, float* __restrict__ c_out,
                                  float* __restrict__ h_next, bf16* __restrict__ hn_hi,
                                  bf16* __restrict__ hn_lo) {
  const long long total = static_cast<long long>(batch) * L4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / L4);
    const bool valid = t < q_len[b];
    const float4 cc = reinterpret_cast<const float4*>(C)[i];
    const float4 hh = reinterpret_cast<const float4*>(h)[i];
    const float4 uu = reinterpret_cast<const float4*>(u)[i];
    float4 c, hn;
    c.x = tanhf(cc.x); c.y = tanhf(cc.y); c.z = tanhf(cc.z); c.w = tanhf(cc.w);
    hn.x = valid ? uu.x * hh.x + (1.0f - uu.x) * c.x : hh.x;
    hn.y = valid ? uu.y * hh.y + (1.0f - uu.y) * c.y : hh.y;
    hn.z = valid ? uu.z * hh.z + (1.0f - uu.z) * c.z : hh.z;
    hn.w = valid ? uu.w * hh.w + (1.0f - uu.w) * c.w : hh.w;
    reinterpret_cast<float4*>(c_out)[i] = c;
    reinterpret_cast<float4*>(h_next)[i] = hn;
    store_planes4(hn_hi, hn_lo, i * 4, hn);
  }
}

// backward through h' = u*h + (1-u)*c, c = tanh(Cpre)
__global__ void gru_bwd_update_kernel(const float* __restrict__ dh, const float* __restrict__ h,
                                      const float* __restrict__ u, const float* __restrict__ c,
                                      const int* __restrict__ q_len, int t, int batch, int L4,
                                      float* __restrict__ du, float* __restrict__ dh_part,
                                      float* __restrict__ dC_f32, bf16* __restrict__ dC_hi,
                                      bf16* __restrict__ dC_lo) {
  const long long total = static_cast<long long>(batch) * L4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / L4);
    const bool valid = t < q_len[b];
    const float4 g = reinterpret_cast<const float4*>(dh)[i];
    float4 o_du = make_float4(0.f, 0.f, 0.f, 0.f), o_dc = o_du, o_dhp = g;
    if (valid) {
      const float4 hh = reinterpret_cast<const float4*>(h)[i];
      const float4 uu = reinterpret_cast<const float4*>(u)[i];
      const float4 cc = reinterpret_cast<const float4*>(c)[i];
      o_du.x = g.x * (hh.x - cc.x); o_du.y = g.y * (hh.y - cc.y);
      o_du.z = g.z * (hh.z - cc.z); o_du.w = g.w * (hh.w - cc.w);
      o_dc.x = g.x * (1.f - uu.x) * (1.f - cc.x * cc.x);
      o_dc.y = g.y * (1.f - uu.y) * (1.f - cc.y * cc.y);
      o_dc.z = g.z * (1.f - uu.z) * (1.f - cc.z * cc.z);
      o_dc.w = g.w * (1.f - uu.w) * (1.f - cc.w * cc.w);
      o_dhp.x = g.x * uu.x; o_dhp.y = g.y * uu.y; o_dhp.z = g.z * uu.z; o_dhp.w = g.w * uu.w;
    }
    reinterpret_cast<float4*>(du)[i] = o_du;
    reinterpret_cast<float4*>(dh_part)[i] = o_dhp;
    if (dC_f32) reinterpret_cast<float4*>(dC_f32)[i] = o_dc;
    store_planes4(dC_hi, dC_lo, i * 4, o_dc);
  }
}

// backward through rh = r*h and the two sigmoids; dRH = dCpre * Wc_h^T
__global__ void gru_bwd_gates_kernel(const float* __restrict__ dRH, const float* __restrict__ du,
                                     const float* __restrict__ h, const float* __restrict__ r,
                                     const float* __restrict__ u, const int* __restrict__ q_len, int t,
                                     int batch, int L4, float* __restrict__ dh_part,
                                     float* __restrict__ dG_f32, bf16* __restrict__ dG_hi,
                                     bf16* __restrict__ dG_lo) {
  const long long total = static_cast<long long>(batch) * L4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / L4, j = i - b * L4;
    const bool valid = t < q_len[b];
    float4 dgr = make_float4(0.f, 0.f, 0.f, 0.f), dgu = dgr;
    if (valid) {
      const float4 g = reinterpret_cast<const float4*>(dRH)[i];
      const float4 hh = reinterpret_cast<const float4*>(h)[i];
      const float4 rr = reinterpret_cast<const float4*>(r)[i];
      const float4 uu = reinterpret_cast<const float4*>(u)[i];
      const float4 d_u = reinterpret_cast<const float4*>(du)[i];
      float4 dp = reinterpret_cast<float4*>(dh_part)[i];
      dp.x += g.x * rr.x; dp.y += g.y * rr.y; dp.z += g.z * rr.z; dp.w += g.w * rr.w;
      reinterpret_cast<float4*>(dh_part)[i] = dp;
      dgr.x = g.x * hh.x * rr.x * (1.f - rr.x); dgr.y = g.y * hh.y * rr.y * (1.f - rr.y);
      dgr.z = g.z * hh.z * rr.z * (1.f - rr.z); dgr.w = g.w * hh.w * rr.w * (1.f - rr.w);
      dgu.x = d_u.x * uu.x * (1.f - uu.x); dgu.y = d_u.y * uu.y * (1.f - uu.y);
      dgu.z = d_u.z * uu.z * (1.f - uu.z); dgu.w = d_u.w * uu.w * (1.f - uu.w);
    }
    const long long o_r = (b * 2 * L4 + j) * 4, o_u = (b * 2 * L4 + L4 + j) * 4;
    if (dG_f32) {
      *reinterpret_cast<float4*>(dG_f32 + o_r) = dgr;
      *reinterpret_cast<float4*>(dG_f32 + o_u) = dgu;
    }
    store_planes4(dG_hi, dG_lo, o_r, dgr);
    store_planes4(dG_hi, dG_lo, o_u, dgu);
  }
}

inline int grid_for(long long work_items, int threads, int cap = 148 * 16) {
  long long g = (work_items + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

VqaStatus split_bf16_launch(const float* src, long long rows, long long cols, long long ld, bf16* hi,
                            bf16* lo, long long ld_out, cudaStream_t s) {
  if ((cols & 3) || (ld & 3) || (ld_out & 3))
    return set_error(VQA_ERR_BAD_SHAPE, "split_bf16: cols and pitches must be multiples of 4");
  if (rows * cols == 0) return VQA_OK;
  launch_pdl(split_bf16_kernel, dim3(grid_for(rows * cols / 4, 256)), dim3(256), 0, s, src, rows, cols / 4, ld, hi, lo,
             ld_out);
  VQA_LAUNCH_CHECK("split_bf16");
  return VQA_OK;
}

namespace {
template <typename... P, typename... A>
VqaStatus gather_exclusive_launch(void (*kern)(P...), int ctas, cudaStream_t s, A... args) {
  constexpr int kExclusiveSmem = 200 * 1024;
  VQA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kExclusiveSmem));
  ctas = (ctas + 1) & ~1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1, ctas);
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = kExclusiveSmem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1;
  at[0].val.clusterDim.y = 2;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  VQA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, args...));
  count_launch();
  return VQA_OK;
}
}  // namespace

VqaStatus gather_features_bf16_launch(const bf16* bank, const int* num_boxes, const long long* image_idx, int batch,
                                      int K, int Dv, bf16* v_hi, int* nbox, long long num_images, cudaStream_t s,
                                      int max_ctas) {
  if (batch == 0) return VQA_OK;
  const long long per_image8 = static_cast<long long>(K) * Dv / 8;
  if (max_ctas > 0)
    return gather_exclusive_launch(gather_features_bf16_kernel, max_ctas < batch ? max_ctas : batch, s, bank, num_boxes,
                                   image_idx, batch, per_image8, v_hi, nbox, num_images);
  int gx = static_cast<int>((per_image8 + 255) / 256);
  if (gx > 8) gx = 8;
  launch_pdl(gather_features_bf16_kernel, dim3(gx, batch), dim3(256), 0, s, bank, num_boxes, image_idx, batch, per_image8,
             v_hi, nbox, num_images);
  VQA_LAUNCH_CHECK("gather_features (bf16 bank)");
  return VQA_OK;
}

VqaStatus gather_features_launch(const float* bank, const int* num_boxes, const long long* image_idx,
                                 int batch, int K, int Dv, bf16* v_hi, bf16* v_lo, int* nbox, long long num_images,
                                 cudaStream_t s, int max_ctas) {
  if (batch == 0) return VQA_OK;
  const long long per_image4 = static_cast<long long>(K) * Dv / 4;
  int gx = static_cast<int>((per_image4 + 255) / 256);
  if (gx > 8) gx = 8;
  int gy = batch;
  if (max_ctas > 0) {
    // background prefetch under the cooperative BPTT kernel: max_ctas CTAs of 1024 threads that each claim a whole
    // SM's shared memory, launched as 2-CTA CLUSTERS so that they fill whole TPCs: the recurrent grid is made of
    // 2-CTA clusters itself, and single CTAs scattered over twenty TPCs would leave it twenty SM pairs short (a
    // cooperative launch then waits for the copy to finish: measured)
    return gather_exclusive_launch(gather_features_kernel, max_ctas < batch ? max_ctas : batch, s, bank, num_boxes, image_idx,
                                   batch, per_image4, v_hi, v_lo, nbox, num_images);
  }
  dim3 grid(gx, gy);
  launch_pdl(gather_features_kernel, dim3(grid), dim3(256), 0, s, bank, num_boxes, image_idx, batch, per_image4, v_hi,
                                              v_lo, nbox, num_images);
  VQA_LAUNCH_CHECK("gather_features");
  return VQA_OK;
}

VqaStatus embed_gather_launch(const float* embed, const int* q_intseq, int batch, int T, int Tstride,
                              int W, int Wpad, int vocab, bf16* e_hi, bf16* e_lo, cudaStream_t s) {
  if (batch * T == 0) return VQA_OK;
  launch_pdl(embed_gather_kernel, dim3(batch * T), dim3(128), 0, s, embed, q_intseq, batch, T, Tstride, W, Wpad, e_hi,
                                                e_lo, vocab);
  VQA_LAUNCH_CHECK("embed_gather");
  return VQA_OK;
}

VqaStatus embed_scatter_add_launch(const float* dE, long long ld_dE, const int* q_intseq,
                                   const int* q_len, int batch, int T, int Tstride, int W, int vocab,
                                   float* d_embed, cudaStream_t s) {
  if (batch * T == 0) return VQA_OK;
  launch_pdl(embed_scatter_add_kernel, dim3(batch * T), dim3(128), 0, s, dE, ld_dE, q_intseq, q_len, batch, T, Tstride, W,
                                                     d_embed, vocab);
  VQA_LAUNCH_CHECK("embed_scatter_add");
  return VQA_OK;
}

VqaStatus input_error_count(unsigned int* out, bool reset) {
  unsigned int v = 0;
  VQA_CUDA_CHECK(cudaMemcpyFromSymbol(&v, g_input_errors, sizeof(v)));   // synchronises with the device
  if (reset && v) {
    const unsigned int zero = 0;
    VQA_CUDA_CHECK(cudaMemcpyToSymbol(g_input_errors, &zero, sizeof(zero)));
  }
  *out = v;
  return VQA_OK;
}

VqaStatus colsum_launch(const float* x, long long rows, long long cols, long long ld, float* out,
                        float* scratch, cudaStream_t s) {
  if (cols == 0) return VQA_OK;
  int rs = static_cast<int>((rows + 255) / 256);
  if (rs > 32) rs = 32;
  if (rs < 1) rs = 1;
  dim3 grid(static_cast<unsigned>((cols + 31) / 32), rs);
  launch_pdl(colsum_partial_kernel, dim3(grid), dim3(32, 32), 0, s, x, rows, cols, ld, scratch);
  VQA_LAUNCH_CHECK("colsum_partial");
  launch_pdl(colsum_final_kernel, dim3(static_cast<unsigned>((cols + 255) / 256)), dim3(256), 0, s, scratch, rs, cols, out);
  VQA_LAUNCH_CHECK("colsum_final");
  return VQA_OK;
}

VqaStatus fill_zero_launch(void* p, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return VQA_OK;
  VQA_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, s));
  return VQA_OK;
}

VqaStatus gru_gates_launch(const float* G, const float* h, int batch, int L, float* r, float* u,
                           bf16* rh_hi, bf16* rh_lo, cudaStream_t s) {
  gru_gates_kernel<<<grid_for(static_cast<long long>(batch) * L / 4, 256), 256, 0, s>>>(
      G, h, batch, L / 4, r, u, rh_hi, rh_lo);
  VQA_LAUNCH_CHECK("gru_gates");
  return VQA_OK;
}

VqaStatus gru_update_launch(const float* C, const float* h, const float* u, const int* q_len, int t,
                            int batch, int L, float* c_out, float* h_next, bf16* hn_hi, bf16* hn_lo,
                            cudaStream_t s) {
  gru_update_kernel<<<grid_for(static_cast<long long>(batch) * L / 4, 256), 256, 0, s>>>(
      C, h, u, q_len, t, batch, L / 4, c_out, h_next, hn_hi, hn_lo);
  VQA_LAUNCH_CHECK("gru_update");
  return VQA_OK;
}

VqaStatus gru_bwd_update_launch(const float* dh, const float* h, const float* u, const float* c,
                                const int* q_len, int t, int batch, int L, float* du, float* dh_part,
                                float* dC_f32, bf16* dC_hi, bf16* dC_lo, cudaStream_t s) {
  gru_bwd_update_kernel<<<grid_for(static_cast<long long>(batch) * L / 4, 256), 256, 0, s>>>(
      dh, h, u, c, q_len, t, batch, L / 4, du, dh_part, dC_f32, dC_hi, dC_lo);
  VQA_LAUNCH_CHECK("gru_bwd_update");
  return VQA_OK;
}

VqaStatus gru_bwd_gates_launch(const float* dRH, const float* du, const float* h, const float* r,
                               const float* u, const int* q_len, int t, int batch, int L,
                               float* dh_part, float* dG_f32, bf16* dG_hi, bf16* dG_lo,
                               cudaStream_t s) {
  gru_bwd_gates_kernel<<<grid_for(static_cast<long long>(batch) * L / 4, 256), 256, 0, s>>>(
      dRH, du, h, r, u, q_len, t, batch, L / 4, dh_part, dG_f32, dG_hi, dG_lo);
  VQA_LAUNCH_CHECK("gru_bwd_gates");
  return VQA_OK;
}

}  // namespace vqa
