// Persistent, software-pipelined forward of the attention block (same math as attn_fwd_kernel in attn.cu:
// vlmap/modules.py:67-97 hadamard_attention + :23-39 attention_pooling + the joint (K, D) LayerNorm + ReLU of
// v_linear_v, vlmap/modules.py:646-649).
//
// attn.cu runs one CTA per sample: slab load -> statistics -> scores -> softmax -> pooling, each step waiting for
// the one before, with only a second co-resident CTA to hide the latencies (measured 1.6 TB/s = 25 % of HBM peak,
// 22 % of the warps active). Here ONE CTA per SM walks its samples and a producer warp keeps the memory system
// busy ahead of the math:
//   * the [K, D] pre-LN slab of sample i + 1 is in flight (bulk async copy into the second slab buffer) while the
//     16 consumer warps work on sample i out of shared memory;
//   * the raw features of sample i stream through a two-slot ring of 8-row chunks (32 KB bulk copies) that the
//     pooling pass consumes, instead of six dependent rounds of global loads per thread.
// ~210 KB are in flight or resident per SM at any time; every HBM byte is still read exactly once.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdlib>

#include "internal.h"
#include "launch.cuh"
#include "philox.cuh"
#include "ptx.cuh"

namespace vqa {

namespace {

constexpr int AP_CONSUMER_WARPS = 16;
constexpr int AP_CONSUMERS = 32 * AP_CONSUMER_WARPS;
constexpr int AP_THREADS = AP_CONSUMERS + 32;
constexpr int AP_VCHUNK_BYTES = 32768;

__device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 1, %0;" ::"n"(AP_CONSUMERS) : "memory"); }
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&x)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    x[2 * j] = f.x;
    x[2 * j + 1] = f.y;
  }
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
// Chan et al. pairwise merge of (count, mean, M2)
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float mb, float m2b) {
  const float nt = n + nb;
  if (nt == 0.f) return;
  const float delta = mb - mean;
  const float f = nb / nt;
  mean += delta * f;
  m2 += m2b + delta * delta * n * f;
  n = nt;
}

struct PipeFwdArgs {
  const bf16* z;
  const float* gamma; const float* beta; const float* hq; const float* att_w; const float* att_b;
  const int* nbox; const bf16* v;
  unsigned long long seed, step;
  float* att; float* pooled; bf16* pooled_bf; float* ln_mean; float* ln_rstd;
  int batch, K, D, Dv, RV;   // RV = feature rows per ring chunk
  float keep;
  uint32_t thr;
};

__global__ void __launch_bounds__(AP_THREADS, 1) attn_fwd_pipe_kernel(PipeFwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int K = a.K, D = a.D, Dv = a.Dv, RV = a.RV;
  const uint32_t zbytes = static_cast<uint32_t>(K) * D * 2;
  const uint32_t vrow = static_cast<uint32_t>(Dv) * 2;
  // layout: [slab 0 | slab 1 | V slot 0 | V slot 1 | cA cB cC | sc | red | barriers]
  uint8_t* zbuf = smem;
  uint8_t* vbuf = smem + 2 * static_cast<size_t>(zbytes);
  float* cA = reinterpret_cast<float*>(vbuf + 2 * AP_VCHUNK_BYTES);
  float* cB = cA + D;
  float* cC = cB + D;
  float* sc = cC + D;
  float* red = sc + ((K + 3) & ~3);
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 64);
  uint64_t* zfull = bars;
  uint64_t* zempty = bars + 2;
  uint64_t* vfull = bars + 4;
  uint64_t* vempty = bars + 6;
  float* poolx = cA;   // partial pooled sums of the second row group (cA / cB are dead by then)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&zfull[i], 1);
      ptx::mbar_init(&zempty[i], AP_CONSUMER_WARPS);
      ptx::mbar_init(&vfull[i], 1);
      ptx::mbar_init(&vempty[i], AP_CONSUMER_WARPS);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  pdl_sync();   // barriers are set up under the previous kernel's tail; z / v / hq are read from here on
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_my = first < a.batch ? (a.batch - first + stride - 1) / stride : 0;

  if (warp == AP_CONSUMER_WARPS) {
    // ===================== producer: one thread =====================
    if (lane == 0) {
      auto load_slab = [&](int i) {
        const int j = i & 1, u = i >> 1;
        ptx::mbar_wait(&zempty[j], (u & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&zfull[j], zbytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.z) + static_cast<size_t>(first + i * stride) * zbytes;
        const uint32_t dst = ptx::smem_u32(zbuf + static_cast<size_t>(j) * zbytes);
        for (uint32_t off = 0; off < zbytes; off += 65536u)
          bulk_load(dst + off, src + off, zbytes - off < 65536u ? zbytes - off : 65536u, &zfull[j]);
      };
      if (n_my > 0) load_slab(0);
      unsigned int vc = 0;
      for (int i = 0; i < n_my; ++i) {
        if (i + 1 < n_my) load_slab(i + 1);
        const int b = first + i * stride;
        int nb = a.nbox[b];
        nb = nb < 0 ? 0 : (nb > K ? K : nb);
        const uint8_t* vsrc = reinterpret_cast<const uint8_t*>(a.v) + static_cast<size_t>(b) * K * vrow;
        for (int k0 = 0; k0 < nb; k0 += RV, ++vc) {
          const int slot = vc & 1;
          const int rows = nb - k0 < RV ? nb - k0 : RV;
          ptx::mbar_wait(&vempty[slot], ((vc >> 1) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&vfull[slot], rows * vrow);
          bulk_load(ptx::smem_u32(vbuf + slot * AP_VCHUNK_BYTES), vsrc + static_cast<size_t>(k0) * vrow, rows * vrow,
                    &vfull[slot]);
        }
      }
    }
    return;
  }

  // ===================== consumers: 16 warps =====================
  const int CH = D >> 3;                 // 16-byte chunks per slab row
  const int nchunks = K * CH;
  const int VCH = Dv >> 3;               // column chunks of the features
  const int tpc = AP_CONSUMERS / VCH;    // threads sharing one column chunk (each takes every tpc-th row)
  const int vcol = tid % VCH, vsub = tid / VCH;
  const float inv_keep = 1.0f / a.keep;
  const float bias = a.att_b[0];
  unsigned int vc = 0;
  // sample-independent coefficients stay in registers for the whole kernel (D <= 2 * AP_CONSUMERS: two columns per
  // thread); the per-sample hq values are fetched at the top of each sample, ahead of the statistics pass that hides
  // their latency (ncu: long-scoreboard stalls of the coefficient set-up)
  const bool coef_regs = D <= 2 * AP_CONSUMERS;
  float g_r[2] = {0.f, 0.f}, be_r[2] = {0.f, 0.f}, w_r[2] = {0.f, 0.f};
  if (coef_regs) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int d = tid + q * AP_CONSUMERS;
      if (d < D) {
        g_r[q] = a.gamma[d];
        be_r[q] = a.beta[d];
        w_r[q] = a.att_w[d] * inv_keep;
      }
    }
  }
  for (int i = 0; i < n_my; ++i) {
    const int b = first + i * stride;
    const int j = i & 1;
    const bf16* zs = reinterpret_cast<const bf16*>(zbuf + static_cast<size_t>(j) * zbytes);
    int nb = a.nbox[b];
    nb = nb < 0 ? 0 : (nb > K ? K : nb);
    float hq_r[2] = {0.f, 0.f};
    if (coef_regs) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int d = tid + q * AP_CONSUMERS;
        if (d < D) hq_r[q] = a.hq[static_cast<long long>(b) * D + d];
      }
    }
    ptx::mbar_wait(&zfull[j], (i >> 1) & 1);

    // ---- statistics over the K*D slab (one pass, Chan merge) ----
    float n = 0.f, mean = 0.f, m2 = 0.f;
    for (int c = tid; c < nchunks; c += AP_CONSUMERS) {
      float x[8];
      unpack8(*reinterpret_cast<const uint4*>(zs + static_cast<size_t>(c) * 8), x);
      float cm = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) cm += x[q];
      cm *= 0.125f;
      float c2 = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) c2 += (x[q] - cm) * (x[q] - cm);
      chan_merge(n, mean, m2, 8.f, cm, c2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float nb2 = __shfl_xor_sync(0xffffffffu, n, o);
      const float mb = __shfl_xor_sync(0xffffffffu, mean, o);
      const float qb = __shfl_xor_sync(0xffffffffu, m2, o);
      chan_merge(n, mean, m2, nb2, mb, qb);
    }
    if (lane == 0) {
      red[3 * warp] = n;
      red[3 * warp + 1] = mean;
      red[3 * warp + 2] = m2;
    }
    cons_bar();   // (also: everybody is done with the previous sample's poolx = cA / cB)
    if (tid == 0) {
      float tn = 0.f, tm = 0.f, tq = 0.f;
      for (int w = 0; w < AP_CONSUMER_WARPS; ++w) chan_merge(tn, tm, tq, red[3 * w], red[3 * w + 1], red[3 * w + 2]);
      red[60] = tm;
      red[61] = 1.0f / sqrtf(tq / tn + 1e-12f);   // biased variance (tf.nn.moments), eps of layers.layer_norm
    }
    cons_bar();
    const float mu = red[60], rstd = red[61];
    if (coef_regs) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int d = tid + q * AP_CONSUMERS;
        if (d < D) {
          const float g = g_r[q] * rstd;
          cA[d] = g;
          cB[d] = be_r[q] - mu * g;
          cC[d] = hq_r[q] * w_r[q];
        }
      }
    } else {
      for (int d = tid; d < D; d += AP_CONSUMERS) {
        const float g = a.gamma[d] * rstd;
        cA[d] = g;
        cB[d] = a.beta[d] - mu * g;
        cC[d] = a.hq[static_cast<long long>(b) * D + d] * a.att_w[d] * inv_keep;
      }
    }
    cons_bar();

    // ---- scores: one warp per box row ----
    for (int k = warp; k < nb; k += AP_CONSUMER_WARPS) {
      const bf16* zr = zs + static_cast<size_t>(k) * D;
      const unsigned long long g0 = (static_cast<unsigned long long>(b) * K + k) * CH;
      float acc = 0.f;
      for (int c = lane; c < CH; c += 32) {
        float x[8];
        unpack8(*reinterpret_cast<const uint4*>(zr + c * 8), x);
        uint32_t bits = 0xFFu;
        if (a.thr < 65536u) bits = philox_keep_bits(philox4x32_10(g0 + c, RNG_STREAM_ATT, a.seed, a.step), a.thr);
        const int d0 = c * 8;
        const float4 a0 = *reinterpret_cast<const float4*>(cA + d0), a1 = *reinterpret_cast<const float4*>(cA + d0 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(cB + d0), b1 = *reinterpret_cast<const float4*>(cB + d0 + 4);
        const float4 w0 = *reinterpret_cast<const float4*>(cC + d0), w1 = *reinterpret_cast<const float4*>(cC + d0 + 4);
        const float ca[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float cb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const float cw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float y = fmaxf(fmaf(x[q], ca[q], cb[q]), 0.f);
          acc += ((bits >> q) & 1u) ? y * cw[q] : 0.f;
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) sc[k] = acc + bias;
    }
    // this warp is done with the slab: hand the buffer back to the producer (sample i + 2 goes there)
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&zempty[j]);
    cons_bar();

    // ---- masked softmax over boxes: exact zeros beyond nbox ----
    if (warp == 0) {
      float mx = -CUDART_INF_F;
      for (int k = lane; k < nb; k += 32) mx = fmaxf(mx, sc[k]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int k = lane; k < nb; k += 32) {
        const float e = expf(sc[k] - mx);
        sc[k] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;   // nb == 0 -> 1/0: TF yields NaN for an all-masked row as well
      for (int k = lane; k < K; k += 32) {
        const float p = k < nb ? sc[k] * inv : (nb == 0 ? CUDART_NAN_F : 0.f);
        sc[k] = p;
        if (a.att) a.att[static_cast<long long>(b) * K + k] = p;
      }
      if (lane == 0) {
        a.ln_mean[b] = mu;
        a.ln_rstd[b] = rstd;
      }
    }
    cons_bar();

    // ---- attended pooling of the raw features out of the ring ----
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int k0 = 0; k0 < nb; k0 += RV, ++vc) {
      const int slot = vc & 1;
      const int rows = nb - k0 < RV ? nb - k0 : RV;
      ptx::mbar_wait(&vfull[slot], (vc >> 1) & 1);
      const uint8_t* vs = vbuf + slot * AP_VCHUNK_BYTES;
      if (vsub < tpc) {
        for (int r = vsub; r < rows; r += tpc) {
          float v[8];
          unpack8(*reinterpret_cast<const uint4*>(vs + static_cast<size_t>(r) * vrow + vcol * 16), v);
          const float ak = sc[k0 + r];
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] = fmaf(ak, v[q], acc[q]);
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&vempty[slot]);
    }
    if (nb == 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = CUDART_NAN_F;
    }
    // combine the tpc row groups of every column chunk (fixed order) and write
    if (tpc > 1) {
      if (vsub > 0 && vsub < tpc) {
        float* px = poolx + (static_cast<size_t>(vsub - 1) * VCH + vcol) * 8;
        *reinterpret_cast<float4*>(px) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(px + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      cons_bar();
    }
    if (vsub == 0) {
      for (int s = 1; s < tpc; ++s) {
        const float* px = poolx + (static_cast<size_t>(s - 1) * VCH + vcol) * 8;
        const float4 p0 = *reinterpret_cast<const float4*>(px), p1 = *reinterpret_cast<const float4*>(px + 4);
        acc[0] += p0.x; acc[1] += p0.y; acc[2] += p0.z; acc[3] += p0.w;
        acc[4] += p1.x; acc[5] += p1.y; acc[6] += p1.z; acc[7] += p1.w;
      }
      const long long o = static_cast<long long>(b) * Dv + vcol * 8;
      if (a.pooled) {
        *reinterpret_cast<float4*>(a.pooled + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(a.pooled + o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      if (a.pooled_bf) {
        __nv_bfloat162 h[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) h[q] = __nv_bfloat162(__float2bfloat16_rn(acc[2 * q]), __float2bfloat16_rn(acc[2 * q + 1]));
        *reinterpret_cast<uint4*>(a.pooled_bf + o) = *reinterpret_cast<uint4*>(h);
      }
    }
    // (the next sample's first cons_bar orders these poolx reads before cA / cB are rewritten)
  }
}

}  // namespace

// the pipelined kernel covers the bf16 mode with one feature plane when both slab buffers, the feature ring and the
// per-column constants fit one SM; everything else stays on attn_fwd_kernel
bool attn_fwd_pipe_supported(int K, int D, int Dv, int precision, bool has_v_lo, size_t* smem_out, int* rv_out) {
  if (precision != VQA_PREC_BF16 || has_v_lo) return false;
  if (getenv("VQA_ATTN_PIPE") && atoi(getenv("VQA_ATTN_PIPE")) == 0) return false;
  const int VCH = Dv >> 3;
  if ((D & 7) || (Dv & 7) || VCH > AP_CONSUMERS || AP_CONSUMERS % VCH != 0) return false;
  const int tpc = AP_CONSUMERS / VCH;
  if (static_cast<size_t>(tpc - 1) * VCH * 8 > 2 * static_cast<size_t>(D)) return false;   // poolx aliases cA | cB
  const size_t vrow = static_cast<size_t>(Dv) * 2;
  const int rv = static_cast<int>(AP_VCHUNK_BYTES / vrow);
  if (rv < 1) return false;
  const size_t zbytes = static_cast<size_t>(K) * D * 2;
  if (zbytes % 128 != 0) return false;
  const size_t smem = 2 * zbytes + 2 * AP_VCHUNK_BYTES + (3 * static_cast<size_t>(D) + ((K + 3) & ~3) + 64) * 4 + 64 + 128;
  if (smem > 227 * 1024) return false;
  *smem_out = smem;
  *rv_out = rv;
  return true;
}

VqaStatus attn_fwd_pipe_launch(const VqaAttnFwd& a, int K, int D, int Dv, float keep, size_t smem, int rv,
                               int num_sms, cudaStream_t s) {
  PipeFwdArgs f{};
  f.z = static_cast<const bf16*>(a.z); f.gamma = a.gamma; f.beta = a.beta; f.hq = a.hq; f.att_w = a.att_w;
  f.att_b = a.att_b; f.nbox = a.nbox; f.v = static_cast<const bf16*>(a.v_hi); f.seed = a.seed; f.step = a.step;
  f.att = a.att; f.pooled = a.pooled; f.pooled_bf = static_cast<bf16*>(a.pooled_hi); f.ln_mean = a.ln_mean;
  f.ln_rstd = a.ln_rstd; f.batch = a.batch; f.K = K; f.D = D; f.Dv = Dv; f.RV = rv; f.keep = keep;
  f.thr = keep_threshold(keep);
  static size_t smem_set = 0;
  if (smem_set < smem) {
    VQA_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    smem_set = smem;
  }
  const int grid = a.batch < num_sms ? a.batch : num_sms;
  launch_pdl(attn_fwd_pipe_kernel, dim3(grid), dim3(AP_THREADS), smem, s, f);
  VQA_LAUNCH_CHECK("attn_fwd (pipelined)");
  return VQA_OK;
}

}  // namespace vqa
