// The memory-bound attention block, one CTA per sample.
//
// forward  (vlmap/modules.py:67-97 hadamard_attention, :23-39 attention_pooling, and the LayerNorm +
//           ReLU of the producing fc_layer, :646-649, which cannot finish in the GEMM epilogue because
//           TF normalises the [K, D] slab of a sample jointly -- SURVEY Q1):
//   stats over the K*D pre-LN projection (one pass, Chan/Welford merge, fp32)
//   hv = relu(LN(z));  F = dropout_0.8(hv * hq);  s_k = F_k . w + b;  s_k = -inf for k >= nbox
//   a = softmax_K(s);  pooled = sum_k a_k V_k  (the RAW region features, not hv)
// backward (hand-derived): da_k = <V_k, dP>; ds = a (da - sum a da);
//   then per element dF, dhv, the ReLU gate, and the joint-(K,D) LayerNorm backward, all from the
//   re-read z; emits dz (GEMM operand planes), dHq and per-CTA partials of dw, db, dgamma, dbeta, dbias.
//
// Access pattern: every global access is a 128-bit load/store of 8 bf16 (or 2x4 fp32) consecutive
// elements, consecutive lanes on consecutive 16-byte chunks; reductions by warp shuffle.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "internal.h"
#include "launch.cuh"
#include "philox.cuh"
#include "ptx.cuh"

namespace vqa {

namespace {

constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;

__device__ __forceinline__ void load8(const bf16* p, float (&x)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    x[2 * j] = f.x;
    x[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&x)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
  x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void ld8f(const float* p, float (&x)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
// features are streamed once per pass: read-only path
__device__ __forceinline__ void load8_planes(const bf16* hi, const bf16* lo, long long off,
                                             float (&x)[8]) {
  load8(hi + off, x);
  if (lo) {
    float y[8];
    load8(lo + off, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
  }
}
__device__ __forceinline__ void store8_planes(bf16* hi, bf16* lo, long long off, const float (&x)[8]) {
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

// 1-D bulk async copy global -> shared (TMA), completion on an mbarrier; size % 16 == 0, 16-byte aligned
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ptx::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
// the [K, D] pre-LN slab of one sample -> shared memory, in pieces of <= 32 KB (one thread issues)
__device__ __forceinline__ void slab_to_smem(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  ptx::mbar_arrive_expect_tx(bar, bytes);
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
    bulk_load(static_cast<uint8_t*>(smem_dst) + off, static_cast<const uint8_t*>(gsrc) + off, n, bar);
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
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float mb,
                                           float m2b) {
  const float nt = n + nb;
  if (nt == 0.f) return;
  const float delta = mb - mean;
  const float f = nb / nt;
  mean += delta * f;
  m2 += m2b + delta * delta * n * f;
  n = nt;
}

// mean / rstd over the K*D slab of one sample; result broadcast to the whole CTA
template <typename ZT>
__device__ __forceinline__ void slab_stats(const ZT* zb, int nchunks, float* red /*[3*ATT_WARPS+2]*/,
                                           float& mean_out, float& rstd_out) {
  float n = 0.f, mean = 0.f, m2 = 0.f;
  // loads go out in batches of U before any arithmetic: the pass is latency-bound otherwise
  constexpr int U = 6;
  for (int c0 = threadIdx.x; c0 < nchunks; c0 += U * ATT_THREADS) {
    float x[U][8];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int c = c0 + i * ATT_THREADS;
      if (c < nchunks) load8(zb + static_cast<long long>(c) * 8, x[i]);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int c = c0 + i * ATT_THREADS;
      if (c < nchunks) {
        float cm = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) cm += x[i][j];
        cm *= 0.125f;
        float c2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) c2 += (x[i][j] - cm) * (x[i][j] - cm);
        chan_merge(n, mean, m2, 8.f, cm, c2);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float nb = __shfl_xor_sync(0xffffffffu, n, o);
    const float mb = __shfl_xor_sync(0xffffffffu, mean, o);
    const float qb = __shfl_xor_sync(0xffffffffu, m2, o);
    chan_merge(n, mean, m2, nb, mb, qb);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[3 * warp] = n;
    red[3 * warp + 1] = mean;
    red[3 * warp + 2] = m2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tn = 0.f, tm = 0.f, tq = 0.f;
    for (int w = 0; w < ATT_WARPS; ++w) chan_merge(tn, tm, tq, red[3 * w], red[3 * w + 1], red[3 * w + 2]);
    const float var = tq / tn;  // biased variance (tf.nn.moments)
    red[3 * ATT_WARPS] = tm;
    red[3 * ATT_WARPS + 1] = 1.0f / sqrtf(var + 1e-12f);
  }
  __syncthreads();
  mean_out = red[3 * ATT_WARPS];
  rstd_out = red[3 * ATT_WARPS + 1];
}

struct FwdArgs {
  const void* z;
  const float* gamma; const float* beta; const float* hq; const float* att_w; const float* att_b;
  const int* nbox; const bf16* v_hi; const bf16* v_lo;
  unsigned long long seed, step;
  float* att; float* pooled; bf16* pooled_hi; bf16* pooled_lo; float* ln_mean; float* ln_rstd;
  const unsigned char* keep_bits;   // [batch*K*D/8] keep bits of this step (NULL: drawn here with Philox)
};

// SLAB: the sample's [K, D] pre-LN slab is brought into shared memory by ONE bulk async copy and both passes
// over it (statistics, scores) read shared memory; the feature slab is prefetched into L2 meanwhile.
template <typename ZT, bool SLAB>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(FwdArgs a, int K, int D, int Dv,
                                                               float keep, uint32_t thr) {
  extern __shared__ __align__(16) float sm[];
  float* cA = sm;           // gamma_d * rstd
  float* cB = cA + D;       // beta_d - mean * rstd * gamma_d
  float* cC = cB + D;       // hq_d * w_d / keep
  float* sc = cC + D;       // [K] scores -> attention
  float* red = sc + K;      // [3*ATT_WARPS + 2]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int CH = D >> 3;
  const ZT* zb = static_cast<const ZT*>(a.z) + static_cast<long long>(b) * K * D;
  pdl_sync();
  if (SLAB) {
    const int head = ((3 * D + K + 3 * ATT_WARPS + 2) * 4 + 15) & ~15;
    uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sm) + head);
    ZT* zs = reinterpret_cast<ZT*>(bar + 2);
    if (tid == 0) {
      ptx::mbar_init(bar, 1);
      ptx::fence_barrier_init();
      slab_to_smem(zs, zb, static_cast<uint32_t>(K) * D * sizeof(ZT), bar);
      // the raw features of this sample are streamed later by the pooling pass: start them towards L2 now
      const uint8_t* vsrc = reinterpret_cast<const uint8_t*>(a.v_hi + static_cast<long long>(b) * K * Dv);
      const uint32_t vbytes = static_cast<uint32_t>(K) * Dv * 2;
      for (uint32_t off = 0; off < vbytes; off += 65536u)
        bulk_prefetch_l2(vsrc + off, vbytes - off < 65536u ? vbytes - off : 65536u);
    }
    __syncthreads();
    ptx::mbar_wait(bar, 0);
    zb = zs;
  }

  float mean, rstd;
  slab_stats<ZT>(zb, K * CH, red, mean, rstd);

  const float inv_keep = 1.0f / keep;
  for (int d = tid; d < D; d += ATT_THREADS) {
    const float g = a.gamma[d] * rstd;
    cA[d] = g;
    cB[d] = a.beta[d] - mean * g;
    cC[d] = a.hq[static_cast<long long>(b) * D + d] * a.att_w[d] * inv_keep;
  }
  int nb = a.nbox[b];
  nb = nb < 0 ? 0 : (nb > K ? K : nb);
  __syncthreads();

  // scores: one warp per box row
  const float bias = a.att_b[0];
  for (int k = warp; k < nb; k += ATT_WARPS) {
    const ZT* zr = zb + static_cast<long long>(k) * D;
    const unsigned long long g0 = (static_cast<unsigned long long>(b) * K + k) * CH;
    float acc = 0.f;
    constexpr int U2 = 4;
    for (int c0 = lane; c0 < CH; c0 += 32 * U2) {
      float x[U2][8];
#pragma unroll
      for (int i = 0; i < U2; ++i) {
        const int c = c0 + 32 * i;
        if (c < CH) load8(zr + c * 8, x[i]);
      }
#pragma unroll
      for (int i = 0; i < U2; ++i) {
        const int c = c0 + 32 * i;
        if (c < CH) {
          uint32_t bits = 0xFFu;
          if (thr < 65536u)
            bits = a.keep_bits ? static_cast<uint32_t>(__ldg(a.keep_bits + g0 + c))
                               : philox_keep_bits(philox4x32_10(g0 + c, RNG_STREAM_ATT, a.seed, a.step), thr);
          const int d0 = c * 8;
          const float4 a0 = *reinterpret_cast<const float4*>(cA + d0), a1 = *reinterpret_cast<const float4*>(cA + d0 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(cB + d0), b1 = *reinterpret_cast<const float4*>(cB + d0 + 4);
          const float4 w0 = *reinterpret_cast<const float4*>(cC + d0), w1 = *reinterpret_cast<const float4*>(cC + d0 + 4);
          const float ca[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float cb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          const float cw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float y = fmaxf(fmaf(x[i][j], ca[j], cb[j]), 0.f);
            acc += ((bits >> j) & 1u) ? y * cw[j] : 0.f;
          }
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) sc[k] = acc + bias;
  }
  __syncthreads();

  // masked softmax over boxes (tf.where(mask, s, -inf) -> tf.nn.softmax): exact zeros beyond nbox
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
    const float inv = 1.0f / sum;  // nb == 0 -> 1/0: TF yields NaN for an all-masked row as well
    for (int k = lane; k < K; k += 32) {
      const float p = k < nb ? sc[k] * inv : (nb == 0 ? CUDART_NAN_F : 0.f);
      sc[k] = p;
      if (a.att) a.att[static_cast<long long>(b) * K + k] = p;
    }
  }
  __syncthreads();

  // attended pooling of the raw features
  const long long vb = static_cast<long long>(b) * K * Dv;
  for (int c = tid; c < (Dv >> 3); c += ATT_THREADS) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    int k = 0;
    constexpr int UP = 6;
    for (; k + UP <= nb; k += UP) {
      float v[UP][8];
#pragma unroll
      for (int i = 0; i < UP; ++i)
        load8_planes(a.v_hi, a.v_lo, vb + static_cast<long long>(k + i) * Dv + c * 8, v[i]);
#pragma unroll
      for (int i = 0; i < UP; ++i) {
        const float ai = sc[k + i];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(ai, v[i][j], acc[j]);
      }
    }
    for (; k < nb; ++k) {
      float v0[8];
      load8_planes(a.v_hi, a.v_lo, vb + static_cast<long long>(k) * Dv + c * 8, v0);
      const float a0 = sc[k];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(a0, v0[j], acc[j]);
    }
    if (nb == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = CUDART_NAN_F;
    }
    const long long o = static_cast<long long>(b) * Dv + c * 8;
    if (a.pooled) {
      *reinterpret_cast<float4*>(a.pooled + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(a.pooled + o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (a.pooled_hi) store8_planes(a.pooled_hi, a.pooled_lo, o, acc);
  }
  if (tid == 0) {
    a.ln_mean[b] = mean;
    a.ln_rstd[b] = rstd;
  }
}

struct BwdArgs {
  const void* z;
  const float* gamma; const float* beta; const float* hq; const float* att_w;
  const int* nbox; const bf16* v_hi; const bf16* v_lo;
  unsigned long long seed, step;
  const float* att; const float* ln_mean; const float* ln_rstd; const float* d_pooled;
  bf16* dz_hi; bf16* dz_lo; float* d_hq;
  float* part;  // [batch, 4*D + 8]: T*hq/keep (dw) | dgamma | dbeta | dbias | db
  const unsigned char* keep_bits;   // [batch*K*D/8] the forward's keep bits (NULL: regenerated here with Philox)
  unsigned long long* trace;        // optional [batch][8] globaltimer stamps
  int vring_rows;                   // > 0: the <V, dP> pass streams the feature rows through a 3-slot ring of this many rows
                                    // laid over the (not yet loaded) slab buffer (bulk copies) instead of global loads
  // optional tail (qv_z != NULL): q_linear_v's ReLU / LayerNorm backward on this sample's d_hq row -- the CTA holds the
  // whole row, so the row kernel between this kernel and the q-projection data gradient disappears from the path
  const float* qv_z; const float* qv_gamma; const float* qv_mean; const float* qv_rstd;
  bf16* qv_dz_hi; bf16* qv_dz_lo; float* qv_dz_f32; float* qv_dgamma_part; float* qv_dbeta_part;
};
constexpr int AB_VSLOTS = 3;
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define AB_TRACE(slot)                                                                     \
  do {                                                                                     \
    if (a.trace && threadIdx.x == 0) a.trace[static_cast<size_t>(blockIdx.x) * 8 + (slot)] = gtimer_ns(); \
  } while (0)

// NCOL = column chunks (of 8) owned per thread: D <= 2048 -> 1, D <= 4096 -> 2
template <typename ZT, int NCOL, bool SLAB>
__global__ void __launch_bounds__(ATT_THREADS, SLAB ? 2 : 1) attn_bwd_kernel(BwdArgs a, int K, int D, int Dv,
                                                                             float keep, uint32_t thr) {
  extern __shared__ __align__(16) float sm[];
  const int CH = D >> 3;
  float* sdP = sm;                  // [Dv]
  float* ds = sdP + Dv;             // [K] (first da, then ds)
  float* red = ds + K;              // [64]
  float* colacc = red + 64;         // [3][ATT_THREADS][8*NCOL] per-thread column accumulators
  unsigned char* flags = reinterpret_cast<unsigned char*>(colacc + 3 * ATT_THREADS * 8 * NCOL);  // [K*CH]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const ZT* zb = static_cast<const ZT*>(a.z) + static_cast<long long>(b) * K * D;
  uint64_t* zbar = nullptr;
  uint64_t* vfull = nullptr;
  uint64_t* vempty = nullptr;
  ZT* zs_mem = nullptr;
  const ZT* zb_global = zb;
  pdl_sync();
  if (SLAB) {
    // the slab copy is in flight while the feature rows are reduced against dP below
    const size_t head = ((static_cast<size_t>(Dv) + K + 64 + 3 * ATT_THREADS * 8 * NCOL) * 4 +
                         static_cast<size_t>(K) * CH + 15) & ~static_cast<size_t>(15);
    zbar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sm) + head);
    ZT* zs = reinterpret_cast<ZT*>(zbar + 8);
    vfull = zbar + 1;
    vempty = zbar + 1 + AB_VSLOTS;
    zs_mem = zs;
    if (tid == 0) {
      ptx::mbar_init(zbar, a.keep_bits ? 2 : 1);
      for (int i = 0; i < AB_VSLOTS; ++i) {
        ptx::mbar_init(&vfull[i], 1);
        ptx::mbar_init(&vempty[i], ATT_WARPS);
      }
      ptx::fence_barrier_init();
      if (a.vring_rows > 0) {
        // the slab buffer first serves as a ring for the feature rows of the <V, dP> pass; the slab itself is fetched
        // when that pass is over (it is needed from the column pass on)
        int nb0 = a.nbox[b];
        nb0 = nb0 < 0 ? 0 : (nb0 > K ? K : nb0);
        const uint32_t vrow = static_cast<uint32_t>(Dv) * 2;
        const uint8_t* vsrc = reinterpret_cast<const uint8_t*>(a.v_hi) + static_cast<size_t>(b) * K * vrow;
        for (int ch = 0; ch < AB_VSLOTS && ch * a.vring_rows < nb0; ++ch) {
          const int rows = nb0 - ch * a.vring_rows < a.vring_rows ? nb0 - ch * a.vring_rows : a.vring_rows;
          ptx::mbar_arrive_expect_tx(&vfull[ch], rows * vrow);
          bulk_load(reinterpret_cast<uint8_t*>(zs) + static_cast<size_t>(ch) * a.vring_rows * vrow,
                    vsrc + static_cast<size_t>(ch) * a.vring_rows * vrow, rows * vrow, &vfull[ch]);
        }
      } else {
        slab_to_smem(zs, zb, static_cast<uint32_t>(K) * D * sizeof(ZT), zbar);
        if (a.keep_bits) {
          // the forward's keep bits of this slab land in `flags`; the column pass below turns them into gate flags in place
          ptx::mbar_arrive_expect_tx(zbar, static_cast<uint32_t>(K) * CH);
          bulk_load(flags, a.keep_bits + static_cast<size_t>(b) * K * CH, static_cast<uint32_t>(K) * CH, zbar);
        }
      }
    }
    zb_global = zb;
    zb = zs;
  }
  AB_TRACE(0);
  const float mean = a.ln_mean[b], rstd = a.ln_rstd[b];
  int nb = a.nbox[b];
  nb = nb < 0 ? 0 : (nb > K ? K : nb);

  for (int d = tid * 4; d < Dv; d += ATT_THREADS * 4)
    *reinterpret_cast<float4*>(sdP + d) =
        *reinterpret_cast<const float4*>(a.d_pooled + static_cast<long long>(b) * Dv + d);
  __syncthreads();
  AB_TRACE(1);

  // da_k = <V_k, dP>: one warp per box row
  const long long vb = static_cast<long long>(b) * K * Dv;
  const bool vring = SLAB && a.vring_rows > 0;
  if (vring) {
    const int RV = a.vring_rows;
    const uint32_t vrow = static_cast<uint32_t>(Dv) * 2;
    const uint8_t* vsrc = reinterpret_cast<const uint8_t*>(a.v_hi) + static_cast<size_t>(b) * K * vrow;
    uint8_t* ringb = reinterpret_cast<uint8_t*>(zs_mem);
    const int nch = (nb + RV - 1) / RV;
    for (int k = nb + tid; k < K; k += ATT_THREADS) ds[k] = 0.f;
    for (int ch = 0; ch < nch; ++ch) {
      const int slot = ch % AB_VSLOTS;
      const int rows = nb - ch * RV < RV ? nb - ch * RV : RV;
      ptx::mbar_wait(&vfull[slot], (ch / AB_VSLOTS) & 1);
      const uint8_t* vs = ringb + static_cast<size_t>(slot) * RV * vrow;
      for (int r = warp; r < rows; r += ATT_WARPS) {
        float acc = 0.f;
        for (int c = lane; c < (Dv >> 3); c += 32) {
          uint4 raw;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                       : "r"(ptx::smem_u32(vs + static_cast<size_t>(r) * vrow + c * 16)));
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&raw);
          const float2 f0 = __bfloat1622float2(hh[0]), f1 = __bfloat1622float2(hh[1]);
          const float2 f2 = __bfloat1622float2(hh[2]), f3 = __bfloat1622float2(hh[3]);
          const float4 p0 = *reinterpret_cast<const float4*>(sdP + c * 8);
          const float4 p1 = *reinterpret_cast<const float4*>(sdP + c * 8 + 4);
          acc = fmaf(f0.x, p0.x, acc); acc = fmaf(f0.y, p0.y, acc);
          acc = fmaf(f1.x, p0.z, acc); acc = fmaf(f1.y, p0.w, acc);
          acc = fmaf(f2.x, p1.x, acc); acc = fmaf(f2.y, p1.y, acc);
          acc = fmaf(f3.x, p1.z, acc); acc = fmaf(f3.y, p1.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) ds[ch * RV + r] = acc;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&vempty[slot]);
      // refill: the thread that issues waits until all eight warps have left this slot (completion ch / AB_VSLOTS + 1)
      if (tid == ATT_THREADS - 1 && (ch + AB_VSLOTS) * RV < nb) {
        ptx::mbar_wait(&vempty[slot], (ch / AB_VSLOTS) & 1);
        const int nxt = ch + AB_VSLOTS;
        const int nrows = nb - nxt * RV < RV ? nb - nxt * RV : RV;
        ptx::fence_proxy_async();   // the generic-proxy reads of this slot precede the async-proxy refill
        ptx::mbar_arrive_expect_tx(&vfull[slot], nrows * vrow);
        bulk_load(ringb + static_cast<size_t>(slot) * RV * vrow, vsrc + static_cast<size_t>(nxt) * RV * vrow, nrows * vrow, &vfull[slot]);
      }
    }
    __syncthreads();   // every warp is done with the ring: the slab (and its keep bits) may land on it
    if (tid == 0) {
      ptx::fence_proxy_async();
      slab_to_smem(zs_mem, zb_global, static_cast<uint32_t>(K) * D * sizeof(ZT), zbar);
      if (a.keep_bits) {
        ptx::mbar_arrive_expect_tx(zbar, static_cast<uint32_t>(K) * CH);
        bulk_load(flags, a.keep_bits + static_cast<size_t>(b) * K * CH, static_cast<uint32_t>(K) * CH, zbar);
      }
    }
  } else
  for (int k = warp; k < K; k += ATT_WARPS) {
    float acc = 0.f;
    if (k < nb && !a.v_lo) {
      // single bf16 plane: keep the raw 16-byte words in registers (4 per chunk instead of 8 floats) so that EIGHT loads
      // per lane are in flight -- the pass is bound by memory latency at 16 warps per SM (ncu: long scoreboard)
      constexpr int UR = 8;
      const bf16* vrow = a.v_hi + vb + static_cast<long long>(k) * Dv;
      for (int c0 = lane; c0 < (Dv >> 3); c0 += 32 * UR) {
        uint4 raw[UR];
#pragma unroll
        for (int i = 0; i < UR; ++i) {
          const int c = c0 + 32 * i;
          if (c < (Dv >> 3)) raw[i] = __ldg(reinterpret_cast<const uint4*>(vrow + c * 8));
        }
#pragma unroll
        for (int i = 0; i < UR; ++i) {
          const int c = c0 + 32 * i;
          if (c < (Dv >> 3)) {
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&raw[i]);
            const float2 f0 = __bfloat1622float2(hh[0]), f1 = __bfloat1622float2(hh[1]);
            const float2 f2 = __bfloat1622float2(hh[2]), f3 = __bfloat1622float2(hh[3]);
            const float4 p0 = *reinterpret_cast<const float4*>(sdP + c * 8);
            const float4 p1 = *reinterpret_cast<const float4*>(sdP + c * 8 + 4);
            acc = fmaf(f0.x, p0.x, acc); acc = fmaf(f0.y, p0.y, acc);
            acc = fmaf(f1.x, p0.z, acc); acc = fmaf(f1.y, p0.w, acc);
            acc = fmaf(f2.x, p1.x, acc); acc = fmaf(f2.y, p1.y, acc);
            acc = fmaf(f3.x, p1.z, acc); acc = fmaf(f3.y, p1.w, acc);
          }
        }
      }
      acc = warp_sum(acc);
    } else if (k < nb) {
      constexpr int UB = 4;
      for (int c0 = lane; c0 < (Dv >> 3); c0 += 32 * UB) {
        float v[UB][8];
#pragma unroll
        for (int i = 0; i < UB; ++i) {
          const int c = c0 + 32 * i;
          if (c < (Dv >> 3)) load8_planes(a.v_hi, a.v_lo, vb + static_cast<long long>(k) * Dv + c * 8, v[i]);
        }
#pragma unroll
        for (int i = 0; i < UB; ++i) {
          const int c = c0 + 32 * i;
          if (c < (Dv >> 3)) {
            const float4 p0 = *reinterpret_cast<const float4*>(sdP + c * 8);
            const float4 p1 = *reinterpret_cast<const float4*>(sdP + c * 8 + 4);
            acc = fmaf(v[i][0], p0.x, acc); acc = fmaf(v[i][1], p0.y, acc);
            acc = fmaf(v[i][2], p0.z, acc); acc = fmaf(v[i][3], p0.w, acc);
            acc = fmaf(v[i][4], p1.x, acc); acc = fmaf(v[i][5], p1.y, acc);
            acc = fmaf(v[i][6], p1.z, acc); acc = fmaf(v[i][7], p1.w, acc);
          }
        }
      }
      acc = warp_sum(acc);
    }
    if (lane == 0) ds[k] = acc;
  }
  __syncthreads();
  AB_TRACE(2);
  // ds_k = a_k (da_k - sum_j a_j da_j); masked slots have a_k = 0 -> ds_k = 0
  if (warp == 0) {
    float dot = 0.f;
    for (int k = lane; k < nb; k += 32) dot += a.att[static_cast<long long>(b) * K + k] * ds[k];
    dot = warp_sum(dot);
    float dbs = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float v = k < nb ? a.att[static_cast<long long>(b) * K + k] * (ds[k] - dot) : 0.f;
      ds[k] = v;
      dbs += v;
    }
    dbs = warp_sum(dbs);
    if (lane == 0) red[32] = dbs;
  }
  __syncthreads();

  AB_TRACE(3);
  if (SLAB) ptx::mbar_wait(zbar, 0);  // (the mbarrier init was published by the __syncthreads above)
  AB_TRACE(4);
  // column-owner mapping: thread (tc, tr) owns column chunks tc (+ CW) and walks rows tr, tr+RP, ...
  const int CW = CH < ATT_THREADS ? CH : ATT_THREADS;
  const int RP = ATT_THREADS / CW;
  const int tc = tid % CW, tr = tid / CW;
  const bool active = tr < RP;

  float accT[NCOL][8], accU[NCOL][8], accV[NCOL][8];
#pragma unroll
  for (int i = 0; i < NCOL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) accT[i][j] = accU[i][j] = accV[i][j] = 0.f;

  const float mr = mean * rstd;
  if (active) {
#pragma unroll
    for (int i = 0; i < NCOL; ++i) {
      const int c = tc + i * CW;
      if (c >= CH) break;
      float g[8], bt[8];
      ld8f(a.gamma + c * 8, g);   // two 128-bit loads per array instead of eight scalar ones
      ld8f(a.beta + c * 8, bt);
      constexpr int UD = SLAB ? 2 : 3;
      for (int k0 = tr; k0 < nb; k0 += RP * UD) {
        float xb[UD][8];
        uint32_t kbits[UD];
#pragma unroll
        for (int u = 0; u < UD; ++u) {
          const int k = k0 + u * RP;
          kbits[u] = 0xFFu;
          if (k < nb) {
            load8(zb + static_cast<long long>(k) * D + c * 8, xb[u]);
            if (thr < 65536u && a.keep_bits) kbits[u] = SLAB ? flags[k * CH + c] : __ldg(a.keep_bits + (static_cast<unsigned long long>(b) * K + k) * CH + c);
          }
        }
#pragma unroll
        for (int u = 0; u < UD; ++u) {
          const int k = k0 + u * RP;
          if (k >= nb) break;
          uint32_t bits = kbits[u];
          if (thr < 65536u && !a.keep_bits)
            bits = philox_keep_bits(
                philox4x32_10((static_cast<unsigned long long>(b) * K + k) * CH + c, RNG_STREAM_ATT,
                              a.seed, a.step),
                thr);
          const float dsk = ds[k];
          uint32_t fl = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = fmaf(xb[u][j], rstd, -mr);
            const float y = fmaf(xh, g[j], bt[j]);
            const bool pos = y > 0.f;
            const bool m = (bits >> j) & 1u;
            accT[i][j] += (m && pos) ? dsk * y : 0.f;
            if (m && pos) {
              fl |= 1u << j;
              accU[i][j] = fmaf(dsk, xh, accU[i][j]);
              accV[i][j] += dsk;
            }
          }
          flags[k * CH + c] = static_cast<unsigned char>(fl);
        }
      }
    }
  }
  AB_TRACE(5);
  // combine the RP row groups deterministically through shared memory
#pragma unroll
  for (int i = 0; i < NCOL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      colacc[(0 * ATT_THREADS + tid) * 8 * NCOL + i * 8 + j] = accT[i][j];
      colacc[(1 * ATT_THREADS + tid) * 8 * NCOL + i * 8 + j] = accU[i][j];
      colacc[(2 * ATT_THREADS + tid) * 8 * NCOL + i * 8 + j] = accV[i][j];
    }
  __syncthreads();
  const float inv_keep = 1.0f / keep;
  float s1 = 0.f, s2 = 0.f;  // sum dxhat, sum dxhat * xhat over the slab
  float* part = a.part + static_cast<long long>(b) * (4 * D + 8);
  float G[NCOL][8];
#pragma unroll
  for (int i = 0; i < NCOL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) G[i][j] = 0.f;
  if (active && tr == 0) {
#pragma unroll
    for (int i = 0; i < NCOL; ++i) {
      const int c = tc + i * CW;
      if (c >= CH) break;
      float w8[8], hq8[8], gm8[8];
      ld8f(a.att_w + c * 8, w8);
      ld8f(a.hq + static_cast<long long>(b) * D + c * 8, hq8);
      ld8f(a.gamma + c * 8, gm8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float T = 0.f, U = 0.f, Vv = 0.f;
        for (int r = 0; r < RP; ++r) {
          const int t2 = r * CW + tc;
          T += colacc[(0 * ATT_THREADS + t2) * 8 * NCOL + i * 8 + j];
          U += colacc[(1 * ATT_THREADS + t2) * 8 * NCOL + i * 8 + j];
          Vv += colacc[(2 * ATT_THREADS + t2) * 8 * NCOL + i * 8 + j];
        }
        const int d = c * 8 + j;
        const float w = w8[j], hq = hq8[j], gm = gm8[j];
        const float whk = w * hq * inv_keep;
        a.d_hq[static_cast<long long>(b) * D + d] = T * w * inv_keep;  // sum_k dF * hv
        part[d] = T * hq * inv_keep;                                   // dw partial
        part[D + d] = whk * U;                                         // dgamma partial
        part[2 * D + d] = whk * Vv;                                    // dbeta partial
        const float Gd = whk * gm;                                     // dxhat = ds_k * Gd * flag
        G[i][j] = Gd;
        s1 = fmaf(Gd, Vv, s1);
        s2 = fmaf(Gd, U, s2);
      }
    }
  }
  // every thread of a column needs G: recompute it for the other row groups
  if (active && tr != 0) {
#pragma unroll
    for (int i = 0; i < NCOL; ++i) {
      const int c = tc + i * CW;
      if (c >= CH) break;
      float w8[8], hq8[8], gm8[8];
      ld8f(a.att_w + c * 8, w8);
      ld8f(a.hq + static_cast<long long>(b) * D + c * 8, hq8);
      ld8f(a.gamma + c * 8, gm8);
#pragma unroll
      for (int j = 0; j < 8; ++j) G[i][j] = w8[j] * hq8[j] * inv_keep * gm8[j];
    }
  }
  // q_linear_v backward, part 1 (optional): d x_hat of this sample's question row and its two row sums. The values go to
  // the colacc entries only this thread has read (row group 0 of its own column), which nobody touches before the end.
  float q1 = 0.f, q2 = 0.f;
  float qmean = 0.f, qrstd = 0.f;
  if (a.qv_z) {
    qmean = a.qv_mean[b];
    qrstd = a.qv_rstd[b];
    if (active && tr == 0) {
#pragma unroll
      for (int i = 0; i < NCOL; ++i) {
        const int c = tc + i * CW;
        if (c >= CH) break;
        const long long o = static_cast<long long>(b) * D + c * 8;
        float dq8[8], hq8[8], zq8[8], gq8[8], dg8[8], db8[8];
        {   // written above by this thread: plain (coherent) loads, not the read-only path
          const float4 d0 = *reinterpret_cast<const float4*>(a.d_hq + o);
          const float4 d1 = *reinterpret_cast<const float4*>(a.d_hq + o + 4);
          dq8[0] = d0.x; dq8[1] = d0.y; dq8[2] = d0.z; dq8[3] = d0.w; dq8[4] = d1.x; dq8[5] = d1.y; dq8[6] = d1.z; dq8[7] = d1.w;
        }
        ld8f(a.hq + o, hq8);
        ld8f(a.qv_z + o, zq8);
        ld8f(a.qv_gamma + c * 8, gq8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (zq8[j] - qmean) * qrstd;
          const float dy = hq8[j] > 0.f ? dq8[j] : 0.f;   // hq = relu(pre): hq > 0 <=> pre > 0
          const float dxh = dy * gq8[j];
          dg8[j] = dy * xh;
          db8[j] = dy;
          colacc[(0 * ATT_THREADS + tc) * 8 * NCOL + i * 8 + j] = dxh;
          colacc[(1 * ATT_THREADS + tc) * 8 * NCOL + i * 8 + j] = xh;
          q1 += dxh;
          q2 = fmaf(dxh, xh, q2);
        }
        if (a.qv_dgamma_part) {
          *reinterpret_cast<float4*>(a.qv_dgamma_part + o) = make_float4(dg8[0], dg8[1], dg8[2], dg8[3]);
          *reinterpret_cast<float4*>(a.qv_dgamma_part + o + 4) = make_float4(dg8[4], dg8[5], dg8[6], dg8[7]);
          *reinterpret_cast<float4*>(a.qv_dbeta_part + o) = make_float4(db8[0], db8[1], db8[2], db8[3]);
          *reinterpret_cast<float4*>(a.qv_dbeta_part + o + 4) = make_float4(db8[4], db8[5], db8[6], db8[7]);
        }
      }
    }
    q1 = warp_sum(q1);
    q2 = warp_sum(q2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    red[warp] = s1;
    red[8 + warp] = s2;
    red[40 + warp] = q1;
    red[48 + warp] = q2;
  }
  __syncthreads();
  if (tid == 0) {
    float t1 = 0.f, t2 = 0.f, u1 = 0.f, u2 = 0.f;
    for (int w = 0; w < ATT_WARPS; ++w) {
      t1 += red[w];
      t2 += red[8 + w];
      u1 += red[40 + w];
      u2 += red[48 + w];
    }
    const float inv_n = 1.0f / (static_cast<float>(K) * static_cast<float>(D));
    red[16] = t1 * inv_n;
    red[17] = t2 * inv_n;
    red[56] = u1 / static_cast<float>(D);
    red[57] = u2 / static_cast<float>(D);
    part[4 * D] = red[32];  // d att_b partial
  }
  __syncthreads();
  const float m1 = red[16], m2 = red[17];
  AB_TRACE(6);
  // q_linear_v backward, part 2: dz = rstd (d x_hat - mean(d x_hat) - x_hat mean(d x_hat x_hat)) over the row of D
  if (a.qv_z && active && tr == 0) {
    const float mq1 = red[56], mq2 = red[57];
#pragma unroll
    for (int i = 0; i < NCOL; ++i) {
      const int c = tc + i * CW;
      if (c >= CH) break;
      const long long o = static_cast<long long>(b) * D + c * 8;
      float dzq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dxh = colacc[(0 * ATT_THREADS + tc) * 8 * NCOL + i * 8 + j];
        const float xh = colacc[(1 * ATT_THREADS + tc) * 8 * NCOL + i * 8 + j];
        dzq[j] = qrstd * (dxh - mq1 - xh * mq2);
      }
      if (a.qv_dz_f32) {
        *reinterpret_cast<float4*>(a.qv_dz_f32 + o) = make_float4(dzq[0], dzq[1], dzq[2], dzq[3]);
        *reinterpret_cast<float4*>(a.qv_dz_f32 + o + 4) = make_float4(dzq[4], dzq[5], dzq[6], dzq[7]);
      }
      if (a.qv_dz_hi) store8_planes(a.qv_dz_hi, a.qv_dz_lo, o, dzq);
    }
  }

  // dz = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)) for ALL K rows (padded rows are part
  // of the LayerNorm slab and receive gradient through the statistics)
  float accB[NCOL][8];
#pragma unroll
  for (int i = 0; i < NCOL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) accB[i][j] = 0.f;
  if (active) {
#pragma unroll
    for (int i = 0; i < NCOL; ++i) {
      const int c = tc + i * CW;
      if (c >= CH) break;
      constexpr int UF = SLAB ? 2 : 3;
      for (int k0 = tr; k0 < K; k0 += RP * UF) {
        float xb[UF][8];
#pragma unroll
        for (int u = 0; u < UF; ++u) {
          const int k = k0 + u * RP;
          if (k < K) load8(zb + static_cast<long long>(k) * D + c * 8, xb[u]);
        }
#pragma unroll
        for (int u = 0; u < UF; ++u) {
          const int k = k0 + u * RP;
          if (k >= K) break;
          float dz[8];
          const uint32_t fl = k < nb ? flags[k * CH + c] : 0u;
          const float dsk = ds[k];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = fmaf(xb[u][j], rstd, -mr);
            const float dxh = ((fl >> j) & 1u) ? dsk * G[i][j] : 0.f;
            dz[j] = rstd * (dxh - m1 - xh * m2);
            accB[i][j] += dz[j];
          }
          store8_planes(a.dz_hi, a.dz_lo, (static_cast<long long>(b) * K + k) * D + c * 8, dz);
        }
      }
    }
  }
  __syncthreads();  // colacc reuse
  AB_TRACE(7);
#pragma unroll
  for (int i = 0; i < NCOL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) colacc[tid * 8 * NCOL + i * 8 + j] = accB[i][j];
  __syncthreads();
  if (active && tr == 0) {
#pragma unroll
    for (int i = 0; i < NCOL; ++i) {
      const int c = tc + i * CW;
      if (c >= CH) break;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
        for (int r = 0; r < RP; ++r) s += colacc[(r * CW + tc) * 8 * NCOL + i * 8 + j];
        part[3 * D + c * 8 + j] = s;  // d(bias of the projection) partial
      }
    }
  }
}

}  // namespace

size_t attn_bwd_partial_floats(int batch, int D) {
  // [batch, 4D+8] per-CTA partials + [4D+8] reduced + [32, 4D+8] scratch of the two-pass column sum
  return static_cast<size_t>(batch + 1 + 32) * (4 * static_cast<size_t>(D) + 8);
}

VqaStatus attn_fwd_launch(const VqaAttnFwd& a, int K, int D, int Dv, int precision, float keep,
                          cudaStream_t s) {
  if (a.batch == 0) return VQA_OK;
  if (!a.z || !a.gamma || !a.beta || !a.hq || !a.att_w || !a.att_b || !a.nbox || !a.v_hi ||
      !a.ln_mean || !a.ln_rstd)
    return set_error(VQA_ERR_BAD_ARG, "vqa_attn_fwd: null argument");
  if ((D & 7) || (Dv & 7)) return set_error(VQA_ERR_BAD_SHAPE, "vqa_attn_fwd: D, Dv must be multiples of 8");
  {
    // the persistent pipelined kernel (attn_pipe.cu) whenever its buffers fit one SM
    size_t psmem = 0;
    int rv = 0;
    const bool mask = a.keep_bits != nullptr && keep_threshold(keep) < 65536u;
    // (with a bit plane that cannot travel by bulk copy the kernel draws the bits itself: same bits)
    const bool mask_ok = mask && (static_cast<size_t>(K) * (D >> 3)) % 16 == 0 && (reinterpret_cast<uintptr_t>(a.keep_bits) & 15) == 0;
    if (attn_fwd_pipe_supported(K, D, Dv, precision, a.v_lo != nullptr, mask_ok, &psmem, &rv)) {
      if (!mask_ok) {
        VqaAttnFwd a2 = a;
        a2.keep_bits = nullptr;
        static int num_sms2 = 0;
        if (num_sms2 == 0) {
          int dev = 0;
          VQA_CUDA_CHECK(cudaGetDevice(&dev));
          VQA_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms2, cudaDevAttrMultiProcessorCount, dev));
        }
        return attn_fwd_pipe_launch(a2, K, D, Dv, keep, psmem, rv, num_sms2, s);
      }
      static int num_sms = 0;
      if (num_sms == 0) {
        int dev = 0;
        VQA_CUDA_CHECK(cudaGetDevice(&dev));
        VQA_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
      }
      return attn_fwd_pipe_launch(a, K, D, Dv, keep, psmem, rv, num_sms, s);
    }
  }
  FwdArgs f;
  f.z = a.z; f.gamma = a.gamma; f.beta = a.beta; f.hq = a.hq; f.att_w = a.att_w; f.att_b = a.att_b;
  f.nbox = a.nbox; f.v_hi = static_cast<const bf16*>(a.v_hi); f.v_lo = static_cast<const bf16*>(a.v_lo);
  f.seed = a.seed; f.step = a.step; f.att = a.att; f.pooled = a.pooled;
  f.pooled_hi = static_cast<bf16*>(a.pooled_hi); f.pooled_lo = static_cast<bf16*>(a.pooled_lo);
  f.ln_mean = a.ln_mean; f.ln_rstd = a.ln_rstd;
  // one-CTA-per-sample kernel: it would fetch the plane byte by byte from global memory inside its score loop, which
  // measured slower than drawing the bits (the ALUs are idle there): the plane is used only where it arrives by bulk copy
  f.keep_bits = getenv("VQA_ATTN_PLANE_GLOBAL") ? a.keep_bits : nullptr;
  const size_t head = (3 * static_cast<size_t>(D) + K + 3 * ATT_WARPS + 2) * sizeof(float);
  const uint32_t thr = keep_threshold(keep);
  const size_t slab = static_cast<size_t>(K) * D * (precision == VQA_PREC_FP32 ? 4 : 2);
  const size_t smem_slab = ((head + 15) & ~static_cast<size_t>(15)) + 16 + slab;
  // two CTAs per SM must still fit, and the slab copy needs 16-byte granularity
  const bool use_slab = smem_slab <= 110 * 1024 && (slab & 15) == 0 && !a.v_lo;
  auto launch = [&](auto kern, size_t smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    launch_pdl(kern, dim3(a.batch), dim3(ATT_THREADS), smem, s, f, K, D, Dv, keep, thr);
  };
  if (precision == VQA_PREC_FP32) {
    if (use_slab) launch(attn_fwd_kernel<float, true>, smem_slab);
    else launch(attn_fwd_kernel<float, false>, head);
  } else {
    if (use_slab) launch(attn_fwd_kernel<bf16, true>, smem_slab);
    else launch(attn_fwd_kernel<bf16, false>, head);
  }
  VQA_LAUNCH_CHECK("attn_fwd");
  return VQA_OK;
}

template <typename ZT, int NCOL>
static cudaError_t launch_bwd(const BwdArgs& g, int batch, int K, int D, int Dv, float keep,
                              uint32_t thr, cudaStream_t s) {
  const size_t head = (static_cast<size_t>(Dv) + K + 64 + 3 * ATT_THREADS * 8 * NCOL) * sizeof(float) +
                      static_cast<size_t>(K) * (D >> 3);
  const size_t slab = static_cast<size_t>(K) * D * sizeof(ZT);
  const size_t smem_slab = ((head + 15) & ~static_cast<size_t>(15)) + 64 + slab;
  const bool use_slab = smem_slab <= 113 * 1024 && (slab & 15) == 0;
  if (head > 220 * 1024) return cudaErrorInvalidValue;
  BwdArgs g2 = g;
  // feature rows of the <V, dP> pass through a ring laid over the slab buffer: single bf16 plane, whole rows per slot
  g2.vring_rows = 0;
  static const bool vring_off = getenv("VQA_ATTN_BWD_VRING") != nullptr && atoi(getenv("VQA_ATTN_BWD_VRING")) == 0;
  if (use_slab && !g.v_lo && !vring_off) {
    const size_t vrow = static_cast<size_t>(Dv) * 2;
    const int rows = static_cast<int>((slab / AB_VSLOTS) / vrow);
    if (rows >= 1 && (reinterpret_cast<uintptr_t>(g.v_hi) & 15) == 0) g2.vring_rows = rows;
  }
  // the bit plane travels by bulk copy into `flags` (slab mode, 16-byte granularity); otherwise the kernel redraws the bits
  const size_t flags_off = (static_cast<size_t>(Dv) + K + 64 + 3 * ATT_THREADS * 8 * NCOL) * sizeof(float);
  if (!use_slab || (static_cast<size_t>(K) * (D >> 3)) % 16 != 0 || (flags_off & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(g.keep_bits) & 15) != 0 || thr >= 65536u)
    g2.keep_bits = nullptr;
  cudaError_t e;
  if (use_slab) {
    e = cudaFuncSetAttribute(attn_bwd_kernel<ZT, NCOL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    launch_pdl(attn_bwd_kernel<ZT, NCOL, true>, dim3(batch), dim3(ATT_THREADS), smem_slab, s, g2, K, D, Dv, keep, thr);
  } else {
    e = cudaFuncSetAttribute(attn_bwd_kernel<ZT, NCOL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    launch_pdl(attn_bwd_kernel<ZT, NCOL, false>, dim3(batch), dim3(ATT_THREADS), head, s, g2, K, D, Dv, keep, thr);
  }
  return cudaGetLastError();
}

VqaStatus attn_bwd_launch(const VqaAttnBwd& a, int K, int D, int Dv, int precision, float keep,
                          float* partials, cudaStream_t s, cudaStream_t reduce_stream, cudaEvent_t kernel_done,
                          const AttnQvBwd* qv) {
  if (a.batch == 0) return VQA_OK;
  if (!a.z || !a.gamma || !a.beta || !a.hq || !a.att_w || !a.nbox || !a.v_hi || !a.att || !a.ln_mean ||
      !a.ln_rstd || !a.d_pooled || !a.dz_hi || !a.d_hq || !partials)
    return set_error(VQA_ERR_BAD_ARG, "vqa_attn_bwd: null argument");
  if ((D & 7) || (Dv & 7) || D > 4096)
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_attn_bwd: D, Dv multiples of 8, D <= 4096");
  BwdArgs g;
  g.z = a.z; g.gamma = a.gamma; g.beta = a.beta; g.hq = a.hq; g.att_w = a.att_w; g.nbox = a.nbox;
  g.v_hi = static_cast<const bf16*>(a.v_hi); g.v_lo = static_cast<const bf16*>(a.v_lo);
  g.seed = a.seed; g.step = a.step; g.att = a.att; g.ln_mean = a.ln_mean; g.ln_rstd = a.ln_rstd;
  g.d_pooled = a.d_pooled; g.dz_hi = static_cast<bf16*>(a.dz_hi); g.dz_lo = static_cast<bf16*>(a.dz_lo);
  g.d_hq = a.d_hq; g.part = partials; g.keep_bits = a.keep_bits;
  g.qv_z = nullptr; g.qv_gamma = g.qv_mean = g.qv_rstd = nullptr;
  g.qv_dz_hi = g.qv_dz_lo = nullptr; g.qv_dz_f32 = g.qv_dgamma_part = g.qv_dbeta_part = nullptr;
  if (qv && qv->z && qv->gamma && qv->mean && qv->rstd && (qv->dz_hi || qv->dz_f32)) {
    g.qv_z = qv->z; g.qv_gamma = qv->gamma; g.qv_mean = qv->mean; g.qv_rstd = qv->rstd;
    g.qv_dz_hi = qv->dz_hi; g.qv_dz_lo = qv->dz_lo; g.qv_dz_f32 = qv->dz_f32;
    if (qv->dgamma_part && qv->dbeta_part) { g.qv_dgamma_part = qv->dgamma_part; g.qv_dbeta_part = qv->dbeta_part; }
  }
  g.trace = g_gru_trace ? g_gru_trace + 98304 : nullptr;
  const uint32_t thr = keep_threshold(keep);
  cudaError_t e;
  const bool two = D > 2048;
  if (precision == VQA_PREC_FP32)
    e = two ? launch_bwd<float, 2>(g, a.batch, K, D, Dv, keep, thr, s)
            : launch_bwd<float, 1>(g, a.batch, K, D, Dv, keep, thr, s);
  else
    e = two ? launch_bwd<bf16, 2>(g, a.batch, K, D, Dv, keep, thr, s)
            : launch_bwd<bf16, 1>(g, a.batch, K, D, Dv, keep, thr, s);
  if (e != cudaSuccess) return set_cuda_error(e, "attn_bwd launch");
  count_launch();
  // reduce the per-sample partials: [dw | dgamma | dbeta | dbias | db] -- nothing downstream but the optimizer
  // reads them, so the caller may move this off the critical path
  if (reduce_stream && kernel_done) {
    VQA_CUDA_CHECK(cudaEventRecord(kernel_done, s));
    VQA_CUDA_CHECK(cudaStreamWaitEvent(reduce_stream, kernel_done, 0));
    s = reduce_stream;
  }
  const int width = 4 * D + 8;
  float* reduced = partials + static_cast<size_t>(a.batch) * width;
  VQA_TRY(colsum_launch(partials, a.batch, width, width, reduced, reduced + width, s));
  struct { float* dst; int off; int n; } outs[5] = {
      {a.d_att_w, 0, D}, {a.d_gamma, D, D}, {a.d_beta, 2 * D, D}, {a.d_bias, 3 * D, D}, {a.d_att_b, 4 * D, 1}};
  for (auto& o : outs)
    if (o.dst)
      VQA_CUDA_CHECK(cudaMemcpyAsync(o.dst, reduced + o.off, sizeof(float) * o.n, cudaMemcpyDeviceToDevice, s));
  return VQA_OK;
}

}  // namespace vqa
