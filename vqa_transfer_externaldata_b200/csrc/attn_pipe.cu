// Persistent, software-pipelined forward of the attention block (same math as attn_fwd_kernel in attn.cu:
// vlmap/modules.py:67-97 hadamard_attention + :23-39 attention_pooling + the joint (K, D) LayerNorm + ReLU of
// v_linear_v, vlmap/modules.py:646-649).
//
// ONE CTA per SM walks its samples; a producer thread keeps the memory system busy ahead of 16 consumer warps.
// Round 2 (profiles/r02_attn_phase_trace.md: in-kernel globaltimer stamps of the round-1 kernel: statistics 3.1 us,
// scores 6.4 us, soft-max 0.8 us, pooling 3.8 us per sample) rebuilt three things:
//   * the whole feature block of a sample ([K, Dv] bf16 = 144 KB at cfg1) is RESIDENT before its pooling pass starts:
//     four 9-row chunks (36 KB bulk copies) arrive while the statistics / score passes run, instead of a two-slot
//     ring whose third to fifth chunk each cost the pooling pass one TMA round trip (the slab is single-buffered to
//     pay for it: the next sample's slab is fetched during soft-max + pooling, when this sample's is dead);
//   * the per-column coefficients live in shared memory as float4 arrays indexed by 16-byte chunk (consecutive lanes
//     -> consecutive 16 bytes: conflict-free) and there are TWO of them, not three: relu(x A + B) C = sign(C)
//     relu(x A|C| + B|C|), the signs travel as one byte per chunk. The score pass was bound by shared-memory
//     bandwidth: six 2-way-conflicted 16-byte coefficient loads per chunk against one of data;
//   * the statistics are two streaming passes over the slab in shared memory (sum, then sum of squared deviations:
//     independent accumulators, no serial Chan merges, every thread finishes the 16-value combine itself).
// Every HBM byte is still read exactly once.
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
constexpr int AP_VSLOTS = 4;
constexpr int AP_DEFAULT_VAR = 10;   // measured best of the sixteen on one box (profiles/r02_attn_phase_trace.md)

__device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 1, %0;" ::"n"(AP_CONSUMERS) : "memory"); }
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
// 16-byte shared-memory load (the compiler splits a uint4 load through a byte pointer with a run-time pitch into four
// 4-way bank-conflicted LDS.32: measured 2.8 us instead of 0.7 us per sample in the pooling pass)
__device__ __forceinline__ uint4 lds128(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ptx::smem_u32(p)));
  return v;
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

struct PipeFwdArgs {
  const bf16* z;
  const float* gamma; const float* beta; const float* hq; const float* att_w; const float* att_b;
  const int* nbox; const bf16* v;
  unsigned long long seed, step;
  float* att; float* pooled; bf16* pooled_bf; float* ln_mean; float* ln_rstd;
  int batch, K, D, Dv, RV;   // RV = feature rows per chunk (AP_VSLOTS chunks cover K rows whenever they fit)
  float keep;
  uint32_t thr;
  const unsigned char* keep_bits;   // [batch*K*D/8] keep bits of this step (NULL: drawn here with Philox)
  unsigned long long* trace;        // optional [grid][8 samples][8] globaltimer stamps (scripts/gpu_attn_trace.py)
};

__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define AP_TRACE(slot)                                                                              \
  do {                                                                                              \
    if (a.trace && tid == 0 && i < 8) a.trace[(static_cast<size_t>(blockIdx.x) * 8 + i) * 8 + (slot)] = gtimer_ns(); \
  } while (0)

// shared-memory plan (bytes), shared by the kernel and the host-side fit test
struct PipePlan {
  uint32_t zbytes, mbytes, vchunk, off_m, off_v, off_ca, off_cb, off_sgn, off_sc, off_red, off_bar, total;
};
__host__ __device__ inline PipePlan pipe_plan(int K, int D, int Dv, int RV, bool mask) {
  PipePlan p;
  p.zbytes = static_cast<uint32_t>(K) * D * 2;
  p.mbytes = mask ? static_cast<uint32_t>(K) * (D >> 3) : 0u;   // keep bits of the slab: one byte per 16-byte chunk
  p.vchunk = static_cast<uint32_t>(RV) * Dv * 2;
  p.off_m = (p.zbytes + 15u) & ~15u;
  p.off_v = (p.off_m + p.mbytes + 127u) & ~127u;
  p.off_ca = p.off_v + AP_VSLOTS * p.vchunk;
  p.off_cb = p.off_ca + static_cast<uint32_t>(D) * 4;
  p.off_sgn = p.off_cb + static_cast<uint32_t>(D) * 4;
  p.off_sc = p.off_sgn + ((static_cast<uint32_t>(D >> 3) + 15u) & ~15u);
  p.off_red = p.off_sc + ((static_cast<uint32_t>(K) + 3u) & ~3u) * 4 * (1u + static_cast<uint32_t>(D >> 8));   // sc + [D / 256] partials
  p.off_bar = p.off_red + 64 * 4;
  p.total = p.off_bar + (2 + 2 * AP_VSLOTS) * 8;
  return p;
}

// VAR: code-shape switches measured against each other on one box (scripts/gpu_attn_trace.py, VQA_ATTN_VAR):
//   bit 0  statistics passes issue three slab loads before any arithmetic (otherwise one per iteration)
//   bit 1  score pass: warp = (column block, row group) with its coefficients in registers (otherwise a warp per row)
//   bit 2  pooling pass: four feature rows in flight per thread (otherwise one)
//   bit 3  the Philox path is compiled out (the keep bits always arrive as the bit plane)
template <int VAR>
__global__ void __launch_bounds__(AP_THREADS, 1) attn_fwd_pipe_kernel(PipeFwdArgs a) {
  constexpr bool V_STATS = (VAR & 1) != 0, V_COLS = (VAR & 2) != 0, V_POOL = (VAR & 4) != 0, V_NOPHILOX = (VAR & 8) != 0;
  extern __shared__ __align__(128) uint8_t smem[];
  const int K = a.K, D = a.D, Dv = a.Dv, RV = a.RV;
  const PipePlan pl = pipe_plan(K, D, Dv, RV, a.keep_bits != nullptr);
  const uint32_t zbytes = pl.zbytes;
  const uint32_t vrow = static_cast<uint32_t>(Dv) * 2;
  const int CH = D >> 3;                 // 16-byte chunks per slab row
  uint8_t* zbuf = smem;
  const unsigned char* mbuf = smem + pl.off_m;
  uint8_t* vbuf = smem + pl.off_v;
  float4* cA4 = reinterpret_cast<float4*>(smem + pl.off_ca);   // [2][CH]: elements 0..3 / 4..7 of chunk c
  float4* cB4 = reinterpret_cast<float4*>(smem + pl.off_cb);
  float* cAf = reinterpret_cast<float*>(cA4);
  float* cBf = reinterpret_cast<float*>(cB4);
  unsigned char* sgn = smem + pl.off_sgn;                       // [CH] sign bits of hq w / keep, bit q = element 8 c + q
  float* sc = reinterpret_cast<float*>(smem + pl.off_sc);
  float* red = reinterpret_cast<float*>(smem + pl.off_red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.off_bar);
  uint64_t* zfull = bars;
  uint64_t* zempty = bars + 1;
  uint64_t* vfull = bars + 2;
  uint64_t* vempty = bars + 2 + AP_VSLOTS;
  float* poolx = cAf;   // partial pooled sums of the other row groups (the coefficients are dead by then)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    ptx::mbar_init(zfull, 1);
    ptx::mbar_init(zempty, AP_CONSUMER_WARPS);
    for (int i = 0; i < AP_VSLOTS; ++i) {
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
    // Completion n (1-based) of an mbarrier is awaited with parity (n - 1) & 1.
    if (lane == 0) {
      unsigned int vc = 0;
      for (int i = 0; i < n_my; ++i) {
        const int b = first + i * stride;
        // the slab buffer is free once every consumer warp has finished the scores of sample i - 1
        if (i > 0) ptx::mbar_wait(zempty, (i - 1) & 1);
        ptx::mbar_arrive_expect_tx(zfull, zbytes + pl.mbytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.z) + static_cast<size_t>(b) * zbytes;
        const uint32_t dst = ptx::smem_u32(zbuf);
        for (uint32_t off = 0; off < zbytes; off += 65536u)
          bulk_load(dst + off, src + off, zbytes - off < 65536u ? zbytes - off : 65536u, zfull);
        if (pl.mbytes)   // the slab's dropout keep bits ride with it (vqa_keep_bits plane)
          bulk_load(ptx::smem_u32(mbuf), a.keep_bits + static_cast<size_t>(b) * pl.mbytes, pl.mbytes, zfull);
        // the features of sample i: chunk slots free up as the pooling pass of sample i - 1 drains them
        int nb = a.nbox[b];
        nb = nb < 0 ? 0 : (nb > K ? K : nb);
        const uint8_t* vsrc = reinterpret_cast<const uint8_t*>(a.v) + static_cast<size_t>(b) * K * vrow;
        for (int k0 = 0; k0 < nb; k0 += RV, ++vc) {
          const int slot = vc % AP_VSLOTS;
          const unsigned int use = vc / AP_VSLOTS;
          const int rows = nb - k0 < RV ? nb - k0 : RV;
          if (use > 0) ptx::mbar_wait(&vempty[slot], (use - 1) & 1);
          ptx::mbar_arrive_expect_tx(&vfull[slot], rows * vrow);
          bulk_load(ptx::smem_u32(vbuf + static_cast<size_t>(slot) * pl.vchunk), vsrc + static_cast<size_t>(k0) * vrow,
                    rows * vrow, &vfull[slot]);
        }
      }
    }
    return;
  }

  // ===================== consumers: 16 warps =====================
  const int nchunks = K * CH;
  const int QN = CH >> 5;                // 32-chunk column blocks of a slab row (attn_fwd_pipe_supported: divides 16)
  const int RGN = AP_CONSUMER_WARPS / QN;   // row groups of the score pass
  const int KP = (K + 3) & ~3;
  float* scp = sc + KP;                  // [QN][KP] partial scores per column block
  const int VCH = Dv >> 3;               // column chunks of the features
  const int tpc = AP_CONSUMERS / VCH;    // threads sharing one column chunk (each takes every tpc-th row)
  const int vcol = tid % VCH, vsub = tid / VCH;
  const float inv_keep = 1.0f / a.keep;
  const float inv_n = 1.0f / (static_cast<float>(K) * static_cast<float>(D));
  const float bias = a.att_b[0];
  unsigned int vc = 0;
  // sample-independent coefficients stay in registers for the whole kernel (D <= 2 * AP_CONSUMERS: two columns per
  // thread); the per-sample hq values are fetched at the top of each sample, ahead of the statistics passes
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
    const bf16* zs = reinterpret_cast<const bf16*>(zbuf);
    int nb = a.nbox[b];
    nb = nb < 0 ? 0 : (nb > K ? K : nb);
    float hq_r[2] = {0.f, 0.f};
    AP_TRACE(0);
    if (coef_regs) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int d = tid + q * AP_CONSUMERS;
        if (d < D) hq_r[q] = a.hq[static_cast<long long>(b) * D + d];
      }
    }
    ptx::mbar_wait(zfull, i & 1);
    AP_TRACE(1);

    // ---- statistics over the K*D slab: mean, then the sum of squared deviations (tf.nn.moments) ----
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    // (loads go out in batches of three before any arithmetic: the volatile shared-memory loads are not reordered
    // by the compiler, and one load per iteration leaves the pass latency-bound)
    constexpr int SU = V_STATS ? 3 : 1;
    for (int c0 = tid; c0 < nchunks; c0 += SU * AP_CONSUMERS) {
      uint4 u[SU];
#pragma unroll
      for (int j = 0; j < SU; ++j) {
        const int c = c0 + j * AP_CONSUMERS;
        u[j] = c < nchunks ? lds128(zs + static_cast<size_t>(c) * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int j = 0; j < SU; ++j) {
        float x[8];
        unpack8(u[j], x);
        s0 += x[0] + x[4];
        s1 += x[1] + x[5];
        s2 += x[2] + x[6];
        s3 += x[3] + x[7];
      }
    }
    float tot = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) red[warp] = tot;
    cons_bar();   // (also: everybody is done with the previous sample's poolx = the coefficient arrays)
    tot = 0.f;
#pragma unroll
    for (int w = 0; w < AP_CONSUMER_WARPS; ++w) tot += red[w];
    const float mu = tot * inv_n;
    s0 = s1 = s2 = s3 = 0.f;
    for (int c0 = tid; c0 < nchunks; c0 += SU * AP_CONSUMERS) {
      uint4 u[SU];
      bool ok[SU];
#pragma unroll
      for (int j = 0; j < SU; ++j) {
        const int c = c0 + j * AP_CONSUMERS;
        ok[j] = c < nchunks;
        u[j] = ok[j] ? lds128(zs + static_cast<size_t>(c) * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int j = 0; j < SU; ++j) {
        if (!ok[j]) continue;
        float x[8];
        unpack8(u[j], x);
#pragma unroll
        for (int q = 0; q < 8; q += 4) {
          const float d0 = x[q] - mu, d1 = x[q + 1] - mu, d2 = x[q + 2] - mu, d3 = x[q + 3] - mu;
          s0 = fmaf(d0, d0, s0);
          s1 = fmaf(d1, d1, s1);
          s2 = fmaf(d2, d2, s2);
          s3 = fmaf(d3, d3, s3);
        }
      }
    }
    tot = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) red[16 + warp] = tot;
    cons_bar();
    tot = 0.f;
#pragma unroll
    for (int w = 0; w < AP_CONSUMER_WARPS; ++w) tot += red[16 + w];
    const float rstd = 1.0f / sqrtf(tot * inv_n + 1e-12f);   // biased variance, eps of layers.layer_norm

    // ---- per-column coefficients: relu(x g rstd + (beta - mu g rstd)) * C  =  sign(C) relu(x A' + B') ----
    for (int base = 0; base < D; base += AP_CONSUMERS) {
      const int d = base + tid;
      float A = 0.f, Bc = 0.f, C = 0.f;
      if (d < D) {
        float g, be, cc;
        if (coef_regs) {   // (no run-time index into the register arrays: that would move them to local memory)
          const bool hi = base != 0;
          g = hi ? g_r[1] : g_r[0]; be = hi ? be_r[1] : be_r[0]; cc = hi ? hq_r[1] * w_r[1] : hq_r[0] * w_r[0];
        } else {
          g = a.gamma[d]; be = a.beta[d];
          cc = a.hq[static_cast<long long>(b) * D + d] * a.att_w[d] * inv_keep;
        }
        g *= rstd;
        const float ac = fabsf(cc);
        A = g * ac;
        Bc = (be - mu * g) * ac;
        C = cc;
        const int pidx = ((d & 7) >> 2) * (CH * 4) + (d >> 3) * 4 + (d & 3);
        cAf[pidx] = A;
        cBf[pidx] = Bc;
      }
      const unsigned int neg = __ballot_sync(0xffffffffu, C < 0.f);
      if ((lane & 7) == 0 && d < D) sgn[d >> 3] = static_cast<unsigned char>((neg >> lane) & 0xFFu);
    }
    cons_bar();
    AP_TRACE(2);

    if (!V_COLS) {
      // ---- scores: one warp per box row ----
      for (int k = warp; k < nb; k += AP_CONSUMER_WARPS) {
        const bf16* zr = zs + static_cast<size_t>(k) * D;
        const unsigned long long g0 = (static_cast<unsigned long long>(b) * K + k) * CH;
        const unsigned char* mrow = mbuf + static_cast<size_t>(k) * CH;   // this row's keep bits (arrived with the slab)
        float accp = 0.f, accn = 0.f;
        for (int c = lane; c < CH; c += 32) {
          float x[8];
          unpack8(lds128(zr + c * 8), x);
          uint32_t bits = 0xFFu;
          if (V_NOPHILOX || a.keep_bits) bits = mrow[c];
          else if (a.thr < 65536u) bits = philox_keep_bits(philox4x32_10(g0 + c, RNG_STREAM_ATT, a.seed, a.step), a.thr);
          const uint32_t sg = sgn[c];
          const uint32_t pos = bits & ~sg, ngt = bits & sg;
          const float4 a0 = cA4[c], a1 = cA4[CH + c];
          const float4 b0 = cB4[c], b1 = cB4[CH + c];
          const float ca[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float cb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float y = fmaxf(fmaf(x[q], ca[q], cb[q]), 0.f);
            accp += ((pos >> q) & 1u) ? y : 0.f;
            accn += ((ngt >> q) & 1u) ? y : 0.f;
          }
        }
        const float acc = warp_sum(accp - accn);
        if (lane == 0) {
          scp[k] = acc;
          for (int q = 1; q < QN; ++q) scp[q * KP + k] = 0.f;
        }
      }
    } else
    // ---- scores: warp = (32-chunk column block qd, row group): the coefficients of its chunk stay in registers for
    // all its rows; every (row, column block) item is one 16-byte load + one keep byte per lane; 36 rows x 4 blocks over
    // 16 warps = 9 items each (a warp per row left 4 warps with 3 rows and 12 with 2). Partial scores per column block.
    {
      const int qd = warp % QN, rg = warp / QN;
      const int c = qd * 32 + lane;
      const float4 a0 = cA4[c], a1 = cA4[CH + c];
      const float4 b0 = cB4[c], b1 = cB4[CH + c];
      const float ca[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float cb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const uint32_t sg = sgn[c];
      constexpr int RU = 3;
      for (int k0 = rg; k0 < nb; k0 += RU * RGN) {
        uint4 u[RU];
        uint32_t bits[RU];
#pragma unroll
        for (int j = 0; j < RU; ++j) {
          const int k = k0 + j * RGN;
          bits[j] = 0xFFu;
          if (k < nb) {
            u[j] = lds128(zs + static_cast<size_t>(k) * D + c * 8);
            const unsigned long long g0 = (static_cast<unsigned long long>(b) * K + k) * CH;
            if (V_NOPHILOX || a.keep_bits) bits[j] = mbuf[static_cast<size_t>(k) * CH + c];
            else if (a.thr < 65536u) bits[j] = philox_keep_bits(philox4x32_10(g0 + c, RNG_STREAM_ATT, a.seed, a.step), a.thr);
          }
        }
#pragma unroll
        for (int j = 0; j < RU; ++j) {
          const int k = k0 + j * RGN;
          if (k >= nb) break;
          float x[8];
          unpack8(u[j], x);
          const uint32_t pos = bits[j] & ~sg, ngt = bits[j] & sg;
          float accp = 0.f, accn = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float y = fmaxf(fmaf(x[q], ca[q], cb[q]), 0.f);
            accp += ((pos >> q) & 1u) ? y : 0.f;
            accn += ((ngt >> q) & 1u) ? y : 0.f;
          }
          const float acc = warp_sum(accp - accn);
          if (lane == 0) scp[qd * KP + k] = acc;
        }
      }
    }
    // this warp is done with the slab: the producer may fetch the next sample's
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(zempty);
    cons_bar();
    AP_TRACE(3);

    // ---- masked softmax over boxes: exact zeros beyond nbox ----
    if (warp == 0) {
      for (int k = lane; k < nb; k += 32) {
        float v = bias;
        for (int q = 0; q < QN; ++q) v += scp[q * KP + k];   // fixed order
        sc[k] = v;
      }
      __syncwarp();
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
    AP_TRACE(4);

    // ---- attended pooling of the raw features (resident by now) ----
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int k0 = 0; k0 < nb; k0 += RV, ++vc) {
      const int slot = vc % AP_VSLOTS;
      const int rows = nb - k0 < RV ? nb - k0 : RV;
      ptx::mbar_wait(&vfull[slot], (vc / AP_VSLOTS) & 1);
      const uint8_t* vs = vbuf + static_cast<size_t>(slot) * pl.vchunk;
      if (vsub < tpc) {
        constexpr int PU = V_POOL ? 4 : 1;
        for (int r0 = vsub; r0 < rows; r0 += PU * tpc) {
          uint4 u[PU];
          float ak[PU];
#pragma unroll
          for (int j = 0; j < PU; ++j) {
            const int r = r0 + j * tpc;
            ak[j] = 0.f;
            u[j] = make_uint4(0u, 0u, 0u, 0u);
            if (r < rows) {
              u[j] = lds128(vs + static_cast<size_t>(r) * vrow + vcol * 16);
              ak[j] = sc[k0 + r];
            }
          }
#pragma unroll
          for (int j = 0; j < PU; ++j) {
            float v[8];
            unpack8(u[j], v);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = fmaf(ak[j], v[q], acc[q]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&vempty[slot]);
    }
    AP_TRACE(5);
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
    // (the next sample's first cons_bar orders these poolx reads before the coefficient arrays are rewritten)
    AP_TRACE(6);
  }
}

}  // namespace

// the pipelined kernel covers the bf16 mode with one feature plane when both slab buffers, the feature ring and the
// per-column constants fit one SM; everything else stays on attn_fwd_kernel
bool attn_fwd_pipe_supported(int K, int D, int Dv, int precision, bool has_v_lo, bool mask, size_t* smem_out, int* rv_out) {
  if (precision != VQA_PREC_BF16 || has_v_lo) return false;
  if (getenv("VQA_ATTN_PIPE") && atoi(getenv("VQA_ATTN_PIPE")) == 0) return false;
  const int VCH = Dv >> 3;
  if ((D & 7) || (Dv & 7) || VCH > AP_CONSUMERS || AP_CONSUMERS % VCH != 0) return false;
  const int tpc = AP_CONSUMERS / VCH;
  if (static_cast<size_t>(tpc - 1) * VCH * 8 > 2 * static_cast<size_t>(D)) return false;   // poolx aliases the coefficients
  if ((static_cast<size_t>(K) * D * 2) % 16 != 0) return false;
  const int QN = D >> 8;   // 32-chunk column blocks per slab row: the score pass gives each of its 16 warps one of them
  if ((D & 255) || QN < 1 || QN > AP_CONSUMER_WARPS || AP_CONSUMER_WARPS % QN != 0) return false;
  // rows per feature chunk: the AP_VSLOTS chunks hold the whole [K, Dv] block when that fits next to the slab,
  // fewer rows (a ring that is refilled during the pooling pass) otherwise
  if (mask && (static_cast<size_t>(K) * (D >> 3)) % 16 != 0) return false;   // the bit plane arrives by bulk copy
  int rv = (K + AP_VSLOTS - 1) / AP_VSLOTS;
  while (rv >= 1 && pipe_plan(K, D, Dv, rv, mask).total + 128 > 227u * 1024u) --rv;
  if (rv < 1) return false;
  // a ring much shorter than the block would bring back the round trips this kernel exists to avoid
  if (rv * AP_VSLOTS * 2 < K) return false;
  *smem_out = pipe_plan(K, D, Dv, rv, mask).total + 128;
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
  f.keep_bits = f.thr < 65536u ? a.keep_bits : nullptr;
  f.trace = g_gru_trace ? g_gru_trace + 65536 : nullptr;   // (shared debugging buffer: vqa_internal_set_gru_trace)
  static const int var_env = getenv("VQA_ATTN_VAR") ? atoi(getenv("VQA_ATTN_VAR")) : AP_DEFAULT_VAR;
  // bits 0 (batched statistics loads) and 2 (four pooling rows in flight) measured slower or equal in every combination
  // (38.7 us for VAR 10 against 41.0 / 41.0 / 42.0 / 42.0 for VAR 0 / 1 / 4 / 15): only bits 1 and 3 are instantiated
  int var = var_env & 10;
  if (!f.keep_bits) var &= 7;   // the Philox path is needed
  using Kern = void (*)(PipeFwdArgs);
  static const Kern kerns[16] = {attn_fwd_pipe_kernel<0>, nullptr, attn_fwd_pipe_kernel<2>, nullptr, nullptr, nullptr, nullptr, nullptr,
                                 attn_fwd_pipe_kernel<8>, nullptr, attn_fwd_pipe_kernel<10>, nullptr, nullptr, nullptr, nullptr, nullptr};
  static size_t smem_set[16] = {};
  if (smem_set[var] < smem) {
    VQA_CUDA_CHECK(cudaFuncSetAttribute(kerns[var], cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    smem_set[var] = smem;
  }
  const int grid = a.batch < num_sms ? a.batch : num_sms;
  launch_pdl(kerns[var], dim3(grid), dim3(AP_THREADS), smem, s, f);
  VQA_LAUNCH_CHECK("attn_fwd (pipelined)");
  return VQA_OK;
}

}  // namespace vqa
