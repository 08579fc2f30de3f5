// modules.fc_layer on a rank-2 input as ONE kernel: the matrix product on tcgen05 and the layer's whole tail --
// bias, LayerNorm over the row, ReLU / tanh, the q (.) v Hadamard partner, dropout -- in its epilogue
// (vlmap/modules.py:616-650; call sites vqa/model_vlmap_answer.py:142-181: q_linear_v, q_linear_l, pooled_linear_l,
// joint_fc), and the mirror image for the backward pass: the data-gradient product with dropout / Hadamard / activation /
// LayerNorm backward in its epilogue.
//
// LayerNorm needs the statistics of a whole output row (N = 1024 / 2048 fp32: more than the 512 TMEM columns of one SM),
// and a [512, N] product only engages the machine if the N axis is spread over many SMs (every CTA streams its slice of
// the weights once; an SM ingests ~140 GB/s). So the CTAs that share a row tile form ONE thread-block cluster along N
// (N / BN = 16 CTAs: a non-portable cluster size) and exchange per-row partial statistics through distributed shared
// memory: every CTA reduces its BN columns per row (forward: mean and centred sum of squares, combined with the
// parallel-variance formula -- the two-pass numerics of tf.nn.moments; backward: the two row sums of the LayerNorm
// gradient), pushes 8 bytes per row into the statistics table of every CTA of the cluster (st.shared::cluster), one
// barrier.cluster later each CTA holds all partials and finishes its own columns. Nothing but the layer's outputs
// touches global memory: the fp32 pre-activation round trip and the row kernels of rows.cu disappear from the path.
//
// Tile plans (launch_plan): bn columns per CTA (cluster = N / bn CTAs), 128- or 64-row tiles (the MMA is always M = 128:
// with 64-row tiles the upper half of the A tile is whatever follows it in shared memory and the upper 64 accumulator
// lanes are never read -- rows of a product are independent), A by multicast or privately. Measured on B200
// (scripts/gpu_linear_ln_bench.py, profiles/r02_linear_ln.md; M 512, back-to-back launches): N 1024 / K 1024: 10.5 us
// against 14.9 for GEMM + row kernel (K 2048: 13.2 / 17.7; N 2048 / K 1024: 15.6 / 18.7); backward N 1024 / K 2048: 18.5
// / 18.7, N 2048 / K 3000: 32.6 / 31.0. The device co-schedules 7 clusters of 16, so 64-row tiles (8 clusters) take two
// waves (19.4 us); 8-CTA clusters of 128 columns double the epilogue per warp (14.9 us). Default: bn 64 (128 for
// N 2048), 128-row tiles, multicast. Warp 0 = producer, warp 1 = MMA issuer, warps 2..5 = epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "internal.h"
#include "launch.cuh"
#include "philox.cuh"
#include "ptx.cuh"

namespace vqa {

namespace {

constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int LL_THREADS = 192;
constexpr int MAX_CL = 16;

struct LlArgs {
  int M, N, K;
  const float* bias;
  const float* gamma;
  const float* beta;
  const float* mul;
  float* z;        // forward: written; backward: read
  float* mean;     // forward: written (cluster rank 0); backward: read
  float* rstd;
  float* y;
  float* out_f32;
  __nv_bfloat16* out_hi;
  float* raw;
  float* dz_f32;
  __nv_bfloat16* dz_hi;
  float* dgamma_part;
  float* dbeta_part;
  unsigned long long seed, step;
  unsigned int site, thr;
  float inv_keep;
  int act;
};

template <int BN, int KBS, int A_ROWS>
struct LlCfg {
  static constexpr int A_TILE = A_ROWS * 128;   // bytes of one k-block of A (64 bf16 per row)
  static constexpr int B_TILE = BN * 128;
  static constexpr int STAGE = KBS * (A_TILE + B_TILE);
  static constexpr int STAGES = (192 * 1024) / STAGE;
  static constexpr int PIPE = STAGES * STAGE;
  static constexpr int STG_LD = BN + 4;              // floats: padded staging row
  static constexpr int STG_BYTES = A_ROWS * STG_LD * 4;  // one staging plane (the backward epilogue uses two)
  static constexpr int STATS_BYTES = MAX_CL * 128 * 8;
  static constexpr int ROW_BYTES = 128 * 8;
  static constexpr int SMEM = PIPE + STATS_BYTES + 2 * ROW_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(2 * STG_BYTES <= PIPE, "staging planes live in the drained pipeline buffers");
  static_assert(STAGES >= 2, "pipeline depth");
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}

__device__ __forceinline__ void ld8g(const float* p, float (&x)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
  x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void st8g(float* p, const float (&x)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
}
__device__ __forceinline__ void st8bf(__nv_bfloat16* p, const float (&x)[8]) {
  __nv_bfloat162 h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __nv_bfloat162(__float2bfloat16_rn(x[2 * j]), __float2bfloat16_rn(x[2 * j + 1]));
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<uint4*>(h);
}

// A (shared by the whole cluster) arrives by TMA multicast: a stage's A tiles (KBS k-blocks x A_ROWS rows) are cut into
// one slice per CTA (`rps` rows of one k-block: a 2-D box of tm_a), every CTA fetches its slice ONCE and the hardware
// delivers it to the same offset of all CTAs of the cluster -- the cluster sits in one GPC, and sixteen private copies of
// the A tile through that GPC's port were what bounded the first version of this kernel (15 -> 22 us from K 1024 to
// 2048). A stage is therefore refilled only when EVERY CTA has consumed it: the MMA issuer's commit is multicast to the
// empty barriers (count = cluster size) of all CTAs. tm_b: this CTA's own weight columns as one multi-k-block box
// ({64 k, BN rows, KBS} when K-major, the {64 mn, 64 k, BN / 64, KBS} block map when MN-major); tm_b2: per-k-block 2-D
// boxes for the group that holds a partial k-block (K-major B whose K is not a multiple of 64: the answer dimension).
template <int BN, int KBS, int A_ROWS, bool B_MN, bool BWD>
__global__ void __launch_bounds__(LL_THREADS) linear_ln_kernel(
    const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
    const __grid_constant__ CUtensorMap tm_b2, const LlArgs g, const int rps, const int mc) {
  using Cfg = LlCfg<BN, KBS, A_ROWS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float2* stats = reinterpret_cast<float2*>(smem + Cfg::PIPE);                       // [MAX_CL][128], written by the peers
  float2* row_a = reinterpret_cast<float2*>(smem + Cfg::PIPE + Cfg::STATS_BYTES);    // [128] combined statistics
  float2* row_p = row_a + 128;                                                       // [128] this CTA's partials (backward)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::PIPE + Cfg::STATS_BYTES + 2 * Cfg::ROW_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cl = gridDim.x;                       // the cluster spans the whole N axis
  const uint32_t rank = ptx::cluster_ctarank();   // == blockIdx.x
  const int m0 = blockIdx.y * A_ROWS;
  const int n0 = static_cast<int>(rank) * BN;
  const int M = g.M, N = g.N, K = g.K;
  const int num_kb = (K + BK - 1) / BK;

  const uint16_t cl_mask = static_cast<uint16_t>((1u << cl) - 1u);
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_b);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], mc ? cl : 1);   // multicast: one arrival from the MMA issuer of every CTA of the cluster
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_arrive();   // (every CTA of the cluster is running before anyone writes into its shared memory)
  pdl_sync();
  cluster_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    const int spk = (mc && cl > KBS) ? cl / KBS : 1;  // slices per k-block
    const int n_slices = KBS * spk;                    // per stage; multicast: CTA j fetches slices j, j + cl, ...
    const int sl0 = mc ? static_cast<int>(rank) : 0, sl_step = mc ? cl : 1;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; kb += KBS) {
      ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sa = smem + stage * Cfg::STAGE;
      uint8_t* sb = sa + KBS * Cfg::A_TILE;
      if (ptx::elect_one()) {
        const bool tail = !B_MN && (K % BK) != 0 && kb + KBS > K / BK;
        const int nblk = num_kb - kb < KBS ? num_kb - kb : KBS;
        // A: always the whole stage (slices beyond K are zero-filled by the tensor map and still count their bytes)
        ptx::mbar_arrive_expect_tx(&full_bar[stage], KBS * Cfg::A_TILE + (tail ? nblk : KBS) * Cfg::B_TILE);
        for (int sl = sl0; sl < n_slices; sl += sl_step) {
          const int i = sl / spk, r0 = (sl % spk) * rps;
          uint8_t* dst = sa + i * Cfg::A_TILE + r0 * 128;
          if (mc) ptx::tma_load_2d_multicast(dst, &tm_a, &full_bar[stage], (kb + i) * BK, m0 + r0, cl_mask);
          else ptx::tma_load_2d(dst, &tm_a, &full_bar[stage], (kb + i) * BK, m0 + r0);
        }
        if (!tail) {
          if (B_MN) ptx::tma_load_4d(sb, &tm_b, &full_bar[stage], 0, 0, n0 >> 6, kb);
          else ptx::tma_load_3d(sb, &tm_b, &full_bar[stage], 0, n0, kb);
        } else {
          for (int i = 0; i < nblk; ++i)
            ptx::tma_load_2d(sb + i * Cfg::B_TILE, &tm_b2, &full_bar[stage], (kb + i) * BK, n0);
        }
      }
      __syncwarp();
      if (++stage == Cfg::STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, BN, false, B_MN);
    constexpr uint32_t B_LBO = B_MN ? 8192 : 16, B_STEP = B_MN ? 2048 : 32;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; kb += KBS) {
      ptx::mbar_wait(&full_bar[stage], phase);
      ptx::tc_fence_after();
      const uint32_t st = ptx::smem_u32(smem + stage * Cfg::STAGE);
      if (ptx::elect_one()) {
#pragma unroll
        for (int i = 0; i < KBS; ++i) {
          if (kb + i >= num_kb) break;
          const uint32_t sa = st + i * Cfg::A_TILE;
          const uint32_t sb = st + KBS * Cfg::A_TILE + i * Cfg::B_TILE;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            const uint64_t da = ptx::make_smem_desc_sw128(sa + kk * 32, 16, 1024);
            const uint64_t db = ptx::make_smem_desc_sw128(sb + kk * B_STEP, B_LBO, 1024);
            ptx::umma_f16(tmem_base, da, db, idesc, (kb | i | kk) != 0);
          }
        }
        // the stage is free once the MMAs of EVERY CTA have read it (its A tiles are shared by multicast)
        if (mc) ptx::umma_commit_multicast(&empty_bar[stage], cl_mask);
        else ptx::umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == Cfg::STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (ptx::elect_one()) ptx::umma_commit(tmem_full_bar);
    __syncwarp();
  }

  // ===================== epilogue, part 1: this CTA's columns =====================
  const int q = warp & 3;                                   // TMEM lanes [32 q, 32 q + 32)
  const bool epi = warp >= 2 && q * 32 < A_ROWS;             // 64-row tiles: the upper accumulator lanes are not rows of ours
  constexpr int LPR = BN / 8;                                // lanes per row in the coalesced passes (8 columns per lane)
  constexpr int RPI = 32 / LPR;                              // rows per warp iteration
  const int cg = lane % LPR;
  const int col = n0 + cg * 8;
  float* stg0 = reinterpret_cast<float*>(smem);
  float* stg1 = reinterpret_cast<float*>(smem + Cfg::STG_BYTES);
  const uint32_t stats_addr = ptx::smem_u32(stats);
  float gm[8], bt[8];
  if (epi) {
    ld8g(g.gamma + col, gm);
    ld8g(g.beta + col, bt);
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    // every MMA has retired => every TMA load was consumed: the pipeline buffers are free for staging
    float* my_row = stg0 + (q * 32 + lane) * Cfg::STG_LD;
    float p0, p1;
    if (!BWD) {
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + c) + j);
          float4 t = make_float4(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y,
                                 __uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
          sum += (t.x + t.y) + (t.z + t.w);
          reinterpret_cast<float4*>(my_row + c)[j] = t;
        }
      }
      const float mean_c = sum * (1.0f / BN);
      float m2 = 0.f;
#pragma unroll 4
      for (int j = 0; j < BN / 4; ++j) {
        const float4 t = reinterpret_cast<const float4*>(my_row)[j];
        const float a = t.x - mean_c, b = t.y - mean_c, c = t.z - mean_c, d = t.w - mean_c;
        m2 += (a * a + b * b) + (c * c + d * d);
      }
      p0 = mean_c;
      p1 = m2;
    } else {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          reinterpret_cast<float4*>(my_row + c)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
      __syncwarp();
      // coalesced pass: dropout / Hadamard / activation backward, d x_hat, the two row sums of this CTA's columns
#pragma unroll 1
      for (int it = 0; it < 32 / RPI; ++it) {
        const int rl = q * 32 + it * RPI + lane / LPR;
        const int row = m0 + rl;
        const bool ok = row < M;
        float d[8], xh[8];
        ld8g(stg0 + rl * Cfg::STG_LD + cg * 8, d);
        float s1 = 0.f, s2 = 0.f;
        if (ok) {
          const long long o = static_cast<long long>(row) * N + col;
          if (g.raw) st8g(g.raw + o, d);
          if (g.thr < 65536u) {
            const uint32_t bits = philox_keep_bits(
                philox4x32_10(static_cast<unsigned long long>(row) * (N >> 3) + (col >> 3), g.site, g.seed, g.step), g.thr);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = ((bits >> j) & 1u) ? d[j] * g.inv_keep : 0.f;
          } else if (g.inv_keep != 1.0f) {
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] *= g.inv_keep;
          }
          if (g.mul) {
            float m[8];
            ld8g(g.mul + o, m);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] *= m[j];
          }
          float z[8], dg[8], db[8];
          ld8g(g.z + o, z);
          const float mean = g.mean[row], rstd = g.rstd[row];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            xh[j] = (z[j] - mean) * rstd;
            const float ypre = fmaf(xh[j], gm[j], bt[j]);
            float dy;
            if (g.act == 1) {
              const float t = tanhf(ypre);
              dy = d[j] * (1.0f - t * t);
            } else {
              dy = ypre > 0.f ? d[j] : 0.f;
            }
            dg[j] = dy * xh[j];
            db[j] = dy;
            d[j] = dy * gm[j];
            s1 += d[j];
            s2 = fmaf(d[j], xh[j], s2);
          }
          if (g.dgamma_part) {
            st8g(g.dgamma_part + o, dg);
            st8g(g.dbeta_part + o, db);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = xh[j] = 0.f;
        }
        st8g(stg0 + rl * Cfg::STG_LD + cg * 8, d);
        st8g(stg1 + rl * Cfg::STG_LD + cg * 8, xh);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (cg == 0) row_p[rl] = make_float2(s1, s2);
      }
      __syncwarp();
      const float2 pp = row_p[q * 32 + lane];
      p0 = pp.x;
      p1 = pp.y;
    }
    // this row's partials -> the statistics table of every CTA of the cluster
    const uint32_t slot = stats_addr + (rank * 128u + static_cast<uint32_t>(q * 32 + lane)) * 8u;
    for (int dst = 0; dst < cl; ++dst) st_cluster_f2(ptx::mapa_u32(slot, static_cast<uint32_t>(dst)), p0, p1);
  }
  cluster_arrive();
  cluster_wait();

  // ===================== epilogue, part 2: whole-row statistics, outputs =====================
  if (epi) {
    const int r = q * 32 + lane;
    if (!BWD) {
      float msum = 0.f;
      for (int c = 0; c < cl; ++c) msum += stats[c * 128 + r].x;
      const float mean = msum / cl;
      float m2 = 0.f, dev = 0.f;
      for (int c = 0; c < cl; ++c) {
        const float2 t = stats[c * 128 + r];
        m2 += t.y;
        dev += (t.x - mean) * (t.x - mean);
      }
      const float var = (m2 + BN * dev) / N;
      const float rstd = 1.0f / sqrtf(var + 1e-12f);
      row_a[r] = make_float2(mean, rstd);
      if (rank == 0 && m0 + r < M) {
        g.mean[m0 + r] = mean;
        g.rstd[m0 + r] = rstd;
      }
    } else {
      float s1 = 0.f, s2 = 0.f;
      for (int c = 0; c < cl; ++c) {
        const float2 t = stats[c * 128 + r];
        s1 += t.x;
        s2 += t.y;
      }
      row_a[r] = make_float2(s1 / N, s2 / N);
    }
    __syncwarp();
#pragma unroll 1
    for (int it = 0; it < 32 / RPI; ++it) {
      const int rl = q * 32 + it * RPI + lane / LPR;
      const int row = m0 + rl;
      if (row >= M) continue;
      const long long o = static_cast<long long>(row) * N + col;
      const float2 st = row_a[rl];
      float x[8];
      ld8g(stg0 + rl * Cfg::STG_LD + cg * 8, x);
      if (!BWD) {
        if (g.z) st8g(g.z + o, x);
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float ypre = fmaf((x[j] - st.x) * st.y, gm[j], bt[j]);
          y[j] = g.act == 1 ? tanhf(ypre) : fmaxf(ypre, 0.f);
        }
        if (g.y) st8g(g.y + o, y);
        if (g.mul) {
          float m[8];
          ld8g(g.mul + o, m);
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] *= m[j];
        }
        if (g.thr < 65536u) {
          const uint32_t bits = philox_keep_bits(
              philox4x32_10(static_cast<unsigned long long>(row) * (N >> 3) + (col >> 3), g.site, g.seed, g.step), g.thr);
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = ((bits >> j) & 1u) ? y[j] * g.inv_keep : 0.f;
        } else if (g.inv_keep != 1.0f) {
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] *= g.inv_keep;
        }
        if (g.out_f32) st8g(g.out_f32 + o, y);
        if (g.out_hi) st8bf(g.out_hi + o, y);
      } else {
        float xh[8], dz[8];
        ld8g(stg1 + rl * Cfg::STG_LD + cg * 8, xh);
        const float rstd = g.rstd[row];
#pragma unroll
        for (int j = 0; j < 8; ++j) dz[j] = rstd * (x[j] - st.x - xh[j] * st.y);
        if (g.dz_f32) st8g(g.dz_f32 + o, dz);
        if (g.dz_hi) st8bf(g.dz_hi + o, dz);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, BN);
  }
}

template <int BN, int KBS, int A_ROWS, bool B_MN, bool BWD>
struct LlLaunch {
  using Cfg = LlCfg<BN, KBS, A_ROWS>;
  static auto kern() { return linear_ln_kernel<BN, KBS, A_ROWS, B_MN, BWD>; }
  static cudaError_t config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at, int cl, int tiles, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(kern(), cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(kern(), cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3(cl, tiles, 1);
    cfg->blockDim = dim3(LL_THREADS);
    cfg->dynamicSmemBytes = Cfg::SMEM;
    cfg->stream = s;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg->attrs = at;
    cfg->numAttrs = 2;
    return cudaSuccess;
  }
  // how many clusters of `cl` CTAs the device co-schedules (cached per cluster size; 0 = the launch is impossible)
  static int max_clusters(int cl) {
    static int cache[MAX_CL + 1] = {};
    static bool known[MAX_CL + 1] = {};
    if (known[cl]) return cache[cl];
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (config(&cfg, at, cl, 1, nullptr) != cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&n, kern(), &cfg) != cudaSuccess) {
      (void)cudaGetLastError();
      n = 0;
    }
    known[cl] = true;
    cache[cl] = n;
    return n;
  }
  static cudaError_t launch(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& b2, const LlArgs& g, int cl,
                            int rps, int mc, cudaStream_t s) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[2];
    cudaError_t e = config(&cfg, at, cl, (g.M + A_ROWS - 1) / A_ROWS, s);
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, kern(), a, b, b2, g, rps, mc);
  }
};

template <int BN, int KBS, int A_ROWS, bool B_MN, bool BWD>
VqaStatus launch_cfg(const LinearLn& d, const LlArgs& g, int cl, int mc, cudaStream_t s, bool* launched) {
  using LL = LlLaunch<BN, KBS, A_ROWS, B_MN, BWD>;
  if (LL::max_clusters(cl) < 1) return VQA_OK;   // the device cannot place this cluster: the caller takes the unfused kernels
  mc = (mc && cl > 1) ? 1 : 0;
  const int spk = (mc && cl > KBS) ? cl / KBS : 1;
  const int rps = A_ROWS / spk;   // rows per A slice (>= 8: whole 1024-byte swizzle atoms)
  CUtensorMap a, b, b2;
  bool ok = cached_tmap(&a, d.a, d.K, d.M, d.lda, 64, rps);
  if (B_MN) {
    ok = ok && cached_tmap_mnblocks(&b, d.b, d.N, d.K, d.ldb, BN / 64, KBS);
    b2 = b;
  } else {
    ok = ok && cached_tmap_kblocks(&b, d.b, d.K, d.N, d.ldb, BN, KBS) && cached_tmap(&b2, d.b, d.K, d.N, d.ldb, 64, BN);
  }
  if (!ok) return set_error(VQA_ERR_CUDA, "linear_ln: cuTensorMapEncodeTiled failed");
  cudaError_t e = LL::launch(a, b, b2, g, cl, rps, mc, s);
  if (e != cudaSuccess) return set_cuda_error(e, "linear_ln launch");
  count_launch();
  *launched = true;
  return VQA_OK;
}

// tile plan: columns per CTA (the cluster is N / bn CTAs), rows per tile, A by multicast or privately
struct LlPlan {
  int bn, rows, mc;
};

template <bool B_MN, bool BWD>
VqaStatus launch_plan(const LinearLn& d, const LlArgs& g, const LlPlan& p, cudaStream_t s, bool* launched) {
  const int cl = d.N / p.bn;
  if (p.bn == 64 && p.rows == 64) return launch_cfg<64, 4, 64, B_MN, BWD>(d, g, cl, p.mc, s, launched);
  if (p.bn == 64) return launch_cfg<64, 4, 128, B_MN, BWD>(d, g, cl, p.mc, s, launched);
  if (p.bn == 128 && p.rows == 64) return launch_cfg<128, 2, 64, B_MN, BWD>(d, g, cl, p.mc, s, launched);
  if (p.bn == 128) return launch_cfg<128, 2, 128, B_MN, BWD>(d, g, cl, p.mc, s, launched);
  return launch_cfg<256, 2, 64, B_MN, BWD>(d, g, cl, p.mc, s, launched);   // 256 columns: 64-row tiles only (staging)
}

bool plan_ok(const LinearLn& d, const LlPlan& p) {
  if (p.bn != 64 && p.bn != 128 && p.bn != 256) return false;
  if (p.rows != 64 && !(p.rows == 128 && p.bn != 256)) return false;
  const int c = d.N / p.bn;
  return d.N % p.bn == 0 && c >= 1 && c <= MAX_CL && (c & (c - 1)) == 0;
}

}  // namespace

bool linear_ln_enabled() {
  static const bool on = getenv("VQA_LINEAR_LN") == nullptr || atoi(getenv("VQA_LINEAR_LN")) != 0;
  return on;
}

VqaStatus linear_ln_launch(const LinearLn& d, cudaStream_t s, bool* launched) {
  *launched = false;
  if (!linear_ln_enabled()) return VQA_OK;
  if (d.M <= 0 || d.N <= 0 || d.K <= 0 || !d.a || !d.b || !d.gamma || !d.beta) return VQA_OK;
  if ((d.lda & 7) || (d.ldb & 7) || (d.N & 63)) return VQA_OK;
  // the cluster covers the row: N = bn x (1, 2, 4, 8 or 16 CTAs). Default: as many CTAs as a cluster holds.
  LlPlan plan{0, 128, 1};
  for (int cand : {64, 128, 256}) {
    plan.bn = cand;
    if (cand == 256) plan.rows = 64;
    if (plan_ok(d, plan)) break;
    plan.bn = 0;
  }
  static const bool tune = getenv("VQA_LINEAR_LN_TUNE") != nullptr;   // experiments: plan from the environment, per call
  if (tune) {
    LlPlan t = plan;
    if (const char* e = getenv("VQA_LINEAR_LN_BN")) t.bn = atoi(e);
    if (const char* e = getenv("VQA_LINEAR_LN_ROWS")) t.rows = atoi(e);
    if (const char* e = getenv("VQA_LINEAR_LN_MC")) t.mc = atoi(e);
    if (plan_ok(d, t)) plan = t;
    else plan.bn = 0;
  }
  if (!plan.bn) return VQA_OK;
  if (!d.backward) {
    if (!d.b_mn_major || (d.K % BK) != 0 || !d.bias || !d.mean || !d.rstd) return VQA_OK;
  } else {
    if (d.b_mn_major || (d.K & 7) || !d.z || !d.mean || !d.rstd) return VQA_OK;
  }
  LlArgs g{};
  g.M = d.M; g.N = d.N; g.K = d.K;
  g.bias = d.bias; g.gamma = d.gamma; g.beta = d.beta; g.mul = d.mul;
  g.z = d.z; g.mean = d.mean; g.rstd = d.rstd; g.y = d.y; g.out_f32 = d.out_f32;
  g.out_hi = static_cast<__nv_bfloat16*>(d.out_hi);
  g.raw = d.raw; g.dz_f32 = d.dz_f32; g.dz_hi = static_cast<__nv_bfloat16*>(d.dz_hi);
  g.dgamma_part = d.dgamma_part; g.dbeta_part = d.dbeta_part;
  g.seed = d.seed; g.step = d.step; g.site = d.stream_id;
  g.thr = keep_threshold(d.keep); g.inv_keep = 1.0f / d.keep; g.act = d.act;
  if (!d.backward) return launch_plan<true, false>(d, g, plan, s, launched);
  return launch_plan<false, true>(d, g, plan, s, launched);
}

}  // namespace vqa
