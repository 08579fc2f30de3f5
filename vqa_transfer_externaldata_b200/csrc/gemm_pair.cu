// Persistent CTA-pair GEMM (tcgen05.mma.cta_group::2): D[M,N] = A[M,K] * B[N,K]^T (+bias) (+addend), bf16
// operands, fp32 accumulation in TMEM. This is the kernel behind the large contractions of the answer model:
// the region-feature projection v_linear_v and its weight gradient (vqa/model_vlmap_answer.py:126-129), the
// hoisted x-projections, weight gradients and embedding gradient of the GRU (vlmap/modules.py:124-140) and the
// wider heads.
//
// Why pairs: a single CTA ingests at most ~67 GB/s from L2 (measured, profiles/r01_launch_summary_v3.md), which
// caps 128 x 256 tiles at ~870 TFLOP/s. In a pair each CTA loads only ITS 128 rows of A and HALF of the B
// columns for a 256 x BN tile: 32 KB instead of 48 KB per 4.2 MFLOP.
//
// Structure (per CTA, 10 warps): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only; all role loops are
// warp-uniform with one elected lane issuing), warps 2..9 = epilogue (two warps per TMEM lane quarter, each
// taking half of the columns). Tiles are visited persistently (item = tile + split * tiles, round-robin over the
// pairs); the accumulator is double-buffered in TMEM so the epilogue of item i overlaps the main loop of i + 1.
// Split-K (weight gradients: few output tiles, long K) reduces in a FIXED order -- split s adds its partial after
// split s - 1, handed over through a per-tile semaphore -- so results do not depend on scheduling.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "internal.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace vqa {

namespace {

constexpr int P_BK = 64;
constexpr int P_THREADS = 320;
constexpr int P_EPI_THREADS = 256;

template <int BN>
struct PairCfg {
  static constexpr int A_TILE = 128 * P_BK * 2;          // this CTA's 128 rows of A
  static constexpr int B_TILE = (BN / 2) * P_BK * 2;     // this CTA's half of the B columns
  // a stage holds TWO k-blocks of each operand: [A kb0 | A kb1 | B kb0 | B kb1]. An SM completes only ~4 TMA
  // operations per microsecond whatever their size (scripts/probes/tma_ingest.cu), so each operand arrives as ONE
  // 3-D / 4-D box of two k-block tiles (32 KB) instead of 2 - 4 separate boxes per k-block.
  static constexpr int KBS = 2;
  static constexpr int STAGE = KBS * (A_TILE + B_TILE);
  static constexpr int STAGES = (BN == 256) ? 3 : 4;     // 192 KB in flight
  static constexpr int STG_CHUNK = 32 * 32 * 4;          // one 32 x 32 fp32 chunk (the source box of a TMA store)
  static constexpr int STG_WARP = STG_CHUNK;             // one staging chunk per epilogue warp
  static constexpr int SMEM = STAGES * STAGE + 8 * STG_WARP + 1024 + 256;
};

struct PairArgs {
  const float* bias;
  const float* addend;
  long long ld_addend;
  float* out_f32;
  long long ld_f32;
  bf16* out_bf;
  long long ld_bf;
  int M, N, K;
  int splits;
  int big_a, big_b;    // operand arrives as one multi-k-block box per stage (needs K % 64 == 0), else 2-D boxes per k-block
  unsigned int* sem;   // [tiles] split hand-over counters (zero on entry, zero again on exit)
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(P_EPI_THREADS) : "memory"); }

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(P_THREADS, 1) gemm_pair_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                                 const __grid_constant__ CUtensorMap tm_b,
                                                                 const __grid_constant__ CUtensorMap tm_a2,
                                                                 const __grid_constant__ CUtensorMap tm_b2,
                                                                 const __grid_constant__ CUtensorMap tm_of,
                                                                 const __grid_constant__ CUtensorMap tm_ob,
                                                                 PairArgs g) {
  using Cfg = PairCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t stg_all = ptx::smem_u32(smem + Cfg::STAGES * Cfg::STAGE);   // 1024-byte aligned (swizzle atoms)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE + 8 * Cfg::STG_WARP);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_nt = (g.N + BN - 1) / BN;
  const int tiles = ((g.M + 255) / 256) * num_nt;
  const int items = tiles * g.splits;
  const int num_kb = (g.K + P_BK - 1) / P_BK;
  const int num_kp = (num_kb + Cfg::KBS - 1) / Cfg::KBS;   // stages (k-block pairs) over K

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(g.big_a ? &tm_a2 : &tm_a);
    ptx::prefetch_tensormap(g.big_b ? &tm_b2 : &tm_b);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 2);   // one arrival per CTA of the pair
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, 2 * BN);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // the peer's barriers exist before anything is signalled at them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // prologue done under the previous kernel's tail; global memory (operands, semaphores) from here on

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int w = pair; w < items; w += num_pairs) {
      const int split = w / tiles, tile = w - split * tiles;
      const int m0 = (tile / num_nt) * 256 + static_cast<int>(rank) * 128;
      const int n0 = (tile % num_nt) * BN + static_cast<int>(rank) * (BN / 2);
      const int kp0 = static_cast<int>(static_cast<long long>(num_kp) * split / g.splits);
      const int kp1 = static_cast<int>(static_cast<long long>(num_kp) * (split + 1) / g.splits);
      for (int kp = kp0; kp < kp1; ++kp) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE;
        uint8_t* sb = sa + Cfg::KBS * Cfg::A_TILE;
        const uint32_t lf = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);   // the leader's barrier
        const int kb = kp * Cfg::KBS;
        if (ptx::elect_one()) {
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE);
          if (g.big_a) {
            if (A_MN) ptx::tma_load_4d_pair(sa, &tm_a2, lf, 0, 0, m0 >> 6, kb);
            else ptx::tma_load_3d_pair(sa, &tm_a2, lf, 0, m0, kb);
          } else {
#pragma unroll
            for (int i = 0; i < Cfg::KBS; ++i) {   // k-blocks beyond K are zero-filled by the tensor map
              const int k0 = (kb + i) * P_BK;
              if (A_MN) {
                ptx::tma_load_2d_pair(sa + i * Cfg::A_TILE, &tm_a, lf, m0, k0);
                ptx::tma_load_2d_pair(sa + i * Cfg::A_TILE + 8192, &tm_a, lf, m0 + 64, k0);
              } else {
                ptx::tma_load_2d_pair(sa + i * Cfg::A_TILE, &tm_a, lf, k0, m0);
              }
            }
          }
          if (g.big_b) {
            if (B_MN) ptx::tma_load_4d_pair(sb, &tm_b2, lf, 0, 0, n0 >> 6, kb);
            else ptx::tma_load_3d_pair(sb, &tm_b2, lf, 0, n0, kb);
          } else {
#pragma unroll
            for (int i = 0; i < Cfg::KBS; ++i) {
              const int k0 = (kb + i) * P_BK;
              if (B_MN) {
#pragma unroll
                for (int j = 0; j < BN / 128; ++j)
                  ptx::tma_load_2d_pair(sb + i * Cfg::B_TILE + j * 8192, &tm_b, lf, n0 + 64 * j, k0);
              } else {
                ptx::tma_load_2d_pair(sb + i * Cfg::B_TILE, &tm_b, lf, k0, n0);
              }
            }
          }
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(256, BN, A_MN, B_MN);
      constexpr uint32_t A_LBO = A_MN ? 8192 : 16, A_STEP = A_MN ? 2048 : 32;
      constexpr uint32_t B_LBO = B_MN ? 8192 : 16, B_STEP = B_MN ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = pair; w < items; w += num_pairs, ++it) {
        const int split = w / tiles;
        const int kp0 = static_cast<int>(static_cast<long long>(num_kp) * split / g.splits);
        const int kp1 = static_cast<int>(static_cast<long long>(num_kp) * (split + 1) / g.splits);
        const int acc = it & 1;
        ptx::mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);   // both epilogues drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kp = kp0; kp < kp1; ++kp) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE);
          const uint32_t sb = sa + Cfg::KBS * Cfg::A_TILE;
          if (ptx::elect_one()) {
#pragma unroll
            for (int i = 0; i < Cfg::KBS; ++i) {
#pragma unroll
              for (int kk = 0; kk < P_BK / 16; ++kk) {
                const uint64_t da = ptx::make_smem_desc_sw128(sa + i * Cfg::A_TILE + kk * A_STEP, A_LBO, 1024);
                const uint64_t db = ptx::make_smem_desc_sw128(sb + i * Cfg::B_TILE + kk * B_STEP, B_LBO, 1024);
                ptx::umma_f16_pair(d_tmem, da, db, idesc, (kp > kp0) || (i > 0) || (kk > 0));
              }
            }
            ptx::umma_commit_pair(&empty_bar[stage], 3);   // frees this stage in BOTH CTAs
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (ptx::elect_one()) ptx::umma_commit_pair(&tmem_full[acc], 3);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (both CTAs) =====================
    // tcgen05.ld gives thread = row. Each warp stages 32 x 32 chunks in shared memory (the 128-byte-swizzled box
    // layout of the output tensor map) and hands them to the TMA: one bulk store per chunk instead of 32 row
    // stores, tails clipped by the tensor map, and split > 0 becomes a reduce-add performed at L2.
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;    // which half of the columns
    const bool leader_thread = threadIdx.x == 64;
    const uint32_t stg = stg_all + (warp - 2) * Cfg::STG_WARP;
    if (lane == 0) {
      if (g.out_f32) ptx::prefetch_tensormap(&tm_of);
      if (g.out_bf) ptx::prefetch_tensormap(&tm_ob);
    }
    int it = 0;
    for (int w = pair; w < items; w += num_pairs, ++it) {
      const int split = w / tiles, tile = w - split * tiles;
      const int row0 = (tile / num_nt) * 256 + static_cast<int>(rank) * 128 + q * 32;
      const int row = row0 + lane;
      const int n0 = (tile % num_nt) * BN;
      const int acc = it & 1;
      ptx::mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      ptx::tc_fence_after();
      if (g.splits > 1 && split > 0) {
        // fixed-order reduction: wait until both CTAs of split - 1 have added their partials
        if (leader_thread) {
          const unsigned int target = 2u * static_cast<unsigned int>(split);
          const long long t0 = clock64();
          while (ld_acquire(g.sem + tile) < target) {
            if (clock64() - t0 > 4000000000LL) __trap();
          }
          ptx::fence_proxy_async_full();
        }
        epi_bar();
      }
      const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
      constexpr int NCH = BN / 64;   // 32-column chunks per warp
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        const int c = half * (BN / 2) + ch * 32;
        const int col0 = n0 + c;
        uint32_t v[32];
        ptx::tmem_ld_32x32(t_row + c, v);
        ptx::tmem_ld_wait();
        if (ch == NCH - 1) {
          // everything of this accumulator is in registers: hand it back to the MMA issuer
          ptx::tc_fence_before();
          epi_bar();
          if (leader_thread) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tmem_empty[acc]), 0));
        }
        if (col0 >= g.N) continue;   // warp-uniform: nothing of this chunk is inside the output
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        if (split == 0) {
          if (g.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (col0 + 4 * j < g.N) {   // N % 4 == 0
                const float4 bv = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + 4 * j));
                x[4 * j] += bv.x; x[4 * j + 1] += bv.y; x[4 * j + 2] += bv.z; x[4 * j + 3] += bv.w;
              }
            }
          }
          if (g.addend && row < g.M) {
            const float* ap = g.addend + static_cast<long long>(row) * g.ld_addend + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (col0 + 4 * j < g.N) {
                const float4 av = *reinterpret_cast<const float4*>(ap + 4 * j);
                x[4 * j] += av.x; x[4 * j + 1] += av.y; x[4 * j + 2] += av.z; x[4 * j + 3] += av.w;
              }
            }
          }
        }
        // one staging chunk per warp: the previous store must have finished READING it (its writes may still be in
        // flight) before it is overwritten
        if (g.out_f32) {
          if (lane == 0) ptx::bulk_wait_read<0>();
          __syncwarp();
          const uint32_t sb = stg;
          // row r of the box at r * 128 bytes, its 16-byte chunk j at position j ^ (r & 7): SWIZZLE_128B
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ptx::st_shared_v4(sb + lane * 128 + ((j ^ (lane & 7)) << 4), x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (split == 0) ptx::tma_store_2d(&tm_of, sb, col0, row0);
            else ptx::tma_reduce_add_2d(&tm_of, sb, col0, row0);
            ptx::bulk_commit();
          }
        }
        if (g.out_bf) {
          if (lane == 0) ptx::bulk_wait_read<0>();
          __syncwarp();
          const uint32_t sb = stg;
          // bf16 box: 64-byte rows, not swizzled
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 p0(__float2bfloat16_rn(x[8 * j]), __float2bfloat16_rn(x[8 * j + 1]));
            __nv_bfloat162 p1(__float2bfloat16_rn(x[8 * j + 2]), __float2bfloat16_rn(x[8 * j + 3]));
            __nv_bfloat162 p2(__float2bfloat16_rn(x[8 * j + 4]), __float2bfloat16_rn(x[8 * j + 5]));
            __nv_bfloat162 p3(__float2bfloat16_rn(x[8 * j + 6]), __float2bfloat16_rn(x[8 * j + 7]));
            ptx::st_shared_v4_b32(sb + lane * 64 + j * 16, *reinterpret_cast<uint32_t*>(&p0),
                                  *reinterpret_cast<uint32_t*>(&p1), *reinterpret_cast<uint32_t*>(&p2),
                                  *reinterpret_cast<uint32_t*>(&p3));
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tm_ob, sb, col0, row0);
            ptx::bulk_commit();
          }
        }
      }
      if (g.splits > 1) {
        // partials of this split are performed (not merely read) before the next split is let in
        if (lane == 0) {
          ptx::bulk_wait<0>();
          ptx::fence_proxy_async_full();
          __threadfence();
        }
        epi_bar();
        if (leader_thread) {
          const unsigned int old = atomicAdd(g.sem + tile, 1u);
          if (old + 1u == 2u * static_cast<unsigned int>(g.splits)) g.sem[tile] = 0u;   // last one in: ready for reuse
        }
      }
    }
    if (lane == 0) ptx::bulk_wait<0>();   // all stores of this warp are performed before the CTA may exit
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // no CTA leaves while its peer may still read its tiles or signal its barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 2 * BN);
  }
}

template <int BN, bool A_MN, bool B_MN>
cudaError_t launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                        const CUtensorMap& tof, const CUtensorMap& tob, const PairArgs& g, int pairs, cudaStream_t s) {
  using Cfg = PairCfg<BN>;
  auto kern = gemm_pair_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(P_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kern, ta, tb, ta2, tb2, tof, tob, g);
}

template <int BN>
cudaError_t launch_pair_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb,
                              const CUtensorMap& ta2, const CUtensorMap& tb2, const CUtensorMap& tof,
                              const CUtensorMap& tob, const PairArgs& g, int pairs, cudaStream_t s) {
  if (a_mn) {
    if (b_mn) return launch_pair<BN, true, true>(ta, tb, ta2, tb2, tof, tob, g, pairs, s);
    return launch_pair<BN, true, false>(ta, tb, ta2, tb2, tof, tob, g, pairs, s);
  }
  if (b_mn) return launch_pair<BN, false, true>(ta, tb, ta2, tb2, tof, tob, g, pairs, s);
  return launch_pair<BN, false, false>(ta, tb, ta2, tb2, tof, tob, g, pairs, s);
}

}  // namespace

// Is the pair kernel the better choice for this problem, and with which tile width / split count?
// (single bf16 plane only: the hi + lo split-precision path stays on the single-CTA kernel)
bool gemm_pair_plan(const VqaGemmDesc& d, int num_sms, const GemmCtx* ctx, int narrow, int* bn_out, int* splits_out) {
  static const int mode = getenv("VQA_GEMM_PAIR") ? atoi(getenv("VQA_GEMM_PAIR")) : 1;
  const bool forced = d.block_n < 0;   // block_n = -128 / -256: the caller insists on the pair kernel
  if (!forced && (mode == 0 || d.block_n != 0)) return false;
  if (d.a_lo || d.b_lo || d.out_lo) return false;
  // TMA stores: 16-byte aligned bases and pitches
  if (d.out_hi && ((d.ld_bf & 7) || (reinterpret_cast<uintptr_t>(d.out_hi) & 15))) return false;
  if (d.out_f32 && (reinterpret_cast<uintptr_t>(d.out_f32) & 15)) return false;
  if (forced && d.block_n != -128 && d.block_n != -256) return false;
  // M <= 512 with a short K (the heads): 128 x 64 single-CTA tiles put more SMs to work than two rows of pair
  // tiles (measured); a long K (weight gradients) is split instead
  if (!forced && mode == 1 && !(d.M >= 1024 || (d.M >= 256 && d.K >= 4096))) return false;
  if (!forced && (d.M < 256 || d.N < 128 || d.K < 256)) return false;
  const int pairs = num_sms / 2;
  const int bn = forced ? -d.block_n : (d.N >= 256 ? 256 : 128);
  const long long tiles = static_cast<long long>((d.M + 255) / 256) * ((d.N + bn - 1) / bn);
  const int num_kb = (d.K + P_BK - 1) / P_BK;
  int splits = 1;
  if (!narrow && !d.out_hi && d.out_f32 && ctx && ctx->sem && tiles < pairs && tiles <= ctx->region_elems) {
    splits = static_cast<int>(pairs / tiles);
    // (a very long K over a handful of tiles -- the x-row weight gradients of 5120 GRU sequences, K = 51200 -- is worth
    // eight hand-overs per tile; the heads' K <= 18432 keeps four)
    const int cap = num_kb >= 512 ? 8 : 4;
    if (splits > cap) splits = cap;
    while (splits > 1 && num_kb / splits < 8) --splits;   // keep >= 8 k-blocks per item
  }
  // M = 512 heads with a short K: the 128 x 64 single-CTA tiles start sooner than 2 x 256-row pair tiles fill
  if (!forced && !narrow && mode == 1 && tiles * splits < pairs / 2) return false;
  *bn_out = bn;
  *splits_out = splits;
  return true;
}

VqaStatus gemm_pair_launch(const VqaGemmDesc& d, int num_sms, int bn, int splits, GemmCtx* ctx, cudaStream_t stream) {
  CUtensorMap ta, tb;
  const bool ok = (d.a_mn_major ? cached_tmap(&ta, d.a_hi, d.M, d.K, d.lda, 64, 64)
                                : cached_tmap(&ta, d.a_hi, d.K, d.M, d.lda, 64, 128)) &&
                  (d.b_mn_major ? cached_tmap(&tb, d.b_hi, d.N, d.K, d.ldb, 64, 64)
                                : cached_tmap(&tb, d.b_hi, d.K, d.N, d.ldb, 64, bn / 2));
  // multi-k-block boxes (one TMA operation per operand per stage) whenever nothing inside a k-block needs clipping
  CUtensorMap ta2 = ta, tb2 = tb;
  const bool big = (d.K % 64) == 0 && !getenv("VQA_PAIR_SMALL_BOXES");
  const bool big_a = big && (d.a_mn_major ? cached_tmap_mnblocks(&ta2, d.a_hi, d.M, d.K, d.lda, 2, 2)
                                          : cached_tmap_kblocks(&ta2, d.a_hi, d.K, d.M, d.lda, 128, 2));
  const bool big_b = big && (d.b_mn_major ? cached_tmap_mnblocks(&tb2, d.b_hi, d.N, d.K, d.ldb, bn / 128, 2)
                                          : cached_tmap_kblocks(&tb2, d.b_hi, d.K, d.N, d.ldb, bn / 2, 2));
  CUtensorMap tof = ta, tob = ta;   // placeholders when an output is absent (never dereferenced)
  const bool ok2 = (!d.out_f32 || cached_tmap_kind(&tof, d.out_f32, 1, d.N, d.M, d.ld_f32, 32, 32)) &&
                   (!d.out_hi || cached_tmap_kind(&tob, d.out_hi, 2, d.N, d.M, d.ld_bf, 32, 32));
  if (!ok || !ok2) return set_error(VQA_ERR_CUDA, "vqa_gemm: cuTensorMapEncodeTiled failed");
  PairArgs g{};
  g.bias = d.bias; g.addend = d.addend; g.ld_addend = d.ld_addend;
  g.out_f32 = d.out_f32; g.ld_f32 = d.ld_f32;
  g.out_bf = static_cast<bf16*>(d.out_hi); g.ld_bf = d.ld_bf;
  g.M = d.M; g.N = d.N; g.K = d.K; g.splits = splits;
  g.big_a = big_a; g.big_b = big_b;

  const long long tiles = static_cast<long long>((d.M + 255) / 256) * ((d.N + bn - 1) / bn);
  g.sem = nullptr;
  if (splits > 1) {
    // concurrent launches (forked streams) get different regions of the semaphore ring
    const int region = ctx->next_region;
    ctx->next_region = (region + 1) % ctx->regions;
    g.sem = ctx->sem + static_cast<long long>(region) * ctx->region_elems;
  }
  const long long items = tiles * splits;
  int pairs = num_sms / 2;
  if (items < pairs) pairs = static_cast<int>(items);
  const bool amn = d.a_mn_major != 0, bmn = d.b_mn_major != 0;
  cudaError_t e = bn == 256 ? launch_pair_major<256>(amn, bmn, ta, tb, ta2, tb2, tof, tob, g, pairs, stream)
                            : launch_pair_major<128>(amn, bmn, ta, tb, ta2, tb2, tof, tob, g, pairs, stream);
  if (e != cudaSuccess) return set_cuda_error(e, "vqa_gemm (pair) launch");
  count_launch();
  return VQA_OK;
}

}  // namespace vqa
