// tcgen05 / TMEM / TMA GEMM for the dense contractions of the answer model
// (layers.fully_connected call sites: vlmap/modules.py:616-626, 634-641; GRUCell matmuls :131-135;
//  and their dgrad / wgrad counterparts that tf.gradients derives).
//
//   D[M,N] = A[M,K] * B[N,K]^T  (+ bias[N]) (+ addend[M,N])      fp32 accumulation in TMEM
//
// * operands are bf16 planes. PREC_BF16: one plane per operand. PREC_FP32: two planes (hi, lo) per
//   operand and three MMAs per k-step (hi*hi + hi*lo + lo*hi) -> ~2^-16 relative operand error.
// * each operand may be stored K-major (contraction index contiguous) or MN-major (contraction index
//   strided) -- forward uses TF's [in,out] weights as MN-major B, dgrad uses the same buffer as
//   K-major B, wgrad uses activations as MN-major A and B. No transposed copies exist anywhere.
// * one 128 x BN output tile per CTA; warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
//   warps 2..5 = epilogue (TMEM -> registers -> shared staging -> coalesced 128-bit global stores).
//   Two CTAs co-reside per SM (<= 113 KB shared, <= 256 TMEM columns each), so one CTA's epilogue
//   overlaps the other's main loop.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <unordered_map>

#include "internal.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace vqa {

namespace {

constexpr int BM = 128;  // UMMA M (cta_group::1)
constexpr int BK = 64;   // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

struct EpilogueArgs {
  const float* bias;
  const float* addend;
  long long ld_addend;
  float* out_f32;
  long long ld_f32;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  long long ld_bf;
  int b_early;   // the B operand (weights) is not written by the in-stream predecessor: its first stages may be fetched
                 // before the programmatic dependency resolves (multi-k-block stages only)
};

// KBS > 1 (bf16 mode only): a stage holds KBS k-blocks of each operand, [A kb0..kb(KBS-1) | B kb0..], and each
// operand arrives as ONE multi-k-block TMA box (cached_tmap_kblocks / cached_tmap_mnblocks): an SM completes only
// ~4 TMA operations per microsecond whatever their size, which is what bounds the latency of the M = 512 heads.
template <int BN, int SPLIT, int KBS = 1>
struct GemmCfg {
  static constexpr int A_TILE = BM * BK * 2;                 // bytes per plane
  static constexpr int B_TILE = BN * BK * 2;
  static constexpr int PLANES = (SPLIT == 3) ? 2 : 1;
  static constexpr int STAGE_BYTES = KBS * PLANES * (A_TILE + B_TILE);
  // keep <= ~110 KB for the bf16 path so two CTAs fit one SM; the split path takes the SM alone
  static constexpr int STAGES = KBS > 1 ? 2
                                : (SPLIT == 3) ? (BN == 256 ? 2 : 3)
                                               : (BN == 256 ? 4 : (BN == 128 ? 3 : 4));
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int STG_LD = BN + 4;                      // floats, padded staging row
  static constexpr int STAGING_BYTES = 4 * 32 * STG_LD * 4;  // 4 epilogue warps x 32 rows
  static constexpr int MAIN_BYTES = PIPE_BYTES > STAGING_BYTES ? PIPE_BYTES : STAGING_BYTES;
  static constexpr int SMEM_BYTES = MAIN_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// KBS > 1: tm_a_lo / tm_b_lo carry the multi-k-block maps of the (single) bf16 plane
template <int BN, int SPLIT, bool A_MN, bool B_MN, int KBS = 1>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_bf16_tcgen05_kernel(
    const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
    const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
    EpilogueArgs ep, int M, int N, int K) {
  static_assert(KBS == 1 || SPLIT == 1, "multi-k-block stages exist for the single-plane mode only");
  using Cfg = GemmCfg<BN, SPLIT, KBS>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::MAIN_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a_hi);
    ptx::prefetch_tensormap(&tm_b_hi);
    if (SPLIT == 3) {
      ptx::prefetch_tensormap(&tm_a_lo);
      ptx::prefetch_tensormap(&tm_b_lo);
    }
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
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
  // everything above overlapped the previous kernel's tail; operands / bias / addend are read after the dependency wait --
  // except the first stages of a B operand the caller declared stable (weights), which the producer requests before it
  int pre = 0;
  if (KBS > 1 && warp == 0 && ep.b_early) {
    for (int kb = 0; kb < num_kb && pre < Cfg::STAGES; kb += KBS, ++pre) {
      if ((K % BK) != 0 && kb + KBS > K / BK) break;   // (the group with the partial k-block takes the ordinary path)
      if (ptx::elect_one()) {
        uint8_t* sb = smem + pre * Cfg::STAGE_BYTES + KBS * Cfg::A_TILE;
        ptx::mbar_arrive_expect_tx(&full_bar[pre], Cfg::STAGE_BYTES);
        if (B_MN) ptx::tma_load_4d(sb, &tm_b_lo, &full_bar[pre], 0, 0, n0 >> 6, kb);
        else ptx::tma_load_3d(sb, &tm_b_lo, &full_bar[pre], 0, n0, kb);
      }
      __syncwarp();
    }
  }
  pdl_sync();

  if (warp == 0) {
    // ===================== TMA producer =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; kb += KBS) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
        const int k0 = kb * BK;
        if (KBS > 1 && kb / KBS < pre) {
          // barrier armed and B requested above: only A is missing
          if (ptx::elect_one()) {
            if (A_MN) ptx::tma_load_4d(st, &tm_a_lo, &full_bar[stage], 0, 0, m0 >> 6, kb);
            else ptx::tma_load_3d(st, &tm_a_lo, &full_bar[stage], 0, m0, kb);
          }
        } else
        if (ptx::elect_one()) {
        // K-major operands whose K is not a multiple of 64 (the answer dimension, 3000): the k-block boxes cannot clip
        // inside a k-block, so the group that holds the partial k-block arrives as per-k-block 2-D boxes (whose inner
        // extent is K: the tail is zero-filled); k-blocks beyond the last are neither loaded nor multiplied
        const bool tail = KBS > 1 && (K % BK) != 0 && kb + KBS > K / BK;
        const int nblk = num_kb - kb < KBS ? num_kb - kb : KBS;
        const uint32_t bytes_a = (!A_MN && tail) ? nblk * Cfg::A_TILE : KBS * Cfg::A_TILE;
        const uint32_t bytes_b = (!B_MN && tail) ? nblk * Cfg::B_TILE : KBS * Cfg::B_TILE;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], KBS > 1 ? bytes_a + bytes_b : Cfg::STAGE_BYTES);
        if (KBS > 1) {
          uint8_t* sa = st;
          uint8_t* sb = st + KBS * Cfg::A_TILE;
          if (A_MN) ptx::tma_load_4d(sa, &tm_a_lo, &full_bar[stage], 0, 0, m0 >> 6, kb);
          else if (!tail) ptx::tma_load_3d(sa, &tm_a_lo, &full_bar[stage], 0, m0, kb);
          else for (int i = 0; i < nblk; ++i) ptx::tma_load_2d(sa + i * Cfg::A_TILE, &tm_a_hi, &full_bar[stage], k0 + i * BK, m0);
          if (B_MN) ptx::tma_load_4d(sb, &tm_b_lo, &full_bar[stage], 0, 0, n0 >> 6, kb);
          else if (!tail) ptx::tma_load_3d(sb, &tm_b_lo, &full_bar[stage], 0, n0, kb);
          else for (int i = 0; i < nblk; ++i) ptx::tma_load_2d(sb + i * Cfg::B_TILE, &tm_b_hi, &full_bar[stage], k0 + i * BK, n0);
        } else
#pragma unroll
        for (int p = 0; p < Cfg::PLANES; ++p) {
          const CUtensorMap* ta = p ? &tm_a_lo : &tm_a_hi;
          const CUtensorMap* tb = p ? &tm_b_lo : &tm_b_hi;
          uint8_t* sa = st + p * Cfg::A_TILE;
          uint8_t* sb = st + Cfg::PLANES * Cfg::A_TILE + p * Cfg::B_TILE;
          if (A_MN) {
            // stored [K, M]: box = 64 (m) x 64 (k); two boxes cover 128 m
            ptx::tma_load_2d(sa, ta, &full_bar[stage], m0, k0);
            ptx::tma_load_2d(sa + 8192, ta, &full_bar[stage], m0 + 64, k0);
          } else {
            // stored [M, K]: box = 64 (k) x 128 (m)
            ptx::tma_load_2d(sa, ta, &full_bar[stage], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              ptx::tma_load_2d(sb + j * 8192, tb, &full_bar[stage], n0 + 64 * j, k0);
          } else {
            ptx::tma_load_2d(sb, tb, &full_bar[stage], k0, n0);
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
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, A_MN, B_MN);
      // K-major: 8-row groups are 1024 B apart (SBO); a K=16 step advances 32 B inside the row.
      // MN-major: 8-k groups are 1024 B apart (SBO); 64-element MN chunks are 8192 B apart (LBO);
      //           a K=16 step advances two 8-k groups = 2048 B.
      constexpr uint32_t A_LBO = A_MN ? 8192 : 16, A_STEP = A_MN ? 2048 : 32;
      constexpr uint32_t B_LBO = B_MN ? 8192 : 16, B_STEP = B_MN ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; kb += KBS) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t st = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
        if (ptx::elect_one()) {
#pragma unroll
        for (int i = 0; i < KBS; ++i) {
        if (KBS > 1 && kb + i >= num_kb) break;   // (a partial last group: nothing was loaded beyond the last k-block)
        const uint32_t sa_hi = st + i * Cfg::A_TILE, sa_lo = st + Cfg::A_TILE;
        const uint32_t sb_hi = st + KBS * Cfg::PLANES * Cfg::A_TILE + i * Cfg::B_TILE, sb_lo = sb_hi + Cfg::B_TILE;
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint64_t da_hi = ptx::make_smem_desc_sw128(sa_hi + kk * A_STEP, A_LBO, 1024);
          const uint64_t db_hi = ptx::make_smem_desc_sw128(sb_hi + kk * B_STEP, B_LBO, 1024);
          ptx::umma_f16(tmem_base, da_hi, db_hi, idesc, (kb | i | kk) != 0);
          if (SPLIT == 3) {
            const uint64_t da_lo = ptx::make_smem_desc_sw128(sa_lo + kk * A_STEP, A_LBO, 1024);
            const uint64_t db_lo = ptx::make_smem_desc_sw128(sb_lo + kk * B_STEP, B_LBO, 1024);
            ptx::umma_f16(tmem_base, da_hi, db_lo, idesc, 1);
            ptx::umma_f16(tmem_base, da_lo, db_hi, idesc, 1);
          }
        }
        }
        ptx::umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs retire
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (ptx::elect_one()) ptx::umma_commit(tmem_full_bar);  // accumulator complete
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int q = warp & 3;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    // all MMAs have retired => pipeline smem is free: reuse it as staging
    float* stg = reinterpret_cast<float*>(smem) + q * 32 * Cfg::STG_LD;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, v);
      ptx::tmem_ld_wait();
      float4* dst = reinterpret_cast<float4*>(stg + lane * Cfg::STG_LD + c);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    }
    __syncwarp();
    // coalesced write-out: the warp walks its 32 rows; each lane owns NV fixed groups of 4 columns.
    // Rows go in batches of RB: all global loads of a batch (addend) are issued before its stores, so
    // the load latencies overlap instead of serialising row by row (the output may alias the addend
    // element-for-element, which keeps the compiler from reordering them on its own).
    constexpr int NV = (BN + 127) / 128;          // float4 groups per lane per row
    constexpr int RB = (NV == 1) ? 16 : 8;        // rows per batch
    float4 bias_v[NV];
    bool col_ok[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c4 = lane * 4 + v * 128;
      const int col = n0 + c4;
      col_ok[v] = (c4 < BN) && (col < N);  // N % 4 == 0 is enforced by the launcher
      bias_v[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col_ok[v] && ep.bias) bias_v[v] = *reinterpret_cast<const float4*>(ep.bias + col);
    }
    const int row0 = m0 + q * 32;
#pragma unroll 1
    for (int rb = 0; rb < 32; rb += RB) {
      if (row0 + rb >= M) break;
      float4 x[RB][NV];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int row = row0 + rb + r;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c4 = lane * 4 + v * 128;
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col_ok[v]) {
            t = *reinterpret_cast<const float4*>(stg + (rb + r) * Cfg::STG_LD + c4);
            t.x += bias_v[v].x; t.y += bias_v[v].y; t.z += bias_v[v].z; t.w += bias_v[v].w;
            if (ep.addend && row < M) {
              const float4 a = *(reinterpret_cast<const float4*>(
                  ep.addend + static_cast<long long>(row) * ep.ld_addend + n0 + c4));
              t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
            }
          }
          x[r][v] = t;
        }
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int row = row0 + rb + r;
        if (row >= M) break;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (!col_ok[v]) continue;
          const int col = n0 + lane * 4 + v * 128;
          const float4 t = x[r][v];
          if (ep.out_f32)
            *reinterpret_cast<float4*>(ep.out_f32 + static_cast<long long>(row) * ep.ld_f32 + col) = t;
          if (ep.out_hi) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(t.x), h1 = __float2bfloat16_rn(t.y),
                                h2 = __float2bfloat16_rn(t.z), h3 = __float2bfloat16_rn(t.w);
            const long long o = static_cast<long long>(row) * ep.ld_bf + col;
            __nv_bfloat162 p0(h0, h1), p1(h2, h3);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&p0);
            pk.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(ep.out_hi + o) = pk;
            if (ep.out_lo) {
              __nv_bfloat162 q0(__float2bfloat16_rn(t.x - __bfloat162float(h0)),
                                __float2bfloat16_rn(t.y - __bfloat162float(h1)));
              __nv_bfloat162 q1(__float2bfloat16_rn(t.z - __bfloat162float(h2)),
                                __float2bfloat16_rn(t.w - __bfloat162float(h3)));
              pk.x = *reinterpret_cast<uint32_t*>(&q0);
              pk.y = *reinterpret_cast<uint32_t*>(&q1);
              *reinterpret_cast<uint2*>(ep.out_lo + o) = pk;
            }
          }
        }
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

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D tensor map: inner (contiguous) extent `inner`, outer extent `outer`, row pitch in elements
// kind: 0 = bf16 / 128-byte swizzle (MMA operands), 1 = fp32 / 128-byte swizzle, 2 = bf16 / no swizzle (stores)
bool make_tmap(CUtensorMap* tm, const void* base, int kind, uint64_t inner, uint64_t outer,
               uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  const uint64_t esz = kind == 1 ? 4 : 2;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * esz};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  kind == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct TmapKey {
  const void* base;
  uint64_t inner, outer, pitch;
  uint32_t bi, bo;
  int kind;
  bool operator==(const TmapKey& o) const {
    return base == o.base && inner == o.inner && outer == o.outer && pitch == o.pitch &&
           bi == o.bi && bo == o.bo && kind == o.kind;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    h = h * 1000003u ^ k.inner;
    h = h * 1000003u ^ k.outer;
    h = h * 1000003u ^ k.pitch;
    h = h * 1000003u ^ (static_cast<size_t>(k.bi) << 16 | k.bo);
    h = h * 1000003u ^ static_cast<size_t>(k.kind);
    return h;
  }
};

}  // namespace

// descriptors are pure functions of (pointer, shape, box, kind): cache them (encode costs ~1 us each)
bool cached_tmap_kind(CUtensorMap* out, const void* base, int kind, uint64_t inner, uint64_t outer, uint64_t pitch,
                      uint32_t bi, uint32_t bo) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{base, inner, outer, pitch, bi, bo, kind};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return true;
  }
  CUtensorMap tm;
  if (!make_tmap(&tm, base, kind, inner, outer, pitch, bi, bo)) return false;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, tm);
  *out = tm;
  return true;
}
bool cached_tmap(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch,
                 uint32_t bi, uint32_t bo) {
  return cached_tmap_kind(out, base, 0, inner, outer, pitch, bi, bo);
}

// K-major bf16 operand [rows, K] seen as {64 k, rows, K / 64 k-blocks} (strides: pitch, 128 bytes), box
// {64, box_rows, box_kb}: ONE TMA operation brings box_kb consecutive k-block tiles, each in the 128-byte-swizzled
// layout the MMA descriptors expect, back to back. An SM completes only ~4 TMA operations per microsecond whatever
// their size (scripts/probes/tma_ingest.cu: 8 KB -> 33 GB/s, 32 KB -> 124 GB/s, 64 KB -> 141 GB/s per SM), so
// operand tiles have to arrive in few, large operations.
// MN-major bf16 operand stored [K, MN] (MN contiguous) seen as {64 mn, 64 k, MN / 64 mn-blocks, K / 64 k-blocks}
// (strides: pitch, 128 bytes, 64 * pitch), box {64, 64, box_mnb, box_kb}: ONE operation brings box_kb k-block tiles,
// each made of box_mnb swizzled [64 k x 64 mn] blocks 8 KB apart -- the layout the MN-major MMA descriptors use.
// Requires K % 64 == 0 (nothing clips k inside a block); mn beyond MN reads neighbouring data, which only feeds
// output rows / columns that are never stored.
bool cached_tmap_mnblocks(CUtensorMap* out, const void* base, uint64_t MN, uint64_t K, uint64_t pitch,
                          uint32_t box_mnb, uint32_t box_kb) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{base, MN, K, pitch, box_mnb, box_kb, 4};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return true;
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {64, 64, (MN + 63) / 64, K / 64};
  cuuint64_t strides[3] = {pitch * 2, 128, 64 * pitch * 2};
  cuuint32_t box[4] = {64, 64, box_mnb, box_kb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap tm;
  if (fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, tm);
  *out = tm;
  return true;
}

bool cached_tmap_kblocks(CUtensorMap* out, const void* base, uint64_t K, uint64_t rows, uint64_t pitch,
                         uint32_t box_rows, uint32_t box_kb) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{base, K, rows, pitch, box_rows, box_kb, 3};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return true;
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {64, rows, (K + 63) / 64};
  cuuint64_t strides[2] = {pitch * 2, 128};
  cuuint32_t box[3] = {64, box_rows, box_kb};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap tm;
  if (fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, tm);
  *out = tm;
  return true;
}

namespace {

template <int BN, int SPLIT, bool A_MN, bool B_MN, int KBS = 1>
cudaError_t launch_one(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi,
                       const CUtensorMap& tb_lo, const EpilogueArgs& ep, int M, int N, int K,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN, SPLIT, KBS>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, SPLIT, A_MN, B_MN, KBS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
  return launch_pdl(kern, grid, dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, ta_hi, ta_lo, tb_hi, tb_lo, ep, M, N, K);
}

template <int BN, int SPLIT, int KBS = 1>
cudaError_t launch_major(bool a_mn, bool b_mn, const CUtensorMap& ta_hi, const CUtensorMap& ta_lo,
                         const CUtensorMap& tb_hi, const CUtensorMap& tb_lo, const EpilogueArgs& ep,
                         int M, int N, int K, cudaStream_t s) {
  if (a_mn) {
    if (b_mn) return launch_one<BN, SPLIT, true, true, KBS>(ta_hi, ta_lo, tb_hi, tb_lo, ep, M, N, K, s);
    return launch_one<BN, SPLIT, true, false, KBS>(ta_hi, ta_lo, tb_hi, tb_lo, ep, M, N, K, s);
  }
  if (b_mn) return launch_one<BN, SPLIT, false, true, KBS>(ta_hi, ta_lo, tb_hi, tb_lo, ep, M, N, K, s);
  return launch_one<BN, SPLIT, false, false, KBS>(ta_hi, ta_lo, tb_hi, tb_lo, ep, M, N, K, s);
}

}  // namespace

int pick_block_n(int M, int N, int num_sms) {
  // fill the machine first: take the widest tile that still yields >= one CTA per SM
  const long long tm = (M + BM - 1) / BM;
  for (int bn : {256, 128}) {
    const long long tiles = tm * ((N + bn - 1) / bn);
    if (tiles >= num_sms) return bn;
  }
  return 64;
}

VqaStatus gemm_launch(const VqaGemmDesc& d, int num_sms, cudaStream_t stream, GemmCtx* ctx, int narrow_flags) {
  const int narrow = narrow_flags & 1;
  static const bool b_early_env = getenv("VQA_GEMM_B_EARLY") == nullptr || atoi(getenv("VQA_GEMM_B_EARLY")) != 0;
  const bool b_early = (narrow_flags & 2) != 0 && b_early_env;   // bit 1: stable B operand (see EpilogueArgs::b_early)
  if (!d.a_hi || !d.b_hi) return set_error(VQA_ERR_BAD_ARG, "vqa_gemm: null operand");
  if (d.M <= 0 || d.N <= 0 || d.K <= 0) return set_error(VQA_ERR_BAD_SHAPE, "vqa_gemm: empty problem");
  if ((d.N & 3) || (d.lda & 7) || (d.ldb & 7))
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_gemm: N %% 4, lda %% 8, ldb %% 8 must be 0");
  if ((d.out_f32 && (d.ld_f32 & 3)) || (d.out_hi && (d.ld_bf & 3)) || (d.addend && (d.ld_addend & 3)))
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_gemm: output pitches must be multiples of 4");
  const bool split = (d.a_lo != nullptr) || (d.b_lo != nullptr);
  if (split && (!d.a_lo || !d.b_lo))
    return set_error(VQA_ERR_BAD_ARG, "vqa_gemm: split precision needs both lo planes");
  static const int m512_bn = getenv("VQA_M512_BN") ? atoi(getenv("VQA_M512_BN")) : 0;   // experiment: head GEMMs on the pair kernel
  if (m512_bn < 0 && d.block_n == 0 && !split && d.M >= 256 && d.M <= 512 && d.out_f32 && !d.out_hi && d.K >= 1024 && d.N >= 512) {
    VqaGemmDesc f = d;
    f.block_n = m512_bn;
    int pbn = 0, psplits = 1;
    if (gemm_pair_plan(f, num_sms, ctx, narrow, &pbn, &psplits)) return gemm_pair_launch(f, num_sms, pbn, psplits, ctx, stream);
  }
  if (d.block_n <= 0) {
    int pbn = 0, psplits = 1;
    if (gemm_pair_plan(d, num_sms, ctx, narrow, &pbn, &psplits)) return gemm_pair_launch(d, num_sms, pbn, psplits, ctx, stream);
    if (d.block_n < 0) return set_error(VQA_ERR_BAD_ARG, "vqa_gemm: the CTA-pair kernel takes one bf16 plane per operand and block_n -128 / -256");
  }
  int bn = d.block_n ? d.block_n : pick_block_n(d.M, d.N, num_sms);
  if (split && bn == 256) bn = 128;
  // latency-bound shapes (a wave or less of tiles): multi-k-block stages, 4 k-blocks per TMA operation with
  // 128 x 64 tiles, 2 with 128 x 128 tiles when 64-wide tiles would need a second wave
  static const bool big_off = getenv("VQA_GEMM_SMALL_BOXES") != nullptr;
  int kbs = 1;
  if (!split && !d.block_n && !big_off && d.K >= 256 && ((d.K % 64) == 0 || (!d.a_mn_major && !d.b_mn_major))) {
    const long long tm = (d.M + BM - 1) / BM;
    const long long t64 = tm * ((d.N + 63) / 64), t128 = tm * ((d.N + 127) / 128);
    if (t64 <= num_sms) { bn = 64; kbs = 4; }
    else if (t128 <= num_sms) { bn = 128; kbs = 2; }
  }
  if (bn != 64 && bn != 128 && bn != 256) return set_error(VQA_ERR_BAD_ARG, "vqa_gemm: block_n");

  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  bool ok = true;
  auto map_a = [&](CUtensorMap* tm, const void* p) {
    return d.a_mn_major ? cached_tmap(tm, p, d.M, d.K, d.lda, 64, 64)
                        : cached_tmap(tm, p, d.K, d.M, d.lda, 64, BM);
  };
  auto map_b = [&](CUtensorMap* tm, const void* p) {
    return d.b_mn_major ? cached_tmap(tm, p, d.N, d.K, d.ldb, 64, 64)
                        : cached_tmap(tm, p, d.K, d.N, d.ldb, 64, bn);
  };
  ok = ok && map_a(&ta_hi, d.a_hi) && map_b(&tb_hi, d.b_hi);
  if (split) {
    ok = ok && map_a(&ta_lo, d.a_lo) && map_b(&tb_lo, d.b_lo);
  } else if (kbs > 1) {
    ok = ok && (d.a_mn_major ? cached_tmap_mnblocks(&ta_lo, d.a_hi, d.M, d.K, d.lda, 2, kbs)
                             : cached_tmap_kblocks(&ta_lo, d.a_hi, d.K, d.M, d.lda, BM, kbs)) &&
         (d.b_mn_major ? cached_tmap_mnblocks(&tb_lo, d.b_hi, d.N, d.K, d.ldb, bn / 64, kbs)
                       : cached_tmap_kblocks(&tb_lo, d.b_hi, d.K, d.N, d.ldb, bn, kbs));
  } else {
    ta_lo = ta_hi;
    tb_lo = tb_hi;
  }
  if (!ok) return set_error(VQA_ERR_CUDA, "vqa_gemm: cuTensorMapEncodeTiled failed");

  EpilogueArgs ep;
  ep.bias = d.bias;
  ep.addend = d.addend;
  ep.ld_addend = d.ld_addend;
  ep.out_f32 = d.out_f32;
  ep.ld_f32 = d.ld_f32;
  ep.out_hi = static_cast<__nv_bfloat16*>(d.out_hi);
  ep.out_lo = static_cast<__nv_bfloat16*>(d.out_lo);
  ep.ld_bf = d.ld_bf;
  ep.b_early = (b_early && kbs > 1) ? 1 : 0;

  cudaError_t e;
  const bool amn = d.a_mn_major != 0, bmn = d.b_mn_major != 0;
  if (split) {
    if (bn == 64) e = launch_major<64, 3>(amn, bmn, ta_hi, ta_lo, tb_hi, tb_lo, ep, d.M, d.N, d.K, stream);
    else e = launch_major<128, 3>(amn, bmn, ta_hi, ta_lo, tb_hi, tb_lo, ep, d.M, d.N, d.K, stream);
  } else {
    if (kbs == 4) e = launch_major<64, 1, 4>(amn, bmn, ta_hi, ta_lo, tb_hi, tb_lo, ep, d.M, d.N, d.K, stream);
    else if (kbs == 2) e = launch_major<128, 1, 2>(amn, bmn, ta_hi, ta_lo, tb_hi, tb_lo, ep, d.M, d.N, d.K, stream);
    else if (bn == 64) e = launch_major<64, 1>(amn, bmn, ta_hi, ta_lo, tb_hi, tb_lo, ep, d.M, d.N, d.K, stream);
    else if (bn == 128) e = launch_major<128, 1>(amn, bmn, ta_hi, ta_lo, tb_hi, tb_lo, ep, d.M, d.N, d.K, stream);
    else e = launch_major<256, 1>(amn, bmn, ta_hi, ta_lo, tb_hi, tb_lo, ep, d.M, d.N, d.K, stream);
  }
  if (e != cudaSuccess) return set_cuda_error(e, "vqa_gemm launch");
  count_launch();
  return VQA_OK;
}

}  // namespace vqa
