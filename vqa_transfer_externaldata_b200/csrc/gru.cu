// Persistent recurrent kernels for the GRU question encoder (vlmap/modules.py:124-140: GRUCell under
// dynamic_rnn) -- forward over T steps and back-propagation through time -- each as ONE cooperative launch.
//
// A step is two DEPENDENT [B, L] x [L, *] matmuls with element-wise gate math in between. Design:
//   * unit-slice ownership: CTA (s, m) owns hidden units [32 s, 32 s + 32) for batch rows [128 m, 128 m + 128).
//     Everything element-wise about a unit (r, u, c, h and their gradients) is computed by the same lane of
//     the same warp in every step, so the recurrent state (h, u / dh, du) lives in REGISTERS for the whole
//     sequence and never round-trips through memory.
//   * the CTA's weight slice (96 rows x L, bf16, 192 KB at L = 1024) is loaded into shared memory ONCE and
//     stays resident; per phase only the activation operand (h_t / r.h / dC / dG, produced by the other CTAs
//     of the same row tile) streams in through a TMA ring, so L2 traffic per phase is the A operand only.
//   * per phase: TMA (128B-swizzled tiles) -> tcgen05.mma (M 128, N 64 | 32, bf16, fp32 accumulate in TMEM)
//     -> tcgen05.ld -> swizzled shared-memory transpose -> gate math with lane = unit (all global accesses
//     coalesced, operands of the epilogue prefetched before the MMA finishes) -> release on a per-row-tile
//     counter. Row tiles are independent: only the CTAs sharing rows synchronise.
//   * bias gradients are accumulated in registers across all steps (no fp32 copies of dG / dC exist).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "internal.h"
#include "ptx.cuh"

namespace vqa {

unsigned long long* g_gru_trace = nullptr;   // debugging aid, see vqa_internal_set_gru_trace

namespace {

constexpr int R_BM = 128;   // batch rows per CTA (UMMA M)
constexpr int R_JN = 32;    // hidden units per CTA
constexpr int R_BK = 64;    // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int R_STAGES = 4;
constexpr int R_A_TILE = R_BM * R_BK * 2;      // 16 KB activation tile
constexpr int R_W_TILE = R_JN * R_BK * 2;      // 4 KB streamed weight tile (32 rows)
constexpr int R_STAGE = R_A_TILE + R_W_TILE;   // 20 KB
constexpr int R_RING = R_STAGES * R_STAGE;     // 80 KB: the ring, reused as the epilogue's transpose buffer
constexpr int R_EPI_WARPS = 16;                // 4 per TMEM lane quarter, 8 rows each
constexpr int R_THREADS = 64 + 32 * R_EPI_WARPS;
constexpr int R_TMEM_COLS = 128;               // gates accumulator at column 0 (64 wide), candidate at 64 (32 wide)

// resident weights: forward = the gate columns (64 rows x L), BPTT = the Wg rows (32 rows x 2L): 128 L bytes
__host__ __device__ constexpr int wres_bytes(int L) { return 2 * R_JN * L * 2; }
__host__ __device__ constexpr int smem_bytes(int L) { return wres_bytes(L) + R_RING + 128 + 1024; }

struct GruArgs {
  int B;        // rows per time block (row stride between steps)
  int row0;     // first batch row of this launch
  int row_end;  // one past the last batch row of this launch
  int L, T;
  const int* q_len;
  unsigned int* counter;   // [gridDim.y] phase counters, zeroed before launch
  // forward
  const float* xg;   // [T*B, 2L] x-part of the gate pre-activations (+bias)
  const float* xc;   // [T*B, L]
  int x_bf;          // xg / xc hold bf16 values (same element indexing)
  float* h_f32;      // [(T+1)*B, L]
  bf16* h_bf;        // [(T+1)*B, L]
  bf16* rh_bf;       // [T*B, L]
  float* r; float* u; float* c;  // [T*B, L]
  // backward
  const float* dq;   // [B, L] gradient of the final state
  const float* dq2;  // optional second addend
  bf16* dG_bf;       // [T*B, 2L]
  bf16* dC_bf;       // [T*B, L]
  float* bias_part;  // [ceil(B/128), 3L] per-row-tile partial sums of (d gates_bias | d candidate_bias)
  unsigned long long* trace;  // optional [num_ctas, num_phases, 4] globaltimer stamps (scripts/gpu_gru_trace.py)
};

// one hoisted x-projection value: fp32 array, or the same array holding bf16 values
__device__ __forceinline__ float ldx(const float* base, long long idx, int is_bf) {
  if (!is_bf) return __ldg(base + idx);
  return __bfloat162float(__ldg(reinterpret_cast<const bf16*>(base) + idx));
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define GRU_TRACE(p, k)                                                                                   \
  do {                                                                                                    \
    if (g.trace)                                                                                          \
      g.trace[((static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * num_phases + (p)) * 8 + (k)] = \
          gtimer();                                                                                       \
  } while (0)

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded spin: a lost arrival becomes a trap (launch failure), never a hung GPU
__device__ __forceinline__ void wait_counter(const unsigned int* p, unsigned int target) {
  if (ld_acquire_u32(p) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_u32(p) < target) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_all() {  // all epilogue warps
  asm volatile("bar.sync 1, %0;" ::"n"(32 * R_EPI_WARPS) : "memory");
}
__device__ __forceinline__ void epi_bar_quarter(int q) {  // the 4 warps sharing a TMEM lane quarter
  asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
}
// the persistent kernels exist in bf16 mode only: MUFU.TANH (2^-11 relative) is far inside its operand rounding
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_relaxed_add(unsigned int* p, unsigned int v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Transpose buffer of one lane quarter: 32 rows x NCH 16-byte chunks of fp32; chunk index XOR (row & 7) keeps
// both the row-per-lane float4 writes and the unit-per-lane scalar reads bank-conflict free.
// stage_cols<NC>: this warp's 32 TMEM lanes (rows) x NC consecutive fp32 columns starting at column col0.
template <int NCH, int NC>
__device__ __forceinline__ void stage_cols(uint32_t taddr, float* stg, int lane, int col0) {
  uint32_t v[NC];
  if constexpr (NC == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr + col0)
        : "memory");
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr + col0)
                 : "memory");
  }
  ptx::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < NC / 4; ++j) {
    const int ch = col0 / 4 + j;
    const int phys = (ch & ~7) | ((ch ^ lane) & 7);
    *reinterpret_cast<float4*>(stg + (lane * NCH + phys) * 4) =
        make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                    __uint_as_float(v[4 * j + 3]));
  }
}
template <int NCH>
__device__ __forceinline__ float stg_read(const float* stg, int row, int col) {
  const int ch = col >> 2;
  const int phys = (ch & ~7) | ((ch ^ row) & 7);
  return stg[(row * NCH + phys) * 4 + (col & 3)];
}

// MODE 0: forward.  phase 2t  : G = h_t Wg_h (+xg) -> r, u, r.h        (skipped matmul at t = 0: h_0 = 0)
//                   phase 2t+1: C = (r.h) Wc_h (+xc) -> c, h_{t+1}
// MODE 1: BPTT.     phase 0   : element-wise head of step T-1 from dq -> du, dC_{T-1}, dh_part
//                   phase 1+2i: dRH = dC_t Wc_h^T -> dG_t, dh_part           (t = T-1-i)
//                   phase 2+2i: dh  = dG_t Wg_h^T + dh_part -> head of step t-1   (not run for t = 0)
// Weights: the larger matrix of the pair stays resident in shared memory (forward: the gate columns, used by
// kind 0; BPTT: the Wg rows, used by kind 1); the smaller one (32 rows) streams through the ring next to A.
// CL > 1: the CL CTAs of a cluster own adjacent unit slices of the SAME row tile, so they consume identical
// activation tiles: each loads 1/CL of every tile and multicasts it to all (L2 reads of A drop CL-fold; the
// kernel is bound by exactly that traffic). Stage hand-back is cluster-wide: every consumer's tcgen05.commit
// arrives on the empty barrier of every CTA.
template <int MODE, int CL>
__global__ void __launch_bounds__(R_THREADS, 1) gru_persistent_kernel(
    const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
    const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1, GruArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int L = g.L, T = g.T, B = g.B;
  uint8_t* wres = smem;                       // resident weight tiles
  uint8_t* ring = smem + wres_bytes(L);       // ring / transpose buffer
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + R_RING);
  uint64_t* empty_bar = full_bar + R_STAGES;
  uint64_t* tmem_full_bar = empty_bar + R_STAGES;
  uint64_t* w_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x, mi = blockIdx.y;
  const int j0 = slice * R_JN;
  const int m0 = g.row0 + mi * R_BM;
  const unsigned int nslices = gridDim.x;
  unsigned int* counter = g.counter + mi;
  const int KB = L / R_BK;  // k-blocks over L
  constexpr int RES_KIND = (MODE == 0) ? 0 : 1;  // the phase kind whose weights are resident
  const int num_phases = 2 * T;                  // BPTT: 1 head + 2T - 1 matmul phases

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a0);
    ptx::prefetch_tensormap(&tm_a1);
    ptx::prefetch_tensormap(&tm_w0);
    ptx::prefetch_tensormap(&tm_w1);
    for (int s = 0; s < R_STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], CL);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::mbar_init(w_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, R_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();  // peers' barriers exist before anything is multicast at them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = (CL > 1) ? ptx::cluster_ctarank() : 0;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1);

  // phase p -> (has matmul, which operand / weight set, time step)
  auto phase_info = [&](int p, bool& mm, int& kind, int& t) {
    if (MODE == 0) {
      t = p >> 1;
      kind = p & 1;
      mm = t > 0;  // h_0 = 0: both products of step 0 vanish
    } else {
      if (p == 0) { mm = false; kind = 1; t = T; return; }  // head only; "t" = T means dh comes from dq
      const int i = (p - 1) >> 1;
      t = T - 1 - i;
      kind = (p - 1) & 1;
      mm = true;
    }
  };

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    {
      // resident weights, once
      if (lane == 0) {
        ptx::mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(wres_bytes(L)));
        if (MODE == 0) {
          for (int kb = 0; kb < KB; ++kb) ptx::tma_load_2d(wres + kb * 8192, &tm_w0, w_bar, kb * R_BK, slice * 96);
        } else {
          for (int kb = 0; kb < 2 * KB; ++kb) ptx::tma_load_2d(wres + kb * 4096, &tm_w1, w_bar, kb * R_BK, j0);
        }
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase_bit = 0;
      for (int p = 0; p < num_phases; ++p) {
        bool mm; int kind, t;
        phase_info(p, mm, kind, t);
        if (!mm) continue;
        const bool streamed = kind != RES_KIND;
        const CUtensorMap* ta = kind ? &tm_a1 : &tm_a0;
        const CUtensorMap* tw = (MODE == 0) ? &tm_w1 : &tm_w0;
        const int wrow = (MODE == 0) ? slice * 96 + 64 : j0;
        const int nkb = (MODE == 1 && kind == 1) ? 2 * KB : KB;
        const int arow = t * B + m0;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase_bit ^ 1);
          uint8_t* st = ring + stage * R_STAGE;
          if (kb == 0) {
            // the A operand of this phase was written by the epilogues of phase p-1 (all CTAs of this row
            // tile); the ring itself doubles as their transpose buffer, so nothing may land in it earlier
            wait_counter(counter, static_cast<unsigned int>(p) * nslices);
            fence_proxy_async_all();
            if (lane == 0) GRU_TRACE(p, 0);
            __syncwarp();
          }
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], streamed ? R_STAGE : R_A_TILE);
            if (streamed) ptx::tma_load_2d(st + R_A_TILE, tw, &full_bar[stage], kb * R_BK, wrow);
            if (CL > 1)
              ptx::tma_load_2d_multicast(st + crank * (R_A_TILE / CL), ta, &full_bar[stage], kb * R_BK,
                                         arow + static_cast<int>(crank) * (R_BM / CL), kMask);
            else
              ptx::tma_load_2d(st, ta, &full_bar[stage], kb * R_BK, arow);
          }
          __syncwarp();
          if (++stage == R_STAGES) {
            stage = 0;
            phase_bit ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    {
      constexpr uint32_t idesc64 = ptx::make_idesc_bf16(R_BM, 64, false, false);
      constexpr uint32_t idesc32 = ptx::make_idesc_bf16(R_BM, 32, false, false);
      ptx::mbar_wait(w_bar, 0);
      int stage = 0;
      uint32_t phase_bit = 0;
      for (int p = 0; p < num_phases; ++p) {
        bool mm; int kind, t;
        phase_info(p, mm, kind, t);
        if (!mm) continue;
        const bool streamed = kind != RES_KIND;
        const bool wide = (MODE == 0 && kind == 0);  // N = 64 (r | u columns)
        const uint32_t idesc = wide ? idesc64 : idesc32;
        const int nkb = (MODE == 1 && kind == 1) ? 2 * KB : KB;
        const uint32_t wtile = wide ? 8192 : 4096;
        const uint32_t d_tmem = tmem_base + ((MODE == 0 && kind == 1) ? 64 : 0);
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase_bit);
          ptx::tc_fence_after();
          if (kb == 0 && lane == 0) GRU_TRACE(p, 1);
          const uint32_t sa = ptx::smem_u32(ring + stage * R_STAGE);
          const uint32_t sb = streamed ? sa + R_A_TILE : ptx::smem_u32(wres) + kb * wtile;
          if (ptx::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < R_BK / 16; ++kk) {
              const uint64_t da = ptx::make_smem_desc_sw128(sa + kk * 32, 16, 1024);
              const uint64_t db = ptx::make_smem_desc_sw128(sb + kk * 32, 16, 1024);
              ptx::umma_f16(d_tmem, da, db, idesc, (kb | kk) != 0);
            }
            if (CL > 1) ptx::umma_commit_multicast(&empty_bar[stage], kMask);
            else ptx::umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == R_STAGES) {
            stage = 0;
            phase_bit ^= 1;
          }
        }
        if (ptx::elect_one()) ptx::umma_commit(tmem_full_bar);
        __syncwarp();
        // the accumulator is overwritten only after the next counter wait, which this CTA's own epilogue
        // reaches after draining TMEM: no tmem_empty barrier needed
      }
    }
  } else {
    // ===================== epilogue warps: lane = hidden unit, 8 rows per warp =========================
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int sub = (warp - 2) >> 2;              // which 8 rows of the quarter / which column group to stage
    const int rbase = m0 + q * 32 + sub * 8;      // first batch row of this warp
    const int unit = j0 + lane;
    float* stg = reinterpret_cast<float*>(ring) + q * (32 * 64);  // 8 KB per quarter
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int my_row = rbase + (lane & 7);
    const int my_len = my_row < g.row_end ? g.q_len[my_row] : 0;   // lanes 0..7 hold the lengths of the 8 rows
    const bool leader = threadIdx.x == 64;
    uint32_t tfull_phase = 0;
    constexpr int NR = 8;

    if (MODE == 0) {
      float h[NR], u[NR];
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) h[rr] = 0.f, u[rr] = 0.f;
      for (int p = 0; p < num_phases; ++p) {
        bool mm; int kind, t;
        phase_info(p, mm, kind, t);
        const long long tb = static_cast<long long>(t) * B;
        float sv[NR];   // r (kind 0) or c (kind 1) of this phase
        if (kind == 0) {
          float xr[NR], xu[NR], ar[NR], au[NR];
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            const int row = rbase + rr;
            xr[rr] = xu[rr] = ar[rr] = au[rr] = 0.f;
            if (row < g.row_end) {
              const long long xo = (tb + row) * 2 * L + unit;
              xr[rr] = ldx(g.xg, xo, g.x_bf);
              xu[rr] = ldx(g.xg, xo + L, g.x_bf);
            }
          }
          if (mm) {
            ptx::mbar_wait(tmem_full_bar, tfull_phase);
            tfull_phase ^= 1;
            ptx::tc_fence_after();
            if (leader) GRU_TRACE(p, 2);
            stage_cols<16, 16>(t_lane, stg, lane, sub * 16);
            ptx::tc_fence_before();
            epi_bar_quarter(q);
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) {
              ar[rr] = stg_read<16>(stg, sub * 8 + rr, lane);
              au[rr] = stg_read<16>(stg, sub * 8 + rr, 32 + lane);
            }
            if (leader) GRU_TRACE(p, 4);
          }
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            sv[rr] = sigm(ar[rr] + xr[rr]);
            u[rr] = sigm(au[rr] + xu[rr]);
          }
          // the operand the other CTAs wait for goes out first; r / u (kept for BPTT) after the arrival
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            const int row = rbase + rr;
            if (row < g.row_end) g.rh_bf[(tb + row) * L + unit] = __float2bfloat16_rn(sv[rr] * h[rr]);
          }
          if (leader) GRU_TRACE(p, 5);
        } else {
          float xc[NR], ac[NR];
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            const int row = rbase + rr;
            ac[rr] = 0.f;
            xc[rr] = (row < g.row_end) ? ldx(g.xc, (tb + row) * L + unit, g.x_bf) : 0.f;
          }
          if (mm) {
            ptx::mbar_wait(tmem_full_bar, tfull_phase);
            tfull_phase ^= 1;
            ptx::tc_fence_after();
            if (leader) GRU_TRACE(p, 2);
            stage_cols<8, 8>(t_lane + 64, stg, lane, sub * 8);
            ptx::tc_fence_before();
            epi_bar_quarter(q);
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) ac[rr] = stg_read<8>(stg, sub * 8 + rr, lane);
          }
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            sv[rr] = tanh_fast(ac[rr] + xc[rr]);
            const bool valid = t < __shfl_sync(0xffffffffu, my_len, rr);
            h[rr] = valid ? u[rr] * h[rr] + (1.0f - u[rr]) * sv[rr] : h[rr];
          }
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            const int row = rbase + rr;
            if (row < g.row_end) g.h_bf[(tb + B + row) * L + unit] = __float2bfloat16_rn(h[rr]);
          }
        }
        // publish (the pattern of a grid barrier: CTA barrier, then ONE thread fences and releases): the
        // stores of all epilogue threads become visible device-wide and to the async proxy (TMA) of other SMs
        epi_bar_all();
        if (leader) {
          GRU_TRACE(p, 6);
          fence_acq_rel_gpu();
          fence_proxy_async_all();
          GRU_TRACE(p, 7);
          red_relaxed_add(counter, 1u);
          GRU_TRACE(p, 3);
        }
        // what only BPTT reads goes out off the critical path
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
          const int row = rbase + rr;
          if (row < g.row_end) {
            const long long o = (tb + row) * L + unit;
            if (kind == 0) {
              g.r[o] = sv[rr];
              g.u[o] = u[rr];
            } else {
              g.c[o] = sv[rr];
              g.h_f32[o + static_cast<long long>(B) * L] = h[rr];
            }
          }
        }
      }
    } else {
      float dhp[NR], du[NR];   // dh_part, du of the current step, rows of this warp
      float db_r = 0.f, db_u = 0.f, db_c = 0.f;
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) dhp[rr] = 0.f, du[rr] = 0.f;
      for (int p = 0; p < num_phases; ++p) {
        bool mm; int kind, t;
        phase_info(p, mm, kind, t);
        if (kind == 0) {
          // dRH = acc ;  dG_r = dRH h r (1-r) ; dG_u = du u (1-u) ; dh_part += dRH r
          const long long tb = static_cast<long long>(t) * B;
          float hh[NR], rv[NR], uv[NR], acc[NR];
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            const int row = rbase + rr;
            hh[rr] = rv[rr] = uv[rr] = 0.f;
            if (row < g.row_end) {
              const long long o = (tb + row) * L + unit;
              hh[rr] = __ldg(g.h_f32 + o);
              rv[rr] = __ldg(g.r + o);
              uv[rr] = __ldg(g.u + o);
            }
          }
          ptx::mbar_wait(tmem_full_bar, tfull_phase);
          tfull_phase ^= 1;
          ptx::tc_fence_after();
          if (leader) GRU_TRACE(p, 2);
          stage_cols<8, 8>(t_lane, stg, lane, sub * 8);
          ptx::tc_fence_before();
          epi_bar_quarter(q);
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) acc[rr] = stg_read<8>(stg, sub * 8 + rr, lane);
          float dgr[NR], dgu[NR];
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            const float drh = acc[rr];   // zero for steps beyond the question length (dC is zero there)
            dgr[rr] = drh * hh[rr] * rv[rr] * (1.0f - rv[rr]);
            dgu[rr] = du[rr] * uv[rr] * (1.0f - uv[rr]);
            dhp[rr] = fmaf(drh, rv[rr], dhp[rr]);
          }
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            const int row = rbase + rr;
            if (row < g.row_end) {
              db_r += dgr[rr];
              db_u += dgu[rr];
              const long long o = (tb + row) * 2 * L + unit;
              g.dG_bf[o] = __float2bfloat16_rn(dgr[rr]);
              g.dG_bf[o + L] = __float2bfloat16_rn(dgu[rr]);
            }
          }
        } else {
          // dh = acc + dh_part (or dq) ; element-wise head of step tp = t - 1
          const int tp = t - 1;
          const long long tb = static_cast<long long>(tp) * B;
          float hh[NR], uv[NR], cv[NR], acc[NR];
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            const int row = rbase + rr;
            hh[rr] = uv[rr] = cv[rr] = acc[rr] = 0.f;
            if (row < g.row_end) {
              const long long o = (tb + row) * L + unit;
              hh[rr] = __ldg(g.h_f32 + o);
              uv[rr] = __ldg(g.u + o);
              cv[rr] = __ldg(g.c + o);
              if (!mm) {
                dhp[rr] = __ldg(g.dq + static_cast<long long>(row) * L + unit);
                if (g.dq2) dhp[rr] += __ldg(g.dq2 + static_cast<long long>(row) * L + unit);
              }
            }
          }
          if (mm) {
            ptx::mbar_wait(tmem_full_bar, tfull_phase);
            tfull_phase ^= 1;
            ptx::tc_fence_after();
            if (leader) GRU_TRACE(p, 2);
            stage_cols<8, 8>(t_lane, stg, lane, sub * 8);
            ptx::tc_fence_before();
            epi_bar_quarter(q);
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) acc[rr] = stg_read<8>(stg, sub * 8 + rr, lane);
          }
          float dcv[NR];
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            const float dh = acc[rr] + dhp[rr];
            const bool pvalid = tp < __shfl_sync(0xffffffffu, my_len, rr);
            dcv[rr] = pvalid ? dh * (1.0f - uv[rr]) * (1.0f - cv[rr] * cv[rr]) : 0.f;
            du[rr] = pvalid ? dh * (hh[rr] - cv[rr]) : 0.f;
            dhp[rr] = pvalid ? dh * uv[rr] : dh;
          }
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            const int row = rbase + rr;
            if (row < g.row_end) {
              db_c += dcv[rr];
              g.dC_bf[(tb + row) * L + unit] = __float2bfloat16_rn(dcv[rr]);
            }
          }
        }
        epi_bar_all();
        if (leader) {
          fence_acq_rel_gpu();
          fence_proxy_async_all();
          red_relaxed_add(counter, 1u);
          GRU_TRACE(p, 3);
        }
      }
      // bias gradients: sum the 16 warps of this CTA (fixed order), one partial row per row tile
      float* red = reinterpret_cast<float*>(ring);  // the ring is free now
      const int e = warp - 2;
      red[(e * 3 + 0) * 32 + lane] = db_r;
      red[(e * 3 + 1) * 32 + lane] = db_u;
      red[(e * 3 + 2) * 32 + lane] = db_c;
      epi_bar_all();
      if (e == 0) {
        // two partial rows per row tile (the CTA-pair variant fills both; here the second one is zero)
        float* out = g.bias_part + static_cast<long long>(m0 / R_BM) * 2 * 3 * L;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < R_EPI_WARPS; ++w) s += red[(w * 3 + k) * 32 + lane];
          out[k * L + unit] = s;
          out[3 * L + k * L + unit] = 0.f;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();  // no CTA leaves while peers may still signal its barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, R_TMEM_COLS);
  }
}

// WT_pack[s][rr][k], s = unit slice, k = input unit (K-major B operand of the forward matmuls):
//   rr in [0,32): Wg_h[k, 32 s + rr] (reset gate) ; [32,64): Wg_h[k, L + 32 s + rr - 32] (update gate) ;
//   [64,96): Wc_h[k, 32 s + rr - 64] (candidate).   wg_h / wc_h: bf16 rows W.. of the TF kernels, [L, 2L] / [L, L].
__global__ void gru_pack_weights_kernel(const bf16* __restrict__ wg_h, const bf16* __restrict__ wc_h, int L,
                                        bf16* __restrict__ out) {
  __shared__ bf16 tile[32][33];
  const int s = blockIdx.x;          // slice
  const int part = blockIdx.y;       // 0 r, 1 u, 2 c
  const int k0 = blockIdx.z * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const bf16* src = part == 2 ? wc_h : wg_h;
  const int ld = part == 2 ? L : 2 * L;
  const int col0 = (part == 1 ? L : 0) + 32 * s;
  for (int i = ty; i < 32; i += 8) tile[i][tx] = src[static_cast<long long>(k0 + i) * ld + col0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    out[(static_cast<long long>(s) * 96 + part * 32 + i) * L + k0 + tx] = tile[tx][i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

bool encode_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch,
                 uint32_t bi, uint32_t bo) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return false;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch * 2};
  cuuint32_t box[2] = {bi, bo};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool g_pair_refused = false;
bool g_fwd_was_pair = false;   // the forward recurrence left r / u / c / h_t in the pair kernels' private layout

constexpr int R_CL = 4;       // cluster size of the multicast variant
// Measured on B200 (profiles/r01_gru_trace.md): multicasting the activation tiles inside 4-CTA clusters leaves the
// per-phase main loop at 5.1 us -- the bound is the ~50 GB/s each SM can ingest, not the L2 read rate -- so the
// cluster variant is opt-in (VQA_GRU_CLUSTER=1) and the plain variant is the default.
bool g_cluster_off = getenv("VQA_GRU_CLUSTER") == nullptr;

template <int MODE, int CL>
cudaError_t launch_one(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w0, const CUtensorMap& w1,
                       GruArgs& a, dim3 grid, int smem, cudaStream_t s) {
  auto kern = gru_persistent_kernel<MODE, CL>;
  static int smem_set = 0;
  if (smem_set < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  void* args[] = {const_cast<CUtensorMap*>(&a0), const_cast<CUtensorMap*>(&a1), const_cast<CUtensorMap*>(&w0),
                  const_cast<CUtensorMap*>(&w1), &a};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(R_THREADS);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  // cooperative: fails instead of deadlocking if the grid cannot be co-resident
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = CL;
  at[1].val.clusterDim.y = 1;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = CL > 1 ? 2 : 1;
  return cudaLaunchKernelExC(&cfg, reinterpret_cast<void*>(kern), args);
}

// a0 / a1: the two activation operands ([rows, inner] bf16, row pitch = inner); w0 / w1: weight tensor maps
template <int MODE>
VqaStatus launch_persistent(const void* a0_ptr, uint64_t a0_inner, uint64_t a0_rows, const void* a1_ptr,
                            uint64_t a1_inner, uint64_t a1_rows, const CUtensorMap& w0, const CUtensorMap& w1,
                            GruArgs g, int num_sms, cudaStream_t s) {
  const int smem = smem_bytes(g.L);
  const int slices = g.L / R_JN;
  const int tiles_per_launch = num_sms / slices;  // co-resident row tiles
  const int Bn = g.row_end;
  for (int row0 = 0; row0 < Bn; row0 += tiles_per_launch * R_BM) {
    GruArgs a = g;
    a.trace = g_gru_trace ? g_gru_trace + (MODE ? 1 : 0) * (1 << 17) : nullptr;
    a.row0 = row0;
    a.row_end = Bn < row0 + tiles_per_launch * R_BM ? Bn : row0 + tiles_per_launch * R_BM;
    const int mt = (a.row_end - row0 + R_BM - 1) / R_BM;
    VQA_CUDA_CHECK(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int) * mt, s));
    bool done = false;
    if (!g_cluster_off && slices % R_CL == 0) {
      CUtensorMap a0, a1;
      if (!encode_bf16(&a0, a0_ptr, a0_inner, a0_rows, a0_inner, 64, R_BM / R_CL) ||
          !encode_bf16(&a1, a1_ptr, a1_inner, a1_rows, a1_inner, 64, R_BM / R_CL))
        return set_error(VQA_ERR_CUDA, "gru_persistent: cuTensorMapEncodeTiled failed");
      cudaError_t e = launch_one<MODE, R_CL>(a0, a1, w0, w1, a, dim3(slices, mt), smem, s);
      if (e == cudaSuccess) done = true;
      else {
        cudaGetLastError();  // the cluster variant cannot be scheduled on this device / partition
        g_cluster_off = true;
        if (getenv("VQA_VERBOSE")) fprintf(stderr, "[vqa] GRU cluster launch refused (%s): using the unicast variant\n", cudaGetErrorString(e));
      }
    }
    if (!done) {
      CUtensorMap a0, a1;
      if (!encode_bf16(&a0, a0_ptr, a0_inner, a0_rows, a0_inner, 64, R_BM) ||
          !encode_bf16(&a1, a1_ptr, a1_inner, a1_rows, a1_inner, 64, R_BM))
        return set_error(VQA_ERR_CUDA, "gru_persistent: cuTensorMapEncodeTiled failed");
      VQA_CUDA_CHECK((launch_one<MODE, 1>(a0, a1, w0, w1, a, dim3(slices, mt), smem, s)));
    }
    count_launch();
  }
  return VQA_OK;
}

}  // namespace

// debugging aid (not part of the ABI header): device buffer of 2 x 2^17 u64 receiving per-phase time stamps
extern "C" __attribute__((visibility("default"))) void vqa_internal_set_gru_trace(void* dev_ptr) {
  g_gru_trace = static_cast<unsigned long long*>(dev_ptr);
}

// which recurrent kernels ran last: bit 0 = CTA-pair kernels, bit 1 = single-CTA kernels, bit 8 = a pair launch was
// refused earlier in this process (sticky)
extern "C" __attribute__((visibility("default"))) int32_t vqa_gru_kernel_path(void) {
  return (g_fwd_was_pair ? 1 : 2) | (g_pair_refused ? 256 : 0);
}

bool gru_persistent_supported(int B, int L, int precision, int num_sms) {
  (void)B;
  if (precision != VQA_PREC_BF16) return false;
  if (L % 64 != 0 || L < 64) return false;
  if (smem_bytes(L) > 227 * 1024) return false;  // the weight slice must fit shared memory (L <= 1024)
  return L / R_JN <= num_sms;
}

size_t gru_pack_elems(int L) { return static_cast<size_t>(3) * L * L; }
size_t gru_bias_part_floats(int B, int L) { return static_cast<size_t>(2 * ((B + R_BM - 1) / R_BM) + 2) * 3 * L; }
int gru_bias_part_rows(int B) { return 2 * ((B + R_BM - 1) / R_BM); }

VqaStatus gru_pack_weights_launch(const bf16* wg_h, const bf16* wc_h, int L, bf16* out, cudaStream_t s) {
  gru_pack_weights_kernel<<<dim3(L / 32, 3, L / 32), dim3(32, 8), 0, s>>>(wg_h, wc_h, L, out);
  VQA_LAUNCH_CHECK("gru_pack_weights");
  return VQA_OK;
}

// the CTA-pair kernels (gru_pair.cu) first; if their cluster + cooperative launch is refused on this device /
// partition, the single-CTA kernels below take over for the rest of the process
static bool try_pair(cudaError_t e, const char* what) {
  if (e == cudaSuccess) {
    count_launch();
    return true;
  }
  cudaGetLastError();
  // LOUD: the single-CTA kernels are ~2x slower per phase (profiles/r01_launch_summary_v3.md); a process that lost
  // the pair kernels says so on stderr, in vqa_last_error() and in vqa_gru_kernel_path() (bench.py prints it)
  fprintf(stderr, "[vqa] WARNING %s: CTA-pair recurrent kernel launch refused (%s): falling back to the slower single-CTA "
          "kernels for the rest of this process\n", what, cudaGetErrorString(e));
  set_error(VQA_OK, "%s: CTA-pair recurrent kernel launch refused (%s); single-CTA fallback in use", what, cudaGetErrorString(e));
  g_pair_refused = true;
  return false;
}

VqaStatus gru_fwd_persistent_launch(const GruFwdPersistent& a, int num_sms, cudaStream_t s) {
  const int B = a.B, L = a.L, T = a.T;
  g_fwd_was_pair = false;
  if (!g_pair_refused && gru_pair_supported(B, L, num_sms) && try_pair(gru_pair_fwd(a, num_sms, s), "gru forward")) {
    g_fwd_was_pair = true;
    return VQA_OK;
  }
  CUtensorMap tm_wg, tm_wc;
  const uint64_t prow = static_cast<uint64_t>(L / R_JN) * 96;
  bool ok = encode_bf16(&tm_wg, a.w_pack, L, prow, L, 64, 64) &&
            encode_bf16(&tm_wc, a.w_pack, L, prow, L, 64, 32);
  if (!ok) return set_error(VQA_ERR_CUDA, "gru_fwd_persistent: cuTensorMapEncodeTiled failed");
  GruArgs g{};
  g.B = B; g.row0 = 0; g.row_end = B; g.L = L; g.T = T; g.q_len = a.q_len; g.counter = a.counter;
  g.xg = a.xg; g.xc = a.xc; g.x_bf = a.x_bf16; g.h_f32 = a.h_f32; g.h_bf = a.h_bf; g.rh_bf = a.rh_bf; g.r = a.r; g.u = a.u; g.c = a.c;
  return launch_persistent<0>(a.h_bf, L, static_cast<uint64_t>(T + 1) * B, a.rh_bf, L, static_cast<uint64_t>(T) * B,
                              tm_wg, tm_wc, g, num_sms, s);
}

VqaStatus gru_bwd_persistent_launch(const GruBwdPersistent& a, int num_sms, cudaStream_t s) {
  const int B = a.B, L = a.L, T = a.T;
  if (g_fwd_was_pair) {
    // the saved state is in the pair kernels' layout: only the pair BPTT kernel can read it
    cudaError_t e = gru_pair_bwd(a, num_sms, s);
    if (e != cudaSuccess) return set_cuda_error(e, "gru BPTT (pair) launch");
    count_launch();
    return VQA_OK;
  }
  CUtensorMap tm_wc, tm_wg;
  bool ok =  // weights in TF layout [in, out]: row = input unit (the N of these products), contiguous = output
            // column (their K): K-major B operands as they are
            encode_bf16(&tm_wc, a.wc_h, L, L, L, 64, 32) &&
            encode_bf16(&tm_wg, a.wg_h, 2 * L, L, 2 * L, 64, 32);
  if (!ok) return set_error(VQA_ERR_CUDA, "gru_bwd_persistent: cuTensorMapEncodeTiled failed");
  GruArgs g{};
  g.B = B; g.row0 = 0; g.row_end = B; g.L = L; g.T = T; g.q_len = a.q_len; g.counter = a.counter;
  g.h_f32 = const_cast<float*>(a.h_f32); g.r = const_cast<float*>(a.r); g.u = const_cast<float*>(a.u);
  g.c = const_cast<float*>(a.c); g.dq = a.dq; g.dq2 = a.dq2; g.dG_bf = a.dG_bf; g.dC_bf = a.dC_bf; g.bias_part = a.bias_part;
  return launch_persistent<1>(a.dC_bf, L, static_cast<uint64_t>(T) * B, a.dG_bf, 2 * L, static_cast<uint64_t>(T) * B,
                              tm_wc, tm_wg, g, num_sms, s);
}

}  // namespace vqa
