// CTA-pair variant of the persistent GRU kernels (forward recurrence and BPTT of vlmap/modules.py:124-140, one
// cooperative launch each). Same algorithm and phase structure as gru.cu; what changes is the tiling:
//
//   gru.cu      : CTA = 128 rows x 32 units, M = 128 MMAs of one CTA. Every phase the CTA pulls its 128 rows of
//                 the activation operand (256 KB at L = 1024) plus a streamed weight tile from L2, and an SM ingests
//                 only ~67 GB/s (measured: profiles/r01_launch_summary_v3.md) -> 3.8 us per phase in the main loop.
//   this file   : a PAIR of CTAs (tcgen05 cta_group::2, M = 128) = 128 rows x 64 units. Each CTA supplies 64 rows of
//                 the activation operand (128 KB per phase) and HALF of the weight columns of the pair. The main loop
//                 turned out to be bound by bytes IN FLIGHT (ring depth x tile / ~0.9 us TMA round trip under load),
//                 so the larger weight matrix stays resident (128 KB) and the ring is 8 stages deep (96 KB).
//                 The operand tiles the other CTAs wait for (r.h, h, dG, dC) leave through ONE TMA store of full
//                 128-byte rows followed by its completion wait, instead of scattered 16-byte stores and a
//                 device-wide fence (2.6 us -> see profiles/).
//
// Accumulator layout of M = 128 / cta_group::2 (each CTA holds 64 rows x N columns, folded "2 x 2": columns
// [0, N/2) on TMEM lanes 0..63, columns [N/2, N) on lanes 64..127, both in TMEM columns [0, N/2)). The weight rows
// are ordered so that the first half of the N columns belongs to the first 32 units of the pair and the second half
// to the other 32; a thread therefore owns ONE batch row and 8 units, reads its accumulators straight from TMEM
// (no shared-memory transpose), and keeps the recurrent state (h, u / dh, du) of those 8 elements in registers for
// the whole sequence.
//
// Dependencies: a CTA's activation rows are produced by the 16 CTAs of the same (row tile, rank); they synchronise
// through one global counter per (row tile, rank).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "internal.h"
#include "ptx.cuh"

namespace vqa {

namespace {

constexpr int Q_ROWS = 64;                      // batch rows per CTA
constexpr int Q_UNITS = 32;                     // hidden units whose weights this CTA holds (64 per pair)
constexpr int Q_BK = 64;
constexpr int Q_KBS = 4;                        // k-blocks per stage = per TMA operation (fewer when L / 64 is not a multiple)
constexpr int Q_A_TILE = Q_ROWS * Q_BK * 2;     // 8 KB activation tile of one k-block
constexpr int Q_W_TILE = Q_UNITS * Q_BK * 2;    // 4 KB streamed weight tile (32 rows) of one k-block
// The 96 KB ring has two geometries (every phase starts at stage 0 with all stages free):
//   phases whose weights STREAM : 2 stages of 48 KB = [kbs activation tiles (32 KB) | kbs weight tiles (16 KB)]
//   phases whose weights are RESIDENT: 3 stages of 32 KB, activation tiles only. With two stages only 64 KB of the
//     128 KB (256 KB in the BPTT's K = 2L phase) operand were in flight per ~0.9 us TMA round trip: the phase's loads
//     took two (four) dependent round trips, 2.1 (4.1) us of every phase (profiles/r02_gru_phase_trace.txt).
//
// SWAP (the two-wave variant): the SMALLER matrix is resident instead (forward: candidate rows, BPTT: Wc_h rows; 64 KB),
// which frees 64 KB: the ring grows to 128 KB (resident phases: 4 stages of 32 KB, the whole 128 KB operand in flight;
// streamed phases: 2 stages of 64 KB = [32 KB activation | up to 32 KB weight tiles]) and each wave gets a DEDICATED
// 16 KB staging buffer for its TMA stores. The number of TMA operations per time step is unchanged in the forward kernel
// (8 + 4 instead of 4 + 8) and grows from 16 to 20 in the BPTT.
constexpr int Q_STAGES = 4;                                // barrier slots (the geometries below use 2, 3 or 4 of them)
constexpr int Q_STAGE_S = Q_KBS * (Q_A_TILE + Q_W_TILE);   // 48 KB streamed-phase stage
constexpr int Q_STAGE_R = Q_KBS * Q_A_TILE;                // 32 KB resident-phase stage
constexpr int Q_RING = 2 * Q_STAGE_S;                      // = 3 * Q_STAGE_R = 96 KB
static_assert(Q_RING == 3 * Q_STAGE_R, "ring geometries");
constexpr int Q_STG_BYTES = 2 * Q_A_TILE;                  // 16 KB: the (up to) two staged operand tiles of a phase
constexpr int Q_EPI_WARPS = 16;
constexpr int Q_THREADS = 64 + 32 * Q_EPI_WARPS;   // producer, MMA issuer, 16 epilogue warps
constexpr int Q_TMEM_COLS = 128;                // forward: gates at column 0 (64 wide), candidate at 64 (32 wide)
constexpr int NU = 8;                           // units per epilogue thread

// resident weights: forward = this CTA's gate rows (64 x L), BPTT = its Wg_h rows (32 x 2L): 128 L bytes
__host__ __device__ constexpr int q_wres_bytes(int L, bool swap = false) { return (swap ? 1 : 2) * Q_UNITS * L * 2; }
__host__ __device__ constexpr int q_ring_bytes(bool swap) { return swap ? 4 * Q_STAGE_R : Q_RING; }
__host__ __device__ constexpr int q_smem_bytes(int L, bool swap = false, int waves = 1) {
  return q_wres_bytes(L, swap) + q_ring_bytes(swap) + (swap ? waves * Q_STG_BYTES : 0) + 256 + 1024;
}

struct PairGruArgs {
  int B, row_end, L, T;
  int row0;                // first batch row of this launch (batches beyond one co-resident wave run as consecutive launches)
  int wave_rows;           // WAVES = 2: rows between the two row groups a CTA pair alternates between (tiles per wave * 128)
  int priv_layout;         // B % 64 == 0: saved state in the lane-contiguous private layout
  const int* q_len;
  unsigned int* counter;   // [2 * row tiles]
  const float* xg; const float* xc;
  int x_bf;                // xg / xc are bf16 arrays (same element indexing)
  float* h_f32; bf16* h_bf; bf16* rh_bf;
  float* r; float* u; float* c;
  const float* dq; const float* dq2;
  bf16* dG_bf; bf16* dC_bf;
  float* bias_part;        // [2 * row tiles, 3L]
  int kbs;                 // k-blocks per stage / TMA operation (4, 2 or 1)
  int dbg;                 // timing experiments only (results wrong): 1 = no MMA issued, 2 = no TMA issued
  int tma_out;             // B % 64 == 0: operand tiles leave through TMA stores (the tensor map of the phase that reads them)
  unsigned long long* trace;   // optional [num_ctas, num_phases, 8] globaltimer stamps (scripts/gpu_gru_trace.py)
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define GRU_TRACEW(w, p, k)                                                                                    \
  do {                                                                                                          \
    if (g.trace)                                                                                                \
      g.trace[(((static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * WAVES + (w)) * num_phases + (p)) * 8 + (k)] = \
          gtimer();                                                                                             \
  } while (0)


__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_counter(const unsigned int* p, unsigned int target) {
  if (ld_acquire_u32(p) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_u32(p) < target) {
    if (clock64() - t0 > 4000000000LL) __trap();   // a lost arrival becomes a launch failure, never a hung GPU
  }
}
__device__ __forceinline__ void epi_bar_all() { asm volatile("bar.sync 1, %0;" ::"n"(32 * Q_EPI_WARPS) : "memory"); }
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ void red_relaxed_add(unsigned int* p, unsigned int v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// this warp's 32 TMEM lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[NU]) {
  uint32_t u[NU];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
  ptx::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < NU; ++j) v[j] = __uint_as_float(u[j]);
}
__device__ __forceinline__ void load8f(const float* p, float (&v)[NU]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// 8 consecutive x-projection values starting at element `idx`: fp32 array, or the same array holding bf16 values
__device__ __forceinline__ void load8x(const float* base, long long idx, int is_bf, float (&v)[NU]) {
  if (!is_bf) {
    load8f(base + idx, v);
    return;
  }
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + idx));
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = __uint_as_float(w[j] << 16);
    v[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ void store8f(float* p, const float (&v)[NU]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8bf(bf16* p, const float (&v)[NU]) {
  __nv_bfloat162 a(__float2bfloat16_rn(v[0]), __float2bfloat16_rn(v[1]));
  __nv_bfloat162 b(__float2bfloat16_rn(v[2]), __float2bfloat16_rn(v[3]));
  __nv_bfloat162 c(__float2bfloat16_rn(v[4]), __float2bfloat16_rn(v[5]));
  __nv_bfloat162 d(__float2bfloat16_rn(v[6]), __float2bfloat16_rn(v[7]));
  uint4 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b);
  pk.z = *reinterpret_cast<uint32_t*>(&c); pk.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = pk;
}

// 8 bf16 values -> this thread's 16-byte chunk of a staged 64 x 64 tile (128-byte rows, SWIZZLE_128B)
__device__ __forceinline__ void stage8bf(uint32_t tile, int r, int chunk, const float (&v)[NU]) {
  __nv_bfloat162 a(__float2bfloat16_rn(v[0]), __float2bfloat16_rn(v[1]));
  __nv_bfloat162 b(__float2bfloat16_rn(v[2]), __float2bfloat16_rn(v[3]));
  __nv_bfloat162 c(__float2bfloat16_rn(v[4]), __float2bfloat16_rn(v[5]));
  __nv_bfloat162 d(__float2bfloat16_rn(v[6]), __float2bfloat16_rn(v[7]));
  ptx::st_shared_v4_b32(tile + r * 128 + ((chunk ^ (r & 7)) << 4), *reinterpret_cast<uint32_t*>(&a),
                        *reinterpret_cast<uint32_t*>(&b), *reinterpret_cast<uint32_t*>(&c),
                        *reinterpret_cast<uint32_t*>(&d));
}

// MODE 0: forward.  phase 2t  : G = h_t Wg_h (+xg) -> r, u, r.h        (no matmul at t = 0: h_0 = 0)
//                   phase 2t+1: C = (r.h) Wc_h (+xc) -> c, h_{t+1}
// MODE 1: BPTT.     phase 0   : element-wise head of step T-1 from dq -> du, dC_{T-1}, dh_part
//                   phase 1+2i: dRH = dC_t Wc_h^T -> dG_t, dh_part           (t = T-1-i)
//                   phase 2+2i: dh  = dG_t Wg_h^T + dh_part -> head of step t-1   (not run for t = 0)
// tm_a0 / tm_a1: activation operands of phase kind 0 / 1 (k-block boxes: 64 k x 64 rows x kbs k-blocks);
// tm_o0 / tm_o1: the tensors phase kind 0 / 1 PRODUCES (= the operand of the other kind), 64 x 64 boxes for TMA stores;
// tm_w0 / tm_w1: this CTA's weight rows of kind 0 / 1 (forward: packed slices, box 64 x 64 and 64 x 32;
//                BPTT: Wc_h rows and Wg_h rows in TF layout, box 64 x 32 both).
// WAVES = 2: every CTA pair serves TWO independent row groups ("waves") with the same resident weights and alternates
// between them phase by phase: while the epilogue warps turn wave A's accumulator into the next operand tile and the
// other CTAs' arrivals trickle in, the producer and MMA warps already run wave B's phase. One wave leaves the tensor
// pipe and the load path idle for ~60 % of a phase (profiles/r02_gru_phase_trace.txt: 2.6 us of loads + MMAs in a 5.9 us
// phase); two waves fill that gap. Each wave has its own accumulator columns, accumulator barrier, counters and
// TMA-store staging buffer; the ring is shared (a phase starts once the previous phase's stages are consumed).
// A first version published the operand tiles with plain stores + a gpu-scope fence (no room for staging next to
// 128 KB of weights): MEMBAR.ALL.GPU cost 1.5 - 3.3 us per phase and stalled the other wave's loads -- hence SWAP.
template <int MODE, int WAVES, bool SWAP>
__global__ void __launch_bounds__(Q_THREADS, 1) gru_pair_kernel(
    const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
    const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1,
    const __grid_constant__ CUtensorMap tm_o0, const __grid_constant__ CUtensorMap tm_o1, PairGruArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int L = g.L, T = g.T, B = g.B;
  const int KB = L / Q_BK;
  // the phase kind whose weights are resident: the larger matrix (forward kind 0 = gate rows, 64 x L: KB tiles of 8 KB;
  // BPTT kind 1 = Wg_h rows, 32 x 2L: 2 KB tiles of 4 KB), or with SWAP the smaller one (forward kind 1 = candidate rows,
  // BPTT kind 0 = Wc_h rows: 32 x L). The other matrix streams through the ring next to the activation tiles.
  constexpr int RES_KIND = SWAP ? ((MODE == 0) ? 1 : 0) : ((MODE == 0) ? 0 : 1);
  constexpr int NST_RES = SWAP ? 4 : 3;                               // ring stages of a resident-weight phase
  constexpr uint32_t PITCH_S = SWAP ? 2 * Q_STAGE_R : Q_STAGE_S;      // stage pitch of a streamed-weight phase
  auto wtile = [](int kind) -> uint32_t { return (MODE == 0 && kind == 0) ? 8192u : 4096u; };   // weight bytes per k-block
  uint8_t* wres = smem;
  uint8_t* ring = smem + q_wres_bytes(L, SWAP);
  uint8_t* stg = ring + q_ring_bytes(SWAP);           // SWAP: [WAVES][16 KB] TMA-store staging
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg + (SWAP ? WAVES * Q_STG_BYTES : 0));
  uint64_t* empty_bar = full_bar + Q_STAGES;
  uint64_t* tmem_full_bar = empty_bar + Q_STAGES;   // [2]: one per wave
  uint64_t* w_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int slice32 = blockIdx.x;               // 32-unit slice whose weights this CTA holds (= 2 * pair + rank)
  const int mi = blockIdx.y;
  const int j0 = slice32 * Q_UNITS;
  int m0w[WAVES];                                // first batch row of this CTA in each wave
  unsigned int* counterw[WAVES];
#pragma unroll
  for (int w = 0; w < WAVES; ++w) {
    m0w[w] = g.row0 + w * g.wave_rows + mi * 128 + static_cast<int>(rank) * Q_ROWS;
    counterw[w] = g.counter + 2 * (w * static_cast<int>(gridDim.y) + mi) + rank;
  }
  const unsigned int nprod = gridDim.x >> 1;    // CTAs producing this CTA's activation rows (same row tile and rank)
  const int num_phases = 2 * T;
  constexpr int TM_WAVE = 128;                  // accumulator columns per wave

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a0);
    ptx::prefetch_tensormap(&tm_a1);
    ptx::prefetch_tensormap(&tm_w0);
    ptx::prefetch_tensormap(&tm_w1);
    for (int s = 0; s < Q_STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(&tmem_full_bar[0], 1);
    ptx::mbar_init(&tmem_full_bar[1], 1);
    ptx::mbar_init(w_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, Q_TMEM_COLS * WAVES);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto phase_info = [&](int p, bool& mm, int& kind, int& t) {
    if (MODE == 0) {
      t = p >> 1;
      kind = p & 1;
      mm = t > 0;
    } else {
      if (p == 0) { mm = false; kind = 1; t = T; return; }
      const int i = (p - 1) >> 1;
      t = T - 1 - i;
      kind = (p - 1) & 1;
      mm = true;
    }
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const uint32_t lw = ptx::mapa_u32(ptx::smem_u32(w_bar), 0);
    if (lane == 0) {
      // resident weights of BOTH CTAs complete on the leader's barrier (it issues the MMAs that read them)
      if (rank == 0) ptx::mbar_arrive_expect_tx(w_bar, 2u * static_cast<uint32_t>(q_wres_bytes(L, SWAP)));
      // tm_w0 / tm_w1 are the weights of phase kind 0 / 1: the resident one as 2-D boxes of one k-block, the streamed one
      // as boxes of g.kbs k-blocks. Forward rows of a slice in the packed matrix: [r rows 32 | u rows 32 | c rows 32]
      // (gate tile of a k-block = [r | u] = 8 KB, candidate tile 4 KB); BPTT: rows j0.. of Wc_h (kind 0) / Wg_h (kind 1).
      const CUtensorMap* tres = RES_KIND == 0 ? &tm_w0 : &tm_w1;
      const int res_row = (MODE == 0) ? slice32 * 96 + (RES_KIND == 0 ? 0 : 64) : j0;
      const int res_nkb = (MODE == 1 && RES_KIND == 1) ? 2 * KB : KB;
      for (int kb = 0; kb < res_nkb; ++kb)
        ptx::tma_load_2d_pair(wres + kb * wtile(RES_KIND), tres, lw, kb * Q_BK, res_row);
    }
    __syncwarp();
    uint32_t uses[Q_STAGES] = {0u, 0u, 0u, 0u};   // how often each stage has been filled so far (barrier phase parity)
    uint32_t gslot = 0;                            // SWAP: boxes issued so far (slot = gslot % 4)
    for (int pw = 0; pw < num_phases * WAVES; ++pw) {
      const int p = pw / WAVES, w = pw - p * WAVES;
      bool mm; int kind, t;
      phase_info(p, mm, kind, t);
      if (!mm) continue;
      const bool streamed = kind != RES_KIND;
      const CUtensorMap* ta = kind ? &tm_a1 : &tm_a0;
      const CUtensorMap* tw = kind ? &tm_w1 : &tm_w0;
      const int wrow = (MODE == 0) ? slice32 * 96 + (kind == 0 ? 0 : 64) : j0;
      const int nkb = (MODE == 1 && kind == 1) ? 2 * KB : KB;
      const int arow = t * B + (w == 0 ? m0w[0] : m0w[WAVES - 1]);
      unsigned int* counter = w == 0 ? counterw[0] : counterw[WAVES - 1];
      const int kbs = g.kbs;
      const uint32_t a_bytes = kbs * Q_A_TILE, w_bytes = kbs * wtile(kind);
      if constexpr (SWAP) {
        // ONE continuous ring of four 32 KB slots across phases and waves: every box (activation or weight tiles of
        // g.kbs k-blocks) takes the next slot, so the loads of the next phase / wave start as soon as slots free up and
        // its counter allows -- nothing drains between phases. A streamed phase issues weight box, activation box,
        // weight box, ... (the first weight box goes out before the counter wait).
        auto issue = [&](const CUtensorMap* map, int row, int kb, uint32_t bytes) {
          const uint32_t slot = gslot & 3u, use = gslot >> 2;
          ++gslot;
          ptx::mbar_wait(&empty_bar[slot], (use & 1u) ^ 1u);
          const uint32_t lf = ptx::mapa_u32(ptx::smem_u32(&full_bar[slot]), 0);
          if (ptx::elect_one()) {
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[slot], 2 * bytes);
            ptx::tma_load_3d_pair(ring + slot * Q_STAGE_R, map, lf, 0, row, kb);
          }
          __syncwarp();
        };
        for (int kb = 0; kb < nkb; kb += kbs) {
          if (streamed) issue(tw, wrow, kb, w_bytes);
          if (kb == 0) {
            wait_counter(counter, static_cast<unsigned int>(p) * nprod);
            ptx::fence_proxy_async_full();
            if (lane == 0) GRU_TRACEW(w, p, 0);
            __syncwarp();
          }
          issue(ta, arow, kb, a_bytes);
        }
        continue;
      }
      const int nstages = streamed ? 2 : NST_RES;
      const uint32_t pitch = streamed ? PITCH_S : Q_STAGE_R;
      // the two geometries overlap: before anything of this phase is written (the early weight load below included),
      // every stage the previous phase filled must have been consumed by its MMAs
#pragma unroll
      for (int q = 0; q < Q_STAGES; ++q)
        if (uses[q]) ptx::mbar_wait(&empty_bar[q], (uses[q] - 1u) & 1u);
      int stage = 0;
      for (int kb = 0; kb < nkb; kb += kbs) {
        uint32_t use = 0;
#pragma unroll
        for (int q = 0; q < Q_STAGES; ++q) if (q == stage) { use = uses[q]; uses[q] = use + 1; }
        ptx::mbar_wait(&empty_bar[stage], (use & 1u) ^ 1u);
        const uint32_t lf = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
        uint8_t* st = ring + stage * pitch;
        const bool first = kb == 0;
        // the streamed weights do not depend on the other CTAs: for the first stage they go out BEFORE the counter
        // wait (stage 0 doubles as the epilogue's TMA-store staging, but only its activation area)
        if (ptx::elect_one() && g.dbg != 2) {
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * (a_bytes + (streamed ? w_bytes : 0)));
          if (streamed) ptx::tma_load_3d_pair(st + Q_KBS * Q_A_TILE, tw, lf, 0, wrow, kb);
        }
        __syncwarp();
        if (first) {
          // this phase's operand rows were written by the epilogues of phase p - 1 of the CTAs sharing our rows
          wait_counter(counter, static_cast<unsigned int>(p) * nprod);
          ptx::fence_proxy_async_full();
          if (lane == 0) GRU_TRACEW(w, p, 0);
          __syncwarp();
        }
        if (ptx::elect_one()) {
          if (g.dbg == 2) {
            if (rank == 0) ptx::mbar_arrive(&full_bar[stage]);
          } else {
            ptx::tma_load_3d_pair(st, ta, lf, 0, arow, kb);
          }
        }
        __syncwarp();
        if (++stage == nstages) stage = 0;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA of the pair) =====================
    if (rank == 0) {
      constexpr uint32_t idesc128 = ptx::make_idesc_bf16(128, 128, false, false);   // forward gates: r|u of 64 units
      constexpr uint32_t idesc64 = ptx::make_idesc_bf16(128, 64, false, false);
      ptx::mbar_wait(w_bar, 0);
      uint32_t uses[Q_STAGES] = {0u, 0u, 0u, 0u};
      uint32_t gslot = 0;
      for (int pw = 0; pw < num_phases * WAVES; ++pw) {
        const int p = pw / WAVES, w = pw - p * WAVES;
        bool mm; int kind, t;
        phase_info(p, mm, kind, t);
        if (!mm) continue;
        const bool streamed = kind != RES_KIND;
        const bool wide = (MODE == 0 && kind == 0);
        const uint32_t idesc = wide ? idesc128 : idesc64;
        const int nkb = (MODE == 1 && kind == 1) ? 2 * KB : KB;
        const uint32_t wt = wtile(kind);
        const uint32_t d_tmem = tmem_base + w * TM_WAVE + ((MODE == 0 && kind == 1) ? 64 : 0);
        const int kbs = g.kbs;
        if constexpr (SWAP) {
          for (int kb = 0; kb < nkb; kb += kbs) {
            uint32_t wslot = 0;
            if (streamed) {
              wslot = gslot & 3u;
              ptx::mbar_wait(&full_bar[wslot], (gslot >> 2) & 1u);
              ++gslot;
            }
            const uint32_t aslot = gslot & 3u;
            ptx::mbar_wait(&full_bar[aslot], (gslot >> 2) & 1u);
            ++gslot;
            ptx::tc_fence_after();
            if (kb == 0 && lane == 0) GRU_TRACEW(w, p, 1);
            const uint32_t sa0 = ptx::smem_u32(ring + aslot * Q_STAGE_R);
            const uint32_t sw0 = ptx::smem_u32(ring + wslot * Q_STAGE_R);
            if (ptx::elect_one()) {
              for (int i = 0; i < kbs; ++i) {
                const uint32_t sa = sa0 + i * Q_A_TILE;
                const uint32_t sb = streamed ? sw0 + i * wt : ptx::smem_u32(wres) + (kb + i) * wt;
#pragma unroll
                for (int kk = 0; kk < Q_BK / 16; ++kk) {
                  const uint64_t da = ptx::make_smem_desc_sw128(sa + kk * 32, 16, 1024);
                  const uint64_t db = ptx::make_smem_desc_sw128(sb + kk * 32, 16, 1024);
                  ptx::umma_f16_pair(d_tmem, da, db, idesc, (kb | i | kk) != 0);
                }
              }
              ptx::umma_commit_pair(&empty_bar[aslot], 3);
              if (streamed) ptx::umma_commit_pair(&empty_bar[wslot], 3);
            }
            __syncwarp();
          }
          if (ptx::elect_one()) ptx::umma_commit_pair(&tmem_full_bar[w], 3);
          __syncwarp();
          continue;
        }
        const int nstages = streamed ? 2 : NST_RES;
        const uint32_t pitch = streamed ? PITCH_S : Q_STAGE_R;
        int stage = 0;
        for (int kb = 0; kb < nkb; kb += kbs) {
          uint32_t use = 0;
#pragma unroll
          for (int q = 0; q < Q_STAGES; ++q) if (q == stage) { use = uses[q]; uses[q] = use + 1; }
          ptx::mbar_wait(&full_bar[stage], use & 1u);
          ptx::tc_fence_after();
          if (kb == 0 && lane == 0) GRU_TRACEW(w, p, 1);
          const uint32_t sa0 = ptx::smem_u32(ring + stage * pitch);
          const uint32_t sw0 = sa0 + Q_KBS * Q_A_TILE;
          if (ptx::elect_one()) {
            if (g.dbg != 1)
            for (int i = 0; i < kbs; ++i) {
              const uint32_t sa = sa0 + i * Q_A_TILE;
              const uint32_t sb = streamed ? sw0 + i * wt : ptx::smem_u32(wres) + (kb + i) * wt;
#pragma unroll
              for (int kk = 0; kk < Q_BK / 16; ++kk) {
                const uint64_t da = ptx::make_smem_desc_sw128(sa + kk * 32, 16, 1024);
                const uint64_t db = ptx::make_smem_desc_sw128(sb + kk * 32, 16, 1024);
                ptx::umma_f16_pair(d_tmem, da, db, idesc, (kb | i | kk) != 0);
              }
            }
            ptx::umma_commit_pair(&empty_bar[stage], 3);
          }
          __syncwarp();
          if (++stage == nstages) stage = 0;
        }
        if (ptx::elect_one()) ptx::umma_commit_pair(&tmem_full_bar[w], 3);
        __syncwarp();
        // the accumulators are overwritten only after both CTAs' next counter waits, which their own epilogues
        // reach after draining TMEM: no tmem_empty barrier needed
      }
    }
  } else {
    // ===================== epilogue warps: thread = (batch row, 8 units), for each wave =====================
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int sub = (warp - 2) >> 2;                 // which 8 of the 32 units of the half
    const int uhalf = q >> 1;                        // lanes 0..63: first 32 units of the pair, 64..127: the others
    const int unit = (slice32 & ~1) * Q_UNITS + uhalf * 32 + sub * NU;   // first of this thread's 8 units
    const bool leader = threadIdx.x == 64;
    const int srow = (q & 1) * 32 + lane;         // row inside the CTA's 64-row tile
    const int schunk = uhalf * 4 + sub;           // 16-byte chunk inside the 128-byte row of 64 units
    const int ucol = (slice32 >> 1) * 64;         // first unit of the pair
    // operand tiles leave by TMA store when the staging buffer is free between a wave's phases: one wave owning the
    // ring (staging inside it), or the dedicated per-wave buffers of SWAP
    static_assert(WAVES == 1 || SWAP, "two waves need the dedicated staging buffers");
    const bool tma_out = g.tma_out != 0;
    const bool priv_layout = g.priv_layout != 0;
    // !SWAP: the A areas of ring stage 0 (idle between the last MMA of a phase and the next phase's first load, which
    // waits for this CTA's own arrival below)
    int roww[WAVES], lenw[WAVES];
    bool okw[WAVES];
    uint32_t tlanew[WAVES], tfullw[WAVES], stg0w[WAVES], stg1w[WAVES];
#pragma unroll
    for (int w = 0; w < WAVES; ++w) {
      roww[w] = m0w[w] + (q & 1) * 32 + lane;
      okw[w] = roww[w] < g.row_end;
      lenw[w] = okw[w] ? g.q_len[roww[w]] : 0;
      tlanew[w] = tmem_base + w * TM_WAVE + (static_cast<uint32_t>(q * 32) << 16) + sub * NU;
      tfullw[w] = 0;
      stg0w[w] = SWAP ? ptx::smem_u32(stg + w * Q_STG_BYTES) : ptx::smem_u32(ring);
      stg1w[w] = stg0w[w] + Q_A_TILE;
    }
    // fp32 state kept for BPTT (r, u, c, h_t): with full 64-row tiles it lives in a layout private to these two
    // kernels, [t][64-row tile][unit / 8][row % 64][unit % 8], in which the 32 bytes of the lanes of a warp are
    // consecutive (1 KB per warp access instead of 32 sectors 4 KB apart). The final state h_T stays row-major:
    // it is the model's `condition` output.
    const int nt64 = (B + Q_ROWS - 1) / Q_ROWS;
    auto priv = [&](int tblock, int m0, int row) -> long long {
      if (priv_layout)
        return ((static_cast<long long>(tblock) * nt64 + (m0 >> 6)) * (L >> 3) + (unit >> 3)) * (Q_ROWS * NU) + srow * NU;
      return (static_cast<long long>(tblock) * B + row) * L + unit;
    };
    // publish: CTA barrier, then ONE thread hands the staged tile(s) to the TMA, waits for the writes to be
    // performed, fences and releases the counter
    auto publish = [&](int p, int w, bool mm, const CUtensorMap* tm, int ntiles, int col1, long long grow, int m0, unsigned int* counter) {
      const uint32_t stg0 = stg0w[w], stg1 = stg1w[w];
      if (tma_out) ptx::fence_proxy_async();
      if (leader) GRU_TRACEW(w, p, 5);
      epi_bar_all();
      if (leader) {
        GRU_TRACEW(w, p, 6);
        if (tma_out) {
          // The tile(s) leave through the async proxy and bulk_wait<0> returns once the writes are PERFORMED (at
          // L2, where the consumers' TMA loads read them), so the relaxed arrival below is ordered after them by
          // this thread's program order alone. No gpu-scope fence: it would also wait for the scattered fp32
          // stores of r / u / c (needed only by the BPTT launch), which was 2.6 us of every phase.
          if (m0 < g.row_end) {   // (B % 64 == 0: a CTA's 64 rows are all valid or all beyond the batch)
            ptx::tma_store_2d(tm, stg0, ucol, static_cast<int>(grow));
            if (ntiles == 2) ptx::tma_store_2d(tm, stg1, col1, static_cast<int>(grow));
            ptx::bulk_commit();
            ptx::bulk_wait<0>();
          }
          GRU_TRACEW(w, p, 4);
        } else {
          GRU_TRACEW(w, p, 4);
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          ptx::fence_proxy_async_full();
        }
        GRU_TRACEW(w, p, 7);
        red_relaxed_add(counter, 1u);
        GRU_TRACEW(w, p, 3);
      }
      // phases without a matmul wait for nothing before their staging writes: hold everybody until the leader's store
      // has left the staging buffer (only the first one or two phases of a launch)
      if (!mm && tma_out) epi_bar_all();
    };

    if (MODE == 0) {
      float hw[WAVES][NU], uw[WAVES][NU];
#pragma unroll
      for (int w = 0; w < WAVES; ++w)
#pragma unroll
        for (int j = 0; j < NU; ++j) hw[w][j] = 0.f, uw[w][j] = 0.f;
      for (int p = 0; p < num_phases; ++p) {
        bool mm; int kind, t;
        phase_info(p, mm, kind, t);
        const long long tb = static_cast<long long>(t) * B;
#pragma unroll
        for (int w = 0; w < WAVES; ++w) {
          const int row = roww[w], m0 = m0w[w];
          const bool row_ok = okw[w];
          float (&h)[NU] = hw[w];
          float (&u)[NU] = uw[w];
          float sv[NU];   // r (kind 0) or c (kind 1)
          if (kind == 0) {
            float xr[NU], xu[NU], ar[NU], au[NU];
#pragma unroll
            for (int j = 0; j < NU; ++j) xr[j] = xu[j] = ar[j] = au[j] = 0.f;
            if (row_ok) {
              const long long xo = (tb + row) * 2 * L + unit;
              load8x(g.xg, xo, g.x_bf, xr);
              load8x(g.xg, xo + L, g.x_bf, xu);
            }
            if (mm) {
              ptx::mbar_wait(&tmem_full_bar[w], tfullw[w]);
              tfullw[w] ^= 1;
              ptx::tc_fence_after();
              if (leader) GRU_TRACEW(w, p, 2);
              tmem_ld8(tlanew[w], ar);
              tmem_ld8(tlanew[w] + 32, au);
              ptx::tc_fence_before();
            }
            float rh[NU];
#pragma unroll
            for (int j = 0; j < NU; ++j) {
              sv[j] = sigm(ar[j] + xr[j]);
              u[j] = sigm(au[j] + xu[j]);
              rh[j] = sv[j] * h[j];
            }
            // the operand the other CTAs wait for goes out first; r / u (kept for BPTT) after the arrival
            if (tma_out) stage8bf(stg0w[w], srow, schunk, rh);
            else if (row_ok) store8bf(g.rh_bf + (tb + row) * L + unit, rh);
          } else {
            float xc[NU], ac[NU];
#pragma unroll
            for (int j = 0; j < NU; ++j) xc[j] = ac[j] = 0.f;
            if (row_ok) load8x(g.xc, (tb + row) * L + unit, g.x_bf, xc);
            if (mm) {
              ptx::mbar_wait(&tmem_full_bar[w], tfullw[w]);
              tfullw[w] ^= 1;
              ptx::tc_fence_after();
              if (leader) GRU_TRACEW(w, p, 2);
              tmem_ld8(tlanew[w] + 64, ac);
              ptx::tc_fence_before();
            }
            const bool valid = t < lenw[w];
#pragma unroll
            for (int j = 0; j < NU; ++j) {
              sv[j] = tanh_fast(ac[j] + xc[j]);
              h[j] = valid ? u[j] * h[j] + (1.0f - u[j]) * sv[j] : h[j];
            }
            if (tma_out) stage8bf(stg0w[w], srow, schunk, h);
            else if (row_ok) store8bf(g.h_bf + (tb + B + row) * L + unit, h);
          }
          // r.h is the operand of the candidate phase (tm_a1), h_{t+1} the operand of the next gate phase (tm_a0)
          publish(p, w, mm, kind == 0 ? &tm_o0 : &tm_o1, 1, 0, kind == 0 ? tb + m0 : tb + B + m0, m0, counterw[w]);
          // what only BPTT reads goes out off the critical path
          if (row_ok) {
            const long long o = priv(t, m0, row);
            if (kind == 0) {
              store8f(g.r + o, sv);
              store8f(g.u + o, u);
            } else {
              store8f(g.c + o, sv);
              if (t + 1 == T) store8f(g.h_f32 + (tb + B + row) * L + unit, h);
              else store8f(g.h_f32 + priv(t + 1, m0, row), h);
            }
          }
        }
      }
    } else {
      float dhpw[WAVES][NU], duw[WAVES][NU];
      float db_r[NU], db_u[NU], db_c[NU];   // bias gradients: both waves' rows add into the same partial row
#pragma unroll
      for (int j = 0; j < NU; ++j) db_r[j] = db_u[j] = db_c[j] = 0.f;
#pragma unroll
      for (int w = 0; w < WAVES; ++w)
#pragma unroll
        for (int j = 0; j < NU; ++j) dhpw[w][j] = duw[w][j] = 0.f;
      for (int p = 0; p < num_phases; ++p) {
        bool mm; int kind, t;
        phase_info(p, mm, kind, t);
#pragma unroll
        for (int w = 0; w < WAVES; ++w) {
          const int row = roww[w], m0 = m0w[w];
          const bool row_ok = okw[w];
          float (&dhp)[NU] = dhpw[w];
          float (&du)[NU] = duw[w];
          if (kind == 0) {
            // dRH = acc ;  dG_r = dRH h r (1-r) ; dG_u = du u (1-u) ; dh_part += dRH r
            const long long tb = static_cast<long long>(t) * B;
            float hh[NU], rv[NU], uv[NU], acc[NU];
#pragma unroll
            for (int j = 0; j < NU; ++j) hh[j] = rv[j] = uv[j] = acc[j] = 0.f;
            if (row_ok) {
              const long long o = priv(t, m0, row);
              if (t > 0) load8f(g.h_f32 + o, hh);   // h_0 = 0 (block 0 is never written in the private layout)
              load8f(g.r + o, rv);
              load8f(g.u + o, uv);
            }
            ptx::mbar_wait(&tmem_full_bar[w], tfullw[w]);
            tfullw[w] ^= 1;
            ptx::tc_fence_after();
            if (leader) GRU_TRACEW(w, p, 2);
            tmem_ld8(tlanew[w], acc);
            ptx::tc_fence_before();
            float dgr[NU], dgu[NU];
#pragma unroll
            for (int j = 0; j < NU; ++j) {
              const float drh = acc[j];   // zero for steps beyond the question length (dC is zero there)
              dgr[j] = drh * hh[j] * rv[j] * (1.0f - rv[j]);
              dgu[j] = du[j] * uv[j] * (1.0f - uv[j]);
              dhp[j] = fmaf(drh, rv[j], dhp[j]);
            }
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < NU; ++j) {
                db_r[j] += dgr[j];
                db_u[j] += dgu[j];
              }
              if (!tma_out) {
                const long long o = (tb + row) * 2 * L + unit;
                store8bf(g.dG_bf + o, dgr);
                store8bf(g.dG_bf + o + L, dgu);
              }
            }
            if (tma_out) {
              stage8bf(stg0w[w], srow, schunk, dgr);
              stage8bf(stg1w[w], srow, schunk, dgu);
            }
            publish(p, w, mm, &tm_o0, 2, L + ucol, tb + m0, m0, counterw[w]);   // dG_t = [dG_r | dG_u]: operand of the next phase
          } else {
            // dh = acc + dh_part (or dq) ; element-wise head of step tp = t - 1
            const int tp = t - 1;
            const long long tb = static_cast<long long>(tp) * B;
            float hh[NU], uv[NU], cv[NU], acc[NU];
#pragma unroll
            for (int j = 0; j < NU; ++j) hh[j] = uv[j] = cv[j] = acc[j] = 0.f;
            if (row_ok) {
              const long long o = priv(tp, m0, row);
              if (tp > 0) load8f(g.h_f32 + o, hh);
              load8f(g.u + o, uv);
              load8f(g.c + o, cv);
              if (!mm) {
                load8f(g.dq + static_cast<long long>(row) * L + unit, dhp);
                if (g.dq2) {
                  float d2[NU];
                  load8f(g.dq2 + static_cast<long long>(row) * L + unit, d2);
#pragma unroll
                  for (int j = 0; j < NU; ++j) dhp[j] += d2[j];
                }
              }
            }
            if (mm) {
              ptx::mbar_wait(&tmem_full_bar[w], tfullw[w]);
              tfullw[w] ^= 1;
              ptx::tc_fence_after();
              if (leader) GRU_TRACEW(w, p, 2);
              tmem_ld8(tlanew[w], acc);
              ptx::tc_fence_before();
            }
            const bool pvalid = tp < lenw[w];
            float dcv[NU];
#pragma unroll
            for (int j = 0; j < NU; ++j) {
              const float dh = acc[j] + dhp[j];
              dcv[j] = pvalid ? dh * (1.0f - uv[j]) * (1.0f - cv[j] * cv[j]) : 0.f;
              du[j] = pvalid ? dh * (hh[j] - cv[j]) : 0.f;
              dhp[j] = pvalid ? dh * uv[j] : dh;
            }
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < NU; ++j) db_c[j] += dcv[j];
              if (!tma_out) store8bf(g.dC_bf + (tb + row) * L + unit, dcv);
            }
            if (tma_out) stage8bf(stg0w[w], srow, schunk, dcv);
            publish(p, w, mm, &tm_o1, 1, 0, tb + m0, m0, counterw[w]);          // dC_{t-1}: operand of the next phase
          }
        }
      }
      // bias gradients: sum over the 32 rows of the warp (fixed butterfly order), then over the two row-warps
      // that share these units; one partial row of [3L] per CTA (the second wave's row is written as zeros)
#pragma unroll
      for (int j = 0; j < NU; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          db_r[j] += __shfl_xor_sync(0xffffffffu, db_r[j], o);
          db_u[j] += __shfl_xor_sync(0xffffffffu, db_u[j], o);
          db_c[j] += __shfl_xor_sync(0xffffffffu, db_c[j], o);
        }
      }
      float* red = reinterpret_cast<float*>(ring);   // the ring is free now: [16 warps][3][8]
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          red[((warp - 2) * 3 + 0) * NU + j] = db_r[j];
          red[((warp - 2) * 3 + 1) * NU + j] = db_u[j];
          red[((warp - 2) * 3 + 2) * NU + j] = db_c[j];
        }
      }
      epi_bar_all();
      // warps with (q & 1) == 0 combine with their partner (q | 1, same sub) and write
      if ((q & 1) == 0 && lane < 3 * NU) {
        const int k = lane / NU, j = lane % NU;
        const float s = red[((warp - 2) * 3 + k) * NU + j] + red[((warp - 2 + 1) * 3 + k) * NU + j];
#pragma unroll
        for (int w = 0; w < WAVES; ++w) {
          const int tile = (g.row0 + w * g.wave_rows) / 128 + mi;
          if (tile * 128 < ((g.B + 127) / 128) * 128) {
            float* out = g.bias_part + static_cast<long long>(2 * tile + rank) * 3 * L;
            out[k * L + unit + j] = w == 0 ? s : 0.f;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, Q_TMEM_COLS * WAVES);
  }
}

template <int MODE, int WAVES, bool SWAP>
cudaError_t launch_pair_gru(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w0, const CUtensorMap& w1,
                            const CUtensorMap& o0, const CUtensorMap& o1, PairGruArgs& a, dim3 grid, int smem,
                            cudaStream_t s) {
  auto kern = gru_pair_kernel<MODE, WAVES, SWAP>;
  static int smem_set = 0;
  if (smem_set < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  void* args[] = {const_cast<CUtensorMap*>(&a0), const_cast<CUtensorMap*>(&a1), const_cast<CUtensorMap*>(&w0),
                  const_cast<CUtensorMap*>(&w1), const_cast<CUtensorMap*>(&o0), const_cast<CUtensorMap*>(&o1), &a};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(Q_THREADS);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeCooperative;   // fails instead of deadlocking if the grid cannot be co-resident
  at[0].val.cooperative = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = 2;
  at[1].val.clusterDim.y = 1;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  return cudaLaunchKernelExC(&cfg, reinterpret_cast<void*>(kern), args);
}

int pick_kbs(int L) {
  const int KB = L / Q_BK;
  return KB % 4 == 0 ? 4 : (KB % 2 == 0 ? 2 : 1);
}

bool g_pair_off = getenv("VQA_GRU_PAIR") != nullptr && atoi(getenv("VQA_GRU_PAIR")) == 0;

}  // namespace

// testing aid (not part of the ABI header): 0 = single-CTA kernels of gru.cu, 1 = the CTA-pair kernels
extern "C" __attribute__((visibility("default"))) void vqa_internal_set_gru_pair(int on) { g_pair_off = on == 0; }

bool gru_pair_supported(int B, int L, int num_sms) {
  if (g_pair_off) return false;
  if (L % 64 != 0 || L < 64) return false;
  if (q_smem_bytes(L) > 227 * 1024 || q_smem_bytes(L, true, 2) > 227 * 1024) return false;
  (void)B;   // any batch: row tiles beyond one co-resident wave run as consecutive launches (launch_chunks)
  return L / Q_UNITS <= num_sms;
}

// Launch plan. tiles_max = num_sms / (L / 32) co-resident row tiles of 128 rows (4 at L = 1024).
//   waves2 (default for batches beyond one wave: the 5120 blank sequences of the pre-training graph, BASELINE config 5's
//   large inference batches): every launch serves 2 x tiles_max row tiles, each CTA pair alternating between two row
//   groups (gru_pair_kernel<MODE, 2>); a remainder of at most tiles_max tiles runs as one wave.
//   VQA_GRU_WAVES = 1 forces single waves, = 2 also splits a batch of ONE wave into two half-waves on half of the SMs
//   (the recurrent kernels then leave the other half of the machine to concurrent kernels).
int g_waves_mode = getenv("VQA_GRU_WAVES") ? atoi(getenv("VQA_GRU_WAVES")) : 0;
// testing aid (not part of the ABI header): the VQA_GRU_WAVES policy at run time
extern "C" __attribute__((visibility("default"))) void vqa_internal_set_gru_waves(int mode) { g_waves_mode = mode; }

// w0 / w1: weight maps of the single-wave kernels (larger matrix resident); sw0 / sw1: those of the two-wave kernels
// (smaller matrix resident, the larger one streamed in k-block boxes)
template <int MODE>
cudaError_t launch_chunks(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w0, const CUtensorMap& w1,
                          const CUtensorMap& sw0, const CUtensorMap& sw1, const CUtensorMap& o0, const CUtensorMap& o1,
                          PairGruArgs& g, int num_sms, cudaStream_t s) {
  const int slices = g.L / Q_UNITS;
  int tiles_max = num_sms / slices;
  if (tiles_max > 16) tiles_max = 16;   // counters: [2 waves][tiles][2 ranks] of a 64-entry array
  g.priv_layout = g.tma_out;
  bool first = true;
  for (int row0 = 0; row0 < g.B;) {
    const int left = g.B - row0;
    const int tiles_left = (left + 127) / 128;
    int waves = 1, tiles = tiles_left < tiles_max ? tiles_left : tiles_max;
    if (g_waves_mode != 1 && tiles_left > tiles_max) {
      waves = 2;
      tiles = (tiles_left + 1) / 2 < tiles_max ? (tiles_left + 1) / 2 : tiles_max;
    } else if (g_waves_mode == 2 && tiles_left >= 2) {
      waves = 2;
      tiles = (tiles_left + 1) / 2;
    }
    cudaError_t e = cudaMemsetAsync(g.counter, 0, sizeof(unsigned int) * 2 * waves * tiles, s);
    if (e != cudaSuccess) return e;
    g.row0 = row0;
    g.wave_rows = tiles * 128;
    static const bool swap1 = getenv("VQA_GRU_SWAP1") != nullptr && atoi(getenv("VQA_GRU_SWAP1")) != 0;   // experiment
    e = waves == 2 ? launch_pair_gru<MODE, 2, true>(a0, a1, sw0, sw1, o0, o1, g, dim3(slices, tiles), q_smem_bytes(g.L, true, 2), s)
        : swap1    ? launch_pair_gru<MODE, 1, true>(a0, a1, sw0, sw1, o0, o1, g, dim3(slices, tiles), q_smem_bytes(g.L, true, 1), s)
                   : launch_pair_gru<MODE, 1, false>(a0, a1, w0, w1, o0, o1, g, dim3(slices, tiles), q_smem_bytes(g.L), s);
    if (e != cudaSuccess) return e;
    if (!first) count_launch();   // (the caller counts the first)
    first = false;
    row0 += waves * tiles * 128;
  }
  return cudaSuccess;
}

// returns cudaSuccess, or the launch error (the caller falls back to the single-CTA kernels)
cudaError_t gru_pair_fwd(const GruFwdPersistent& a, int num_sms, cudaStream_t s) {
  const int B = a.B, L = a.L, T = a.T;
  CUtensorMap tm_h, tm_rh, tm_wg, tm_wc, to_rh, to_h, sw_wg, sw_wc;
  const uint64_t prow = static_cast<uint64_t>(L / 32) * 96;
  const int kbs = pick_kbs(L);
  if (!cached_tmap_kblocks(&tm_h, a.h_bf, L, static_cast<uint64_t>(T + 1) * B, L, Q_ROWS, kbs) ||
      !cached_tmap_kblocks(&tm_rh, a.rh_bf, L, static_cast<uint64_t>(T) * B, L, Q_ROWS, kbs) ||
      !cached_tmap(&tm_wg, a.w_pack, L, prow, L, 64, 64) || !cached_tmap_kblocks(&tm_wc, a.w_pack, L, prow, L, 32, kbs) ||
      !cached_tmap_kblocks(&sw_wg, a.w_pack, L, prow, L, 64, kbs) || !cached_tmap(&sw_wc, a.w_pack, L, prow, L, 64, 32) ||
      !cached_tmap(&to_rh, a.rh_bf, L, static_cast<uint64_t>(T) * B, L, 64, Q_ROWS) ||
      !cached_tmap(&to_h, a.h_bf, L, static_cast<uint64_t>(T + 1) * B, L, 64, Q_ROWS))
    return cudaErrorInvalidValue;
  PairGruArgs g{};
  g.B = B; g.row_end = B; g.L = L; g.T = T; g.q_len = a.q_len; g.counter = a.counter;
  g.xg = a.xg; g.xc = a.xc; g.x_bf = a.x_bf16; g.h_f32 = a.h_f32; g.h_bf = a.h_bf; g.rh_bf = a.rh_bf; g.r = a.r; g.u = a.u; g.c = a.c;
  g.trace = g_gru_trace;
  g.dbg = getenv("VQA_GRU_DBG") ? atoi(getenv("VQA_GRU_DBG")) : 0;
  g.tma_out = (B % Q_ROWS == 0) ? 1 : 0;
  g.kbs = kbs;
  return launch_chunks<0>(tm_h, tm_rh, tm_wg, tm_wc, sw_wg, sw_wc, to_rh, to_h, g, num_sms, s);
}

cudaError_t gru_pair_bwd(const GruBwdPersistent& a, int num_sms, cudaStream_t s) {
  const int B = a.B, L = a.L, T = a.T;
  CUtensorMap tm_dc, tm_dg, tm_wc, tm_wg, to_dg, to_dc, sw_wc, sw_wg;
  const int kbs = pick_kbs(L);
  if (!cached_tmap_kblocks(&tm_dc, a.dC_bf, L, static_cast<uint64_t>(T) * B, L, Q_ROWS, kbs) ||
      !cached_tmap_kblocks(&tm_dg, a.dG_bf, 2 * L, static_cast<uint64_t>(T) * B, 2 * L, Q_ROWS, kbs) ||
      !cached_tmap_kblocks(&tm_wc, a.wc_h, L, L, L, 32, kbs) || !cached_tmap(&tm_wg, a.wg_h, 2 * L, L, 2 * L, 64, 32) ||
      !cached_tmap(&sw_wc, a.wc_h, L, L, L, 64, 32) || !cached_tmap_kblocks(&sw_wg, a.wg_h, 2 * L, L, 2 * L, 32, kbs) ||
      !cached_tmap(&to_dg, a.dG_bf, 2 * L, static_cast<uint64_t>(T) * B, 2 * L, 64, Q_ROWS) ||
      !cached_tmap(&to_dc, a.dC_bf, L, static_cast<uint64_t>(T) * B, L, 64, Q_ROWS))
    return cudaErrorInvalidValue;
  PairGruArgs g{};
  g.B = B; g.row_end = B; g.L = L; g.T = T; g.q_len = a.q_len; g.counter = a.counter;
  g.h_f32 = const_cast<float*>(a.h_f32); g.r = const_cast<float*>(a.r); g.u = const_cast<float*>(a.u);
  g.c = const_cast<float*>(a.c); g.dq = a.dq; g.dq2 = a.dq2; g.dG_bf = a.dG_bf; g.dC_bf = a.dC_bf; g.bias_part = a.bias_part;
  g.trace = g_gru_trace ? g_gru_trace + (1 << 17) : nullptr;
  g.dbg = getenv("VQA_GRU_DBG") ? atoi(getenv("VQA_GRU_DBG")) : 0;
  g.tma_out = (B % Q_ROWS == 0) ? 1 : 0;
  g.kbs = kbs;
  return launch_chunks<1>(tm_dc, tm_dg, tm_wc, tm_wg, sw_wc, sw_wg, to_dg, to_dc, g, num_sms, s);
}

}  // namespace vqa
