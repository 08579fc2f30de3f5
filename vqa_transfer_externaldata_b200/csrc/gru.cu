// Persistent recurrent kernels for the GRU question encoder (vlmap/modules.py:124-140: GRUCell under
// dynamic_rnn) -- forward over T steps and back-propagation through time -- each as ONE cooperative launch.
//
// Why: a step is two DEPENDENT [B, L] x [L, 2L | L] matmuls with element-wise gate math in between; as
// separate launches each 1-2 GFLOP GEMM costs ~20 us of fixed overhead (launch gap, TMEM alloc, barrier
// init, pipeline fill), 65 % of the whole train step in the first profile. Here every CTA keeps its TMEM
// allocation, mbarrier ring and tensor maps for the whole sequence, owns a fixed 128 x 64 output tile, and
// walks the 2T phases; the gate math is the GEMM epilogue; phases are separated by a device-wide barrier
// (all CTAs are co-resident: cooperative launch, grid <= #SMs).
//
// Per phase and CTA: TMA (128B-swizzled 2-D tiles, 6-stage mbarrier ring) -> tcgen05.mma (128x64x16, bf16,
// fp32 accumulate in TMEM) -> tcgen05.ld -> epilogue math -> global stores -> release on the grid counter.
// Weight tiles of the next phase are requested BEFORE waiting on the barrier (they do not depend on it).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "internal.h"
#include "ptx.cuh"

namespace vqa {

namespace {

constexpr int R_BM = 128, R_BN = 64, R_BK = 64;
constexpr int R_STAGES = 6;
constexpr int R_A_TILE = R_BM * R_BK * 2;  // 16 KB
constexpr int R_B_TILE = R_BN * R_BK * 2;  // 8 KB
constexpr int R_STAGE_BYTES = R_A_TILE + R_B_TILE;
constexpr int R_THREADS = 192;
constexpr int R_SMEM = R_STAGES * R_STAGE_BYTES + 1024 + 256;

struct GruArgs {
  int B, L, T;
  const int* q_len;
  unsigned int* counter;  // device-wide phase counter, zeroed before launch
  // forward
  const float* xg;   // [T*B, 2L] x-part of the gate pre-activations (+bias)
  const float* xc;   // [T*B, L]
  float* h_f32;      // [(T+1)*B, L]
  bf16* h_bf;        // [(T+1)*B, L]
  bf16* rh_bf;       // [T*B, L]
  float* r; float* u; float* c;  // [T*B, L]
  // backward
  float* du;         // [B, L]
  float* dh_part;    // [B, L]
  float* dG_f32;     // [T*B, 2L]
  bf16* dG_bf;       // [T*B, 2L]
  float* dC_f32;     // [T*B, L]
  bf16* dC_bf;       // [T*B, L]
};

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
__device__ __forceinline__ void epi_bar_sync() {  // the 4 epilogue warps only
  asm volatile("bar.sync 1, 128;" ::: "memory");
}
__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + __expf(-x)); }

// thread-private row segment helpers: 16 consecutive floats / bf16 of one row
__device__ __forceinline__ void ld16(const float* p, float (&x)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(p + 4 * j);
    x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
  }
}
__device__ __forceinline__ void st16(float* p, const float (&x)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<float4*>(p + 4 * j) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
}
__device__ __forceinline__ void st16_bf(bf16* p, const float (&x)[16]) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      h[q] = __nv_bfloat162(__float2bfloat16_rn(x[8 * j + 2 * q]), __float2bfloat16_rn(x[8 * j + 2 * q + 1]));
    *reinterpret_cast<uint4*>(p + 8 * j) = *reinterpret_cast<uint4*>(h);
  }
}
// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&x)[16]) {
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]);
}

// MODE 0: forward. phase 2t: G = h_t Wg_h (+xg) -> r,u,rh ; phase 2t+1: C = rh_t Wc_h (+xc) -> c, h_{t+1}
// MODE 1: BPTT.    phase 2i: dRH = dC_t Wc_h^T -> dG_t, dh_part ; phase 2i+1: dh = dG_t Wg_h^T + dh_part ->
//                  (prepare step t-1) du, dC_{t-1}, dh_part          with t = T-1-i
template <int MODE>
__global__ void __launch_bounds__(R_THREADS, 1) gru_persistent_kernel(
    const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
    const __grid_constant__ CUtensorMap tm_b0, const __grid_constant__ CUtensorMap tm_b1, GruArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + R_STAGES * R_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + R_STAGES;
  uint64_t* tmem_full_bar = empty_bar + R_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ni = blockIdx.x, mi = blockIdx.y;
  const int m0 = mi * R_BM, n0 = ni * R_BN;
  const unsigned int ncta = gridDim.x * gridDim.y;
  const int B = g.B, L = g.L, T = g.T;
  const int num_phases = 2 * T;
  // output tiles per phase kind: forward gates 2L/64, everything else L/64
  const int ntiles0 = (MODE == 0) ? (2 * L) / R_BN : L / R_BN;
  const int ntiles1 = L / R_BN;
  const int kb0 = L / R_BK;
  const int kb1 = (MODE == 0) ? L / R_BK : (2 * L) / R_BK;
  constexpr bool B_MN = (MODE == 0);  // forward reads TF [in,out] weights as MN-major B; BPTT as K-major B

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a0);
    ptx::prefetch_tensormap(&tm_a1);
    ptx::prefetch_tensormap(&tm_b0);
    ptx::prefetch_tensormap(&tm_b1);
    for (int s = 0; s < R_STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, R_BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase_bit = 0;
      for (int p = 0; p < num_phases; ++p) {
        const int kind = p & 1;
        if (ni >= (kind ? ntiles1 : ntiles0)) continue;  // idle in this phase
        const int step = p >> 1;
        const int t = (MODE == 0) ? step : (T - 1 - step);
        const CUtensorMap* ta = kind ? &tm_a1 : &tm_a0;
        const CUtensorMap* tb = kind ? &tm_b1 : &tm_b0;
        const int nkb = kind ? kb1 : kb0;
        const int arow = t * B + m0;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase_bit ^ 1);
          uint8_t* st = smem + stage * R_STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], R_STAGE_BYTES);
          // weights first: independent of the previous phase
          if (B_MN) ptx::tma_load_2d(st + R_A_TILE, tb, &full_bar[stage], n0, kb * R_BK);
          else ptx::tma_load_2d(st + R_A_TILE, tb, &full_bar[stage], kb * R_BK, n0);
          if (kb == 0 && p > 0) {
            // the A operand of this phase was written by the epilogues of phase p-1 (all CTAs)
            wait_counter(g.counter, static_cast<unsigned int>(p) * ncta);
            fence_proxy_async_all();
          }
          ptx::tma_load_2d(st, ta, &full_bar[stage], kb * R_BK, arow);
          if (++stage == R_STAGES) {
            stage = 0;
            phase_bit ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(R_BM, R_BN, false, B_MN);
      constexpr uint32_t B_LBO = B_MN ? 8192 : 16, B_STEP = B_MN ? 2048 : 32;
      int stage = 0;
      uint32_t phase_bit = 0;
      for (int p = 0; p < num_phases; ++p) {
        const int kind = p & 1;
        if (ni >= (kind ? ntiles1 : ntiles0)) continue;
        const int nkb = kind ? kb1 : kb0;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase_bit);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * R_STAGE_BYTES);
          const uint32_t sb = sa + R_A_TILE;
#pragma unroll
          for (int kk = 0; kk < R_BK / 16; ++kk) {
            const uint64_t da = ptx::make_smem_desc_sw128(sa + kk * 32, 16, 1024);
            const uint64_t db = ptx::make_smem_desc_sw128(sb + kk * B_STEP, B_LBO, 1024);
            ptx::umma_f16(tmem_base, da, db, idesc, (kb | kk) != 0);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == R_STAGES) {
            stage = 0;
            phase_bit ^= 1;
          }
        }
        ptx::umma_commit(tmem_full_bar);
        // the accumulator is overwritten by the next active phase only after the device-wide barrier,
        // which this CTA's own epilogue reaches after draining TMEM: no tmem_empty barrier needed
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int row = m0 + q * 32 + lane;     // sample index
    const bool row_ok = row < B;
    const int et = threadIdx.x - 64;        // 0..127 within the epilogue group
    uint32_t tfull_phase = 0;
    for (int p = 0; p < num_phases; ++p) {
      const int kind = p & 1;
      const int step = p >> 1;
      const int t = (MODE == 0) ? step : (T - 1 - step);
      const bool active = ni < (kind ? ntiles1 : ntiles0);
      // nobody may arrive for phase p before every CTA has arrived for phase p-1 (monotonic counter)
      if (p > 0) {
        if (lane == 0) wait_counter(g.counter, static_cast<unsigned int>(p) * ncta);
        __syncwarp();
      }
      if (active) {
        ptx::mbar_wait(tmem_full_bar, tfull_phase);
        tfull_phase ^= 1;
        ptx::tc_fence_after();
        const long long trow = static_cast<long long>(t) * B + row;
        const long long brow = static_cast<long long>(row) * L;
        const bool valid = row_ok && (t < g.q_len[row_ok ? row : 0]);
        const bool pvalid = row_ok && ((t - 1) < g.q_len[row_ok ? row : 0]);
        const uint32_t trow_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < R_BN; c0 += 16) {
          float acc[16];
          tmem_ld16(trow_addr + c0, acc);  // warp-collective: all lanes, also for masked rows
          if (!row_ok) continue;
          const int col = n0 + c0;
          if (MODE == 0) {
            if (kind == 0) {
              float x[16];
              ld16(g.xg + trow * 2 * L + col, x);
              if (n0 < L) {  // reset-gate columns: r = sigmoid(.), rh = r * h_t
                float hh[16];
                ld16(g.h_f32 + trow * L + col, hh);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  x[j] = sigm(acc[j] + x[j]);
                  hh[j] *= x[j];
                }
                st16(g.r + trow * L + col, x);
                st16_bf(g.rh_bf + trow * L + col, hh);
              } else {       // update-gate columns
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = sigm(acc[j] + x[j]);
                st16(g.u + trow * L + (col - L), x);
              }
            } else {
              float x[16], hh[16], uu[16];
              ld16(g.xc + trow * L + col, x);
              ld16(g.h_f32 + trow * L + col, hh);
              ld16(g.u + trow * L + col, uu);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float cc = tanhf(acc[j] + x[j]);
                x[j] = cc;
                hh[j] = valid ? uu[j] * hh[j] + (1.0f - uu[j]) * cc : hh[j];
              }
              st16(g.c + trow * L + col, x);
              st16(g.h_f32 + (trow + B) * L + col, hh);
              st16_bf(g.h_bf + (trow + B) * L + col, hh);
            }
          } else {
            if (kind == 0) {
              // dRH = acc. dr = dRH*h ; dh_part += dRH*r ; dG = [dr r(1-r), du u(1-u)]
              float hh[16], rr[16], dp[16];
              ld16(g.h_f32 + trow * L + col, hh);
              ld16(g.r + trow * L + col, rr);
              ld16(g.dh_part + brow + col, dp);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float d = valid ? acc[j] : 0.f;
                dp[j] += d * rr[j];
                hh[j] = d * hh[j] * rr[j] * (1.f - rr[j]);
              }
              st16(g.dh_part + brow + col, dp);
              st16(g.dG_f32 + trow * 2 * L + col, hh);
              st16_bf(g.dG_bf + trow * 2 * L + col, hh);
              float uu[16], dd[16];
              ld16(g.u + trow * L + col, uu);
              ld16(g.du + brow + col, dd);
#pragma unroll
              for (int j = 0; j < 16; ++j) dd[j] = valid ? dd[j] * uu[j] * (1.f - uu[j]) : 0.f;
              st16(g.dG_f32 + trow * 2 * L + L + col, dd);
              st16_bf(g.dG_bf + trow * 2 * L + L + col, dd);
            } else {
              // dh_t = acc + dh_part ; then the element-wise head of step t-1
              float dp[16];
              ld16(g.dh_part + brow + col, dp);
#pragma unroll
              for (int j = 0; j < 16; ++j) dp[j] += acc[j];
              if (t > 0) {
                const long long prow = trow - B;
                float hh[16], uu[16], cc[16], o_du[16];
                ld16(g.h_f32 + prow * L + col, hh);
                ld16(g.u + prow * L + col, uu);
                ld16(g.c + prow * L + col, cc);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float gdh = dp[j];
                  o_du[j] = pvalid ? gdh * (hh[j] - cc[j]) : 0.f;
                  hh[j] = pvalid ? gdh * (1.f - uu[j]) * (1.f - cc[j] * cc[j]) : 0.f;  // dC_{t-1}
                  dp[j] = pvalid ? gdh * uu[j] : gdh;                                   // dh_part
                }
                st16(g.du + brow + col, o_du);
                st16(g.dC_f32 + prow * L + col, hh);
                st16_bf(g.dC_bf + prow * L + col, hh);
              }
              st16(g.dh_part + brow + col, dp);
            }
          }
        }
        ptx::tc_fence_before();
      }
      // publish: generic-proxy stores -> visible device-wide and to the async proxy (TMA) of other SMs
      __threadfence();
      fence_proxy_async_all();
      epi_bar_sync();
      if (et == 0) red_release_add(g.counter, 1u);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, R_BN);
  }
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

template <int MODE>
VqaStatus launch_persistent(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0,
                            const CUtensorMap& b1, const GruArgs& g, dim3 grid, cudaStream_t s) {
  auto kern = gru_persistent_kernel<MODE>;
  static bool set = false;
  if (!set) {
    VQA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM));
    set = true;
  }
  VQA_CUDA_CHECK(cudaMemsetAsync(g.counter, 0, sizeof(unsigned int), s));
  void* args[] = {const_cast<CUtensorMap*>(&a0), const_cast<CUtensorMap*>(&a1), const_cast<CUtensorMap*>(&b0),
                  const_cast<CUtensorMap*>(&b1), const_cast<GruArgs*>(&g)};
  // cooperative launch: fails instead of deadlocking if the grid cannot be co-resident
  VQA_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), grid, dim3(R_THREADS), args, R_SMEM, s));
  count_launch();
  return VQA_OK;
}

}  // namespace

bool gru_persistent_supported(int B, int L, int precision, int num_sms) {
  if (precision != VQA_PREC_BF16) return false;
  if (L % 64 != 0 || L < 64) return false;
  const int mt = (B + R_BM - 1) / R_BM;
  return mt * ((2 * L) / R_BN) <= num_sms;
}

VqaStatus gru_fwd_persistent_launch(const GruFwdPersistent& a, cudaStream_t s) {
  const int B = a.B, L = a.L, T = a.T;
  CUtensorMap tm_h, tm_rh, tm_wg, tm_wc;
  bool ok = encode_bf16(&tm_h, a.h_bf, L, static_cast<uint64_t>(T + 1) * B, L, 64, R_BM) &&
            encode_bf16(&tm_rh, a.rh_bf, L, static_cast<uint64_t>(T) * B, L, 64, R_BM) &&
            // weights: TF [in, out] rows W.. (the h part), read as MN-major B: inner = out columns
            encode_bf16(&tm_wg, a.wg_h, 2 * L, L, 2 * L, 64, 64) &&
            encode_bf16(&tm_wc, a.wc_h, L, L, L, 64, 64);
  if (!ok) return set_error(VQA_ERR_CUDA, "gru_fwd_persistent: cuTensorMapEncodeTiled failed");
  GruArgs g{};
  g.B = B; g.L = L; g.T = T; g.q_len = a.q_len; g.counter = a.counter; g.xg = a.xg; g.xc = a.xc;
  g.h_f32 = a.h_f32; g.h_bf = a.h_bf; g.rh_bf = a.rh_bf; g.r = a.r; g.u = a.u; g.c = a.c;
  dim3 grid((2 * L) / R_BN, (B + R_BM - 1) / R_BM);
  return launch_persistent<0>(tm_h, tm_rh, tm_wg, tm_wc, g, grid, s);
}

VqaStatus gru_bwd_persistent_launch(const GruBwdPersistent& a, cudaStream_t s) {
  const int B = a.B, L = a.L, T = a.T;
  CUtensorMap tm_dc, tm_dg, tm_wc, tm_wg;
  bool ok = encode_bf16(&tm_dc, a.dC_bf, L, static_cast<uint64_t>(T) * B, L, 64, R_BM) &&
            encode_bf16(&tm_dg, a.dG_bf, 2 * L, static_cast<uint64_t>(T) * B, 2 * L, 64, R_BM) &&
            // weights as K-major B: rows = input unit (N'), contiguous = output column (K')
            encode_bf16(&tm_wc, a.wc_h, L, L, L, 64, 64) &&
            encode_bf16(&tm_wg, a.wg_h, 2 * L, L, 2 * L, 64, 64);
  if (!ok) return set_error(VQA_ERR_CUDA, "gru_bwd_persistent: cuTensorMapEncodeTiled failed");
  GruArgs g{};
  g.B = B; g.L = L; g.T = T; g.q_len = a.q_len; g.counter = a.counter;
  g.h_f32 = const_cast<float*>(a.h_f32); g.r = const_cast<float*>(a.r); g.u = const_cast<float*>(a.u);
  g.c = const_cast<float*>(a.c); g.du = a.du; g.dh_part = a.dh_part; g.dG_f32 = a.dG_f32; g.dG_bf = a.dG_bf;
  g.dC_f32 = a.dC_f32; g.dC_bf = a.dC_bf;
  dim3 grid(L / R_BN, (B + R_BM - 1) / R_BM);
  return launch_persistent<1>(tm_dc, tm_dg, tm_wc, tm_wg, g, grid, s);
}

}  // namespace vqa
