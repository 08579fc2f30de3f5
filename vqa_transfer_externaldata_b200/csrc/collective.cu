// Gradient all-reduce over NVSwitch multicast memory (the one collective of the training path).
//
// The flat trainable-gradient buffer of every rank lives in symmetric memory that is also mapped as ONE multicast
// object (torch.distributed._symmetric_memory does the rendezvous and the mapping: plumbing). Each rank owns 1/world
// of the elements:  multimem.ld_reduce pulls that slice from ALL ranks with the sum formed inside the switch, and
// multimem.st broadcasts the result back to all of them. Every GPU therefore moves ~n bytes in and ~n bytes out once
// (NCCL's ring moves 2 (w-1)/w n through w-1 dependent hops: 188 us for the 39 MB of cfg1 on 8 GPUs, and 40 us of
// that is fixed latency), every element is reduced exactly once in a hardware-fixed order, and all ranks receive
// bit-identical sums. The caller brackets the kernel with the symmetric-memory barrier (gradients of all ranks are
// complete before / all slices are broadcast after).
#include <cuda_runtime.h>

#include <cstdlib>

#include "internal.h"

namespace vqa {

namespace {

constexpr int MC_THREADS = 512;
constexpr int MC_UNROLL = 8;

__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// [begin4, end4): this rank's slice in units of float4
__global__ void __launch_bounds__(MC_THREADS) multimem_allreduce_kernel(float* mc, long long begin4, long long end4) {
  const long long stride = static_cast<long long>(gridDim.x) * MC_THREADS;
  long long i = begin4 + static_cast<long long>(blockIdx.x) * MC_THREADS + threadIdx.x;
  for (; i + (MC_UNROLL - 1) * stride < end4; i += MC_UNROLL * stride) {
    float4 v[MC_UNROLL];
#pragma unroll
    for (int u = 0; u < MC_UNROLL; ++u) v[u] = mc_ld_reduce(mc + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < MC_UNROLL; ++u) mc_st(mc + 4 * (i + u * stride), v[u]);
  }
  for (; i < end4; i += stride) mc_st(mc + 4 * i, mc_ld_reduce(mc + 4 * i));
}

// ---- the same exchange with the two cross-rank barriers INSIDE the kernel -------------------------------------------
// flags[0] / flags[1] live in the symmetric buffer behind the gradients (each rank has its own copy; the multicast
// address reaches all of them). Every call adds `world` to both: the host passes the totals after this call.
//   entry : "the gradients of every rank are complete" -- block 0 of every rank adds 1 to flags[0] of ALL ranks
//           (multimem.red, release: ordered after this rank's earlier kernels, whose writes a kernel boundary has
//           already made visible), every block waits until its rank's copy has reached the total;
//   exit  : "every slice has been broadcast" -- blocks count themselves on a local counter after their last
//           multimem.st (+ fence); the last one adds 1 to flags[1] everywhere and waits for the total, so the kernel
//           (and with it the stream) does not complete before every rank's slice has landed here.
// A spin that sees no progress for ~30 s traps (a lost rank becomes a launch failure, not a hung GPU).
__device__ __forceinline__ void mc_red_add_release(unsigned int* p, unsigned int v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until(const unsigned int* p, unsigned int target) {
  const long long t0 = clock64();
  while (static_cast<int>(ld_acquire_sys(p) - target) < 0) {
    if (clock64() - t0 > 60000000000LL) __trap();   // ~30 s: ranks of one job may be seconds apart (host-side stalls of a single rank)
  }
}

__global__ void __launch_bounds__(1024) multimem_allreduce_sync_kernel(float* mc, long long begin4, long long end4,
                                                                       unsigned int* mc_flags, const unsigned int* my_flags,
                                                                       unsigned int* grid_ctr, unsigned int flag_total,
                                                                       unsigned int grid_total) {
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) mc_red_add_release(mc_flags, 1u);
    spin_until(my_flags, flag_total);
  }
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = begin4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (MC_UNROLL - 1) * stride < end4; i += MC_UNROLL * stride) {
    float4 v[MC_UNROLL];
#pragma unroll
    for (int u = 0; u < MC_UNROLL; ++u) v[u] = mc_ld_reduce(mc + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < MC_UNROLL; ++u) mc_st(mc + 4 * (i + u * stride), v[u]);
  }
  for (; i < end4; i += stride) mc_st(mc + 4 * i, mc_ld_reduce(mc + 4 * i));
  __threadfence_system();   // this thread's broadcasts are performed before the block counts itself
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(grid_ctr, 1u);
    if (prev + 1u == grid_total) {   // the last block of this launch
      mc_red_add_release(mc_flags + 1, 1u);
      spin_until(my_flags + 1, flag_total);
    }
  }
}

}  // namespace

// cluster of two CTAs per launch unit, each claiming a whole SM (the shape the cooperative recurrent grid tolerates
// beside it: csrc/elementwise.cu gather_exclusive_launch) when `exclusive`; a plain wide grid otherwise
VqaStatus multimem_allreduce_sync_launch(float* mc, long long n, int rank, int world, unsigned int* mc_flags,
                                         const unsigned int* my_flags, unsigned int* grid_ctr, unsigned int flag_total,
                                         unsigned int* grid_total_io, bool exclusive, int ctas, cudaStream_t s) {
  const long long n4 = n >> 2;
  const long long per = (n4 + world - 1) / world;
  long long begin = per * rank;
  long long end = begin + per < n4 ? begin + per : n4;
  if (begin > end) begin = end;   // a rank without elements still takes part in the two barriers
  const int threads = exclusive ? 1024 : MC_THREADS;
  static const int ctas_env = getenv("VQA_AR_CTAS") ? atoi(getenv("VQA_AR_CTAS")) : 0;   // tuning aid for the wide launches
  if (ctas <= 0) ctas = exclusive ? 20 : (ctas_env > 0 ? ctas_env : 96);
  long long need = (end - begin + threads - 1) / threads;
  if (need < 1) need = 1;
  if (need < ctas) ctas = static_cast<int>(need);
  if (exclusive) ctas = (ctas + 1) & ~1;
  *grid_total_io += static_cast<unsigned int>(ctas);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(threads);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  int nat = 0;
  if (exclusive) {
    constexpr int kExclusiveSmem = 200 * 1024;
    VQA_CUDA_CHECK(cudaFuncSetAttribute(multimem_allreduce_sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kExclusiveSmem));
    cfg.dynamicSmemBytes = kExclusiveSmem;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    nat = 1;
  }
  cfg.attrs = at;
  cfg.numAttrs = nat;
  VQA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, multimem_allreduce_sync_kernel, mc, begin, end, mc_flags, my_flags, grid_ctr,
                                    flag_total, *grid_total_io));
  count_launch();
  return VQA_OK;
}

}  // namespace vqa

using namespace vqa;

extern "C" VQA_API VqaStatus vqa_multimem_all_reduce(void* multicast_ptr, int64_t n, int32_t rank, int32_t world,
                                                     int32_t num_ctas, void* stream) {
  if (!multicast_ptr || n <= 0 || world <= 0 || rank < 0 || rank >= world)
    return set_error(VQA_ERR_BAD_ARG, "vqa_multimem_all_reduce: bad argument");
  if ((n & 3) || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15))
    return set_error(VQA_ERR_BAD_SHAPE, "vqa_multimem_all_reduce: n %% 4 == 0 and a 16-byte aligned buffer are required");
  const long long n4 = n >> 2;
  const long long per = (n4 + world - 1) / world;
  const long long begin = per * rank;
  const long long end = begin + per < n4 ? begin + per : n4;
  if (begin >= end) return VQA_OK;
  int ctas = num_ctas > 0 ? num_ctas : 96;
  const long long need = (end - begin + MC_THREADS - 1) / MC_THREADS;
  if (need < ctas) ctas = static_cast<int>(need);
  multimem_allreduce_kernel<<<ctas, MC_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<float*>(multicast_ptr), begin, end);
  VQA_LAUNCH_CHECK("multimem_all_reduce");
  return VQA_OK;
}
