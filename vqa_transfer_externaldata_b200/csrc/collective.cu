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

}  // namespace

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
