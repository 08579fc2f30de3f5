// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may be
// scheduled while the previous kernel of the stream is still running; everything it does before
// `griddepcontrol.wait` (barrier init, TMEM allocation, descriptor prefetch -- nothing that touches global memory)
// overlaps the predecessor's tail, and the wait returns once the predecessor has completed and flushed. The step is a
// chain of ~60 short kernels, so the launch latency + prologue hidden per boundary adds up.
// Rules kept here: (1) only kernels whose every thread executes pdl_wait() before its first global access are
// launched through launch_pdl; (2) the trigger comes right after the wait, so at most one dependent is ever parked
// behind a running kernel. VQA_PDL=0 turns the attribute off (the device instructions are then no-ops).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace vqa {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_wait();
  pdl_trigger();
}

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("VQA_PDL");
    return !(e && atoi(e) == 0);
  }();
  return on;
}

template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}

}  // namespace vqa
