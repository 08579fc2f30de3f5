// Per-SM ingest rate from L2: 2-D tensor TMA (128-byte-wide swizzled boxes) vs 1-D bulk copies.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ingest tma_ingest.cu -lcuda ; ./tma_ingest
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ void tma2d(void* dst, const void* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma3d(void* dst, const void* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: 2-D tensor boxes of 64 elem (128 B) x rows; mode 1: 1-D bulk of `chunk` bytes; mode 2: 3-D boxes
// {64 elem, rows, nkb k-blocks} of a [R, 1024] bf16 matrix (k-block stride 128 B). STAGES in flight.
template <int STAGES>
__global__ void ingest(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm3, const uint8_t* src, int mode, int rows, int chunk, int nkb, int iters, long long span, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[STAGES];
  uint8_t* buf = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const int bytes = mode == 0 ? rows * 128 : (mode == 2 ? rows * 128 * nkb : chunk);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned long long t0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    // every CTA walks its own region of an L2-resident span
    long long off = ((long long)blockIdx.x * 7919 * bytes) % span;
    int issued = 0, done = 0;
    uint32_t par[STAGES];
    for (int s = 0; s < STAGES; ++s) par[s] = 0;
    while (done < iters) {
      while (issued < iters && issued - done < STAGES) {
        int s = issued % STAGES;
        mbar_expect(&bar[s], bytes);
        if (mode == 0) { long long r = (off / 128) % (span / 128 - rows); tma2d(buf + s * bytes, &tm, &bar[s], 0, (int)r); }
        else if (mode == 2) { long long r = (off / 2048) % (span / 2048 - rows); tma3d(buf + s * bytes, &tm3, &bar[s], 0, (int)r, (int)((issued * nkb) % 16)); }
        else bulk1d(buf + s * bytes, src + (off % (span - bytes)) / 16 * 16, bytes, &bar[s]);
        off += bytes * 131;
        ++issued;
      }
      int s = done % STAGES;
      while (!mbar_try(&bar[s], par[s])) {}
      par[s] ^= 1;
      ++done;
    }
    unsigned long long t1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    out[blockIdx.x] = t1 - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const long long span = 32ll << 20;   // 32 MB: L2 resident
  uint8_t* src; cudaMalloc(&src, span); cudaMemset(src, 1, span);
  unsigned long long* out; cudaMalloc(&out, 148 * 8);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)p;
  auto run = [&](const char* name, int mode, int rows, int chunk, int stages, int grid, int nkb = 1) {
    CUtensorMap tm, tm3;
    {
      cuuint64_t d3[3] = {64, (cuuint64_t)(span / 2048), 16}; cuuint64_t s3[2] = {2048, 128};
      cuuint32_t b3[3] = {64, (cuuint32_t)(rows > 0 ? rows : 8), (cuuint32_t)nkb}; cuuint32_t e3[3] = {1, 1, 1};
      CUresult r3 = enc(&tm3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, src, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (mode == 2 && r3 != CUDA_SUCCESS) { printf("%-28s 3-D encode failed: %d\n", name, (int)r3); return; }
    }
    cuuint64_t dims[2] = {64, (cuuint64_t)(span / 128)}; cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)(rows > 0 ? rows : 8)}; cuuint32_t es[2] = {1, 1};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int bytes = mode == 0 ? rows * 128 : (mode == 2 ? rows * 128 * nkb : chunk);
    const int iters = (8 << 20) / bytes;   // 8 MB per CTA
    const int smem = stages * bytes + 2048;
    auto launch = [&](auto kern) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      for (int rep = 0; rep < 2; ++rep) kern<<<grid, 32, smem>>>(tm, tm3, src, mode, rows, chunk, nkb, iters, span, out);
    };
    if (stages == 2) launch(ingest<2>); else if (stages == 4) launch(ingest<4>); else launch(ingest<8>);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<unsigned long long> h(grid); cudaMemcpy(h.data(), out, grid * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    printf("%-28s grid %3d stages %d bytes/op %6d : %7.1f GB/s per SM, %7.2f TB/s total  (%s)\n", name, grid, stages, bytes,
           (double)iters * bytes / mx, (double)iters * bytes * grid / mx / 1e3, cudaGetErrorString(e));
  };
  for (int grid : {148}) {
    run("tensor3d 128r x 2kb (32 KB)", 2, 128, 0, 4, grid, 2);
    run("tensor3d 128r x 2kb (32 KB)", 2, 128, 0, 2, grid, 2);
    run("tensor3d 128r x 4kb (64 KB)", 2, 128, 0, 2, grid, 4);
    run("tensor3d 64r x 4kb (32 KB)", 2, 64, 0, 4, grid, 4);
    run("tensor3d 64r x 8kb (64 KB)", 2, 64, 0, 2, grid, 8);
    run("tensor3d 32r x 4kb (16 KB)", 2, 32, 0, 4, grid, 4);
    run("bulk1d 64 KB", 1, 0, 65536, 2, grid);
    run("tensor2d 256 rows (32 KB)", 0, 256, 0, 2, grid);
  }
  for (int grid : {148}) {
    run("tensor2d 64 rows (8 KB)", 0, 64, 0, 4, grid);
    run("tensor2d 64 rows (8 KB)", 0, 64, 0, 8, grid);
    run("tensor2d 128 rows (16 KB)", 0, 128, 0, 4, grid);
    run("tensor2d 256 rows (32 KB)", 0, 256, 0, 4, grid);
    run("bulk1d 8 KB", 1, 0, 8192, 4, grid);
    run("bulk1d 8 KB", 1, 0, 8192, 8, grid);
    run("bulk1d 16 KB", 1, 0, 16384, 4, grid);
    run("bulk1d 32 KB", 1, 0, 32768, 4, grid);
    run("bulk1d 2 KB", 1, 0, 2048, 8, grid);
  }
  return 0;
}
