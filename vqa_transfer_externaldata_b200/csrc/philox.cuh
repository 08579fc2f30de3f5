// Counter-based dropout RNG (Philox4x32-10, Salmon et al. 2011).
// tf.nn.dropout (vlmap/modules.py:82, vqa/model_vlmap_answer.py:180) draws keep = floor(p + U[0,1));
// here every group of 8 consecutive elements shares one Philox call keyed by (seed, step, site) with
// the group index as counter: 8 x 16-bit uniforms, keep iff u16 < p * 65536. Forward and backward
// regenerate identical bits, and vqa_dropout_masks() materialises them for parity tests.
#pragma once
#include <cstdint>

namespace vqa {

struct Philox8 {
  uint32_t w[4];  // 8 x 16 bit
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox8 philox4x32_10(uint64_t group, uint32_t site,
                                                          uint64_t seed, uint64_t step) {
  uint32_t c0 = static_cast<uint32_t>(group), c1 = static_cast<uint32_t>(group >> 32);
  uint32_t c2 = site, c3 = static_cast<uint32_t>(step);
  uint32_t k0 = static_cast<uint32_t>(seed);
  uint32_t k1 = static_cast<uint32_t>(seed >> 32) ^ static_cast<uint32_t>(step >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox8 o;
  o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
  return o;
}

__host__ __device__ __forceinline__ uint32_t keep_threshold(float keep) {
  // keep == 1 must keep everything: threshold 65536 > any u16
  float t = keep * 65536.0f;
  if (t > 65536.0f) t = 65536.0f;
  if (t < 0.0f) t = 0.0f;
  return static_cast<uint32_t>(t);
}

// j in [0, 8): the j-th element of the group
__host__ __device__ __forceinline__ bool philox_keep(const Philox8& p, int j, uint32_t thr) {
  const uint32_t u = (p.w[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
  return u < thr;
}

// 8-bit keep mask of a whole group
__host__ __device__ __forceinline__ uint32_t philox_keep_bits(const Philox8& p, uint32_t thr) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) m |= static_cast<uint32_t>(philox_keep(p, j, thr)) << j;
  return m;
}

}  // namespace vqa
