// Device helpers shared by the kernels that touch P8 bf16 activations outside the tensor-core path: 8-channel vector
// (un)packing, activation derivatives and the counter-based dropout mask of the head BatchNorm (src/unet.py:69).
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

namespace abc {

__device__ __forceinline__ void unpack8u(const uint4& u, float* v) {   // bf16 -> fp32 is a 16-bit shift
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
  v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}

__device__ __forceinline__ float act_grad(float pre, int act) {
  if (act == 1) return pre > 0.f ? 1.f : 0.f;
  if (act == 2) return pre > 0.f ? 1.f : 0.01f;
  return 1.f;
}

// counter-based dropout mask: keep iff hash(seed, element) >= p * 2^32 (same function in forward and backward)
__device__ __forceinline__ float drop_scale(unsigned long long seed, unsigned long long idx, float p) {
  if (p <= 0.f) return 1.f;
  unsigned long long z = idx * 0x9E3779B97F4A7C15ull + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = static_cast<float>(static_cast<unsigned>(z >> 40)) * (1.f / 16777216.f);
  return u >= p ? 1.f / (1.f - p) : 0.f;
}

}  // namespace abc
