// Device helpers shared by the kernels that touch P8 bf16 activations outside the tensor-core path: 8-channel vector
// (un)packing, activation derivatives and the counter-based dropout mask of the head BatchNorm (src/unet.py:69).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

// fp32 -> one 16-bit activation value in the storage format of the launch: bf16 (default) or IEEE fp16 (AbcConvDesc.act_fp16;
// saturating, fp16 has no headroom beyond 65504). Returned as raw bits.
__device__ __forceinline__ unsigned short to_act16(float v, bool fp16) {
  if (fp16) return __half_as_ushort(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ uint32_t pack2_act16(float a, float b, bool fp16) {
  return static_cast<uint32_t>(to_act16(a, fp16)) | (static_cast<uint32_t>(to_act16(b, fp16)) << 16);
}
__device__ __forceinline__ uint4 pack8_act16(const float* v, bool fp16) {
  uint4 r;
  if (fp16) {                                   // one (warp-uniform) branch per vector, two straight-line conversion sequences
    __half2 h[4];
    const __half2 hi = __floats2half2_rn(65504.f, 65504.f), lo = __floats2half2_rn(-65504.f, -65504.f);
#pragma unroll
    for (int i = 0; i < 4; ++i)                 // saturate on the packed pair: +-inf -> +-65504 (two instructions per pair)
      h[i] = __hmax2(__hmin2(__floats2half2_rn(v[2 * i], v[2 * i + 1]), hi), lo);
    r.x = *reinterpret_cast<uint32_t*>(&h[0]); r.y = *reinterpret_cast<uint32_t*>(&h[1]);
    r.z = *reinterpret_cast<uint32_t*>(&h[2]); r.w = *reinterpret_cast<uint32_t*>(&h[3]);
  } else {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    r.x = *reinterpret_cast<uint32_t*>(&h[0]); r.y = *reinterpret_cast<uint32_t*>(&h[1]);
    r.z = *reinterpret_cast<uint32_t*>(&h[2]); r.w = *reinterpret_cast<uint32_t*>(&h[3]);
  }
  return r;
}
// element-wise max of two packed pairs in that format
__device__ __forceinline__ uint32_t max2_act16(uint32_t a, uint32_t b, bool fp16) {
  if (fp16) {
    __half2 m = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&m);
  }
  __nv_bfloat162 m = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&m);
}

__device__ __forceinline__ float act_grad(float pre, int act) {
  if (act == 1) return pre > 0.f ? 1.f : 0.f;
  if (act == 2) return pre > 0.f ? 1.f : 0.01f;
  return 1.f;
}

// Counter-based dropout masks for the 8 channels of one P8 vector (same function in forward and backward): two 64-bit
// hashes of (seed, vector index) give eight 16-bit uniforms; channel i is kept iff its uniform >= p * 65536 (p is thereby
// quantised to 2^-16; nn.Dropout's scale 1 / (1 - p) is kept). One hash per FOUR channels instead of one per channel: the
// dropout passes of the 1024-channel head BatchNorm were ALU-bound on the 64-bit multiplies of the per-element hash.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ void drop_scale8(unsigned long long seed, unsigned long long vec_idx, float p, float* m) {
  const unsigned thr = static_cast<unsigned>(p * 65536.f);
  const float keep = 1.f / (1.f - p);
  const unsigned long long h0 = mix64((2 * vec_idx) * 0x9E3779B97F4A7C15ull + seed);
  const unsigned long long h1 = mix64((2 * vec_idx + 1) * 0x9E3779B97F4A7C15ull + seed);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = (static_cast<unsigned>(h0 >> (16 * i)) & 0xffffu) >= thr ? keep : 0.f;
    m[4 + i] = (static_cast<unsigned>(h1 >> (16 * i)) & 0xffffu) >= thr ? keep : 0.f;
  }
}

}  // namespace abc
