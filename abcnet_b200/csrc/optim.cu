// Optimiser step and per-step weight re-layout of the training path.
//
//  * abc_adam_step: torch.optim.Adam(lr, betas, eps, weight_decay) of /root/reference/src/train.py:55,141 for ALL
//    parameter tensors in one launch (a chunk table maps thread blocks to (tensor, offset)), step counter on the device
//    so that the launch can be replayed inside a CUDA graph.
//  * abc_gather_pack: the kernel-ready bf16 weight blocks (and fp32 padded / folded bias vectors) of every convolution of
//    the forward AND backward pass are rebuilt from the fp32 master parameters by ONE gather: out[i] = src[code[i]], where
//    code[i] = (tensor id << 22 | element offset) was derived once by running the host-side packing code on index tensors.
//    Replaces the ~1000 tiny slice / permute / cat / cast kernels per iteration of a PyTorch-level re-pack.
#include <cuda_bf16.h>

#include "common.cuh"

namespace abc {

constexpr uint32_t kCodeZero = 0xFFFFFFFFu;
constexpr int kCodeShift = 22;

template <typename OutT>
__global__ void __launch_bounds__(256) gather_pack_kernel(const float* const* __restrict__ src, const uint32_t* __restrict__ codes,
                                                          OutT* __restrict__ out, int64_t n8) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 c0 = __ldg(reinterpret_cast<const uint4*>(codes) + 2 * i);
    const uint4 c1 = __ldg(reinterpret_cast<const uint4*>(codes) + 2 * i + 1);
    const uint32_t c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = c[k] == kCodeZero ? 0.f : __ldg(src[c[k] >> kCodeShift] + (c[k] & ((1u << kCodeShift) - 1u)));
    if constexpr (sizeof(OutT) == 2) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
      uint4 q;
      q.x = *reinterpret_cast<uint32_t*>(&p0);
      q.y = *reinterpret_cast<uint32_t*>(&p1);
      q.z = *reinterpret_cast<uint32_t*>(&p2);
      q.w = *reinterpret_cast<uint32_t*>(&p3);
      reinterpret_cast<uint4*>(out)[i] = q;
    } else {
      reinterpret_cast<float4*>(out)[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

constexpr int kAdamChunk = 4096;          // elements per thread block

// hyper = {lr, beta1, beta2, eps, weight_decay}; step_dev = number of steps taken so far (this launch is step + 1).
__global__ void __launch_bounds__(256) adam_kernel(float* const* __restrict__ p_ptrs, const float* const* __restrict__ g_ptrs,
                                                   float* const* __restrict__ m_ptrs, float* const* __restrict__ v_ptrs,
                                                   const int64_t* __restrict__ sizes, const int2* __restrict__ chunks,
                                                   const float* __restrict__ hyper, const float* __restrict__ step_dev) {
  const int2 ck = chunks[blockIdx.x];                                  // (tensor, first element / kAdamChunk)
  float* __restrict__ p = p_ptrs[ck.x];
  const float* __restrict__ g = g_ptrs[ck.x];
  float* __restrict__ m = m_ptrs[ck.x];
  float* __restrict__ v = v_ptrs[ck.x];
  const int64_t n = sizes[ck.x];
  const int64_t i0 = static_cast<int64_t>(ck.y) * kAdamChunk;
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float step = *step_dev + 1.f;
  // torch.optim.Adam (capturable): step_size = lr / (1 - b1^t), denom = sqrt(v) / sqrt(1 - b2^t) + eps
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
#pragma unroll 4
  for (int k = threadIdx.x; k < kAdamChunk; k += 256) {
    const int64_t i = i0 + k;
    if (i >= n) break;
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i]);                               // L2 weight decay folded into the gradient
    const float mi = fmaf(1.f - b1, gi - m[i], m[i]);                  // lerp(m, g, 1 - b1)
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
  }
}

__global__ void adam_tick_kernel(float* step_dev) { *step_dev += 1.f; }

}  // namespace abc

extern "C" int abc_gather_pack(const void* src_ptrs, const void* codes, void* out, int64_t n, int out_is_bf16, void* stream) {
  using namespace abc;
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(src_ptrs && codes && out, "abc_gather_pack: null pointer");
  ABC_REQUIRE(n > 0 && n % 8 == 0, "abc_gather_pack: n=%lld must be a positive multiple of 8", static_cast<long long>(n));
  ABC_REQUIRE((reinterpret_cast<uintptr_t>(codes) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "abc_gather_pack: codes / out must be 16-byte aligned");
  const int64_t n8 = n / 8;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int64_t want = (n8 + 255) / 256;
  const int grid = static_cast<int>(want < 8LL * sms ? want : 8LL * sms);
  if (out_is_bf16)
    gather_pack_kernel<__nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float* const*>(src_ptrs), static_cast<const uint32_t*>(codes), static_cast<__nv_bfloat16*>(out), n8);
  else
    gather_pack_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float* const*>(src_ptrs), static_cast<const uint32_t*>(codes), static_cast<float*>(out), n8);
  return launch_check("gather_pack_kernel");
}

extern "C" int abc_adam_chunk_elems(void) { return abc::kAdamChunk; }

extern "C" int abc_adam_step(const void* p_ptrs, const void* g_ptrs, const void* m_ptrs, const void* v_ptrs, const int64_t* sizes,
                             const int32_t* chunks, int n_chunks, const float* hyper, float* step_dev, void* stream) {
  using namespace abc;
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(p_ptrs && g_ptrs && m_ptrs && v_ptrs && sizes && chunks && hyper && step_dev, "abc_adam_step: null pointer");
  ABC_REQUIRE(n_chunks > 0, "abc_adam_step: n_chunks=%d", n_chunks);
  adam_kernel<<<n_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<float* const*>(p_ptrs), static_cast<const float* const*>(g_ptrs), static_cast<float* const*>(m_ptrs),
      static_cast<float* const*>(v_ptrs), sizes, reinterpret_cast<const int2*>(chunks), hyper, step_dev);
  if (int rc = launch_check("adam_kernel")) return rc;
  adam_tick_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(step_dev);
  return launch_check("adam_tick_kernel");
}
