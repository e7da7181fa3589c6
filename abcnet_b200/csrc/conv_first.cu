// First convolution of the U-Net: 1 -> 16 channels on the binary 1 x H x W image, 3x3, pad 1, BatchNorm folded, ReLU.
// Replaces nn.Conv2d(1,16,3,padding=1) + BatchNorm2d + ReLU of inc1 (/root/reference/src/unet.py:12-14 via :83,:101).
// Bandwidth class: reads 4 B/pixel (fp32 image as the reference's DataLoader delivers it, utils.py:80-81) and writes
// 32 B/pixel (two bf16 P8 planes); one thread per pixel, 128-bit coalesced stores.
#include <cuda_bf16.h>

#include "common.cuh"

namespace abc {

template <typename T>
__global__ void __launch_bounds__(256) conv3x3_c1_kernel(const T* __restrict__ img, const float* __restrict__ w,
                                                         const float* __restrict__ b, uint4* __restrict__ out, int H, int W,
                                                         int out_planes, int out_plane_off, int relu) {
  __shared__ float ws[16 * 9 + 16];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < 144) ws[tid] = w[tid];
  if (tid < 16) ws[144 + tid] = b[tid];
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int n = blockIdx.z;
  if (x >= W || y >= H) return;
  const T* im = img + static_cast<size_t>(n) * H * W;
  float t[9];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int yy = y + dy - 1, xx = x + dx - 1;
      t[dy * 3 + dx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? static_cast<float>(__ldg(im + static_cast<size_t>(yy) * W + xx)) : 0.f;
    }
  float v[16];
#pragma unroll
  for (int co = 0; co < 16; ++co) {
    float a = ws[144 + co];
#pragma unroll
    for (int k = 0; k < 9; ++k) a = fmaf(t[k], ws[co * 9 + k], a);
    v[co] = relu ? fmaxf(a, 0.f) : a;
  }
#pragma unroll
  for (int pl = 0; pl < 2; ++pl) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[pl * 8 + 0], v[pl * 8 + 1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(v[pl * 8 + 2], v[pl * 8 + 3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[pl * 8 + 4], v[pl * 8 + 5]);
    __nv_bfloat162 p3 = __floats2bfloat162_rn(v[pl * 8 + 6], v[pl * 8 + 7]);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&p0);
    o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2);
    o.w = *reinterpret_cast<uint32_t*>(&p3);
    out[((static_cast<size_t>(n) * out_planes + out_plane_off + pl) * H + y) * W + x] = o;
  }
}

// Fast variant (W % 4 == 0): one thread computes FOUR horizontally adjacent pixels, so that every shared-memory weight
// fetch (one LDS.128 = four output channels of one tap) feeds 16 FMAs instead of 1 -- the one-pixel kernel above is
// issue-bound on its 144 broadcast LDS per pixel (measured 1.23 ms per 256 x 512^2 batch against an HBM floor of 0.37 ms).
// The summation order per output (bias, then taps 0..8) is the same, so both kernels give identical bits.
template <typename T>
__global__ void __launch_bounds__(256) conv3x3_c1_x4_kernel(const T* __restrict__ img, const float* __restrict__ w,
                                                            const float* __restrict__ b, uint4* __restrict__ out, int H, int W,
                                                            int out_planes, int out_plane_off, int relu) {
  __shared__ __align__(16) float ws[9 * 16 + 16];          // [tap][co], then bias[co]
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < 144) ws[(tid % 9) * 16 + tid / 9] = w[tid];    // w is [co][tap]
  if (tid < 16) ws[144 + tid] = b[tid];
  __syncthreads();
  const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int n = blockIdx.z;
  if (x0 >= W || y >= H) return;
  const T* im = img + static_cast<size_t>(n) * H * W;
  float t[3][6];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int yy = y + dy - 1;
    if (yy >= 0 && yy < H) {
      const T* row = im + static_cast<size_t>(yy) * W + x0;
      if constexpr (sizeof(T) == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row));
        t[dy][1] = v.x; t[dy][2] = v.y; t[dy][3] = v.z; t[dy][4] = v.w;
      } else {
        const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(row));
        t[dy][1] = v.x; t[dy][2] = v.y; t[dy][3] = v.z; t[dy][4] = v.w;
      }
      t[dy][0] = x0 > 0 ? static_cast<float>(__ldg(row - 1)) : 0.f;
      t[dy][5] = x0 + 4 < W ? static_cast<float>(__ldg(row + 4)) : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) t[dy][i] = 0.f;
    }
  }
  const float4* w4 = reinterpret_cast<const float4*>(ws);
#pragma unroll
  for (int pl = 0; pl < 2; ++pl) {
    float acc[4][8];                                          // [pixel][channel of this plane]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 bb = w4[36 + pl * 2 + h];
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        acc[px][h * 4 + 0] = bb.x; acc[px][h * 4 + 1] = bb.y; acc[px][h * 4 + 2] = bb.z; acc[px][h * 4 + 3] = bb.w;
      }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 ww = w4[k * 4 + pl * 2 + h];
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const float v = t[k / 3][px + k % 3];
          acc[px][h * 4 + 0] = fmaf(v, ww.x, acc[px][h * 4 + 0]);
          acc[px][h * 4 + 1] = fmaf(v, ww.y, acc[px][h * 4 + 1]);
          acc[px][h * 4 + 2] = fmaf(v, ww.z, acc[px][h * 4 + 2]);
          acc[px][h * 4 + 3] = fmaf(v, ww.w, acc[px][h * 4 + 3]);
        }
      }
    }
    uint4* o = out + ((static_cast<size_t>(n) * out_planes + out_plane_off + pl) * H + y) * W + x0;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float* a = acc[px];
      if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], 0.f);
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(a[0], a[1]);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(a[2], a[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(a[4], a[5]);
      __nv_bfloat162 p3 = __floats2bfloat162_rn(a[6], a[7]);
      uint4 q;
      q.x = *reinterpret_cast<uint32_t*>(&p0);
      q.y = *reinterpret_cast<uint32_t*>(&p1);
      q.z = *reinterpret_cast<uint32_t*>(&p2);
      q.w = *reinterpret_cast<uint32_t*>(&p3);
      o[px] = q;
    }
  }
}

}  // namespace abc

template <typename T>
static int conv3x3_c1_launch(const T* img, const float* w, const float* b, void* out, int N, int H, int W, int out_planes,
                             int out_plane_off, void* stream, int relu = 1) {
  using namespace abc;
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(img && w && b && out, "abc_conv3x3_c1: null pointer");
  ABC_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535, "abc_conv3x3_c1: bad geometry N=%d H=%d W=%d", N, H, W);
  ABC_REQUIRE(out_plane_off >= 0 && out_plane_off + 2 <= out_planes, "abc_conv3x3_c1: output plane range");
  ABC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "abc_conv3x3_c1: output must be 16-byte aligned");
  dim3 block(32, 8);
  if (W % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0) {
    dim3 grid((W / 4 + 31) / 32, (H + 7) / 8, N);
    conv3x3_c1_x4_kernel<T><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(img, w, b, static_cast<uint4*>(out), H, W,
                                                                                    out_planes, out_plane_off, relu);
    return launch_check("conv3x3_c1_x4_kernel");
  }
  dim3 grid((W + 31) / 32, (H + 7) / 8, N);
  conv3x3_c1_kernel<T><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(img, w, b, static_cast<uint4*>(out), H, W,
                                                                               out_planes, out_plane_off, relu);
  return launch_check("conv3x3_c1_kernel");
}

extern "C" int abc_conv3x3_c1(const float* img, const float* w, const float* b, void* out, int N, int H, int W,
                              int out_planes, int out_plane_off, void* stream) {
  return conv3x3_c1_launch<float>(img, w, b, out, N, H, W, out_planes, out_plane_off, stream);
}

extern "C" int abc_conv3x3_c1_u8(const uint8_t* img, const float* w, const float* b, void* out, int N, int H, int W,
                                 int out_planes, int out_plane_off, void* stream) {
  return conv3x3_c1_launch<uint8_t>(img, w, b, out, N, H, W, out_planes, out_plane_off, stream);
}

// Training mode: raw convolution + bias (no activation); BatchNorm with batch statistics and ReLU follow in abc_bn_act.
extern "C" int abc_conv3x3_c1_raw(const void* img, int img_is_u8, const float* w, const float* b, void* out, int N, int H, int W,
                                  int out_planes, int out_plane_off, void* stream) {
  if (img_is_u8)
    return conv3x3_c1_launch<uint8_t>(static_cast<const uint8_t*>(img), w, b, out, N, H, W, out_planes, out_plane_off, stream, 0);
  return conv3x3_c1_launch<float>(static_cast<const float*>(img), w, b, out, N, H, W, out_planes, out_plane_off, stream, 0);
}
