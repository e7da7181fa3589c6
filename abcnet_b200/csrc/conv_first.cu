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
  dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, N);
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
