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

// Fast variant (W % 4 == 0): one thread computes FOUR horizontally adjacent pixels.
//  * The input is a binarised drawing (utils.py:80-81: values in {0, 1}), so the 16 outputs of a pixel are a function of
//    the 9-bit pattern of its 3x3 neighbourhood: a 512-entry table of finished (bias + taps, ReLU, bf16-packed) outputs
//    is built once per block in shared memory (16 KB) and a pixel costs ~10 integer instructions + two LDS.128 + two
//    STG.128 -- the kernel becomes a pure streaming kernel. Table entries are accumulated in tap order with the same
//    fmaf sequence as the arithmetic path (fmaf(0, w, a) == a), so both paths give identical bits.
//  * A warp that sees any value other than 0 / 1 falls back to the arithmetic path for its pixels (one LDS.128 of
//    weights feeds 16 FMAs), so arbitrary fp32 images still work.
__device__ __forceinline__ uint4 c1_pack8(const float* a) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(a[0], a[1]);
  __nv_bfloat162 p1 = __floats2bfloat162_rn(a[2], a[3]);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(a[4], a[5]);
  __nv_bfloat162 p3 = __floats2bfloat162_rn(a[6], a[7]);
  uint4 q;
  q.x = *reinterpret_cast<uint32_t*>(&p0);
  q.y = *reinterpret_cast<uint32_t*>(&p1);
  q.z = *reinterpret_cast<uint32_t*>(&p2);
  q.w = *reinterpret_cast<uint32_t*>(&p3);
  return q;
}

// Arithmetic path for four pixels (non-binary inputs): one LDS.128 of weights feeds 16 FMAs. Kept out of line so that its
// registers do not burden the table path.
__device__ __noinline__ void c1_fma_path(const float (&t)[3][6], const float4* w4, int relu, uint4* o0, uint4* o1) {
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
    uint4* o = pl ? o1 : o0;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float* a = acc[px];
      if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], 0.f);
      }
      o[px] = c1_pack8(a);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 3) conv3x3_c1_x4_kernel(const T* __restrict__ img, const float* __restrict__ w,
                                                            const float* __restrict__ b, uint4* __restrict__ out, int N, int H, int W,
                                                            int out_planes, int out_plane_off, int relu) {
  __shared__ __align__(16) float ws[9 * 16 + 16];          // [tap][co], then bias[co]
  __shared__ __align__(16) uint4 lut[512][2];              // [pattern][plane]: 8 bf16 channels each
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < 144) ws[(tid % 9) * 16 + tid / 9] = w[tid];    // w is [co][tap]
  if (tid < 16) ws[144 + tid] = b[tid];
  __syncthreads();
#pragma unroll 1
  for (int e = tid; e < 512; e += 256) {
#pragma unroll 1
    for (int pl = 0; pl < 2; ++pl) {
      float v[8];
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const int co = pl * 8 + c8;
        float a = ws[144 + co];
#pragma unroll
        for (int k = 0; k < 9; ++k) a = fmaf(((e >> k) & 1) ? 1.f : 0.f, ws[k * 16 + co], a);
        v[c8] = relu ? fmaxf(a, 0.f) : a;
      }
      lut[e][pl] = c1_pack8(v);
    }
  }
  __syncthreads();
  const int tiles_x = (W / 4 + 31) / 32, tiles_y = (H + 7) / 8;
  const int tiles = tiles_x * tiles_y * N;
  const float4* w4 = reinterpret_cast<const float4*>(ws);
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int rem = tile - n * (tiles_x * tiles_y);
    const int x0 = ((rem % tiles_x) * 32 + threadIdx.x) * 4;
    const int y = (rem / tiles_x) * 8 + threadIdx.y;
    const bool inside = x0 < W && y < H;
    const T* im = img + static_cast<size_t>(n) * H * W;
    float t[3][6];
    bool binary = true;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if (inside && yy >= 0 && yy < H) {
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
#pragma unroll
      for (int i = 0; i < 6; ++i) binary = binary && (t[dy][i] == 0.f || t[dy][i] == 1.f);
    }
    const bool warp_binary = __all_sync(0xffffffffu, binary);
    if (!inside) continue;
    uint4* o0 = out + ((static_cast<size_t>(n) * out_planes + out_plane_off) * H + y) * W + x0;
    uint4* o1 = o0 + static_cast<size_t>(H) * W;
    if (warp_binary) {
      uint32_t rb[3];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        uint32_t r = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) r |= (t[dy][i] != 0.f ? 1u : 0u) << i;
        rb[dy] = r;
      }
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        const uint32_t e = ((rb[0] >> px) & 7u) | (((rb[1] >> px) & 7u) << 3) | (((rb[2] >> px) & 7u) << 6);
        o0[px] = lut[e][0];
        o1[px] = lut[e][1];
      }
      continue;
    }
    c1_fma_path(t, w4, relu, o0, o1);
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
    // persistent blocks (the 512-entry output table is built once per block): a few per SM, looping over 128 x 8 pixel tiles
    const long long tiles = static_cast<long long>((W / 4 + 31) / 32) * ((H + 7) / 8) * N;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const int gx = static_cast<int>(tiles < 8LL * sms ? tiles : 8LL * sms);
    conv3x3_c1_x4_kernel<T><<<gx, block, 0, static_cast<cudaStream_t>(stream)>>>(img, w, b, static_cast<uint4*>(out), N, H, W,
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
