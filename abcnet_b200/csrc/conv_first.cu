// First convolution of the U-Net: 1 -> 16 channels on the binary 1 x H x W image, 3x3, pad 1, BatchNorm folded, ReLU.
// Replaces nn.Conv2d(1,16,3,padding=1) + BatchNorm2d + ReLU of inc1 (/root/reference/src/unet.py:12-14 via :83,:101).
// Bandwidth class: reads 4 B/pixel (fp32 image as the reference's DataLoader delivers it, utils.py:80-81) and writes
// 32 B/pixel (two bf16 P8 planes). conv3x3_c1_rows_kernel (table-driven, one warp per 128-pixel-wide strip) is the product
// path; conv3x3_c1_kernel (one thread per pixel, plain arithmetic) is kept as the readable reference (ABCNET_C1_SIMPLE=1).
#include <cstdlib>
#include <cuda_bf16.h>

#include "common.cuh"
#include "p8_device.cuh"

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
    v[co] = (relu & 1) ? fmaxf(a, 0.f) : a;
  }
#pragma unroll
  for (int pl = 0; pl < 2; ++pl)
    out[((static_cast<size_t>(n) * out_planes + out_plane_off + pl) * H + y) * W + x] = pack8_act16(v + pl * 8, (relu & 2) != 0);
}

// Streaming variant (the default): one warp owns a 128-pixel-wide, kC1Rows-high strip and walks down its rows.
//  * The input is a binarised drawing (utils.py:80-81: values in {0, 1}), so the 16 outputs of a pixel are a function of
//    the 9-bit pattern of its 3x3 neighbourhood: a 512-entry table of finished (bias + taps, ReLU, bf16-packed) outputs
//    is built once per block in shared memory (16 KB). Table entries are accumulated in tap order with the same fmaf
//    sequence as the arithmetic path (fmaf(0, w, a) == a), so both paths give identical bits.
//  * Lane l loads pixels x0 + 32k + l (k = 0..3) of an input row -- four fully coalesced 128-byte requests -- and four
//    ballots turn the row into a 128-bit ink mask held by every lane (plus the two pixels left / right of the strip).
//    Three such row masks (a rotating window) give every lane the 9-bit pattern of its four output pixels by plain shifts.
//  * Stores are lane-contiguous: one STG.128 per (k, plane) writes 512 contiguous bytes. (The earlier four-pixels-per-
//    thread kernel stored 16-byte pieces 64 bytes apart: ncu showed the L2 at 59 % busy for 35 % of the DRAM peak.)
//  * Rows whose 3-row window holds any value other than 0 / 1 take the arithmetic path (same fmaf order), so arbitrary
//    fp32 images still work.
constexpr int kC1Rows = 16;

// `relu` carries two flags in all stem kernels: bit 0 = ReLU, bit 1 = fp16 output instead of bf16 (abc_conv3x3_stem)
__device__ __forceinline__ uint4 c1_pack8(const float* a, int flags) { return pack8_act16(a, (flags & 2) != 0); }

// Arithmetic path for one pixel (non-binary inputs); ws = [tap][co] weights followed by bias[co]. Out of line: rare.
template <typename T>
__device__ __noinline__ void c1_generic_pixel(const T* __restrict__ im, int y, int x, int H, int W, const float* ws, int relu,
                                              uint4* o0, uint4* o1) {
  float t[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
    t[k] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? static_cast<float>(__ldg(im + static_cast<size_t>(yy) * W + xx)) : 0.f;
  }
#pragma unroll
  for (int pl = 0; pl < 2; ++pl) {
    float v[8];
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      const int co = pl * 8 + c8;
      float a = ws[144 + co];
#pragma unroll
      for (int k = 0; k < 9; ++k) a = fmaf(t[k], ws[k * 16 + co], a);
      v[c8] = (relu & 1) ? fmaxf(a, 0.f) : a;
    }
    *(pl ? o1 : o0) = c1_pack8(v, relu);
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 4) conv3x3_c1_rows_kernel(const T* __restrict__ img, const float* __restrict__ w,
                                                              const float* __restrict__ b, uint4* __restrict__ out, int N, int H, int W,
                                                              int out_planes, int out_plane_off, int relu) {
  __shared__ __align__(16) float ws[9 * 16 + 16];          // [tap][co], then bias[co]
  __shared__ __align__(16) uint4 lut[512][2];              // [pattern][plane]: 8 bf16 channels each
  const int tid = threadIdx.x;
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
        v[c8] = (relu & 1) ? fmaxf(a, 0.f) : a;
      }
      lut[e][pl] = c1_pack8(v, relu);
    }
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int tiles_x = (W + 127) >> 7, strips = (H + kC1Rows - 1) / kC1Rows;
  const long long items = static_cast<long long>(N) * strips * tiles_x;
  const size_t plane = static_cast<size_t>(H) * W;
  for (long long it = static_cast<long long>(blockIdx.x) * 8 + warp; it < items; it += static_cast<long long>(gridDim.x) * 8) {
    const int n = static_cast<int>(it / (strips * tiles_x));
    const int rem = static_cast<int>(it - static_cast<long long>(n) * (strips * tiles_x));
    const int x0 = (rem % tiles_x) << 7, y0 = (rem / tiles_x) * kC1Rows;
    const int y1 = min(y0 + kC1Rows, H);
    const T* im = img + static_cast<size_t>(n) * plane;
    uint4* o_img = out + (static_cast<size_t>(n) * out_planes + out_plane_off) * plane;
    const int xe = lane == 0 ? x0 - 1 : x0 + 128;           // lanes 0 / 1 also fetch the pixels beside the strip
    const bool edge_ok = lane < 2 && xe >= 0 && xe < W;
    uint32_t m[3][4], e[3];                                 // ink masks of rows y-1, y, y+1 (bit l of m[r][k] = pixel x0+32k+l)
    bool nb[3];                                             // row holds a value other than 0 / 1
    auto load_row = [&](int yy, uint32_t (&mm)[4], uint32_t& ee, bool& nbin) {
      float v[4], ve = 0.f;
      const bool in = yy >= 0 && yy < H;
      const T* row = im + static_cast<size_t>(in ? yy : 0) * W;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = x0 + 32 * k + lane;
        v[k] = (in && x < W) ? static_cast<float>(__ldg(row + x)) : 0.f;
      }
      if (in && edge_ok) ve = static_cast<float>(__ldg(row + xe));
      bool bad = (ve != 0.f) && (ve != 1.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        mm[k] = __ballot_sync(0xffffffffu, v[k] != 0.f);
        bad = bad || ((v[k] != 0.f) && (v[k] != 1.f));
      }
      ee = __ballot_sync(0xffffffffu, ve != 0.f);
      nbin = __any_sync(0xffffffffu, bad);
    };
    load_row(y0 - 1, m[0], e[0], nb[0]);
    load_row(y0, m[1], e[1], nb[1]);
    for (int y = y0; y < y1; ++y) {
      load_row(y + 1, m[2], e[2], nb[2]);
      uint4* o0 = o_img + static_cast<size_t>(y) * W;
      uint4* o1 = o0 + plane;
      if (!(nb[0] || nb[1] || nb[2])) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t pat = 0;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const uint32_t lo = k == 0 ? (e[r] & 1u) : (m[r][k - 1] >> 31);
            const uint32_t hi = k == 3 ? ((e[r] >> 1) & 1u) : (m[r][k + 1] & 1u);
            // bits (x-1, x, x+1) of the row: the 34-bit string [hi | 32 pixels | lo] shifted right by the lane index
            const uint64_t wide = (static_cast<uint64_t>(hi) << 33) | (static_cast<uint64_t>(m[r][k]) << 1) | lo;
            pat |= (static_cast<uint32_t>(wide >> lane) & 7u) << (3 * r);
          }
          const int x = x0 + 32 * k + lane;
          if (x < W) {
            o0[x] = lut[pat][0];
            o1[x] = lut[pat][1];
          }
        }
      } else {
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          const int x = x0 + 32 * k + lane;
          if (x < W) c1_generic_pixel<T>(im, y, x, H, W, ws, relu, o0 + x, o1 + x);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        m[0][k] = m[1][k];
        m[1][k] = m[2][k];
      }
      e[0] = e[1]; e[1] = e[2];
      nb[0] = nb[1]; nb[1] = nb[2];
    }
  }
}

// General stem: Cin (1..8) real-valued fp32 input channels -> 16 channels. UNet(in_channels=3) of the reference's own
// self-check (/root/reference/src/unet.py:122-134) and any non-binarised input take this path; bandwidth class like the
// binary kernel (4 * Cin B read + 32 B written per pixel). One thread per pixel, weights [16][Cin][9] in shared memory,
// fp32 fmaf accumulation in (ci, ky, kx) order.
__global__ void __launch_bounds__(256) conv3x3_cn_kernel(const float* __restrict__ img, int cin, const float* __restrict__ w,
                                                         const float* __restrict__ b, uint4* __restrict__ out, int H, int W,
                                                         int out_planes, int out_plane_off, int relu) {
  __shared__ float ws[16 * 8 * 9 + 16];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int i = tid; i < 16 * cin * 9; i += 256) ws[i] = w[i];
  if (tid < 16) ws[16 * 8 * 9 + tid] = b[tid];
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int n = blockIdx.z;
  if (x >= W || y >= H) return;
  float v[16];
#pragma unroll
  for (int co = 0; co < 16; ++co) v[co] = ws[16 * 8 * 9 + co];
  for (int ci = 0; ci < cin; ++ci) {
    const float* im = img + (static_cast<size_t>(n) * cin + ci) * H * W;
    float t[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int yy = y + dy - 1, xx = x + dx - 1;
        t[dy * 3 + dx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(im + static_cast<size_t>(yy) * W + xx) : 0.f;
      }
#pragma unroll
    for (int co = 0; co < 16; ++co) {
      float a = v[co];
#pragma unroll
      for (int k = 0; k < 9; ++k) a = fmaf(t[k], ws[(co * cin + ci) * 9 + k], a);
      v[co] = a;
    }
  }
  if (relu & 1) {
#pragma unroll
    for (int co = 0; co < 16; ++co) v[co] = fmaxf(v[co], 0.f);
  }
#pragma unroll
  for (int pl = 0; pl < 2; ++pl)
    out[((static_cast<size_t>(n) * out_planes + out_plane_off + pl) * H + y) * W + x] = pack8_act16(v + pl * 8, (relu & 2) != 0);
}

}  // namespace abc

extern "C" int abc_conv3x3_cn(const float* img, int cin, const float* w, const float* b, void* out, int N, int H, int W,
                              int out_planes, int out_plane_off, int relu, void* stream) {
  using namespace abc;
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(img && w && b && out, "abc_conv3x3_cn: null pointer");
  ABC_REQUIRE(cin >= 1 && cin <= 8, "abc_conv3x3_cn: cin=%d must be in 1..8", cin);
  ABC_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535, "abc_conv3x3_cn: bad geometry N=%d H=%d W=%d", N, H, W);
  ABC_REQUIRE(out_plane_off >= 0 && out_plane_off + 2 <= out_planes, "abc_conv3x3_cn: output plane range");
  ABC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "abc_conv3x3_cn: output must be 16-byte aligned");
  dim3 grid((W + 31) / 32, (H + 7) / 8, N);
  conv3x3_cn_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(img, cin, w, b, static_cast<uint4*>(out), H, W,
                                                                                  out_planes, out_plane_off, relu);
  return launch_check("conv3x3_cn_kernel");
}

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
  static const bool simple = getenv("ABCNET_C1_SIMPLE") != nullptr;      // debugging: the one-thread-per-pixel kernel
  if (!simple) {
    // persistent blocks (the 512-entry output table is built once per block): four per SM, one warp per 128 x 16 strip
    const long long items = static_cast<long long>((W + 127) / 128) * ((H + kC1Rows - 1) / kC1Rows) * N;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const long long want = (items + 7) / 8;
    const int gx = static_cast<int>(want < 4LL * sms ? want : 4LL * sms);
    conv3x3_c1_rows_kernel<T><<<gx, 256, 0, static_cast<cudaStream_t>(stream)>>>(img, w, b, static_cast<uint4*>(out), N, H, W,
                                                                                 out_planes, out_plane_off, relu);
    return launch_check("conv3x3_c1_rows_kernel");
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


extern "C" int abc_conv3x3_stem(const void* img, int img_is_u8, int cin, const float* w, const float* b, void* out, int N, int H, int W,
                                int out_planes, int out_plane_off, int flags, void* stream) {
  ABC_REQUIRE(flags >= 0 && flags <= 3, "abc_conv3x3_stem: flags=%d (bit 0 ReLU, bit 1 fp16 output)", flags);
  if (cin != 1) {
    ABC_REQUIRE(!img_is_u8, "abc_conv3x3_stem: uint8 images are the 1-channel binarised format");
    return abc_conv3x3_cn(static_cast<const float*>(img), cin, w, b, out, N, H, W, out_planes, out_plane_off, flags, stream);
  }
  if (img_is_u8)
    return conv3x3_c1_launch<uint8_t>(static_cast<const uint8_t*>(img), w, b, out, N, H, W, out_planes, out_plane_off, stream, flags);
  return conv3x3_c1_launch<float>(static_cast<const float*>(img), w, b, out, N, H, W, out_planes, out_plane_off, stream, flags);
}
