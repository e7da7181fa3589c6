// Training-mode companions of the convolution kernels (bandwidth class, P8 bf16 layout [N][planes][H][W][8]):
//   abc_bn_stats        per-channel sum / sum of squares of a conv output (batch statistics of nn.BatchNorm2d in
//                       train mode, /root/reference/src/unet.py:13,16,67), fp32 partials, fp64 accumulation
//   abc_bn_finalize     mean / biased var -> (scale, shift, mean, invstd); running-stat update (momentum 0.1, unbiased var)
//   abc_bn_act          a = act(z * scale + shift) [* dropout mask / (1-p)], optional fused MaxPool2d(2) output
//   abc_bn_act_bwd_*    backward of the same chain including max-pool routing (autograd of unet.py:11-18,30,66-69):
//                       reduce: s1 = sum g, s2 = sum g * xhat ; apply: dz = scale * (g - s1/M - xhat * s2/M)
//   abc_nchw_to_p8      fp32 NCHW -> bf16 P8 (zero padded channels), feeds dlogits to the tensor-core kernels
//   abc_channel_sum     per-channel sum of a P8 tensor (bias gradients of convs not followed by BatchNorm)
#include <cuda_bf16.h>

#include "common.cuh"

namespace abc {

struct P8View {
  const uint4* ptr;   // base of the buffer
  int planes;         // planes in the buffer
  int plane_off;      // first plane of the channel range
};

__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == 1) return fmaxf(x, 0.f);
  if (act == 2) return x > 0.f ? x : 0.01f * x;
  return x;
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

// ------------------------------------------------------------------------------------------- statistics
__global__ void __launch_bounds__(256) bn_stats_kernel(P8View z, int N, int HW, double* __restrict__ sum,
                                                       double* __restrict__ sumsq) {
  const int plane = blockIdx.x;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  const long long total = static_cast<long long>(N) * HW;
  for (long long e = static_cast<long long>(blockIdx.y) * 256 + threadIdx.x; e < total; e += static_cast<long long>(gridDim.y) * 256) {
    const int n = static_cast<int>(e / HW);
    const int pix = static_cast<int>(e - static_cast<long long>(n) * HW);
    float v[8];
    unpack8(z.ptr[(static_cast<size_t>(n) * z.planes + z.plane_off + plane) * HW + pix], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i] += v[i];
      q[i] += v[i] * v[i];
    }
  }
  __shared__ double red[16][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    double v = static_cast<double>(i < 8 ? s[i] : q[i - 8]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    const int c = plane * 8 + (threadIdx.x & 7);
    atomicAdd((threadIdx.x < 8 ? sum : sumsq) + c, v);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sum[c] / count;
  double var = sumsq[c] / count - m * m;
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - static_cast<float>(m) * sc;
  mean_out[c] = static_cast<float>(m);
  invstd_out[c] = invstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(m);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// ------------------------------------------------------------------------------------------- forward normalise + act (+ pool)
struct BnActParams {
  P8View z;
  uint4* out;       // may be null
  int out_planes, out_plane_off;
  uint4* pool;      // may be null
  int pool_planes, pool_plane_off;
  int N, H, W, planes;   // planes = C / 8
  const float* scale;
  const float* shift;
  int act;
  float drop_p;
  unsigned long long seed;
};

__global__ void __launch_bounds__(256) bn_act_kernel(const BnActParams p) {
  // one thread per (n, plane, 2x2 pixel block); partial blocks at odd H / W are masked
  const int bw = (p.W + 1) >> 1;
  const int hw2 = ((p.H + 1) >> 1) * bw;
  const long long total = static_cast<long long>(p.N) * p.planes * hw2;
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= total) return;
  const int b = static_cast<int>(t % hw2);
  const long long r = t / hw2;
  const int plane = static_cast<int>(r % p.planes);
  const int n = static_cast<int>(r / p.planes);
  const int by = b / bw, bx = b - by * bw;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = p.scale[plane * 8 + i];
    sh[i] = p.shift[plane * 8 + i];
  }
  float mx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
  const size_t HW = static_cast<size_t>(p.H) * p.W;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int y = 2 * by + (k >> 1), x = 2 * bx + (k & 1);
    if (y >= p.H || x >= p.W) continue;
    const size_t pix = static_cast<size_t>(y) * p.W + x;
    float v[8];
    unpack8(p.z.ptr[(static_cast<size_t>(n) * p.z.planes + p.z.plane_off + plane) * HW + pix], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = act_fwd(v[i] * sc[i] + sh[i], p.act);
      if (p.drop_p > 0.f)
        a *= drop_scale(p.seed, ((static_cast<unsigned long long>(n) * p.planes + plane) * 8 + i) * HW + pix, p.drop_p);
      v[i] = bf16r(a);
      mx[i] = fmaxf(mx[i], v[i]);
    }
    if (p.out) p.out[(static_cast<size_t>(n) * p.out_planes + p.out_plane_off + plane) * HW + pix] = pack8(v);
  }
  if (p.pool)   // even H, W guaranteed by the host check
    p.pool[(static_cast<size_t>(n) * p.pool_planes + p.pool_plane_off + plane) * hw2 + b] = pack8(mx);
}

// ------------------------------------------------------------------------------------------- backward
struct BnActBwdParams {
  P8View z;                 // saved conv output
  P8View dA;                // gradient wrt the full-resolution activation (ptr may be null)
  P8View dP;                // gradient wrt the pooled activation (ptr may be null)
  int N, H, W, planes;
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  int act;
  float drop_p;
  unsigned long long seed;
  double* s1;               // [C] sum g          (reduce: out, apply: in)
  double* s2;               // [C] sum g * xhat
  uint4* dz;                // apply: output
  int dz_planes, dz_plane_off;
  double count;
};

// g (gradient wrt the pre-activation BN output) and xhat for the 2x2 block handled by this thread
__device__ __forceinline__ void bwd_block(const BnActBwdParams& p, int n, int plane, int by, int bx, float (&g)[4][8],
                                          float (&xh)[4][8]) {
  const size_t HW = static_cast<size_t>(p.H) * p.W;
  const int hw2w = p.W >> 1;
  float a[4][8], pre[4][8], dsc[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int y = 2 * by + (k >> 1), x = 2 * bx + (k & 1);
    if (y >= p.H || x >= p.W) {                      // masked pixel of a partial block: contributes nothing
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[k][i] = -INFINITY; pre[k][i] = 0.f; dsc[k][i] = 0.f; g[k][i] = 0.f; xh[k][i] = 0.f;
      }
      continue;
    }
    const size_t pix = static_cast<size_t>(y) * p.W + x;
    float v[8];
    unpack8(p.z.ptr[(static_cast<size_t>(n) * p.z.planes + p.z.plane_off + plane) * HW + pix], v);
    float d[8];
    if (p.dA.ptr) unpack8(p.dA.ptr[(static_cast<size_t>(n) * p.dA.planes + p.dA.plane_off + plane) * HW + pix], d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = plane * 8 + i;
      pre[k][i] = v[i] * p.scale[c] + p.shift[c];
      xh[k][i] = (v[i] - p.mean[c]) * p.invstd[c];
      dsc[k][i] = p.drop_p > 0.f
                      ? drop_scale(p.seed, ((static_cast<unsigned long long>(n) * p.planes + plane) * 8 + i) * HW + pix, p.drop_p)
                      : 1.f;
      a[k][i] = bf16r(act_fwd(pre[k][i], p.act) * dsc[k][i]);
      g[k][i] = p.dA.ptr ? d[i] : 0.f;
    }
  }
  if (p.dP.ptr) {
    float dp[8];
    unpack8(p.dP.ptr[(static_cast<size_t>(n) * p.dP.planes + p.dP.plane_off + plane) * (HW >> 2) + static_cast<size_t>(by) * hw2w + bx], dp);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int best = 0;                                   // first maximum in window order (torch max_pool2d)
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (a[k][i] > a[best][i]) best = k;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k == best) g[k][i] += dp[i];
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) g[k][i] *= act_grad(pre[k][i], p.act) * dsc[k][i];
}

__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(const BnActBwdParams p) {
  const int plane = blockIdx.x;
  const int bw = (p.W + 1) >> 1;
  const int hw2 = ((p.H + 1) >> 1) * bw;
  const long long total = static_cast<long long>(p.N) * hw2;
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  for (long long e = static_cast<long long>(blockIdx.y) * 256 + threadIdx.x; e < total; e += static_cast<long long>(gridDim.y) * 256) {
    const int n = static_cast<int>(e / hw2);
    const int b = static_cast<int>(e - static_cast<long long>(n) * hw2);
    const int by = b / bw, bx = b - by * bw;
    float g[4][8], xh[4][8];
    bwd_block(p, n, plane, by, bx, g, xh);
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s1[i] += g[k][i];
        s2[i] += g[k][i] * xh[k][i];
      }
  }
  __shared__ double red[16][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    double v = static_cast<double>(i < 8 ? s1[i] : s2[i - 8]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    atomicAdd((threadIdx.x < 8 ? p.s1 : p.s2) + plane * 8 + (threadIdx.x & 7), v);
  }
}

__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(const BnActBwdParams p) {
  const int bw = (p.W + 1) >> 1;
  const int hw2 = ((p.H + 1) >> 1) * bw;
  const long long total = static_cast<long long>(p.N) * p.planes * hw2;
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= total) return;
  const int b = static_cast<int>(t % hw2);
  const long long r = t / hw2;
  const int plane = static_cast<int>(r % p.planes);
  const int n = static_cast<int>(r / p.planes);
  const int by = b / bw, bx = b - by * bw;
  float g[4][8], xh[4][8];
  bwd_block(p, n, plane, by, bx, g, xh);
  float m1[8], m2[8], sc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = plane * 8 + i;
    m1[i] = static_cast<float>(p.s1[c] / p.count);
    m2[i] = static_cast<float>(p.s2[c] / p.count);
    sc[i] = p.scale[c];
  }
  const size_t HW = static_cast<size_t>(p.H) * p.W;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int y = 2 * by + (k >> 1), x = 2 * bx + (k & 1);
    if (y >= p.H || x >= p.W) continue;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = sc[i] * (g[k][i] - m1[i] - xh[k][i] * m2[i]);
    p.dz[(static_cast<size_t>(n) * p.dz_planes + p.dz_plane_off + plane) * HW + static_cast<size_t>(y) * p.W + x] = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------- layout / sums
__global__ void __launch_bounds__(256) nchw_to_p8_kernel(const float* __restrict__ src, uint4* __restrict__ dst, int N, int C,
                                                         int HW, int planes) {
  const long long total = static_cast<long long>(N) * planes * HW;
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= total) return;
  const int pix = static_cast<int>(t % HW);
  const long long r = t / HW;
  const int plane = static_cast<int>(r % planes);
  const int n = static_cast<int>(r / planes);
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = plane * 8 + i;
    v[i] = c < C ? src[(static_cast<size_t>(n) * C + c) * HW + pix] : 0.f;
  }
  dst[t] = pack8(v);
}

// dW[16][9], db-free: gradient of the first 1 -> 16 convolution (no dgrad: the image needs no gradient)
template <typename T>
__global__ void __launch_bounds__(256) conv_c1_wgrad_kernel(const T* __restrict__ img, P8View dz, int N, int H, int W,
                                                            float* __restrict__ dw) {
  __shared__ float acc[144];
  if (threadIdx.x < 144) acc[threadIdx.x] = 0.f;
  __syncthreads();
  float loc[144];
#pragma unroll
  for (int i = 0; i < 144; ++i) loc[i] = 0.f;
  const long long total = static_cast<long long>(N) * H * W;
  for (long long e = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256) {
    const int n = static_cast<int>(e / (static_cast<long long>(H) * W));
    const int pix = static_cast<int>(e - static_cast<long long>(n) * H * W);
    const int y = pix / W, x = pix - y * W;
    float t[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int yy = y + dy - 1, xx = x + dx - 1;
        t[dy * 3 + dx] = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                             ? static_cast<float>(img[static_cast<size_t>(n) * H * W + static_cast<size_t>(yy) * W + xx]) : 0.f;
      }
    float g[16];
    unpack8(dz.ptr[(static_cast<size_t>(n) * dz.planes + dz.plane_off) * H * W + pix], g);
    unpack8(dz.ptr[(static_cast<size_t>(n) * dz.planes + dz.plane_off + 1) * H * W + pix], g + 8);
#pragma unroll
    for (int co = 0; co < 16; ++co)
#pragma unroll
      for (int k = 0; k < 9; ++k) loc[co * 9 + k] = fmaf(g[co], t[k], loc[co * 9 + k]);
  }
#pragma unroll
  for (int i = 0; i < 144; ++i) {
    float v = loc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc[i], v);
  }
  __syncthreads();
  if (threadIdx.x < 144) atomicAdd(dw + threadIdx.x, acc[threadIdx.x]);
}

// P8 [N][planes][2H][2W][8] -> 4 phase tensors stacked on the plane axis: dst[N][4*cp][H][W][8], phase = 2*py + px
__global__ void __launch_bounds__(256) deinterleave2_kernel(P8View src, uint4* __restrict__ dst, int N, int cp, int H, int W) {
  const long long total = static_cast<long long>(N) * 4 * cp * H * W;
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= total) return;
  const int x = static_cast<int>(t % W);
  long long r = t / W;
  const int y = static_cast<int>(r % H);
  r /= H;
  const int pl = static_cast<int>(r % (4 * cp));
  const int n = static_cast<int>(r / (4 * cp));
  const int phase = pl / cp, plane = pl - phase * cp;
  const int py = phase >> 1, px = phase & 1;
  dst[t] = src.ptr[((static_cast<size_t>(n) * src.planes + src.plane_off + plane) * (2 * H) + 2 * y + py) * (2 * W) + 2 * x + px];
}

}  // namespace abc

using namespace abc;

static int p8_check(const void* ptr, int planes, int plane_off, int cplanes, const char* what) {
  ABC_REQUIRE(ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s: null or misaligned pointer", what);
  ABC_REQUIRE(plane_off >= 0 && cplanes >= 1 && plane_off + cplanes <= planes, "%s: plane range", what);
  return ABC_OK;
}

extern "C" int abc_bn_stats(const void* z, int N, int H, int W, int planes, int plane_off, int C, double* sum, double* sumsq,
                            void* stream) {
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(C >= 8 && C % 8 == 0 && N > 0 && H > 0 && W > 0 && sum && sumsq, "abc_bn_stats: bad arguments");
  if (int rc = p8_check(z, planes, plane_off, C / 8, "abc_bn_stats")) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ABC_CUDA(cudaMemsetAsync(sum, 0, C * sizeof(double), st));
  ABC_CUDA(cudaMemsetAsync(sumsq, 0, C * sizeof(double), st));
  const long long total = static_cast<long long>(N) * H * W;
  int gy = static_cast<int>((total + 256 * 16 - 1) / (256 * 16));
  const int max_gy = (148 * 8 + C / 8 - 1) / (C / 8);
  if (gy > max_gy) gy = max_gy;
  if (gy < 1) gy = 1;
  P8View v{static_cast<const uint4*>(z), planes, plane_off};
  bn_stats_kernel<<<dim3(C / 8, gy), 256, 0, st>>>(v, N, H * W, sum, sumsq);
  return launch_check("bn_stats_kernel");
}

extern "C" int abc_bn_finalize(const double* sum, const double* sumsq, int C, double count, const float* gamma, const float* beta,
                               float eps, float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                               float* mean, float* invstd, void* stream) {
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(sum && sumsq && gamma && beta && scale && shift && mean && invstd && C > 0 && count > 0, "abc_bn_finalize: bad arguments");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(sum, sumsq, C, count, gamma, beta, eps, momentum,
                                                                                      running_mean, running_var, scale, shift, mean, invstd);
  return launch_check("bn_finalize_kernel");
}

extern "C" int abc_bn_act(const AbcBnActDesc* d, void* stream) {
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(d && d->scale && d->shift, "abc_bn_act: null descriptor / scale / shift");
  ABC_REQUIRE(d->C >= 8 && d->C % 8 == 0 && d->N > 0 && d->H > 0 && d->W > 0, "abc_bn_act: bad geometry (C=%d H=%d W=%d)", d->C, d->H, d->W);
  ABC_REQUIRE(!d->pool || (d->H % 2 == 0 && d->W % 2 == 0), "abc_bn_act: fused max-pool needs even H, W");
  ABC_REQUIRE(d->out || d->pool, "abc_bn_act: no output");
  if (int rc = p8_check(d->z, d->z_planes, d->z_plane_off, d->C / 8, "abc_bn_act(z)")) return rc;
  if (d->out) if (int rc = p8_check(d->out, d->out_planes, d->out_plane_off, d->C / 8, "abc_bn_act(out)")) return rc;
  if (d->pool) if (int rc = p8_check(d->pool, d->pool_planes, d->pool_plane_off, d->C / 8, "abc_bn_act(pool)")) return rc;
  BnActParams p;
  p.z = P8View{static_cast<const uint4*>(d->z), d->z_planes, d->z_plane_off};
  p.out = static_cast<uint4*>(d->out); p.out_planes = d->out_planes; p.out_plane_off = d->out_plane_off;
  p.pool = static_cast<uint4*>(d->pool); p.pool_planes = d->pool_planes; p.pool_plane_off = d->pool_plane_off;
  p.N = d->N; p.H = d->H; p.W = d->W; p.planes = d->C / 8;
  p.scale = d->scale; p.shift = d->shift; p.act = d->act; p.drop_p = d->drop_p; p.seed = d->seed;
  const long long total = static_cast<long long>(d->N) * p.planes * ((d->H + 1) / 2) * ((d->W + 1) / 2);
  bn_act_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return launch_check("bn_act_kernel");
}

extern "C" int abc_bn_act_backward(const AbcBnActBwdDesc* d, void* stream) {
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(d && d->scale && d->shift && d->mean && d->invstd && d->s1 && d->s2 && d->dz, "abc_bn_act_backward: null argument");
  ABC_REQUIRE(d->C >= 8 && d->C % 8 == 0 && d->N > 0 && d->H > 0 && d->W > 0, "abc_bn_act_backward: bad geometry");
  ABC_REQUIRE(!d->dP || (d->H % 2 == 0 && d->W % 2 == 0), "abc_bn_act_backward: pooled gradient needs even H, W");
  ABC_REQUIRE(d->dA || d->dP, "abc_bn_act_backward: no incoming gradient");
  const int cp = d->C / 8;
  if (int rc = p8_check(d->z, d->z_planes, d->z_plane_off, cp, "abc_bn_act_backward(z)")) return rc;
  if (d->dA) if (int rc = p8_check(d->dA, d->dA_planes, d->dA_plane_off, cp, "abc_bn_act_backward(dA)")) return rc;
  if (d->dP) if (int rc = p8_check(d->dP, d->dP_planes, d->dP_plane_off, cp, "abc_bn_act_backward(dP)")) return rc;
  if (int rc = p8_check(d->dz, d->dz_planes, d->dz_plane_off, cp, "abc_bn_act_backward(dz)")) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BnActBwdParams p;
  p.z = P8View{static_cast<const uint4*>(d->z), d->z_planes, d->z_plane_off};
  p.dA = P8View{static_cast<const uint4*>(d->dA), d->dA_planes, d->dA_plane_off};
  p.dP = P8View{static_cast<const uint4*>(d->dP), d->dP_planes, d->dP_plane_off};
  p.N = d->N; p.H = d->H; p.W = d->W; p.planes = cp;
  p.scale = d->scale; p.shift = d->shift; p.mean = d->mean; p.invstd = d->invstd;
  p.act = d->act; p.drop_p = d->drop_p; p.seed = d->seed;
  p.s1 = d->s1; p.s2 = d->s2;
  p.dz = static_cast<uint4*>(d->dz); p.dz_planes = d->dz_planes; p.dz_plane_off = d->dz_plane_off;
  p.count = static_cast<double>(d->N) * d->H * d->W;
  ABC_CUDA(cudaMemsetAsync(d->s1, 0, d->C * sizeof(double), st));
  ABC_CUDA(cudaMemsetAsync(d->s2, 0, d->C * sizeof(double), st));
  const long long blocks2 = static_cast<long long>(d->N) * ((d->H + 1) / 2) * ((d->W + 1) / 2);
  int gy = static_cast<int>((blocks2 + 256 * 4 - 1) / (256 * 4));
  const int max_gy = (148 * 8 + cp - 1) / cp;
  if (gy > max_gy) gy = max_gy;
  if (gy < 1) gy = 1;
  bn_act_bwd_reduce_kernel<<<dim3(cp, gy), 256, 0, st>>>(p);
  if (int rc = launch_check("bn_act_bwd_reduce_kernel")) return rc;
  const long long total = blocks2 * cp;
  bn_act_bwd_apply_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(p);
  return launch_check("bn_act_bwd_apply_kernel");
}

extern "C" int abc_nchw_to_p8(const float* src, void* dst, int N, int C, int H, int W, void* stream) {
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, "abc_nchw_to_p8: bad arguments");
  const int planes = (C + 7) / 8;
  const long long total = static_cast<long long>(N) * planes * H * W;
  nchw_to_p8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<uint4*>(dst), N, C, H * W, planes);
  return launch_check("nchw_to_p8_kernel");
}

extern "C" int abc_channel_sum(const void* x, int N, int H, int W, int planes, int plane_off, int C, double* sum, double* sumsq_scratch,
                               void* stream) {
  return abc_bn_stats(x, N, H, W, planes, plane_off, C, sum, sumsq_scratch, stream);
}

extern "C" int abc_conv3x3_c1_wgrad(const void* img, int img_is_u8, const void* dz, int dz_planes, int dz_plane_off, int N, int H, int W,
                                    float* dw, void* stream) {
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(img && dw && N > 0 && H > 0 && W > 0, "abc_conv3x3_c1_wgrad: bad arguments");
  if (int rc = p8_check(dz, dz_planes, dz_plane_off, 2, "abc_conv3x3_c1_wgrad(dz)")) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ABC_CUDA(cudaMemsetAsync(dw, 0, 144 * sizeof(float), st));
  P8View v{static_cast<const uint4*>(dz), dz_planes, dz_plane_off};
  const int blocks = 148 * 4;
  if (img_is_u8)
    conv_c1_wgrad_kernel<uint8_t><<<blocks, 256, 0, st>>>(static_cast<const uint8_t*>(img), v, N, H, W, dw);
  else
    conv_c1_wgrad_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(img), v, N, H, W, dw);
  return launch_check("conv_c1_wgrad_kernel");
}

extern "C" int abc_deinterleave2(const void* src, int src_planes, int src_plane_off, int C, void* dst, int N, int H, int W, void* stream) {
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(dst && C >= 8 && C % 8 == 0 && N > 0 && H > 0 && W > 0, "abc_deinterleave2: bad arguments");
  if (int rc = p8_check(src, src_planes, src_plane_off, C / 8, "abc_deinterleave2(src)")) return rc;
  P8View v{static_cast<const uint4*>(src), src_planes, src_plane_off};
  const long long total = static_cast<long long>(N) * 4 * (C / 8) * H * W;
  deinterleave2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      v, static_cast<uint4*>(dst), N, C / 8, H, W);
  return launch_check("deinterleave2_kernel");
}
