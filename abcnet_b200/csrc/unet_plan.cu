// Whole-network inference behind ONE C entry point (abc_unet_forward_infer), so that a host written in any
// language can run the U-Net of /root/reference/src/unet.py:77-119 without the Python layer: abc_unet_create folds BatchNorm
// (eval mode, running statistics) and packs every convolution from the raw fp32 state_dict tensors ON THE HOST (10.7 M
// parameters, milliseconds) into caller-owned device memory; abc_unet_forward_infer enqueues the ~45 launches of one forward
// pass (the same kernels, packs and launch parameters as abcnet_b200.UNet.infer: tests assert bit-identical logits).
// No device allocation, no synchronisation except the one-time weight upload in abc_unet_create.
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace abc {

struct PackedConv {
  int64_t w_off = 0, b_off = 0;      // byte offsets into the device weight arena
  int cin = 0, cout = 0, n_tile = 0, ntaps = 0, fold = 1, fold_swap = 0, subpixel = 0;
  int tap_dy[9] = {0}, tap_dx[9] = {0};
};

struct UNetHandle {
  AbcUNetConfig cfg;
  const uint8_t* arena = nullptr;    // device
  int64_t arena_bytes = 0;
  std::map<std::string, PackedConv> convs;
  int64_t stem_w_off = 0, stem_b_off = 0;
};

static const char* kDoubleConvs[] = {"inc1", "inc2", "down1.maxpool_conv.1", "down2.maxpool_conv.1", "inc3", "down3.maxpool_conv.1",
                                     "down4.maxpool_conv.1", "down5.maxpool_conv.1", "up1.conv", "up2.conv", "up3.conv", "dconv1", "dconv2"};
static const char* kShort[] = {"inc1", "inc2", "down1", "down2", "inc3", "down3", "down4", "down5", "up1.conv", "up2.conv", "up3.conv",
                               "dconv1", "dconv2"};
static const int kDcIn[] = {0, 16, 16, 32, 64, 64, 128, 256, 512, 256, 128, 128, 128};     // 0 = in_channels
static const int kDcOut[] = {16, 16, 32, 64, 64, 128, 256, 512, 256, 128, 128, 128, 128};

static bool is_pooled(const std::string& n) {
  return n == "inc2.3" || n == "down1.3" || n == "inc3.3" || n == "down3.3" || n == "down4.3";
}
static int default_n_tile(int cin, int cout) {
  const int c16 = (cout + 15) / 16 * 16;
  if (c16 <= 128) return c16;
  return (cout % 256 == 0 && cin >= 128) ? 256 : 128;
}
static int row_fold_for(int cin, int cout) {
  if (cout == 16 && cin <= 32) return 4;
  if (cout == 32 && cin <= 32) return 2;
  return 1;
}
static int swap_fold_for(int cin, int cout, bool plain) {
  if (!plain) return 0;
  const int J = cout == 64 ? 2 : (cout == 32 ? 4 : 0);
  if (J == 4 && cin > 32) return 0;
  return J;
}
static int head_conv2_n_tile(int h) {
  if (h <= 16) return 16;
  if (h <= 64) return 64;
  if (h <= 128) return 128;
  return ((h + 191) / 192 * 192 <= (h + 127) / 128 * 128) ? 192 : 128;
}
static void phase_taps(int parity, bool crop_first, int (&k)[2], int (&d)[2], int& n) {
  if (crop_first) {
    if (parity == 0) { n = 1; k[0] = 1; d[0] = 0; }
    else { n = 2; k[0] = 0; d[0] = 1; k[1] = 2; d[1] = 0; }
  } else {
    if (parity == 0) { n = 2; k[0] = 0; d[0] = 0; k[1] = 2; d[1] = -1; }
    else { n = 1; k[0] = 1; d[0] = 0; }
  }
}

// A logical weight matrix [ntaps][rows][cin] (fp32, already BatchNorm-folded) -> the kernel's bf16 blocks
// [n_tiles][cin/kc][ntaps][kc/8][n_tile][8] (include/abcnet_b200.h), rows padded with zeros; bias padded likewise.
struct Logical {
  int ntaps, rows, cin;
  std::vector<float> w;      // [ntaps][rows][cin]
  std::vector<float> bias;   // [rows]
  float& at(int t, int r, int c) { return w[(static_cast<size_t>(t) * rows + r) * cin + c]; }
};

static void append_pack(std::vector<uint8_t>& arena, PackedConv& pc, Logical& L, int n_tile, bool fp16) {
  const int kc = L.cin < 64 ? L.cin : 64;
  const int n_tiles = (L.rows + n_tile - 1) / n_tile;
  pc.n_tile = n_tile;
  pc.ntaps = L.ntaps;
  pc.cin = L.cin;
  while (arena.size() % 256) arena.push_back(0);
  pc.w_off = static_cast<int64_t>(arena.size());
  const size_t n_el = static_cast<size_t>(n_tiles) * n_tile * L.cin * L.ntaps;
  arena.resize(arena.size() + n_el * 2);
  uint16_t* dst = reinterpret_cast<uint16_t*>(arena.data() + pc.w_off);
  size_t o = 0;
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int c = 0; c < L.cin / kc; ++c)
      for (int t = 0; t < L.ntaps; ++t)
        for (int p = 0; p < kc / 8; ++p)
          for (int r = 0; r < n_tile; ++r)
            for (int e = 0; e < 8; ++e) {
              const int row = nt * n_tile + r;
              const float v = row < L.rows ? L.at(t, row, c * kc + p * 8 + e) : 0.f;
              if (fp16) {
                const __half hv = __float2half_rn(v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v));
                memcpy(&dst[o++], &hv, 2);
              } else {
                const __nv_bfloat16 bv = __float2bfloat16_rn(v);
                memcpy(&dst[o++], &bv, 2);
              }
            }
  while (arena.size() % 256) arena.push_back(0);
  pc.b_off = static_cast<int64_t>(arena.size());
  arena.resize(arena.size() + static_cast<size_t>(n_tiles) * n_tile * 4);
  float* b = reinterpret_cast<float*>(arena.data() + pc.b_off);
  for (int i = 0; i < n_tiles * n_tile; ++i) b[i] = i < L.rows ? L.bias[i] : 0.f;
}

struct Tensors {
  std::map<std::string, const AbcNamedTensor*> by_name;
  const float* get(const std::string& name, int64_t numel) const {
    auto it = by_name.find(name);
    if (it == by_name.end()) {
      set_error("abc_unet_create: tensor '%s' missing from the state_dict", name.c_str());
      return nullptr;
    }
    if (it->second->numel != numel || it->second->data == nullptr) {
      set_error("abc_unet_create: tensor '%s' has %lld elements, expected %lld", name.c_str(), static_cast<long long>(it->second->numel),
                static_cast<long long>(numel));
      return nullptr;
    }
    return it->second->data;
  }
};

// Inference BatchNorm fold, fp32 in torch's operation order: scale = g / sqrt(var + eps); w' = w * scale;
// b' = (b - mean) * scale + beta.
static bool folded_conv(const Tensors& T, const std::string& conv, const std::string& bn, int cout, int cin, int k,
                        std::vector<float>& w, std::vector<float>& b) {
  const int64_t per = static_cast<int64_t>(cin) * k * k;
  const float* cw = T.get(conv + ".weight", cout * per);
  const float* cb = T.get(conv + ".bias", cout);
  const float* g = T.get(bn + ".weight", cout);
  const float* be = T.get(bn + ".bias", cout);
  const float* rm = T.get(bn + ".running_mean", cout);
  const float* rv = T.get(bn + ".running_var", cout);
  if (!cw || !cb || !g || !be || !rm || !rv) return false;
  w.resize(static_cast<size_t>(cout) * per);
  b.resize(cout);
  for (int co = 0; co < cout; ++co) {
    const float scale = g[co] / sqrtf(rv[co] + 1e-5f);
    for (int64_t i = 0; i < per; ++i) w[co * per + i] = cw[co * per + i] * scale;
    const float t = cb[co] - rm[co];
    b[co] = t * scale + be[co];
  }
  return true;
}

// 3x3 conv [cout][cin][3][3] (folded) -> packed, with the row-folding / operand-swap layout the layer uses
static void pack3x3(std::vector<uint8_t>& arena, PackedConv& pc, const std::vector<float>& w, const std::vector<float>& b, int cout, int cin,
                    bool plain_output, bool fp16) {
  const int js = swap_fold_for(cin, cout, plain_output);
  const int J = js ? js : row_fold_for(cin, cout);
  pc.cout = cout;
  pc.fold = J;
  pc.fold_swap = js ? 1 : 0;
  auto W = [&](int co, int ci, int dy, int dx) { return w[((static_cast<size_t>(co) * cin + ci) * 3 + (dy + 1)) * 3 + (dx + 1)]; };
  Logical L;
  L.cin = cin;
  if (J == 1) {
    L.ntaps = 9;
    L.rows = cout;
    L.w.assign(static_cast<size_t>(9) * cout * cin, 0.f);
    L.bias = b;
    for (int t = 0; t < 9; ++t) {
      pc.tap_dy[t] = t / 3 - 1;
      pc.tap_dx[t] = t % 3 - 1;
      for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci) L.at(t, co, ci) = W(co, ci, t / 3 - 1, t % 3 - 1);
    }
    append_pack(arena, pc, L, default_n_tile(cin, cout), fp16);
    return;
  }
  // Toeplitz expansion along y (AbcConvDesc.row_fold / swap_mn): folded taps t' = 3 r + c, r in [0, J + 2)
  L.ntaps = 3 * (J + 2);
  L.rows = J * cout;
  L.w.assign(static_cast<size_t>(L.ntaps) * L.rows * cin, 0.f);
  L.bias.resize(L.rows);
  for (int t = 0; t < 9; ++t) {                 // the descriptor still carries the plain 3x3 tap set
    pc.tap_dy[t] = t / 3 - 1;
    pc.tap_dx[t] = t % 3 - 1;
  }
  for (int r = 0; r < J + 2; ++r)
    for (int c = 0; c < 3; ++c)
      for (int j = 0; j < J; ++j) {
        const int dy = r - 1 - j;
        if (dy < -1 || dy > 1) continue;
        for (int co = 0; co < cout; ++co) {
          const int row = js ? j * cout + co : ((co / 16) * J + j) * 16 + (co % 16);
          for (int ci = 0; ci < cin; ++ci) L.at(r * 3 + c, row, ci) = W(co, ci, dy, c - 1);
        }
      }
  for (int j = 0; j < J; ++j)
    for (int co = 0; co < cout; ++co) L.bias[js ? j * cout + co : ((co / 16) * J + j) * 16 + (co % 16)] = b[co];
  append_pack(arena, pc, L, J * cout, fp16);
  pc.ntaps = 9;                                  // descriptor ntaps (the kernel derives the folded count)
}

static int64_t align256(int64_t v) { return (v + 255) & ~static_cast<int64_t>(255); }

struct Workspace {
  uint8_t* base;
  int64_t used = 0, cap;
  void* take(int64_t bytes) {
    void* p = base ? base + used : nullptr;
    used += align256(bytes);
    return p;
  }
};

static int64_t p8_bytes(int N, int C, int H, int W) { return static_cast<int64_t>(N) * (C / 8) * H * W * 16; }

}  // namespace abc

using namespace abc;

static int check_cfg(const AbcUNetConfig* c) {
  ABC_REQUIRE(c != nullptr, "abc_unet: null configuration");
  ABC_REQUIRE(c->in_channels >= 1 && c->in_channels <= 8, "abc_unet: in_channels=%d must be in 1..8", c->in_channels);
  ABC_REQUIRE(c->n_heads >= 1 && c->n_heads <= 16, "abc_unet: n_heads=%d must be in 1..16", c->n_heads);
  for (int i = 0; i < c->n_heads; ++i) ABC_REQUIRE(c->heads[i] >= 1 && c->heads[i] <= 4096, "abc_unet: heads[%d]=%d", i, c->heads[i]);
  return ABC_OK;
}

// bytes of caller-owned DEVICE memory for the packed weights: the exact size abc_unet_create fills (same layer rules)
static int64_t pack_bytes(int rows, int cin, int ntaps, int n_tile) {
  const int64_t n_tiles = (rows + n_tile - 1) / n_tile;
  return align256(n_tiles * n_tile * cin * ntaps * 2) + align256(n_tiles * n_tile * 4);
}

extern "C" int64_t abc_unet_wpack_bytes(const AbcUNetConfig* cfg) {
  if (check_cfg(cfg)) return -1;
  int64_t b = 0;
  for (int i = 0; i < 13; ++i) {
    const int cin0 = kDcIn[i] ? kDcIn[i] : cfg->in_channels, cout = kDcOut[i];
    for (int half = 0; half < 2; ++half) {
      const int cin = half ? cout : cin0;
      if (i == 0 && half == 0) {                                   // stem: fp32 [16][cin * 9] + bias
        b += align256(static_cast<int64_t>(16) * cin * 9 * 4) + align256(64);
        continue;
      }
      const std::string name = std::string(kShort[i]) + (half ? ".3" : ".0");
      const int js = swap_fold_for(cin, cout, !is_pooled(name));
      const int J = js ? js : row_fold_for(cin, cout);
      b += J == 1 ? pack_bytes(cout, cin, 9, default_n_tile(cin, cout)) : pack_bytes(J * cout, cin, 3 * (J + 2), J * cout);
    }
  }
  const int up_cin[] = {512, 256, 128};
  for (int u = 0; u < 3; ++u) b += pack_bytes(4 * (up_cin[u] / 2), up_cin[u], 4, (up_cin[u] / 2) % 64 == 0 ? 256 : 128);
  b += pack_bytes(128 * cfg->n_heads, 128, 9, 256);
  for (int i = 0; i < cfg->n_heads; ++i) b += pack_bytes(cfg->heads[i], 128, 1, head_conv2_n_tile(cfg->heads[i]));
  return b + 256;
}

extern "C" int64_t abc_unet_workspace_bytes(const AbcUNetConfig* cfg, int N, int H, int W) {
  if (check_cfg(cfg)) return -1;
  if (N <= 0 || H <= 0 || W <= 0 || H % 32 || W % 32) {
    set_error("abc_unet_workspace_bytes: N=%d H=%d W=%d (H, W multiples of 32)", N, H, W);
    return -1;
  }
  int64_t b = 0;
  auto add = [&](int C, int s) { b += align256(p8_bytes(N, C, H / s, W / s)); };
  add(16, 1); add(16, 1); add(16, 2); add(32, 2); add(32, 4); add(64, 4); add(64, 4); add(128, 4); add(64, 8); add(128, 8); add(256, 8);
  add(128, 16); add(256, 16); add(512, 16); add(256, 32); add(512, 32); add(512, 32); add(256, 16); add(256, 16); add(128, 8); add(128, 8);
  add(128, 4); add(128, 4); add(128 * cfg->n_heads, 4);
  return b + 4096;
}

// Host-only part of abc_unet_create: fold + pack every convolution of the network into `arena` (the bytes that are uploaded) and
// record the per-layer offsets in `h`. No CUDA call.
static int pack_all(const AbcUNetConfig* cfg, const AbcNamedTensor* tensors, int n_tensors, UNetHandle* h, std::vector<uint8_t>& arena) {
  Tensors T;
  for (int i = 0; i < n_tensors; ++i) {
    ABC_REQUIRE(tensors[i].name != nullptr, "abc_unet_create: tensor %d has no name", i);
    std::string nm = tensors[i].name;
    if (nm.rfind("module.", 0) == 0) nm = nm.substr(7);          // DataParallel / DDP checkpoints (train.py:435)
    T.by_name[nm] = &tensors[i];
  }
  h->cfg = *cfg;
  const bool fp16 = cfg->act_fp16 != 0;
  arena.reserve(64 << 20);
  std::vector<float> w, b;
  auto fail = [&]() { return ABC_ERR_INVALID; };
  // ---- DoubleConvs
  for (int i = 0; i < 13; ++i) {
    const int cin0 = kDcIn[i] ? kDcIn[i] : cfg->in_channels, cout = kDcOut[i];
    const std::string pre = std::string(kDoubleConvs[i]) + ".double_conv";
    for (int half = 0; half < 2; ++half) {
      const int cin = half ? cout : cin0;
      const std::string name = std::string(kShort[i]) + (half ? ".3" : ".0");
      if (!folded_conv(T, pre + (half ? ".3" : ".0"), pre + (half ? ".4" : ".1"), cout, cin, 3, w, b)) return fail();
      if (i == 0 && half == 0) {                                   // stem: fp32 [16][cin * 9] + bias, direct kernel
        while (arena.size() % 256) arena.push_back(0);
        h->stem_w_off = static_cast<int64_t>(arena.size());
        arena.resize(arena.size() + w.size() * 4);
        memcpy(arena.data() + h->stem_w_off, w.data(), w.size() * 4);
        while (arena.size() % 256) arena.push_back(0);
        h->stem_b_off = static_cast<int64_t>(arena.size());
        arena.resize(arena.size() + 64);
        memcpy(arena.data() + h->stem_b_off, b.data(), 64);
        continue;
      }
      pack3x3(arena, h->convs[name], w, b, cout, cin, !is_pooled(name), fp16);
    }
  }
  // ---- up-sampling convolutions: the four sub-pixel phases as blocks of the N axis (AbcConvDesc.subpixel)
  const char* ups[] = {"up1", "up2", "up3"};
  const int up_cin[] = {512, 256, 128};
  for (int u = 0; u < 3; ++u) {
    const int cin = up_cin[u], cout = cin / 2;
    const float* uw = T.get(std::string(ups[u]) + ".up.weight", static_cast<int64_t>(cin) * cout * 9);     // [cin][cout][3][3]
    const float* ub = T.get(std::string(ups[u]) + ".up.bias", cout);
    if (!uw || !ub) return fail();
    PackedConv& pc = h->convs[std::string(ups[u]) + ".up"];
    std::vector<std::pair<int, int>> offs;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        bool used = false;
        for (int py = 0; py < 2 && !used; ++py)
          for (int px = 0; px < 2 && !used; ++px) {
            int ky[2], dyv[2], ny, kx[2], dxv[2], nx;
            phase_taps(py, cfg->crop_first != 0, ky, dyv, ny);
            phase_taps(px, cfg->crop_first != 0, kx, dxv, nx);
            for (int a = 0; a < ny; ++a)
              for (int c = 0; c < nx; ++c) used = used || (dyv[a] == dy && dxv[c] == dx);
          }
        if (used) offs.push_back({dy, dx});
      }
    Logical L;
    L.ntaps = static_cast<int>(offs.size());
    L.rows = 4 * cout;
    L.cin = cin;
    L.w.assign(static_cast<size_t>(L.ntaps) * L.rows * cin, 0.f);
    L.bias.resize(L.rows);
    for (int ph = 0; ph < 4; ++ph)
      for (int co = 0; co < cout; ++co) L.bias[ph * cout + co] = ub[co];
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        int ky[2], dyv[2], ny, kx[2], dxv[2], nx;
        phase_taps(py, cfg->crop_first != 0, ky, dyv, ny);
        phase_taps(px, cfg->crop_first != 0, kx, dxv, nx);
        for (int a = 0; a < ny; ++a)
          for (int c = 0; c < nx; ++c) {
            int t = 0;
            while (offs[t].first != dyv[a] || offs[t].second != dxv[c]) ++t;
            for (int co = 0; co < cout; ++co)
              for (int ci = 0; ci < cin; ++ci)
                L.at(t, (2 * py + px) * cout + co, ci) = uw[((static_cast<size_t>(ci) * cout + co) * 3 + ky[a]) * 3 + kx[c]];
          }
      }
    for (int t = 0; t < L.ntaps; ++t) {
      pc.tap_dy[t] = offs[t].first;
      pc.tap_dx[t] = offs[t].second;
    }
    pc.cout = 4 * cout;
    pc.subpixel = cout;
    append_pack(arena, pc, L, cout % 64 == 0 ? 256 : 128, fp16);
  }
  // ---- heads: the conv1 of all heads as one N = 128 * n_heads GEMM; conv2 per head
  {
    const int nh = cfg->n_heads;
    Logical L;
    L.ntaps = 9;
    L.rows = 128 * nh;
    L.cin = 128;
    L.w.assign(static_cast<size_t>(9) * L.rows * 128, 0.f);
    L.bias.resize(L.rows);
    PackedConv& pc = h->convs["heads.conv1"];
    for (int i = 0; i < nh; ++i) {
      const std::string om = "out_modules." + std::to_string(i);
      if (!folded_conv(T, om + ".conv1", om + ".bn", 128, 128, 3, w, b)) return fail();
      for (int co = 0; co < 128; ++co) {
        L.bias[128 * i + co] = b[co];
        for (int ci = 0; ci < 128; ++ci)
          for (int t = 0; t < 9; ++t) L.at(t, 128 * i + co, ci) = w[(static_cast<size_t>(co) * 128 + ci) * 9 + t];
      }
    }
    for (int t = 0; t < 9; ++t) {
      pc.tap_dy[t] = t / 3 - 1;
      pc.tap_dx[t] = t % 3 - 1;
    }
    pc.cout = 128 * nh;
    append_pack(arena, pc, L, 256, fp16);
    for (int i = 0; i < nh; ++i) {
      const int hc = cfg->heads[i];
      const std::string om = "out_modules." + std::to_string(i);
      const float* w2 = T.get(om + ".conv2.weight", static_cast<int64_t>(hc) * 128);
      const float* b2 = T.get(om + ".conv2.bias", hc);
      if (!w2 || !b2) return fail();
      Logical L2;
      L2.ntaps = 1;
      L2.rows = hc;
      L2.cin = 128;
      L2.w.assign(w2, w2 + static_cast<size_t>(hc) * 128);
      L2.bias.assign(b2, b2 + hc);
      PackedConv& p2 = h->convs["heads." + std::to_string(i) + ".conv2"];
      p2.cout = hc;
      append_pack(arena, p2, L2, head_conv2_n_tile(hc), fp16);
    }
  }
  return ABC_OK;
}

extern "C" int abc_unet_create(const AbcUNetConfig* cfg, const AbcNamedTensor* tensors, int n_tensors, void* wpack_dev, int64_t wpack_bytes,
                               void* stream, AbcUNet** out) {
  if (int rc = check_cfg(cfg)) return rc;
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(tensors && n_tensors > 0 && wpack_dev && out, "abc_unet_create: null argument");
  auto* h = new UNetHandle();
  std::vector<uint8_t> arena;
  if (int rc = pack_all(cfg, tensors, n_tensors, h, arena)) {
    delete h;
    return rc;
  }
  if (static_cast<int64_t>(arena.size()) > wpack_bytes) {
    set_error("abc_unet_create: packed weights need %lld bytes, wpack_bytes = %lld", static_cast<long long>(arena.size()),
              static_cast<long long>(wpack_bytes));
    delete h;
    return ABC_ERR_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemcpyAsync(wpack_dev, arena.data(), arena.size(), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);            // the staging vector dies with this call
  if (e != cudaSuccess) {
    set_error("abc_unet_create: weight upload failed: %s", cudaGetErrorString(e));
    delete h;
    return ABC_ERR_CUDA;
  }
  h->arena = static_cast<const uint8_t*>(wpack_dev);
  h->arena_bytes = static_cast<int64_t>(arena.size());
  *out = reinterpret_cast<AbcUNet*>(h);
  return ABC_OK;
}

// The packed weight arena in HOST memory (what abc_unet_create uploads): lets a host inspect / cache the pack, and lets the CPU
// test suite compare the C++ packing with the Python packing byte for byte. No CUDA call; `used` receives the byte count.
extern "C" int abc_unet_pack_host(const AbcUNetConfig* cfg, const AbcNamedTensor* tensors, int n_tensors, void* out, int64_t out_bytes,
                                  int64_t* used) {
  if (int rc = check_cfg(cfg)) return rc;
  ABC_REQUIRE(tensors && n_tensors > 0 && out && used, "abc_unet_pack_host: null argument");
  UNetHandle h;
  std::vector<uint8_t> arena;
  if (int rc = pack_all(cfg, tensors, n_tensors, &h, arena)) return rc;
  *used = static_cast<int64_t>(arena.size());
  ABC_REQUIRE(*used <= out_bytes, "abc_unet_pack_host: %lld bytes needed, %lld given", static_cast<long long>(*used),
              static_cast<long long>(out_bytes));
  memcpy(out, arena.data(), arena.size());
  return ABC_OK;
}

extern "C" int abc_unet_destroy(AbcUNet* net) {
  delete reinterpret_cast<UNetHandle*>(net);
  return ABC_OK;
}

extern "C" int abc_unet_forward_infer(AbcUNet* net, const void* img, int img_is_u8, int N, int H, int W, void* workspace,
                                      int64_t workspace_bytes, void* const* out_ptrs, int logits_layout, void* stream) {
  if (int rc = device_check()) return rc;
  ABC_REQUIRE(net && img && workspace && out_ptrs, "abc_unet_forward_infer: null argument");
  auto* h = reinterpret_cast<UNetHandle*>(net);
  const AbcUNetConfig& cfg = h->cfg;
  ABC_REQUIRE(N > 0 && H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, "abc_unet_forward_infer: N=%d H=%d W=%d (H, W multiples of 32)", N, H, W);
  ABC_REQUIRE(logits_layout == 1 || logits_layout == 2, "abc_unet_forward_infer: logits_layout 1 (NCHW fp32) or 2 (planar-8 fp32)");
  ABC_REQUIRE(!(img_is_u8 && cfg.in_channels != 1), "abc_unet_forward_infer: uint8 images are the 1-channel binarised format");
  const int64_t need = abc_unet_workspace_bytes(&cfg, N, H, W);
  ABC_REQUIRE(workspace_bytes >= need, "abc_unet_forward_infer: workspace of %lld bytes, %lld needed", static_cast<long long>(workspace_bytes),
              static_cast<long long>(need));
  ABC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "abc_unet_forward_infer: workspace must be 256-byte aligned");
  Workspace ws{static_cast<uint8_t*>(workspace), 0, workspace_bytes};
  auto buf = [&](int C, int s) { return ws.take(p8_bytes(N, C, H / s, W / s)); };
  void *a = buf(16, 1), *b = buf(16, 1), *p1 = buf(16, 2), *d1 = buf(32, 2), *p2 = buf(32, 4), *e1 = buf(64, 4), *e2 = buf(64, 4);
  void *cat3 = buf(128, 4), *p3 = buf(64, 8), *f1 = buf(128, 8), *cat2 = buf(256, 8), *p4 = buf(128, 16), *g1 = buf(256, 16);
  void *cat1 = buf(512, 16), *p5 = buf(256, 32), *h1 = buf(512, 32), *h2 = buf(512, 32), *i1 = buf(256, 16), *i2 = buf(256, 16);
  void *j1 = buf(128, 8), *j2 = buf(128, 8), *k1 = buf(128, 4), *k2 = buf(128, 4), *hid = buf(128 * cfg.n_heads, 4);

  // one launch: conv `name` on src [N][in_planes][hh][ww][8] -> dst (P8 plane slot) and / or the fused max-pool output
  auto conv = [&](const char* name, const void* src, int in_planes, int hh, int ww, void* dst, int out_planes, int out_plane_off, int act,
                  void* pool, int pool_planes) -> int {
    auto it = h->convs.find(name);
    ABC_REQUIRE(it != h->convs.end(), "abc_unet_forward_infer: internal: no pack '%s'", name);
    const PackedConv& pc = it->second;
    AbcConvDesc d;
    memset(&d, 0, sizeof(d));
    d.in = src; d.N = N; d.H = hh; d.W = ww; d.in_planes = in_planes; d.in_plane_off = 0; d.cin = pc.cin;
    d.wpack = h->arena + pc.w_off;
    d.bias = reinterpret_cast<const float*>(h->arena + pc.b_off);
    d.cout = pc.cout; d.n_tile = pc.n_tile; d.ntaps = pc.ntaps;
    for (int t = 0; t < 9; ++t) { d.tap_dy[t] = pc.tap_dy[t]; d.tap_dx[t] = pc.tap_dx[t]; }
    d.act = act; d.out_mode = 0;
    d.out = dst; d.out_planes = out_planes; d.out_plane_off = out_plane_off;
    d.out_sy = d.out_sx = pc.subpixel ? 2 : 1;
    d.out_H = hh * d.out_sy; d.out_W = ww * d.out_sx;
    d.pool_out = pool; d.pool_planes = pool_planes;
    d.row_fold = pc.fold; d.subpixel = pc.subpixel;
    d.act_fp16 = cfg.act_fp16 ? 1 : 0;
    d.swap_mn = (!pc.subpixel && pc.n_tile == 128 && (pc.fold == 1 || pc.fold_swap) && dst != nullptr && pool == nullptr) ? 1 : 0;
    return abc_conv_igemm(&d, stream);
  };
#define RUN(call)               \
  do {                          \
    if (int rc_ = (call)) return rc_; \
  } while (0)

  const float* sw = reinterpret_cast<const float*>(h->arena + h->stem_w_off);
  const float* sb = reinterpret_cast<const float*>(h->arena + h->stem_b_off);
  RUN(abc_conv3x3_stem(img, img_is_u8 ? 1 : 0, cfg.in_channels, sw, sb, a, N, H, W, 2, 0, 1 | (cfg.act_fp16 ? 2 : 0), stream));
  RUN(conv("inc1.3", a, 2, H, W, b, 2, 0, 1, nullptr, 0));
  RUN(conv("inc2.0", b, 2, H, W, a, 2, 0, 1, nullptr, 0));
  RUN(conv("inc2.3", a, 2, H, W, nullptr, 0, 0, 1, p1, 2));                       // x1 is never used as a skip
  RUN(conv("down1.0", p1, 2, H / 2, W / 2, d1, 4, 0, 1, nullptr, 0));
  RUN(conv("down1.3", d1, 4, H / 2, W / 2, nullptr, 0, 0, 1, p2, 4));
  RUN(conv("down2.0", p2, 4, H / 4, W / 4, e1, 8, 0, 1, nullptr, 0));
  RUN(conv("down2.3", e1, 8, H / 4, W / 4, e2, 8, 0, 1, nullptr, 0));
  RUN(conv("inc3.0", e2, 8, H / 4, W / 4, e1, 8, 0, 1, nullptr, 0));
  RUN(conv("inc3.3", e1, 8, H / 4, W / 4, cat3, 16, 0, 1, p3, 8));                // x3 -> concat slot [0, 64)
  RUN(conv("down3.0", p3, 8, H / 8, W / 8, f1, 16, 0, 1, nullptr, 0));
  RUN(conv("down3.3", f1, 16, H / 8, W / 8, cat2, 32, 0, 1, p4, 16));
  RUN(conv("down4.0", p4, 16, H / 16, W / 16, g1, 32, 0, 1, nullptr, 0));
  RUN(conv("down4.3", g1, 32, H / 16, W / 16, cat1, 64, 0, 1, p5, 32));
  RUN(conv("down5.0", p5, 32, H / 32, W / 32, h1, 64, 0, 1, nullptr, 0));
  RUN(conv("down5.3", h1, 64, H / 32, W / 32, h2, 64, 0, 1, nullptr, 0));
  RUN(conv("up1.up", h2, 64, H / 32, W / 32, cat1, 64, 32, 0, nullptr, 0));
  RUN(conv("up1.conv.0", cat1, 64, H / 16, W / 16, i1, 32, 0, 1, nullptr, 0));
  RUN(conv("up1.conv.3", i1, 32, H / 16, W / 16, i2, 32, 0, 1, nullptr, 0));
  RUN(conv("up2.up", i2, 32, H / 16, W / 16, cat2, 32, 16, 0, nullptr, 0));
  RUN(conv("up2.conv.0", cat2, 32, H / 8, W / 8, j1, 16, 0, 1, nullptr, 0));
  RUN(conv("up2.conv.3", j1, 16, H / 8, W / 8, j2, 16, 0, 1, nullptr, 0));
  RUN(conv("up3.up", j2, 16, H / 8, W / 8, cat3, 16, 8, 0, nullptr, 0));
  RUN(conv("up3.conv.0", cat3, 16, H / 4, W / 4, k1, 16, 0, 1, nullptr, 0));
  RUN(conv("up3.conv.3", k1, 16, H / 4, W / 4, k2, 16, 0, 1, nullptr, 0));
  RUN(conv("dconv1.0", k2, 16, H / 4, W / 4, k1, 16, 0, 1, nullptr, 0));
  RUN(conv("dconv1.3", k1, 16, H / 4, W / 4, k2, 16, 0, 1, nullptr, 0));
  RUN(conv("dconv2.0", k2, 16, H / 4, W / 4, k1, 16, 0, 1, nullptr, 0));
  RUN(conv("dconv2.3", k1, 16, H / 4, W / 4, k2, 16, 0, 1, nullptr, 0));          // trunk
  RUN(conv("heads.conv1", k2, 16, H / 4, W / 4, hid, 16 * cfg.n_heads, 0, 2, nullptr, 0));      // BN fold + LeakyReLU(0.01)
  for (int i = 0; i < cfg.n_heads; ++i) {
    const PackedConv& pc = h->convs["heads." + std::to_string(i) + ".conv2"];
    ABC_REQUIRE(out_ptrs[i] != nullptr, "abc_unet_forward_infer: out_ptrs[%d] is null", i);
    AbcConvDesc d;
    memset(&d, 0, sizeof(d));
    d.in = hid; d.N = N; d.H = H / 4; d.W = W / 4; d.in_planes = 16 * cfg.n_heads; d.in_plane_off = 16 * i; d.cin = 128;
    d.wpack = h->arena + pc.w_off;
    d.bias = reinterpret_cast<const float*>(h->arena + pc.b_off);
    d.cout = pc.cout; d.n_tile = pc.n_tile; d.ntaps = 1;
    d.act = 0;
    // one-channel heads are always NCHW (the decoder reads the centre maps densely); planar-8 otherwise when asked for
    d.out_mode = (logits_layout == 2 && pc.cout > 1) ? 2 : 1;
    d.out = out_ptrs[i]; d.out_planes = d.out_mode == 2 ? (pc.cout + 7) / 8 : 0; d.out_plane_off = 0;
    d.out_H = H / 4; d.out_W = W / 4; d.out_sy = d.out_sx = 1;
    d.row_fold = 1;
    d.act_fp16 = cfg.act_fp16 ? 1 : 0;
    RUN(abc_conv_igemm(&d, stream));
  }
#undef RUN
  return ABC_OK;
}
