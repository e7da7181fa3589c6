// A host WITHOUT Python or torch: the whole inference path of the reference (src/unet.py:77-119 forward + src/img2smiles.py:62-193
// decode) through include/abcnet_b200.h only.
//
//   g++ -std=c++17 -O2 examples/native_host.cpp -Iinclude -Labcnet_b200 -labcnet_b200 -L/usr/local/cuda/lib64 -lcudart \
//       -Wl,-rpath,$PWD/abcnet_b200 -o native_host
//   ./native_host weights.bin images.u8 N H W            # prints "image i: <atom peaks> <bond records>" and a logit checksum
//
// weights.bin: for every state_dict entry  int32 name_len | name bytes | int64 numel | numel fp32 values  (the test writes it from
// the same state_dict the Python model loads; any checkpoint reader can produce it). images.u8: N * H * W bytes in {0, 1}.
// tests/test_native_gpu.py::test_native_host_program compiles and runs this file and compares with abcnet_b200.UNet + PeakDecoder.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "abcnet_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    if ((x) != 0) {                                                                \
      fprintf(stderr, "%s failed: %s\n", #x, abc_last_error());                    \
      return 1;                                                                    \
    }                                                                              \
  } while (0)
#define CU(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));              \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

int main(int argc, char** argv) {
  if (argc != 6) {
    fprintf(stderr, "usage: %s weights.bin images.u8 N H W\n", argv[0]);
    return 2;
  }
  const int N = atoi(argv[3]), H = atoi(argv[4]), W = atoi(argv[5]);
  // ---- the checkpoint, as named host tensors
  std::vector<std::string> names;
  std::vector<std::vector<float>> data;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  for (;;) {
    int32_t len;
    if (fread(&len, 4, 1, f) != 1) break;
    std::string nm(len, '\0');
    int64_t numel;
    if (fread(&nm[0], 1, len, f) != static_cast<size_t>(len) || fread(&numel, 8, 1, f) != 1) return 2;
    std::vector<float> v(numel);
    if (fread(v.data(), 4, numel, f) != static_cast<size_t>(numel)) return 2;
    names.push_back(nm);
    data.push_back(std::move(v));
  }
  fclose(f);
  std::vector<AbcNamedTensor> t(names.size());
  for (size_t i = 0; i < names.size(); ++i) t[i] = AbcNamedTensor{names[i].c_str(), data[i].data(), static_cast<int64_t>(data[i].size())};
  std::vector<uint8_t> img(static_cast<size_t>(N) * H * W);
  f = fopen(argv[2], "rb");
  if (!f || fread(img.data(), 1, img.size(), f) != img.size()) return 2;
  fclose(f);

  // ---- UNet(in_channels=1, heads=[1,14,3,2,1,360,60,60]) + load_state_dict + eval()   (train.py:47, img2smiles.py:42-49)
  AbcUNetConfig cfg = {1, 8, {1, 14, 3, 2, 1, 360, 60, 60}, 1};
  cudaStream_t st;
  CU(cudaStreamCreate(&st));
  void *wpack, *ws, *d_img;
  const int64_t wbytes = abc_unet_wpack_bytes(&cfg), wsbytes = abc_unet_workspace_bytes(&cfg, N, H, W);
  CU(cudaMalloc(&wpack, wbytes));
  CU(cudaMalloc(&ws, wsbytes));
  CU(cudaMalloc(&d_img, img.size()));
  CU(cudaMemcpyAsync(d_img, img.data(), img.size(), cudaMemcpyHostToDevice, st));
  AbcUNet* net = nullptr;
  CK(abc_unet_create(&cfg, t.data(), static_cast<int>(t.size()), wpack, wbytes, st, &net));
  // ---- outs = model(imgs): planar-8 fp32 logits for the multi-channel heads (what the decoder reads fastest), NCHW for the centre maps
  const int H4 = H / 4, W4 = W / 4;
  void* outs[8];
  size_t out_floats[8];
  for (int i = 0; i < 8; ++i) {
    const int c = cfg.heads[i];
    out_floats[i] = static_cast<size_t>(N) * (c > 1 ? (c + 7) / 8 * 8 : 1) * H4 * W4;
    CU(cudaMalloc(&outs[i], out_floats[i] * 4));
  }
  CK(abc_unet_forward_infer(net, d_img, /*img_is_u8=*/1, N, H, W, ws, wsbytes, outs, /*planar-8=*/2, st));
  // ---- img2smiles.py:62-193 as one launch
  const int atom_cap = 4096, bond_cap = 16384;
  AbcAtomRec* d_atoms;
  AbcBondRec* d_bonds;
  int32_t* d_counts;
  CU(cudaMalloc(&d_atoms, sizeof(AbcAtomRec) * N * atom_cap));
  CU(cudaMalloc(&d_bonds, sizeof(AbcBondRec) * N * bond_cap));
  CU(cudaMalloc(&d_counts, 16 * N));
  AbcDecodeDesc d = {};
  for (int i = 0; i < 8; ++i) {
    d.maps[i] = static_cast<const float*>(outs[i]);
    if (cfg.heads[i] > 1) d.p8f_mask |= 1 << i;
  }
  d.N = N; d.H = H4; d.W = W4;
  d.c_type = 14; d.c_charge = 3; d.c_hs = 2; d.n_omega = 60; d.n_btype = 6;
  d.thr = -1.0f; d.omega_mode = 0;
  d.atoms = d_atoms; d.atom_cap = atom_cap; d.bonds = d_bonds; d.bond_cap = bond_cap; d.counts = d_counts;
  CK(abc_decode_peaks(&d, st));
  std::vector<int32_t> counts(4 * N);
  std::vector<float> centre(static_cast<size_t>(N) * H4 * W4);
  CU(cudaMemcpyAsync(counts.data(), d_counts, 16 * N, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(centre.data(), outs[0], centre.size() * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int i = 0; i < N; ++i) printf("image %d: %d %d\n", i, counts[4 * i], counts[4 * i + 1]);
  double sum = 0;
  for (float v : centre) sum += v;
  printf("atom-centre logit sum %.6f\n", sum);
  abc_unet_destroy(net);
  return 0;
}
