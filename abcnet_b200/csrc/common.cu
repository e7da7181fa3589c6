// Error state, device probing and launch accounting for the C-ABI.
#include "common.cuh"

namespace abc {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_check() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available (abcnet_b200 has no CPU fallback)");
    return ABC_ERR_NO_DEVICE;
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
    cudaGetLastError();
    set_error("device %d has compute capability major %d; the kernels are built for sm_100a only", dev, major);
    return ABC_ERR_NO_DEVICE;
  }
  return ABC_OK;
}

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

}  // namespace abc

extern "C" {
const char* abc_last_error(void) { return abc::g_err; }
int abc_version(void) { return 100; }
int abc_device_ok(void) { return abc::device_check() == ABC_OK ? 1 : 0; }
int abc_sm_count(void) {
  int n = abc::sm_count();
  return n > 0 ? n : ABC_ERR_NO_DEVICE;
}
int64_t abc_launch_count(void) { return abc::g_launches.load(); }
}
