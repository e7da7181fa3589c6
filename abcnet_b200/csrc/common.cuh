// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/abcnet_b200.h"

namespace abc {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
int device_check();   // ABC_OK or ABC_ERR_NO_DEVICE (message set)
int sm_count();

#define ABC_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::abc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ABC_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define ABC_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::abc::set_error(__VA_ARGS__);  \
      return ABC_ERR_INVALID;         \
    }                                 \
  } while (0)

inline int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return ABC_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ABC_OK;
}

}  // namespace abc
