// Shared host-side helpers for libsgmc_b200 (error reporting, launch counter).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/sgmc_b200.h"

namespace sgmc {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
int option(int which);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

inline int post_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), what);
}

inline int sm_count() {
  static int cached = 0;
  if (!cached) {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached = n > 0 ? n : 148;
  }
  return cached;
}

}  // namespace sgmc

#define SGMC_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::sgmc::set_error(__VA_ARGS__);      \
      return 2;                            \
    }                                      \
  } while (0)
