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
void prof_mark(cudaStream_t stream, int slot);   // SGMC_OPT_STEP_PROFILE, runtime.cu

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

// Launch with programmatic dependent launch (PDL) unless disabled: the grid may
// start while the previous kernel of the stream drains, so every kernel
// launched through here calls pdl_wait() before its first global access.
template <class... P, class... A>
inline void launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem,
                       cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = option(SGMC_OPT_SERIAL_LAUNCH) ? 0 : 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

#ifdef __CUDACC__
// Blocks until the grids this one depends on have completed and their writes
// are visible (no-op when launched without the PDL attribute).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the next PDL grid of the stream start being scheduled.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

}  // namespace sgmc

#define SGMC_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::sgmc::set_error(__VA_ARGS__);      \
      return 2;                            \
    }                                      \
  } while (0)
