// Minibatch rows pulled by the GPU straight out of a HOST-resident data set.
//
// The reference's host loaders gather every cached minibatch with NumPy fancy indexing
// on the host and ship the block through an io_callback (data/numpy_loader.py:342-389,
// data/core.py:664-791).  Here the data set stays where the user allocated it: its pages
// are locked and mapped into the device's address space once (cudaHostRegister), and one
// small kernel per minibatch reads the n drawn rows over the host link by index -- no
// host gather, no staging copy, no host thread on the data path.
// MEASURED (B200, PCIe Gen5 x16, profiles/r02_host_link.md): alone the pull moves
// 43 GB/s (the DMA engine: 53 GB/s), but SM-issued reads of host memory and HBM-bound
// kernels slow each other down badly when they run together (the fused update 20 -> 40 us,
// the pull 97 -> 130..170 us per minibatch; the same with ld.relaxed.sys, plain cached
// loads and cp.async.bulk through shared memory), so inside the scan it is no faster than
// the host-gathered, DMA-copied ring.  It therefore is the OPT-IN source of the host scan
// (SGMC_HOST_PULL=1): useful when the host has no cores to spare.  Every row is d
// contiguous floats (4 KB at d = 1024), read as 16-byte loads with four requests in
// flight per lane.  The kernel needs no shared memory and at most 40 registers per thread
// in CTAs of 128 threads, so up to two of its CTAs fit into what the persistent potential
// kernel leaves free on an SM (576 threads x 96 registers, all of the shared memory) and
// the link stays busy for the whole step.
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"

namespace sgmc {

__device__ __forceinline__ float4 ld_host16(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// One warp per row.  VEC: rows are 16-byte aligned multiples of 4 floats.
constexpr int kPullThreads = 128;
constexpr int kPullInFlight = 4;     // 16-byte loads in flight per lane

template <bool VEC>
__global__ void __launch_bounds__(kPullThreads, 12)
k_pull_rows(const float* __restrict__ X, const float* __restrict__ y,
            const int32_t* __restrict__ idx, int n, int row0, int rows, int d,
            float* __restrict__ dst, float* __restrict__ labels) {
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int n_warps = (int)((gridDim.x * blockDim.x) >> 5);
  // labels first: n scattered 4-byte reads, issued before the row traffic queues up
  if (labels != nullptr)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      labels[i] = y[idx[i]];
  for (int r = warp; r < rows; r += n_warps) {
    const int64_t src_row = idx[row0 + r];
    const float* src = X + src_row * d;
    float* out = dst + (int64_t)(row0 + r) * d;
    if (VEC) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* o4 = reinterpret_cast<float4*>(out);
      const int q = d >> 2;
      for (int j0 = 0; j0 < q; j0 += 32 * kPullInFlight) {
        float4 v[kPullInFlight];
#pragma unroll
        for (int u = 0; u < kPullInFlight; ++u) {
          const int j = j0 + u * 32 + lane;
          if (j < q) v[u] = ld_host16(s4 + j);
        }
#pragma unroll
        for (int u = 0; u < kPullInFlight; ++u) {
          const int j = j0 + u * 32 + lane;
          if (j < q) o4[j] = v[u];
        }
      }
    } else {
      for (int j = lane; j < d; j += 32) out[j] = src[j];
    }
  }
}

}  // namespace sgmc

using namespace sgmc;

extern "C" {

int sgmc_host_register(void* ptr, size_t bytes, void** device_ptr) {
  SGMC_REQUIRE(ptr != nullptr && bytes > 0 && device_ptr != nullptr, "null argument");
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    e = cudaSuccess;
  }
  if (check_cuda(e, "cudaHostRegister")) return 1;
  return check_cuda(cudaHostGetDevicePointer(device_ptr, ptr, 0), "cudaHostGetDevicePointer");
}

int sgmc_host_unregister(void* ptr) {
  cudaError_t e = cudaHostUnregister(ptr);
  if (e == cudaErrorHostMemoryNotRegistered) {
    cudaGetLastError();
    return 0;
  }
  return check_cuda(e, "cudaHostUnregister");
}

int sgmc_pull_rows(void* stream, const float* X_mapped, const float* y_mapped,
                   const int32_t* idx, int64_t batch_size, int64_t row0, int64_t rows,
                   int64_t d, float* dst_rows, float* dst_labels, int n_ctas) {
  SGMC_REQUIRE(X_mapped && idx && dst_rows && batch_size > 0 && d > 0 && row0 >= 0 && rows >= 0 &&
               row0 + rows <= batch_size, "bad pull arguments");
  SGMC_REQUIRE(dst_labels == nullptr || y_mapped != nullptr, "labels need y");
  if (rows == 0 && dst_labels == nullptr) return 0;
  if (n_ctas <= 0) n_ctas = 48;
  const int64_t wpc = kPullThreads / 32;                            // one row per warp
  const int64_t want = (rows + wpc - 1) / wpc > 0 ? (rows + wpc - 1) / wpc : 1;
  const unsigned grid = (unsigned)(want < n_ctas ? want : n_ctas);
  const bool vec = d % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(X_mapped) | reinterpret_cast<uintptr_t>(dst_rows)) & 15u) == 0;
  // An SM changes its L1 / shared-memory split only when it is idle: ask for the split the
  // persistent potential kernel runs with (all shared memory), otherwise this kernel and
  // that one would exclude each other from an SM although both fit.
  static bool carveout_set = false;
  if (!carveout_set) {
    cudaFuncSetAttribute(k_pull_rows<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_pull_rows<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    carveout_set = true;
  }
  if (vec)
    k_pull_rows<true><<<grid, kPullThreads, 0, (cudaStream_t)stream>>>(
        X_mapped, y_mapped, idx, (int)batch_size, (int)row0, (int)rows, (int)d, dst_rows, dst_labels);
  else
    k_pull_rows<false><<<grid, kPullThreads, 0, (cudaStream_t)stream>>>(
        X_mapped, y_mapped, idx, (int)batch_size, (int)row0, (int)rows, (int)d, dst_rows, dst_labels);
  return post_launch("sgmc_pull_rows");
}

}  // extern "C"
