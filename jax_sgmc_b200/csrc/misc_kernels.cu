// Small utility kernels (one-time data-set statistics).
#include "common.cuh"

namespace sgmc {

__global__ void k_absmax(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ out_bits) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));   // m >= 0
}

}  // namespace sgmc

using namespace sgmc;

extern "C" int sgmc_absmax(void* stream, const float* x, int64_t n, float* out) {
  cudaStream_t s = (cudaStream_t)stream;
  if (check_cuda(cudaMemsetAsync(out, 0, 4, s), "memset")) return 1;
  if (n <= 0) return 0;
  const int grid = (int)((n + 255) / 256 > sm_count() * 8 ? sm_count() * 8 : (n + 255) / 256);
  k_absmax<<<grid, 256, 0, s>>>(x, n, reinterpret_cast<uint32_t*>(out));
  return post_launch("sgmc_absmax");
}
