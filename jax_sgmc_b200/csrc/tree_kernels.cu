// Pytree vector-space operations of jax_sgmc/util/tree_util.py:27-133 on the
// chain-batched flat layout f32[C][P]: tree_scale (:83-96), tree_add (:99-110),
// tree_multiply (:58-80) and tree_dot (:120-133, one scalar per chain).  The
// integrators fuse these into their update kernels; the stand-alone versions
// exist for API parity and tests (bit-exact against NumPy: one rounding per op).
#include "common.cuh"

namespace sgmc {

// op 0: out = alpha * x ; op 1: out = x + y ; op 2: out = x * y
__global__ void k_tree_ewise(int op, float* __restrict__ out, float alpha,
                             const float* __restrict__ x, const float* __restrict__ y,
                             int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float a = x[i];
    out[i] = op == 0 ? __fmul_rn(alpha, a) : (op == 1 ? __fadd_rn(a, y[i]) : __fmul_rn(a, y[i]));
  }
}

// out[c] = sum_j x[c][j] * y[c][j]   (one CTA per chain, fp32 tree reduction)
__global__ void __launch_bounds__(256) k_tree_dot(float* __restrict__ out,
                                                  const float* __restrict__ x,
                                                  const float* __restrict__ y, int64_t P) {
  __shared__ float red[8];
  const int64_t c = blockIdx.x;
  float s = 0.f;
  for (int64_t j = threadIdx.x; j < P; j += 256) s = fmaf(x[c * P + j], y[c * P + j], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    out[c] = t;
  }
}

}  // namespace sgmc

using namespace sgmc;

extern "C" {

int sgmc_tree_ewise(void* stream, int op, float* out, float alpha, const float* x,
                    const float* y, int64_t n) {
  SGMC_REQUIRE(op >= 0 && op <= 2, "unknown elementwise op %d", op);
  if (n <= 0) return 0;
  const int grid = (int)((n + 255) / 256 > sm_count() * 16 ? sm_count() * 16 : (n + 255) / 256);
  k_tree_ewise<<<grid, 256, 0, (cudaStream_t)stream>>>(op, out, alpha, x, y, n);
  return post_launch("sgmc_tree_ewise");
}

int sgmc_tree_dot(void* stream, float* out, const float* x, const float* y,
                  int64_t n_chains, int64_t P) {
  if (n_chains <= 0) return 0;
  k_tree_dot<<<(unsigned)n_chains, 256, 0, (cudaStream_t)stream>>>(out, x, y, P);
  return post_launch("sgmc_tree_dot");
}

}  // extern "C"
