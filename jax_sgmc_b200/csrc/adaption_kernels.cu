// Stand-alone adaption.rms_prop kernels (jax_sgmc/adaption.py:225-293) for
// callers that use the (init, update, get) triplet outside the fused pSGLD
// update: update v' = alpha v + (1-alpha) g^2 (:270-272); get G = 1/(lmbd +
// sqrt(v)), sqrt(G) (:289-291; Gamma == 0).  IEEE arithmetic, one rounding per
// operation like the oracle.  HBM-bound elementwise passes (12 B and 12 B per
// parameter); the hot path uses the fused kernel in update_kernels.cu instead.
#include "common.cuh"

namespace sgmc {

__global__ void k_rms_update(float* __restrict__ v, const float* __restrict__ g,
                             int64_t n, float alpha, float one_m) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    v[i] = __fadd_rn(__fmul_rn(alpha, v[i]), __fmul_rn(one_m, __fmul_rn(gi, gi)));
  }
}

__global__ void k_rms_get(const float* __restrict__ v, float* __restrict__ g_inv,
                          float* __restrict__ sqrt_g_inv, int64_t n, float lmbd) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float G = __frcp_rn(__fadd_rn(lmbd, __fsqrt_rn(v[i])));
    g_inv[i] = G;
    sqrt_g_inv[i] = __fsqrt_rn(G);
  }
}

// out = a*x + b*y (small per-chain scalars, e.g. OBABO's 0.5*(U1+U2),
// integrator.py:264)
__global__ void k_axpby(float* __restrict__ out, float a, const float* __restrict__ x,
                        float b, const float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn(a, x[i]), __fmul_rn(b, y[i]));
}

}  // namespace sgmc

using namespace sgmc;

extern "C" {

int sgmc_rms_prop_update(void* stream, float* v, const float* grad, int64_t n,
                         float alpha) {
  if (n <= 0) return 0;
  const int grid = (int)((n + 255) / 256 > sm_count() * 16 ? sm_count() * 16 : (n + 255) / 256);
  k_rms_update<<<grid, 256, 0, (cudaStream_t)stream>>>(v, grad, n, alpha, 1.0f - alpha);
  return post_launch("sgmc_rms_prop_update");
}

int sgmc_rms_prop_get(void* stream, const float* v, float* g_inv, float* sqrt_g_inv,
                      int64_t n, float lmbd) {
  if (n <= 0) return 0;
  const int grid = (int)((n + 255) / 256 > sm_count() * 16 ? sm_count() * 16 : (n + 255) / 256);
  k_rms_get<<<grid, 256, 0, (cudaStream_t)stream>>>(v, g_inv, sqrt_g_inv, n, lmbd);
  return post_launch("sgmc_rms_prop_get");
}

int sgmc_axpby(void* stream, float* out, float a, const float* x, float b,
               const float* y, int64_t n) {
  if (n <= 0) return 0;
  const int grid = (int)((n + 255) / 256 > sm_count() * 16 ? sm_count() * 16 : (n + 255) / 256);
  k_axpby<<<grid, 256, 0, (cudaStream_t)stream>>>(out, a, x, b, y, n);
  return post_launch("sgmc_axpby");
}

}  // extern "C"
