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

// adaption.mass_matrix(diagonal=True).update (adaption.py:343-363): Welford running mean
// and sum of squares of the samples, one chain per row, and -- in the iteration that
// completes the burn in -- the mass matrix itself (:310-313):
//   mean' = ((it-1)/it) mean + (1/it) x;  ssq += (x - mean) (x - mean')
//   it == burn_in:  M^-1 = ssq / it,  M^1/2 = sqrt(it / ssq)
// `it` is the already incremented iteration; 28 B per parameter (36 B when finalising).
__global__ void k_mass_matrix_update(float* __restrict__ mean, float* __restrict__ ssq,
                                     float* __restrict__ m_inv, float* __restrict__ m_sqrt,
                                     const float* __restrict__ x, int64_t n, float it,
                                     float w_old, float w_new, int finalize) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float xi = x[i], m0 = mean[i];
    const float m1 = __fadd_rn(__fmul_rn(w_old, m0), __fmul_rn(w_new, xi));
    const float q = __fadd_rn(ssq[i], __fmul_rn(__fadd_rn(xi, -m0), __fadd_rn(xi, -m1)));
    mean[i] = m1;
    ssq[i] = q;
    if (finalize) {
      m_inv[i] = __fdiv_rn(q, it);
      m_sqrt[i] = __fsqrt_rn(__fdiv_rn(it, q));
    }
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

int sgmc_mass_matrix_update(void* stream, float* mean, float* ssq, float* m_inv, float* m_sqrt,
                            const float* sample, int64_t n, int64_t iteration, int64_t burn_in) {
  SGMC_REQUIRE(mean && ssq && m_inv && m_sqrt && sample && iteration >= 1, "bad arguments");
  if (n <= 0) return 0;
  const int grid = (int)((n + 255) / 256 > sm_count() * 16 ? sm_count() * 16 : (n + 255) / 256);
  const float it = (float)iteration;
  // (iteration - 1) / iteration and 1 / iteration: int32 / int32 -> f32 true division
  k_mass_matrix_update<<<grid, 256, 0, (cudaStream_t)stream>>>(
      mean, ssq, m_inv, m_sqrt, sample, n, it, (float)(iteration - 1) / it, 1.0f / it,
      iteration == burn_in ? 1 : 0);
  return post_launch("sgmc_mass_matrix_update");
}

int sgmc_axpby(void* stream, float* out, float a, const float* x, float b,
               const float* y, int64_t n) {
  if (n <= 0) return 0;
  const int grid = (int)((n + 255) / 256 > sm_count() * 16 ? sm_count() * 16 : (n + 255) / 256);
  k_axpby<<<grid, 256, 0, (cudaStream_t)stream>>>(out, a, x, b, y, n);
  return post_launch("sgmc_axpby");
}

}  // extern "C"
