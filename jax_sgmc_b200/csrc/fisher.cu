// adaption.fisher_information(diagonal=True).get (jax_sgmc/adaption.py:372-457): the
// empirical Fisher noise model of SGHMC [Ahn et al. 2012] for the recognised GLM
// likelihoods, all chains at once.
//
//   m      = grad U / N                                          (:404)
//   ssq_p  = sum_i (d ell_i / d theta_p  -  m_p)^2               (:406-420; the reference's
//            sign convention: the per-observation LIKELIHOOD gradient minus the unscaled
//            POTENTIAL gradient -- reproduced as written)
//   v = ssq / (n - 1);  b = (0.5 eps) v;  correction = friction - b          (:423-427)
//   smallest = min over the chain's parameters of the positive corrections   (:428)
//   positive = correction <= 0 ? smallest : correction                       (:429)
//   noise_scale = sqrt(positive)   (cb_diff_sqrt, multiplies the SGHMC noise, integrator.py:648)
//   scale       = sqrt(friction - positive)                                  (:432-435)
//
// For a GLM the per-observation gradient is a_i x_i on the weights and b_i on the auxiliary
// parameter (logistic: a = b = y - sigmoid(z); gaussian: a = r / s2, b = r^2 / s2 - 1), so
// ssq is one more GEMM-shaped pass over the minibatch: k_fisher_coef (one warp per
// (chain, observation): z, a, b), k_fisher_ssq (a thread owns one parameter of eight chains
// and walks the minibatch; X rows are read coalesced and reused across the eight chains),
// k_fisher_finalize (per-chain minimum + the two square roots).  fp32 FFMA: this runs once
// per leapfrog step of alias.sghmc(adapt_noise_model=True), beside a potential evaluation of
// the same shape.
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "common.cuh"
#include "glm_math.cuh"

namespace sgmc {

struct FisherArgs {
  sgmc_glm_spec spec;
  const float* theta; int64_t C, P;
  const float* X; const float* y; const int32_t* idx; int n;
  const float* grad;
  const float* friction; float friction_scalar;
  float inv_N, inv_nm1, half_eps;
  float* a; float* b;            // [C][n]
  float* corr;                   // [C][P]
  float* noise_scale; float* scale;
};

__global__ void __launch_bounds__(256) k_fisher_coef(const FisherArgs f) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= f.C * f.n) return;
  const int64_t c = w / f.n;
  const int i = (int)(w - c * f.n);
  const int64_t row = f.idx ? f.idx[i] : i;
  const float* x = f.X + row * f.spec.d;
  const float* th = f.theta + c * f.P;
  float z = 0.f;
  for (int j = lane; j < f.spec.d; j += 32) z = fmaf(x[j], th[f.spec.w_off + j], z);
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) z += __shfl_xor_sync(0xffffffffu, z, k);
  if (lane != 0) return;
  GaussConst gc{1.f, 0.f};
  if (f.spec.family == kFamilyGaussian) gc = gauss_const(th[f.spec.aux_off]);
  else if (f.spec.aux_off >= 0) z += th[f.spec.aux_off];
  float ell, dz;
  glm_link(f.spec.family, z, f.y[row], gc, ell, dz);
  f.a[w] = dz;
  if (f.spec.family == kFamilyGaussian) {
    const float r = f.y[row] - z;
    f.b[w] = (r * r) / gc.s2 - 1.0f;              // d ell / d log_sigma
  } else {
    f.b[w] = dz;                                    // d ell / d bias
  }
}

constexpr int kFisherChains = 8;

__global__ void __launch_bounds__(128) k_fisher_ssq(const FisherArgs f) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.y * kFisherChains;
  if (p >= f.P) return;
  const bool is_w = p >= f.spec.w_off && p < f.spec.w_off + f.spec.d;
  const bool is_aux = f.spec.aux_off >= 0 && p == f.spec.aux_off;
  float m[kFisherChains], acc[kFisherChains];
#pragma unroll
  for (int k = 0; k < kFisherChains; ++k) {
    const int64_t c = c0 + k;
    m[k] = c < f.C ? f.grad[c * f.P + p] * f.inv_N : 0.f;      // sample_grad /= N
    acc[k] = 0.f;
  }
  const int j = (int)(p - f.spec.w_off);
  for (int i = 0; i < f.n; ++i) {
    float xv = 0.f;
    if (is_w) xv = f.X[(int64_t)(f.idx ? f.idx[i] : i) * f.spec.d + j];
#pragma unroll
    for (int k = 0; k < kFisherChains; ++k) {
      const int64_t c = c0 + k;
      if (c >= f.C) break;
      float gi = 0.f;                                  // a leaf the likelihood ignores
      if (is_w) gi = f.a[c * f.n + i] * xv;
      else if (is_aux) gi = f.b[c * f.n + i];
      const float t = gi - m[k];
      acc[k] = fmaf(t, t, acc[k]);
    }
  }
  const float fr = f.friction ? f.friction[p] : f.friction_scalar;
#pragma unroll
  for (int k = 0; k < kFisherChains; ++k) {
    const int64_t c = c0 + k;
    if (c >= f.C) break;
    const float v = f.inv_nm1 * acc[k];                // 1 / (n - 1) * ssq
    const float bb = f.half_eps * v;                   // 0.5 * step_size * v
    f.corr[c * f.P + p] = fr - bb;                     // friction - b
  }
}

__global__ void __launch_bounds__(256) k_fisher_finalize(const FisherArgs f) {
  const int64_t c = blockIdx.x;
  const float* corr = f.corr + c * f.P;
  float mn = INFINITY;
  for (int64_t p = threadIdx.x; p < f.P; p += 256) {
    const float v = corr[p];
    if (v > 0.f) mn = fminf(mn, v);                    // min(where(correction <= 0, inf, .))
  }
  __shared__ float red[256];
  red[threadIdx.x] = mn;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] = fminf(red[threadIdx.x], red[threadIdx.x + w]);
    __syncthreads();
  }
  const float smallest = red[0];
  for (int64_t p = threadIdx.x; p < f.P; p += 256) {
    const float v = corr[p];
    const float pos = v <= 0.f ? smallest : v;
    const float fr = f.friction ? f.friction[p] : f.friction_scalar;
    f.noise_scale[c * f.P + p] = __fsqrt_rn(pos);
    f.scale[c * f.P + p] = __fsqrt_rn(fr - pos);       // sqrt(friction - positive_correction)
  }
}

}  // namespace sgmc

using namespace sgmc;

extern "C" {

size_t sgmc_glm_fisher_scratch_floats(int64_t n_chains, int64_t P, int64_t batch_size) {
  return (size_t)(2 * n_chains * batch_size + n_chains * P);
}

int sgmc_glm_fisher_diag(void* stream, const sgmc_glm_spec* spec, const float* theta,
                         int64_t n_chains, int64_t P, const float* X, const float* y,
                         const int32_t* idx, int64_t batch_size, int64_t observation_count,
                         const float* grad, const float* friction, float friction_scalar,
                         float step_size, float* noise_scale, float* scale, float* scratch) {
  SGMC_REQUIRE(spec && theta && X && y && grad && noise_scale && scale && scratch, "null argument");
  SGMC_REQUIRE(batch_size >= 2, "the Fisher estimate needs at least two observations");
  if (n_chains == 0 || P == 0) return 0;
  FisherArgs f{};
  f.spec = *spec; f.theta = theta; f.C = n_chains; f.P = P;
  f.X = X; f.y = y; f.idx = idx; f.n = (int)batch_size; f.grad = grad;
  f.friction = friction; f.friction_scalar = friction_scalar;
  f.inv_N = 1.0f / (float)observation_count;
  f.inv_nm1 = (float)(1.0 / (double)(batch_size - 1));
  f.half_eps = 0.5f * step_size;
  f.a = scratch; f.b = scratch + n_chains * batch_size;
  f.corr = scratch + 2 * n_chains * batch_size;
  f.noise_scale = noise_scale; f.scale = scale;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t warps = n_chains * batch_size;
  k_fisher_coef<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(f);
  if (post_launch("k_fisher_coef")) return 1;
  SGMC_REQUIRE((n_chains + kFisherChains - 1) / kFisherChains <= 65535, "too many chains");
  k_fisher_ssq<<<dim3((unsigned)((P + 127) / 128), (unsigned)((n_chains + kFisherChains - 1) / kFisherChains)),
                 128, 0, s>>>(f);
  if (post_launch("k_fisher_ssq")) return 1;
  k_fisher_finalize<<<(unsigned)n_chains, 256, 0, s>>>(f);
  return post_launch("k_fisher_finalize");
}

}  // extern "C"
