// Per-observation GLM likelihood math shared by the SIMT and tensor-core
// potential kernels.  Restates, for the recognised families, the user
// likelihoods of the reference examples and their reverse-mode derivatives
// (what jax.value_and_grad of potential.minibatch_potential produces,
// jax_sgmc/potential.py:159-214; integrator.py:166,593,792).
#pragma once
#include "common.cuh"

namespace sgmc {

enum : int { kFamilyGaussian = 0, kFamilyLogistic = 1 };
enum : int { kPriorFlat = 0, kPriorGaussian = 1, kPriorInvSigma = 2 };

// Per-chain constants of the gaussian family (examples/quickstart.md:164-169,
// jax.scipy.stats.norm.logpdf): s2 = exp(log_sigma)^2, ln = log(2 pi s2).
struct GaussConst {
  float s2, ln;
};
__device__ __forceinline__ GaussConst gauss_const(float log_sigma) {
  const float sigma = expf(log_sigma);
  GaussConst g;
  g.s2 = sigma * sigma;
  g.ln = logf(6.2831855f * g.s2);
  return g;
}

// ell and d ell / d z for one observation; z = x.w (+ bias).
//   gaussian: ell = (ln + r^2/s2) / -2, r = y - z;   d ell/dz = r / s2
//   logistic: ell = y z - softplus(z);                d ell/dz = y - sigmoid(z)
__device__ __forceinline__ void glm_link(int family, float z, float y,
                                         GaussConst gc, float& ell, float& dz) {
  if (family == kFamilyGaussian) {
    const float r = y - z;
    const float q = (r * r) / gc.s2;
    ell = (gc.ln + q) / -2.0f;
    dz = r / gc.s2;
  } else {
    const float e = expf(-fabsf(z));
    const float sp = fmaxf(z, 0.0f) + log1pf(e);
    ell = y * z - sp;
    const float den = 1.0f + e;
    const float sig = z >= 0.0f ? 1.0f / den : e / den;
    dz = y - sig;
  }
}

}  // namespace sgmc
