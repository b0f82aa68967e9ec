// The per-element arithmetic of the SGLD / pSGLD step (SURVEY Appendix A.2;
// integrator.py:882-914, adaption.py:254-291), shared by the stand-alone fused
// update (update_kernels.cu) and the GEMM2-epilogue update (glm_tc.cu) so both
// produce the same bits from the same gradient and noise.
#pragma once
#include <cuda_runtime.h>

namespace sgmc {

// theta' from (theta, g, v, xi).  RMS: v is updated in place.  FAST: the
// preconditioner uses the SFU approximations (flush-to-zero forms: one MUFU each, no
// denormal fix-up code; v >= 0 and lmbd + sqrt(v) >= lmbd > 0) + FMA contraction (<= 2 ulp);
// otherwise every operation is a separately rounded IEEE op (oracle order).
template <bool RMS, bool FAST>
__device__ __forceinline__ float sgld_one(float t, float g, float& vv, float xi, float ns,
                                          float neg_eps, float alpha, float one_m_alpha,
                                          float lmbd) {
  const float sg = __fmul_rn(neg_eps, g);
  const float sn = __fmul_rn(ns, xi);
  float delta;
  if (RMS && FAST) {
    vv = fmaf(alpha, vv, one_m_alpha * (g * g));
    float s, G, S;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(vv));
    const float den = lmbd + s;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(G) : "f"(den));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(S) : "f"(den));   // sqrt(1/den)
    delta = fmaf(S, sn, G * sg);
  } else if (RMS) {
    vv = __fadd_rn(__fmul_rn(alpha, vv), __fmul_rn(one_m_alpha, __fmul_rn(g, g)));
    const float G = __frcp_rn(__fadd_rn(lmbd, __fsqrt_rn(vv)));
    const float S = __fsqrt_rn(G);
    // (eps*Gamma + G*sg) + S*sn with Gamma == 0; the "0 +" only affects the
    // sign of an exact zero and is dropped.
    delta = __fadd_rn(__fmul_rn(G, sg), __fmul_rn(S, sn));
  } else {
    delta = __fadd_rn(sg, sn);
  }
  return __fadd_rn(t, delta);
}

}  // namespace sgmc
