// Argument block shared by the GLM potential kernels (glm_simt.cu, glm_tc.cu).
#pragma once
#include "common.cuh"
#include "glm_math.cuh"
#include "sgld_split.cuh"

namespace sgmc {

// The SGLD / pSGLD update the caller wants applied right after the gradient
// (sgmc_glm_sgld_step).  The tensor-core path applies it inside GEMM2's
// epilogue when the shapes allow and reports that through *applied.
struct FusedSgld {
  bool requested;
  float* theta_rw;          // f32[C][P], updated in place
  float* v;                 // RMSprop state or null
  const uint32_t* keys_in;
  uint32_t* keys_out;
  float step_size, temperature, alpha, lmbd;
  int layout;
  bool write_grad;
  bool* applied;
};

// sgmc_glm_sgld_step with SGMC_STEP_CARRY*: Theta's tensor-core operand form travels
// from one step's update to the next step's potential inside the workspace.
struct CarryCtx {
  int mode;                 // 1: first step of a carried sequence, 2: carried
  const uint32_t* keys_in;  // chain keys of this step (noise-key cache)
  uint32_t* keys_out;
  int prng_layout;
  bool active;              // set by glm_tc when the carried path was taken
  SgldSplitOut out;         // filled by glm_tc: what this step's update must write
};

struct GlmArgs {
  sgmc_glm_spec spec;
  const float* theta;     // f32[C][P]
  int64_t C, P;
  const float* X;         // f32[N_total][d]
  const float* y;         // f32[N_total]
  const int32_t* idx;     // int32[n] or null (rows 0..n-1); int32[C][n] when idx_stride = n
  int64_t idx_stride;     // 0: one minibatch shared by all chains; n: one per chain
  const float* mask;      // f32[n] or null
  int64_t n, N;
  float cot;              // (-N/n)/T
  float* potential;       // f32[C]
  float* variance;        // f32[C] or null
  float* grad;            // f32[C][P] or null
  float* R;               // f32[C][n] scratch: cot * d ell/dz
  float* ell;             // f32[C][n]
  bool ell_requested;     // the caller passed an ell output buffer
  float* tc_ws;           // extra scratch of the tensor-core path
  FusedSgld fused;
  CarryCtx* carry;        // null: stateless operand preparation
  int x_slot;             // which copy of the minibatch operands in the workspace (0 / 1)
  bool x_prepared;        // slot x_slot was staged by sgmc_glm_prepare_minibatch
  bool only_x;            // stage the minibatch operands into slot x_slot and return
};

// -(d prior / d theta_p) / T for flat parameter index p of chain c.
__device__ __forceinline__ float prior_grad_term(const GlmArgs& a, int64_t c, int p) {
  const sgmc_glm_spec& s = a.spec;
  if (s.prior == kPriorGaussian) {
    if (p >= s.prior_off && p < s.prior_off + s.prior_size) {
      const float inv = 1.0f / (s.prior_scale * s.prior_scale);
      // grad(prior) = -theta*inv  ->  g -= (-theta*inv)/T
      return (a.theta[c * a.P + p] * inv) / s.temperature;
    }
  } else if (s.prior == kPriorInvSigma) {
    if (p == s.prior_off)
      return (1.0f / expf(a.theta[c * a.P + p])) / s.temperature;
  }
  return 0.0f;
}

int glm_finalize(cudaStream_t stream, const GlmArgs& a);
int glm_simt(cudaStream_t stream, const GlmArgs& a);
int glm_tc(cudaStream_t stream, const GlmArgs& a, int path);
size_t glm_tc_workspace_bytes(int64_t n_chains, int64_t batch_size, int64_t d,
                              int path);
int32_t* glm_tc_spare_idx(float* tc_ws, int64_t n_chains, int64_t batch_size, int64_t d);

// which copy of the minibatch operands a call works on (sgmc_glm_prepare_minibatch)
struct XStage {
  int slot = 0;
  bool prepared = false;
  bool only_x = false;
};

}  // namespace sgmc
