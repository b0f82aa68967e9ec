// Carried operand split: what the SGLD / pSGLD update writes for the tensor-core
// potential of the next step (update_kernels.cu: SgldSplitOp; consumer:
// k_prepare_all / k_glm_tc_fused in glm_tc.cu).
#pragma once
#include "common.cuh"

namespace sgmc {

struct SgldSplitOut {
  int fmt;                     // 1: fp16 hi/lo with a per-row power-of-two scale, 2: bf16
  void* th_hi;                 // [C][P] halves
  void* th_lo;                 // [C][P] halves (fmt 1)
  const float* scale;          // f32[C]: scale of this update's split
  uint32_t* amax_bits;         // u32[C]: max |theta'| per row (atomicMax on the bits)
  float* sumsq_part;           // f32[C][tiles_per_chain]: partial sums of theta'^2 (prior range)
  int prior_lo, prior_hi;
  float prior_coef;            // != 0: grad lacks the prior term; the update adds theta * coef
  float* grad_rw;              // the completed gradient is written back here (or null)
  const uint32_t* noise_keys;  // u32[C][2] noise keys derived earlier in the step, or null
  const float* xi;             // f32[C][P]: the step's noise, generated earlier in the step
                               // (k_glm_tc_pair's shadow job), or null: generate it here
};

// warp-tiles per chain of the update kernel for a one-leaf sample of P elements
int sgld_split_tiles_per_chain(int64_t P);

int sgld_update_split(cudaStream_t stream, float* theta, float* v, const float* grad,
                      const uint32_t* keys_in, uint32_t* keys_out, int64_t n_chains, int64_t P,
                      float step_size, float temperature, const float* temp_per_chain,
                      float alpha, float lmbd, int prng_layout, const SgldSplitOut& so);

}  // namespace sgmc
