// Tensor-core (tcgen05) GLM potential path -- placeholder until the tcgen05
// kernels land; fails loudly (no silent fallback to the SIMT path).
#include "glm.cuh"

namespace sgmc {

size_t glm_tc_workspace_bytes(int64_t, int64_t, int64_t, int) { return 0; }

int glm_tc(cudaStream_t, const GlmArgs&, int path) {
  set_error("GLM tensor-core path %d is not built in this library", path);
  return 3;
}

}  // namespace sgmc
