// reSGLD with a temperature ladder sharded over GPUs (BASELINE.json configs[3]).
//
// Generalises the swap step of solver.parallel_tempering.update
// (jax_sgmc/solver.py:273-291) from two chains to R
// replicas.  Replica r (on some rank) runs B independent systems; instead of
// exchanging whole chain states (positions, keys, RMSprop state -- megabytes
// over NVLink) the replicas exchange their TEMPERATURE LABELS: only the
// all-gathered per-replica (U, var) scalars cross NVLink, and every rank
// evaluates the identical decision from the identical key, so the ladder state
// stays replicated without further communication.
//
// holder[t][b] = replica currently running system b at temperature index t.
// Pair p = (t, t+1) is attempted at step k iff R == 2 (every step, the
// reference's behaviour) or p % 2 == k % 2 (deterministic even/odd sweep; an
// extension -- the reference has only two temperatures).  The decision uses the
// reference's arithmetic and its (inverted) predicate: exchange iff
// !(log_u < log_s).
#include "common.cuh"
#include "rng.cuh"

namespace sgmc {

constexpr int kMaxReplicas = 32;

// One thread per system b walks the ladder: decisions of the attempted pairs
// (disjoint within a step, so applying them on the fly equals deciding all
// pairs first), label table update, and the labels of the local replicas
// [r0, r0 + n_local) -- the temperature every system runs at next step and its
// index (0 = the cold one).  All arrays are [*][B]: coalesced over b.
__global__ void k_resgld_ladder_fused(const float* __restrict__ gathered,   // [R][2][B] (U, var)
                                      int32_t* __restrict__ holder,        // [R][B]
                                      float* __restrict__ ssq,             // [R-1][B]
                                      const float* __restrict__ F,         // [B]
                                      const float* __restrict__ temps,     // [R]
                                      const uint32_t* __restrict__ keys_in,   // [R-1][B][2]
                                      uint32_t* __restrict__ keys_out,
                                      int32_t* __restrict__ exchange,      // [R-1][B]
                                      int R, int64_t B, float eta, int parity, int layout,
                                      int r0, int n_local, float* __restrict__ temp_out,
                                      int32_t* __restrict__ tidx_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int32_t h[kMaxReplicas];
#pragma unroll 4
  for (int t = 0; t < R; ++t) h[t] = holder[(int64_t)t * B + b];
  const float Fb = F[b];
  for (int p = 0; p < R - 1; ++p) {
    const int64_t i = (int64_t)p * B + b;
    const bool active = (R == 2) || ((p & 1) == parity);
    const Key k{keys_in[2 * i], keys_in[2 * i + 1]};
    if (!active) {                     // untouched pair: carry state forward
      keys_out[2 * i] = k.k0;
      keys_out[2 * i + 1] = k.k1;
      exchange[i] = 0;
      continue;
    }
    const int r_lo = h[p], r_hi = h[p + 1];
    const float U_n = gathered[((int64_t)r_lo * 2) * B + b];
    const float var_n = gathered[((int64_t)r_lo * 2 + 1) * B + b];
    const float U_h = gathered[((int64_t)r_hi * 2) * B + b];
    const float q = __fadd_rn(__fmul_rn(__fadd_rn(1.0f, -eta), ssq[i]), __fmul_rn(eta, var_n));
    ssq[i] = q;
    const float tau = __fadd_rn(__fdiv_rn(1.0f, temps[p]), -__fdiv_rn(1.0f, temps[p + 1]));
    const float corr = __fdiv_rn(__fmul_rn(tau, q), Fb);
    const float log_s = __fmul_rn(tau, __fadd_rn(__fadd_rn(U_n, -U_h), -corr));
    Key nk, sub;
    split2(k, layout, nk, sub);
    keys_out[2 * i] = nk.k0;
    keys_out[2 * i + 1] = nk.k1;
    const float u = bits_to_uniform(random_word(sub, 0, 1, layout), 0.0f, 1.0f);
    const float log_u = log_libdevice(u);
    const bool swap = !(log_u < log_s);       // the reference's (inverted) predicate
    exchange[i] = swap ? 1 : 0;
    if (swap) {
      h[p] = r_hi;
      h[p + 1] = r_lo;
      holder[(int64_t)p * B + b] = r_hi;
      holder[(int64_t)(p + 1) * B + b] = r_lo;
    }
  }
  for (int l = 0; l < n_local; ++l) {
    int t = 0;
    for (int tt = 0; tt < R; ++tt)
      if (h[tt] == r0 + l) t = tt;
    temp_out[(int64_t)l * B + b] = temps[t];
    tidx_out[(int64_t)l * B + b] = t;
  }
}

}  // namespace sgmc

using namespace sgmc;

extern "C" int sgmc_resgld_ladder_step(void* stream, const float* gathered, int32_t* holder,
                                       float* ssq, const float* F, const float* temps,
                                       const uint32_t* keys_in, uint32_t* keys_out,
                                       int32_t* exchange, int n_replicas, int64_t n_systems,
                                       int64_t step, int first_local_replica,
                                       int n_local_replicas, float* temp_per_chain,
                                       int32_t* temp_index, int prng_layout) {
  SGMC_REQUIRE(keys_in != keys_out, "keys_out must not alias keys_in");
  SGMC_REQUIRE(n_replicas >= 2 && n_systems > 0 && step >= 1, "bad ladder arguments");
  SGMC_REQUIRE(n_replicas <= kMaxReplicas, "at most %d replicas", kMaxReplicas);
  const float eta = 1.0f / (float)step;
  launch_pdl(k_resgld_ladder_fused, dim3((unsigned)((n_systems + 127) / 128)), dim3(128), 0,
             (cudaStream_t)stream, gathered, holder, ssq, F, temps, keys_in, keys_out, exchange,
             n_replicas, n_systems, eta, (int)(step & 1), prng_layout, first_local_replica,
             n_local_replicas, temp_per_chain, temp_index);
  return post_launch("k_resgld_ladder_fused");
}
