// reSGLD swap step of solver.parallel_tempering.update
// (jax_sgmc/solver.py:273-291), batched over S
// independent two-temperature systems.
#include "common.cuh"
#include "rng.cuh"

namespace sgmc {

__global__ void k_resgld_decide(const float* __restrict__ U_n,
                                const float* __restrict__ U_h,
                                const float* __restrict__ var_n,
                                float* __restrict__ ssq,
                                const float* __restrict__ F, float eta,
                                float temps, const uint32_t* __restrict__ keys_in,
                                uint32_t* __restrict__ keys_out,
                                int32_t* __restrict__ exchange, int64_t S,
                                int layout) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  // ssq' = ((1-eta)*ssq) + (eta*var_n)                      solver.py:275-276
  const float q = __fadd_rn(__fmul_rn(__fadd_rn(1.0f, -eta), ssq[s]),
                            __fmul_rn(eta, var_n[s]));
  ssq[s] = q;
  // log_s = temps * (U_n - U_h - temps*ssq'/F)              solver.py:280-281
  const float corr = __fdiv_rn(__fmul_rn(temps, q), F[s]);
  const float log_s =
      __fmul_rn(temps, __fadd_rn(__fadd_rn(U_n[s], -U_h[s]), -corr));
  Key k{keys_in[2 * s], keys_in[2 * s + 1]}, nk, sub;
  split2(k, layout, nk, sub);                               // solver.py:283
  keys_out[2 * s] = nk.k0;
  keys_out[2 * s + 1] = nk.k1;
  const uint32_t w = random_word(sub, 0, 1, layout);        // uniform(split), shape ()
  const float u = bits_to_uniform(w, 0.0f, 1.0f);
  const float log_u = log_libdevice(u);                     // solver.py:284
  // lax.cond(log_u < log_s, keep, swap): exchange iff NOT (log_u < log_s)
  exchange[s] = (log_u < log_s) ? 0 : 1;                    // solver.py:287-291
}

// Metropolis-Hastings accept/reject of solver.sggmc (mode 0, solver.py:524-539)
// and solver.amagold (mode 1, solver.py:381-395) for C chains.
__global__ void k_mh_decide(int mode, float* __restrict__ U_state,
                            const float* __restrict__ U_new, const float* __restrict__ e0,
                            const float* __restrict__ e1, float neg_inv_T,
                            const uint32_t* __restrict__ keys_in,
                            uint32_t* __restrict__ keys_out, int32_t* __restrict__ reject,
                            float* __restrict__ ratio, int64_t n, int layout) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float la;
  if (mode == 0) {
    // -1/T * (U_new - U_old + ke_end - ke_start), capped at 0      solver.py:524-529
    const float s = __fadd_rn(__fadd_rn(__fadd_rn(U_new[c], -U_state[c]), e1[c]), -e0[c]);
    la = __fmul_rn(neg_inv_T, s);
    la = (la <= 0.0f) ? la : 0.0f;
  } else {
    // U_old - U_new + accumulated energy, capped at 0               solver.py:381-383
    la = __fadd_rn(__fadd_rn(U_state[c], -U_new[c]), e1[c]);
    la = (la > 0.0f) ? 0.0f : la;
  }
  Key k{keys_in[2 * c], keys_in[2 * c + 1]}, nk, sub;
  split2(k, layout, nk, sub);                               // key, split = split(key, 2)
  keys_out[2 * c] = nk.k0;
  keys_out[2 * c + 1] = nk.k1;
  const float u = bits_to_uniform(random_word(sub, 0, 1, layout), 0.0f, 1.0f);
  const bool accept = log_libdevice(u) < la;                // lax.cond(log(slice) < log_alpha
  reject[c] = accept ? 0 : 1;
  if (accept) U_state[c] = U_new[c];
  ratio[c] = expf(la);                                      // acceptance_ratio statistic
}

__global__ void k_swap_rows(uint32_t* __restrict__ a, uint32_t* __restrict__ b,
                            const int32_t* __restrict__ exchange,
                            int64_t n_rows, int64_t row_words) {
  const int64_t row = blockIdx.y;
  if (row >= n_rows || exchange[row] == 0) return;
  uint32_t* pa = a + row * row_words;
  uint32_t* pb = b + row * row_words;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < row_words;
       j += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t t = pa[j];
    pa[j] = pb[j];
    pb[j] = t;
  }
}

}  // namespace sgmc

using namespace sgmc;

extern "C" {

int sgmc_resgld_decide_eta(void* stream, const float* U_normal, const float* U_hot,
                           const float* var_normal, float* ssq, const float* F,
                           float eta, float T_normal, float T_hot,
                           const uint32_t* keys_in, uint32_t* keys_out,
                           int32_t* exchange, int64_t n_systems, int prng_layout) {
  SGMC_REQUIRE(keys_in != keys_out, "keys_out must not alias keys_in");
  if (n_systems == 0) return 0;
  const float temps = 1.0f / T_normal - 1.0f / T_hot;       // :279
  k_resgld_decide<<<(unsigned)((n_systems + 127) / 128), 128, 0,
                    (cudaStream_t)stream>>>(
      U_normal, U_hot, var_normal, ssq, F, eta, temps, keys_in, keys_out,
      exchange, n_systems, prng_layout);
  return post_launch("sgmc_resgld_decide");
}

int sgmc_resgld_decide(void* stream, const float* U_normal, const float* U_hot,
                       const float* var_normal, float* ssq, const float* F,
                       int64_t step, float T_normal, float T_hot,
                       const uint32_t* keys_in, uint32_t* keys_out,
                       int32_t* exchange, int64_t n_systems, int prng_layout) {
  SGMC_REQUIRE(step >= 1, "step must be >= 1 (already incremented)");
  // the reference's default sa_schedule, 1 / n (solver.py:221)
  return sgmc_resgld_decide_eta(stream, U_normal, U_hot, var_normal, ssq, F,
                                1.0f / (float)step, T_normal, T_hot, keys_in, keys_out,
                                exchange, n_systems, prng_layout);
}

int sgmc_mh_decide(void* stream, int mode, float* U_state, const float* U_new,
                   const float* e0, const float* e1, float temperature,
                   const uint32_t* keys_in, uint32_t* keys_out, int32_t* reject,
                   float* ratio, int64_t n_chains, int prng_layout) {
  SGMC_REQUIRE(keys_in != keys_out, "keys_out must not alias keys_in");
  SGMC_REQUIRE(mode == 0 || mode == 1, "unknown MH mode %d", mode);
  SGMC_REQUIRE(e1 != nullptr && (mode == 1 || e0 != nullptr), "null energy argument");
  if (n_chains == 0) return 0;
  k_mh_decide<<<(unsigned)((n_chains + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      mode, U_state, U_new, e0, e1, -1.0f / temperature, keys_in, keys_out, reject, ratio,
      n_chains, prng_layout);
  return post_launch("sgmc_mh_decide");
}

int sgmc_swap_rows(void* stream, void* a, void* b, const int32_t* exchange,
                   int64_t n_rows, int64_t row_bytes) {
  SGMC_REQUIRE(row_bytes % 4 == 0, "row_bytes must be a multiple of 4");
  if (n_rows == 0 || row_bytes == 0) return 0;
  const int64_t words = row_bytes / 4;
  unsigned gx = (unsigned)((words + 255) / 256);
  if (gx > 64) gx = 64;
  // gridDim.y holds at most 65535 rows: larger row counts go in slices
  for (int64_t r0 = 0; r0 < n_rows; r0 += 65535) {
    const int64_t nr = n_rows - r0 < 65535 ? n_rows - r0 : 65535;
    k_swap_rows<<<dim3(gx, (unsigned)nr), 256, 0, (cudaStream_t)stream>>>(
        (uint32_t*)a + r0 * words, (uint32_t*)b + r0 * words, exchange + r0, nr, words);
  }
  return post_launch("sgmc_swap_rows");
}

}  // extern "C"
