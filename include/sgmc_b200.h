/* sgmc_b200.h -- C ABI of libsgmc_b200.so
 *
 * B200 (sm_100a) kernels for the data-parallel sampling hot path of
 * tummfm/jax-sgmc.  The reference is pure Python/JAX and has no FFI of its
 * own; these entry points are what an XLA-FFI / ctypes binding for that path
 * binds.  Each entry cites the reference code it
 * replaces (paths relative to the reference repository root).
 *
 * Conventions
 *  - plain pointers and sizes; no torch / jax types.
 *  - every compute entry only ENQUEUES work on `stream` (a cudaStream_t passed
 *    as void*): no allocation, no synchronisation -> CUDA-graph capturable and
 *    usable from an XLA custom call.
 *  - return value: 0 = ok, nonzero = error; message via sgmc_last_error()
 *    (thread local).
 *  - layout: chain-batched flat parameters, f32[C][P] row-major, where P is the
 *    raveled sample (jax.flatten_util.ravel_pytree order) and C the number of
 *    independent chains (the reference's leading list_vmap axis,
 *    jax_sgmc/util/list_map.py:53-56).  PRNG keys are uint32[C][2].
 *  - `leaf_sizes[n_leaves]` describes the pytree leaves in tree_flatten order
 *    (sum = P): integrator.random_tree (jax_sgmc/integrator.py:119-135) draws
 *    one jax.random.normal stream per leaf from split(key, n_leaves).
 *  - prng_layout: 0 = "original" threefry counter layout (reference-era JAX
 *    default), 1 = "partitionable" (jax_threefry_partitionable=True).
 */
#ifndef SGMC_B200_H_
#define SGMC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGMC_MAX_LEAVES 64

/* ---- library / runtime helpers (plumbing for hosts without a CUDA binding) */
const char* sgmc_last_error(void);
int  sgmc_version(void);
int  sgmc_device_count(int* count);
int  sgmc_set_device(int device);
int  sgmc_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                      size_t* total_mem);
int  sgmc_malloc(void** dptr, size_t bytes);
int  sgmc_free(void* dptr);
int  sgmc_host_alloc(void** hptr, size_t bytes);      /* pinned */
int  sgmc_host_free(void* hptr);
int  sgmc_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int  sgmc_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int  sgmc_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
int  sgmc_memset(void* dst, int value, size_t bytes, void* stream);
int  sgmc_stream_create(void** stream);
int  sgmc_stream_destroy(void* stream);
int  sgmc_stream_sync(void* stream);
int  sgmc_device_sync(void);
int  sgmc_event_create(void** event);
int  sgmc_event_destroy(void* event);
int  sgmc_event_record(void* event, void* stream);
int  sgmc_event_sync(void* event);
int  sgmc_stream_wait_event(void* stream, void* event);
int  sgmc_event_elapsed_ms(void* start, void* stop, float* ms);
/* Process-wide options.
 * SGMC_OPT_EXACT_UPDATE_MATH: 0 (default) = the RMSprop preconditioner
 *   arithmetic (sqrt, reciprocal) uses the SFU approximations (<= 2 ulp) and
 *   FMA contraction; 1 = IEEE-exact unfused arithmetic, bit-identical to the
 *   NumPy oracle given the same gradient.  The Gaussian noise is bit-exact in
 *   both modes.
 * SGMC_OPT_SERIAL_LAUNCH: 0 (default) = the step's kernels are launched with
 *   programmatic dependent launch (each kernel's prologue and CTA ramp overlap
 *   the previous kernel's tail; every kernel executes griddepcontrol.wait
 *   before its first global access); 1 = plain stream-serialised launches. */
/* SGMC_OPT_FUSED_STEP_EPILOGUE: 1 = sgmc_glm_sgld_step applies the update inside
 *   the gradient GEMM's epilogue when the shapes allow (same bits; measured
 *   slower than the two-kernel sequence on B200 so far); 0
 *   (default) = potential/gradient kernels followed by the fused update kernel. */
enum { SGMC_OPT_EXACT_UPDATE_MATH = 0, SGMC_OPT_SERIAL_LAUNCH = 1,
       SGMC_OPT_FUSED_STEP_EPILOGUE = 2, SGMC_OPT_COUNT = 4 };
int  sgmc_set_option(int option, int value);
int  sgmc_get_option(int option);
/* counts kernels launched by this library since load (bench "gpu_launches") */
unsigned long long sgmc_launch_count(void);

/* ---- PRNG: jax.random on device (third-party jax/_src/prng.py, random.py;
 *      reference call sites integrator.py:131-133,208,630,736,871;
 *      solver.py:254,283-284; data/numpy_loader.py:132-134) --------------- */

/* random.split(key, num) for C keys: keys_out[C][num][2]. */
int sgmc_prng_split(void* stream, const uint32_t* keys_in, uint32_t* keys_out,
                    int64_t n_keys, int num, int prng_layout);
/* random.bits / uniform / normal of shape (n,) per key: out[C][n]. */
int sgmc_random_bits(void* stream, const uint32_t* keys, uint32_t* out,
                     int64_t n_keys, int64_t n, int prng_layout);
int sgmc_uniform(void* stream, const uint32_t* keys, float* out,
                 int64_t n_keys, int64_t n, float minval, float maxval,
                 int prng_layout);
int sgmc_normal(void* stream, const uint32_t* keys, float* out,
                int64_t n_keys, int64_t n, int prng_layout);
/* integrator.random_tree (integrator.py:119-135): noise[C][P] shaped like the
 * sample, per-leaf keys from split(key, n_leaves). */
int sgmc_normal_like(void* stream, const uint32_t* keys, float* noise,
                     int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                     int prng_layout);
/* random.randint(key, (n,), minval, maxval) int32 (numpy_loader.py:133). */
int sgmc_randint(void* stream, const uint32_t* key, int32_t* out, int64_t n,
                 int32_t minval, int32_t maxval, int prng_layout);

/* ---- minibatch draw: DeviceNumpyDataLoader.get_random_data
 *      (data/numpy_loader.py:128-141): key' , split = split(key);
 *      idx = randint(split, (n,), 0, N).  key_out must not alias key_in. */
int sgmc_minibatch_draw(void* stream, const uint32_t* key_in,
                        uint32_t* key_out, int32_t* idx, int64_t batch_size,
                        int64_t observation_count, int prng_layout);
/* tree_index (data/core.py:642-660): out[n][row_elems] = src[idx[i]][:]. */
int sgmc_gather_rows(void* stream, const float* src, const int32_t* idx,
                     float* out, int64_t n, int64_t row_elems);

/* Synthetic logistic-regression data set generated directly in HBM (bench /
 * test support): kx,kw,ky = split(key,3);
 * X = normal(kx,(N,d))/sqrt(d); w = normal(kw,(d,)); y = uniform(ky) <
 * sigmoid(X w).  `key` is a HOST pointer to uint32[2].  Synchronises. */
int sgmc_synth_logistic_data(void* stream, const uint32_t* key, float* X,
                             float* y, float* w, int64_t N, int64_t d,
                             int prng_layout);

/* ---- fused integrator updates (one elementwise pass, noise in-kernel) ---- */

/* integrator.langevin_diffusion.update_fn (integrator.py:860-922) without
 * adaption:  theta' = theta + ((-eps)*g + sqrt(2*T*eps)*xi);
 * key', split = split(key); xi = random_tree(split, theta).
 * theta updated in place; keys_out must not alias keys_in.
 * temp_per_chain (device f32[C]) overrides `temperature` when non-NULL. */
int sgmc_sgld_update(void* stream, float* theta, const float* grad,
                     const uint32_t* keys_in, uint32_t* keys_out,
                     int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                     float step_size, float temperature,
                     const float* temp_per_chain, int prng_layout);

/* Same with adaption.rms_prop (adaption.py:225-293) fused:
 * v' = alpha*v + (1-alpha)*g*g;  G = 1/(lmbd + sqrt(v'));
 * theta' = theta + ((eps*0 + G*(-eps*g)) + sqrt(G)*(sqrt(2*T*eps)*xi)).
 * theta and v updated in place. */
int sgmc_sgld_rms_update(void* stream, float* theta, float* v,
                         const float* grad, const uint32_t* keys_in,
                         uint32_t* keys_out, int64_t n_chains,
                         const int64_t* leaf_sizes, int n_leaves,
                         float step_size, float temperature,
                         const float* temp_per_chain, float alpha, float lmbd,
                         int prng_layout);

/* Stand-alone adaption.rms_prop (adaption.py:225-293) for users of the
 * (init, update, get) triplet outside the fused update:
 * update: v' = alpha v + (1-alpha) g^2 (in place, :270-272);
 * get   : g_inv = 1/(lmbd + sqrt(v)), sqrt_g_inv = sqrt(g_inv) (:289-291). */
int sgmc_rms_prop_update(void* stream, float* v, const float* grad, int64_t n,
                         float alpha);
int sgmc_rms_prop_get(void* stream, const float* v, float* g_inv,
                      float* sqrt_g_inv, int64_t n, float lmbd);
/* out = a*x + b*y on per-chain scalars (e.g. obabo's potential =
 * 0.5*(U1+U2), integrator.py:264). */
int sgmc_axpby(void* stream, float* out, float a, const float* x, float b,
               const float* y, int64_t n);

/* util.tree_scale / tree_add / tree_multiply / tree_dot
 * (util/tree_util.py:58-133) on the flat chain-batched layout.
 * op 0: out = alpha*x (tree_scale), 1: out = x + y (tree_add),
 * 2: out = x*y (tree_multiply); tree_dot: out[c] = <x[c,:], y[c,:]>. */
int sgmc_tree_ewise(void* stream, int op, float* out, float alpha,
                    const float* x, const float* y, int64_t n);
int sgmc_tree_dot(void* stream, float* out, const float* x, const float* y,
                  int64_t n_chains, int64_t P);

/* integrator.friction_leapfrog (integrator.py:563-765).
 * begin: key', split = split(key); p = sqrt(m) * random_tree(split)  (:736-738)
 *        fused with the first position update theta += eps * (p / m) (:610-612).
 * step : p1 = p + (-eps*C)*(p/m); p2 = p1 + (-eps)*g; key',split = split(key);
 *        p3 = p2 + C*(sqrt(2 eps)*xi)  (:616-655); unless `last`, fused with
 *        the next position update theta += eps * (p3 / m).
 * mass / friction: device f32[P] or NULL (unit mass / scalar friction). */
int sgmc_sghmc_begin(void* stream, float* theta, float* momentum,
                     const uint32_t* keys_in, uint32_t* keys_out,
                     int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                     float step_size, const float* mass, int prng_layout);
int sgmc_sghmc_step(void* stream, float* theta, float* momentum,
                    const float* grad, const uint32_t* keys_in,
                    uint32_t* keys_out, int64_t n_chains,
                    const int64_t* leaf_sizes, int n_leaves, float step_size,
                    float friction_scalar, const float* friction,
                    const float* mass, int last, int prng_layout);

/* integrator.obabo step (integrator.py:203-273) as two passes.
 * Keys: one split(key, 3) per step (:208): pass A consumes split1 and writes
 * key' to keys_out; pass B consumes split2 derived from the SAME keys_in.
 * pass A (after grad at theta): p1 = O(p, xi1); ke_start += 0.5 <p1, p1/m>;
 *        p2 = -(0.5 eps) g1 + p1; theta += eps * (p2/m).
 * pass B (after grad at theta'): p3 = -(0.5 eps) g2 + p2;
 *        ke_end += 0.5 <p3, p3/m>; p4 = O(p3, xi2).
 * O(p, xi) = sqrt(a) p + sqrt((1-a) T) (sqrt(m) xi), a = exp(-friction eps). */
int sgmc_obabo_pass_a(void* stream, float* theta, float* momentum,
                      const float* grad, float* ke_start,
                      const uint32_t* keys_in, uint32_t* keys_out,
                      int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                      float step_size, float temperature, float friction,
                      const float* mass, int prng_layout);
int sgmc_obabo_pass_b(void* stream, float* momentum, const float* grad,
                      float* ke_end, const uint32_t* keys_in,
                      int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                      float step_size, float temperature, float friction,
                      const float* mass, int prng_layout);

/* ---- GLM stochastic potential + gradient
 *      potential.minibatch_potential (potential.py:94-216) evaluated together
 *      with its reverse-mode gradient (integrator.py:166,593,792) for the
 *      recognised GLM families, batched over chains that share the minibatch.
 *
 * family: 0 = gaussian linear regression with learned log_sigma
 *             (examples/quickstart.md:158-176), 1 = logistic regression.
 * prior : 0 = flat, 1 = gaussian N(0, prior_scale^2) on [prior_off,
 *             prior_off+prior_size), 2 = 1/exp(theta[prior_off]) (quickstart).
 * X f32[N_total][d], y f32[N_total] is the full reference data resident in
 * HBM; idx int32[n] selects the minibatch (gather is fused); mask f32[n] or
 * NULL (potential.py:182-185).
 * Outputs: potential[C] = U, variance[C] = var(ell) (integrator.py:880),
 * grad[C][P] = dU/dtheta, ell[C][n] (optional, may be NULL). */
typedef struct {
  int32_t family;
  int32_t d;               /* features */
  int32_t w_off;           /* offset of the weight vector in the flat sample */
  int32_t aux_off;         /* log_sigma (gaussian) / bias (logistic) or -1 */
  int32_t prior;
  int32_t prior_off;
  int32_t prior_size;
  float   prior_scale;
  float   temperature;     /* potential temperature T (potential.py:99) */
  float   x_absmax;        /* max |X| over the data set if known (> 0): lets the
                              tensor-core path pick its fp16 scale without a
                              per-minibatch reduction; 0 = compute per call */
} sgmc_glm_spec;

/* out[0] = max |x[i]| (device scalar; e.g. the data-set bound for
 * sgmc_glm_spec.x_absmax, computed once when a data set is registered). */
int sgmc_absmax(void* stream, const float* x, int64_t n, float* out);

/* `workspace` is device scratch of at least sgmc_glm_workspace_bytes(C, n)
 * bytes (residuals and per-observation likelihoods between the two GEMM-shaped
 * passes); the library never allocates.
 * `path`: 0 = fp32 SIMT kernels (any shape; the precise path),
 *         1 = tcgen05 tensor-core kernels, 3-way split fp16 operands with fp32
 *             accumulation in TMEM ("parity" mode, ~fp32 accuracy),
 *         2 = tcgen05, single-pass bf16 operands ("throughput" mode).
 * Paths 1/2 require d % 64 == 0, n % 64 == 0. */
size_t sgmc_glm_workspace_bytes(int64_t n_chains, int64_t batch_size,
                                int64_t d, int path);
int sgmc_glm_potential_grad(void* stream, const sgmc_glm_spec* spec,
                            const float* theta, int64_t n_chains, int64_t P,
                            const float* X, const float* y, const int32_t* idx,
                            const float* mask, int64_t batch_size,
                            int64_t observation_count, float* potential,
                            float* variance, float* grad, float* ell,
                            void* workspace, size_t workspace_bytes, int path);

/* One whole integrator.langevin_diffusion.update_fn step for the GLM potential
 * (integrator.py:860-922): the minibatch potential + gradient at the current
 * theta (sgmc_glm_potential_grad; `potential` / `variance` describe theta
 * BEFORE the update, as LangevinState.potential / .variance do) followed by
 * sgmc_sgld_rms_update (v != NULL) or sgmc_sgld_update (v == NULL) on that
 * gradient.  With SGMC_OPT_FUSED_STEP_EPILOGUE on the tensor-core paths, when
 * the sample is exactly the weight vector (P == d, d % 256 == 0, n_chains %
 * 128 == 0, original threefry layout, SGMC_OPT_EXACT_UPDATE_MATH off), the
 * update runs inside the epilogue of the gradient GEMM (the Gaussian noise is
 * generated while the tensor pipe works and the gradient never goes to memory
 * unless `grad` is wanted); every other case runs the kernels one after the
 * other.  Same results either way (bit for bit).
 * `grad` (f32[C][P]) must be a valid buffer; with write_grad != 0 it receives
 * dU/dtheta, with write_grad == 0 its contents are unspecified afterwards (the
 * fused kernel then skips the store). */
int sgmc_glm_sgld_step(void* stream, const sgmc_glm_spec* spec, float* theta, float* v,
                       int64_t n_chains, int64_t P, const float* X, const float* y,
                       const int32_t* idx, const float* mask, int64_t batch_size,
                       int64_t observation_count, float* potential, float* variance,
                       float* grad, const uint32_t* keys_in, uint32_t* keys_out,
                       float step_size, float temperature, float alpha, float lmbd,
                       void* workspace, size_t workspace_bytes, int path, int prng_layout,
                       int write_grad);

/* ---- reSGLD: solver.parallel_tempering.update swap step
 *      (solver.py:273-291).  For S systems: ssq' = (1-1/k) ssq + var_n/k;
 *      log_s = tau (U_n - U_h - tau ssq'/F), tau = 1/T_n - 1/T_h;
 *      key', split = split(key); log_u = log(uniform(split));
 *      exchange[s] = !(log_u < log_s)   (the reference's inverted predicate).
 * keys_out must not alias keys_in. */
int sgmc_resgld_decide(void* stream, const float* U_normal, const float* U_hot,
                       const float* var_normal, float* ssq, const float* F,
                       int64_t step, float T_normal, float T_hot,
                       const uint32_t* keys_in, uint32_t* keys_out,
                       int32_t* exchange, int64_t n_systems, int prng_layout);
/* lax.cond swap of whole chain states (solver.py:287-291): rows of a and b
 * (elem_bytes * row_elems each) are exchanged where exchange[s] != 0. */
int sgmc_swap_rows(void* stream, void* a, void* b, const int32_t* exchange,
                   int64_t n_rows, int64_t row_bytes);

/* reSGLD with a ladder of n_replicas temperatures sharded over GPUs
 * (BASELINE.json configs[3]); extends solver.py:273-291 from 2 to R replicas
 * by exchanging temperature LABELS instead of chain states, so that only the
 * all-gathered (U, var) scalars cross NVLink and every rank replays the same
 * decisions from the same keys.
 *   gathered f32[R][2][B]: potential row and variance row of replica r
 *            (the all-gather of every rank's local values, rank-major);
 *   holder   int32[R][B]: replica running system b at temperature index t
 *            (in/out, replicated on every rank);
 *   ssq f32[R-1][B] (in/out), F f32[B], temps f32[R], keys uint32[R-1][B][2];
 *   exchange int32[R-1][B] (out): pair (t, t+1) of system b swapped labels;
 *   temp_per_chain f32[n_local][B], temp_index int32[n_local][B] (out): the
 *            temperature each local replica's systems run at next step.
 * Pair p is attempted every step when R == 2 (the reference), else iff
 * p % 2 == step % 2.  `step` is the already incremented counter (>= 1). */
int sgmc_resgld_ladder_step(void* stream, const float* gathered, int32_t* holder,
                            float* ssq, const float* F, const float* temps,
                            const uint32_t* keys_in, uint32_t* keys_out,
                            int32_t* exchange, int n_replicas, int64_t n_systems,
                            int64_t step, int first_local_replica,
                            int n_local_replicas, float* temp_per_chain,
                            int32_t* temp_index, int prng_layout);

/* ---- NCCL over NVLink (multi-GPU exchange steps).  libnccl.so.2 is resolved
 *      with dlopen at first use; no link-time dependency.  Used by the reSGLD
 *      replica exchange (all-gather of per-replica (U, var), solver.py:273-291
 *      distributed over GPUs) and by the minibatch-sharded gradient all-reduce.
 *      unique_id is the 128-byte ncclUniqueId, created on rank 0 and
 *      distributed by the host (any out-of-band channel). */
int sgmc_nccl_available(void);
int sgmc_nccl_unique_id(void* unique_id_128);
int sgmc_nccl_init(void** comm, const void* unique_id_128, int n_ranks, int rank);
int sgmc_nccl_destroy(void* comm);
int sgmc_nccl_allgather(void* comm, void* stream, const void* send, void* recv,
                        size_t bytes_per_rank);
int sgmc_nccl_allreduce_sum_f32(void* comm, void* stream, const float* send,
                                float* recv, size_t count);

#ifdef __cplusplus
}
#endif
#endif  /* SGMC_B200_H_ */
