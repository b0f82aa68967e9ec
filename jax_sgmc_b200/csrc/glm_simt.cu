// GLM stochastic potential + gradient, fp32 SIMT path (path 0) and the shared
// "finalize" reduction used by every path.
//
// Replaces, for the recognised GLM families, potential.minibatch_potential
// (jax_sgmc/potential.py:94-216) together with the reverse-mode gradient the
// integrators take of it (jax_sgmc/integrator.py:166, :593, :792), batched over
// chains that share one minibatch:
//   pass 1  Z = Theta_w . X_b^T  -> per-observation ell and residual
//           R = cot * d ell/dz, cot = (-N/n)/T (* mask)     [GEMM-shaped, NT]
//   final   U = (-N mean(ell) - prior)/T, var(ell), aux-parameter gradients
//   pass 2  G = R . X_b - grad(prior)/T                     [GEMM-shaped, NN]
// The minibatch gather (data/core.py:642-660) is fused: rows are read through
// idx.  This path is the precise fp32 fallback for arbitrary shapes (e.g. the
// quickstart's d=4); the tensor-core path lives in glm_tc.cu.
#include "glm.cuh"

namespace sgmc {

constexpr int TM = 64, TN = 64, TK = 16;

template <bool NT, class Epi>
__global__ void __launch_bounds__(256)
k_sgemm(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
        int64_t ldb, const int32_t* __restrict__ bidx, int M, int N, int K,
        const Epi epi, int64_t a_zstride, int64_t bidx_zstride) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  // blockIdx.z = chain when every chain has its own minibatch (M = 1 per slice)
  const int z = blockIdx.z;
  A += (int64_t)z * a_zstride;
  if (bidx) bidx += (int64_t)z * bidx_zstride;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += TK) {
    {
      const int r = tid >> 2, kq = (tid & 3) * 4, m = m0 + r;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + kq + q;
        As[kq + q][r] = (m < M && k < K) ? A[(int64_t)m * lda + k] : 0.f;
      }
    }
    if (NT) {
      const int r = tid >> 2, kq = (tid & 3) * 4, nn = n0 + r;
      const int64_t row = nn < N ? (bidx ? (int64_t)bidx[nn] : nn) : 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + kq + q;
        Bs[kq + q][r] = (nn < N && k < K) ? B[row * ldb + k] : 0.f;
      }
    } else {
      const int kr = tid >> 4, nq = (tid & 15) * 4, k = k0 + kr;
      const int64_t row = k < K ? (bidx ? (int64_t)bidx[k] : k) : 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int nn = n0 + nq + q;
        Bs[kr][nq + q] = (k < K && nn < N) ? B[row * ldb + nn] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tx * 4 + j;
      if (nn < N) epi(m + z, nn, acc[i][j]);
    }
  }
}

// pass 1 epilogue: z -> (ell, R)
struct LinkEpi {
  GlmArgs a;
  __device__ void operator()(int c, int i, float z) const {
    const float* th = a.theta + (int64_t)c * a.P;
    GaussConst gc{1.f, 0.f};
    if (a.spec.family == kFamilyGaussian) gc = gauss_const(th[a.spec.aux_off]);
    else if (a.spec.aux_off >= 0) z += th[a.spec.aux_off];
    const float y = a.y[a.idx ? a.idx[(int64_t)c * a.idx_stride + i] : i];
    float ell, dz;
    glm_link(a.spec.family, z, y, gc, ell, dz);
    float cot = a.cot;
    if (a.mask) cot *= a.mask[i];
    a.R[(int64_t)c * a.n + i] = dz * cot;
    a.ell[(int64_t)c * a.n + i] = ell;
  }
};

// pass 2 epilogue: G -> grad (adds -grad(prior)/T)
struct GradEpi {
  GlmArgs a;
  __device__ void operator()(int c, int j, float g) const {
    const int64_t p = (int64_t)c * a.P + a.spec.w_off + j;
    a.grad[p] = g + prior_grad_term(a, c, a.spec.w_off + j);
  }
};

// One CTA per chain: U, var(ell), aux-parameter gradient (log_sigma / bias).
__global__ void __launch_bounds__(256) k_glm_finalize(const GlmArgs a) {
  __shared__ float red[2][8];
  const int c = blockIdx.x;
  const float* ell = a.ell + (int64_t)c * a.n;
  const float* R = a.R + (int64_t)c * a.n;
  const float* th = a.theta + (int64_t)c * a.P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  auto block_sum2 = [&](float x, float y2, float& ox, float& oy) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      x += __shfl_xor_sync(0xffffffffu, x, s);
      y2 += __shfl_xor_sync(0xffffffffu, y2, s);
    }
    __syncthreads();
    if (lane == 0) { red[0][warp] = x; red[1][warp] = y2; }
    __syncthreads();
    ox = 0.f; oy = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { ox += red[0][w]; oy += red[1][w]; }
  };

  // sum(ell) [masked sum when mask given], aux gradient
  GaussConst gc{1.f, 0.f};
  const bool gauss = a.spec.family == kFamilyGaussian;
  if (gauss) gc = gauss_const(th[a.spec.aux_off]);
  float s_ell = 0.f, s_aux = 0.f;
  for (int i = tid; i < a.n; i += 256) {
    const float l = ell[i];
    const float m = a.mask ? a.mask[i] : 1.0f;
    s_ell += l * m;
    if (gauss) {
      // d ell / d log_sigma = r^2/s2 - 1 = (-2 ell - ln) - 1
      s_aux += (a.cot * m) * ((-2.0f * l - gc.ln) - 1.0f);
    } else if (a.spec.aux_off >= 0) {
      s_aux += R[i];
    }
  }
  float sum_ell, sum_aux;
  block_sum2(s_ell, s_aux, sum_ell, sum_aux);
  const float mean = sum_ell / (float)a.n;

  // population variance of ell (integrator.py:880: jnp.var), two-pass
  float s_var = 0.f, s_prior = 0.f;
  {
    float plain = 0.f;
    if (a.mask) {   // variance is over the unmasked likelihoods
      for (int i = tid; i < a.n; i += 256) plain += ell[i];
    }
    float psum, dummy;
    if (a.mask) block_sum2(plain, 0.f, psum, dummy); else psum = sum_ell;
    const float mu = psum / (float)a.n;
    for (int i = tid; i < a.n; i += 256) {
      const float dv = ell[i] - mu;
      s_var += dv * dv;
    }
    if (a.spec.prior == kPriorGaussian) {
      for (int p = a.spec.prior_off + tid; p < a.spec.prior_off + a.spec.prior_size;
           p += 256)
        s_prior += th[p] * th[p];
    }
  }
  float sum_var, sum_prior;
  block_sum2(s_var, s_prior, sum_var, sum_prior);

  if (tid == 0) {
    float L;
    if (a.mask) L = (-(float)a.N / (float)a.n) * sum_ell;      // potential.py:185
    else L = -(float)a.N * mean;                               // potential.py:183
    float prior = 0.f;
    if (a.spec.prior == kPriorGaussian)
      prior = -0.5f * (1.0f / (a.spec.prior_scale * a.spec.prior_scale)) * sum_prior;
    else if (a.spec.prior == kPriorInvSigma)
      prior = 1.0f / expf(th[a.spec.prior_off]);
    a.potential[c] = (L - prior) / a.spec.temperature;         // potential.py:210
    if (a.variance) a.variance[c] = sum_var / (float)a.n;
    if (a.spec.aux_off >= 0 && a.grad)
      a.grad[(int64_t)c * a.P + a.spec.aux_off] =
          sum_aux + prior_grad_term(a, c, a.spec.aux_off);
  }
}

// full-potential helpers: wrapped index / mask vector of the last batch
// (data/core.py:571-572) and the running sum over batches
__global__ void k_wrap_batch(int32_t* __restrict__ idx, float* __restrict__ mask, int64_t first,
                             int64_t n, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = first + i;
  idx[i] = (int32_t)(id % N);
  mask[i] = id < N ? 1.0f : 0.0f;
}
__global__ void k_fill(float* __restrict__ x, float v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}
__global__ void k_axpy_acc(float* __restrict__ total, float a, const float* __restrict__ u,
                           int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) total[i] = __fadd_rn(__fmul_rn(1.0f, total[i]), __fmul_rn(a, u[i]));
}
__global__ void k_full_finish(float* __restrict__ out, const float* __restrict__ total,
                              const float* __restrict__ neg_prior, float inv_T, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fadd_rn(__fmul_rn(inv_T, total[i]), __fmul_rn(inv_T, neg_prior[i]));
}

int glm_finalize(cudaStream_t stream, const GlmArgs& a) {
  k_glm_finalize<<<(unsigned)a.C, 256, 0, stream>>>(a);
  return post_launch("k_glm_finalize");
}

int glm_simt(cudaStream_t stream, const GlmArgs& a) {
  const int C = (int)a.C, n = (int)a.n, d = a.spec.d;
  // one minibatch per chain: a [1 x n] / [1 x d] slice per chain along grid.z
  const bool per_chain = a.idx_stride != 0;
  const int M = per_chain ? 1 : C;
  const unsigned gz = per_chain ? (unsigned)C : 1u;
  SGMC_REQUIRE(gz <= 65535u, "per-chain minibatches: at most 65535 chains per call");
  {
    dim3 grid((n + TN - 1) / TN, (M + TM - 1) / TM, gz);
    k_sgemm<true, LinkEpi><<<grid, 256, 0, stream>>>(
        a.theta + a.spec.w_off, a.P, a.X, d, a.idx, M, n, d, LinkEpi{a},
        per_chain ? a.P : 0, a.idx_stride);
    if (post_launch("k_sgemm<link>")) return 1;
  }
  if (glm_finalize(stream, a)) return 1;
  if (a.grad) {
    dim3 grid((d + TN - 1) / TN, (M + TM - 1) / TM, gz);
    k_sgemm<false, GradEpi><<<grid, 256, 0, stream>>>(
        a.R, a.n, a.X, d, a.idx, M, d, n, GradEpi{a}, per_chain ? a.n : 0, a.idx_stride);
    if (post_launch("k_sgemm<grad>")) return 1;
  }
  return 0;
}

}  // namespace sgmc

using namespace sgmc;

namespace sgmc {
int glm_tc_debug_read(unsigned long long* out);
int glm_pair_timeline_read(unsigned long long* out, int n);
}

extern "C" {

int sgmc_debug_tc_timers(unsigned long long* out8) { return sgmc::glm_tc_debug_read(out8); }
int sgmc_debug_pair_timeline(unsigned long long* out, int n) {
  return sgmc::glm_pair_timeline_read(out, n);
}

size_t sgmc_glm_workspace_bytes(int64_t n_chains, int64_t batch_size, int64_t d,
                                int path) {
  size_t base = (size_t)n_chains * batch_size * sizeof(float) * 2;  // R + ell
  if (path != 0) base += glm_tc_workspace_bytes(n_chains, batch_size, d, path);
  return base + 256;
}

static int glm_dispatch(void* stream, const sgmc_glm_spec* spec, const float* theta,
                        int64_t n_chains, int64_t P, const float* X, const float* y,
                        const int32_t* idx, const float* mask, int64_t batch_size,
                        int64_t observation_count, float* potential, float* variance,
                        float* grad, float* ell, void* workspace, size_t workspace_bytes,
                        int path, const FusedSgld& fused, int64_t idx_stride = 0,
                        CarryCtx* carry = nullptr, XStage xs = XStage{}) {
  SGMC_REQUIRE(spec && X && (xs.only_x || (theta && y && potential)), "null argument");
  SGMC_REQUIRE(spec->family == kFamilyGaussian || spec->family == kFamilyLogistic,
               "unknown GLM family %d", spec->family);
  SGMC_REQUIRE(spec->family != kFamilyGaussian || spec->aux_off >= 0,
               "gaussian family needs aux_off (log_sigma)");
  SGMC_REQUIRE(spec->d > 0 && spec->w_off >= 0 && spec->w_off + spec->d <= P,
               "weight range outside the sample");
  SGMC_REQUIRE(P == spec->d + (spec->aux_off >= 0 ? 1 : 0),
               "sample size %lld != d + aux", (long long)P);
  SGMC_REQUIRE(n_chains > 0 && batch_size > 0 && n_chains < (1ll << 31) &&
               batch_size < (1ll << 31), "bad sizes");
  SGMC_REQUIRE(workspace_bytes >=
               sgmc_glm_workspace_bytes(n_chains, batch_size, spec->d, path),
               "workspace too small");
  SGMC_REQUIRE(path >= 0 && path <= 2, "unknown path %d", path);
  GlmArgs a;
  a.spec = *spec;
  a.theta = theta; a.C = n_chains; a.P = P;
  a.X = X; a.y = y; a.idx = idx; a.mask = mask;
  a.idx_stride = idx_stride;
  a.n = batch_size; a.N = observation_count;
  a.potential = potential; a.variance = variance; a.grad = grad;
  float* ws = reinterpret_cast<float*>(
      (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  a.R = ws;
  a.ell = ell ? ell : ws + (size_t)n_chains * batch_size;
  a.ell_requested = ell != nullptr;
  a.tc_ws = ws + 2 * (size_t)n_chains * batch_size;
  a.fused = fused;
  a.carry = carry;
  a.x_slot = xs.slot; a.x_prepared = xs.prepared; a.only_x = xs.only_x;
  SGMC_REQUIRE(path != 0 || (!xs.prepared && !xs.only_x), "staged minibatches need a tensor-core path");
  // cotangent of every ell_i: (1/T) * (-N) / n    (potential.py:183,210)
  a.cot = (-(float)observation_count / (float)batch_size) / spec->temperature;
  if (path == 0) return glm_simt((cudaStream_t)stream, a);
  return glm_tc((cudaStream_t)stream, a, path);
}

int sgmc_glm_potential_grad(void* stream, const sgmc_glm_spec* spec,
                            const float* theta, int64_t n_chains, int64_t P,
                            const float* X, const float* y, const int32_t* idx,
                            const float* mask, int64_t batch_size,
                            int64_t observation_count, float* potential,
                            float* variance, float* grad, float* ell,
                            void* workspace, size_t workspace_bytes, int path) {
  FusedSgld none{};
  return glm_dispatch(stream, spec, theta, n_chains, P, X, y, idx, mask, batch_size,
                      observation_count, potential, variance, grad, ell, workspace,
                      workspace_bytes, path, none);
}

int sgmc_glm_potential_grad_per_chain(void* stream, const sgmc_glm_spec* spec,
                                      const float* theta, int64_t n_chains, int64_t P,
                                      const float* X, const float* y, const int32_t* idx,
                                      const float* mask, int64_t batch_size,
                                      int64_t observation_count, float* potential,
                                      float* variance, float* grad, float* ell,
                                      void* workspace, size_t workspace_bytes) {
  SGMC_REQUIRE(idx != nullptr, "per-chain minibatches need idx int32[C][n]");
  FusedSgld none{};
  return glm_dispatch(stream, spec, theta, n_chains, P, X, y, idx, mask, batch_size,
                      observation_count, potential, variance, grad, ell, workspace,
                      workspace_bytes, 0, none, batch_size);
}

int sgmc_glm_full_potential(void* stream, const sgmc_glm_spec* spec, const float* theta,
                            int64_t n_chains, int64_t P, const float* X, const float* y,
                            int64_t observation_count, int64_t batch_size, float* potential,
                            float* scratch, int32_t* wrap_idx, float* wrap_mask,
                            void* workspace, size_t workspace_bytes, int path) {
  SGMC_REQUIRE(spec && potential && scratch && wrap_idx && wrap_mask, "null argument");
  SGMC_REQUIRE(batch_size > 0 && observation_count > 0, "bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t C = n_chains, n = batch_size, N = observation_count;
  const int d = spec->d;
  float* total = scratch;            // f32[C]
  float* U_b = scratch + C;          // f32[C]
  const unsigned gc = (unsigned)((C + 255) / 256);
  FusedSgld none{};
  // inner potential: zero prior, T = 1 (potential.py:258-262)
  sgmc_glm_spec inner = *spec;
  inner.prior = kPriorFlat;
  inner.temperature = 1.0f;
  k_fill<<<gc, 256, 0, s>>>(total, 0.0f, C);
  if (post_launch("k_fill")) return 1;
  // the reference maps with masking=True: every batch goes through the masked
  // form -N/n * dot(ell, mask) (potential.py:185), whole batches with mask = 1
  float* ones = wrap_mask;           // f32[n]
  float* last_mask = wrap_mask + n;  // f32[n]
  k_fill<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ones, 1.0f, n);
  if (post_launch("k_fill")) return 1;
  const float unscale = (float)((double)n / (double)N);   // undo N/n (potential.py:264-271)
  const int64_t n_batches = (N + n - 1) / n;
  for (int64_t b = 0; b < n_batches; ++b) {
    const bool whole = (b + 1) * n <= N;
    if (!whole) {                     // last batch: wrap modulo N, mask the overhang
      k_wrap_batch<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(wrap_idx, last_mask, b * n, n, N);
      if (post_launch("k_wrap_batch")) return 1;
    }
    if (int e = glm_dispatch(stream, &inner, theta, C, P, whole ? X + b * n * d : X,
                             whole ? y + b * n : y, whole ? nullptr : wrap_idx,
                             whole ? ones : last_mask, n, N, U_b, nullptr, nullptr, nullptr,
                             workspace, workspace_bytes, path, none))
      return e;
    k_axpy_acc<<<gc, 256, 0, s>>>(total, unscale, U_b, C);
    if (post_launch("k_axpy_acc")) return 1;
  }
  // prior: the potential of an all-masked batch is -prior (T = 1)
  sgmc_glm_spec pr = *spec;
  pr.temperature = 1.0f;
  k_fill<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(last_mask, 0.0f, n);
  if (post_launch("k_fill")) return 1;
  if (int e = glm_dispatch(stream, &pr, theta, C, P, X, y, nullptr, last_mask, n, N, U_b, nullptr,
                           nullptr, nullptr, workspace, workspace_bytes, path, none))
    return e;
  k_full_finish<<<gc, 256, 0, s>>>(potential, total, U_b,
                                   (float)(1.0 / (double)spec->temperature), C);
  return post_launch("k_full_finish");
}

// ---- minibatch rows sharded over ranks (BASELINE.json configs[4], north star (3)) -------
// Rank r evaluates rows [r n/R, (r+1) n/R) of the shared minibatch for ALL chains; the
// partial gradients and likelihood sums are all-reduced over NVLink.  The split is exact:
// with N_r = N / R the local cotangent (-N_r / n_r) / T equals the global one, rank 0
// carries the prior and the other ranks a flat one, so sum_r U_r = U and sum_r g_r = g.
// var(ell) is rebuilt from e1 = sum ell and e2 = sum ell^2 (accumulated around the local
// mean): var = (E2 - E1^2 / n) / n.
__global__ void __launch_bounds__(256) k_row_shard_stats(const float* __restrict__ ell, int n_r,
                                                         const float* __restrict__ U_r,
                                                         float* __restrict__ extras) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[256];
  const int64_t c = blockIdx.x;
  const float* row = ell + c * n_r;
  float s = 0.f;
  for (int i = threadIdx.x; i < n_r; i += 256) s += row[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  const float e1 = red[0];
  const float mean = e1 / (float)n_r;
  __syncthreads();
  float q = 0.f;
  for (int i = threadIdx.x; i < n_r; i += 256) {
    const float dlt = row[i] - mean;
    q = fmaf(dlt, dlt, q);
  }
  red[threadIdx.x] = q;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    extras[c * 3 + 0] = U_r[c];
    extras[c * 3 + 1] = e1;
    extras[c * 3 + 2] = fmaf((float)n_r * mean, mean, red[0]);
  }
}

__global__ void k_row_shard_finalize(const float* __restrict__ extras, float n,
                                     float* __restrict__ potential, float* __restrict__ variance,
                                     int64_t C) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float U = extras[c * 3], e1 = extras[c * 3 + 1], e2 = extras[c * 3 + 2];
  potential[c] = U;
  if (variance) variance[c] = fmaxf(e2 - e1 * e1 / n, 0.f) / n;
}

int sgmc_glm_row_shard_finalize(void* stream, const float* extras, int64_t batch_size,
                                float* potential, float* variance, int64_t n_chains) {
  SGMC_REQUIRE(extras && potential && n_chains > 0 && batch_size > 0, "bad arguments");
  launch_pdl(k_row_shard_finalize, dim3((unsigned)((n_chains + 255) / 256)), dim3(256), 0,
             (cudaStream_t)stream, extras, (float)batch_size, potential, variance, n_chains);
  return post_launch("k_row_shard_finalize");
}

int sgmc_glm_potential_grad_row_sharded(void* stream, const sgmc_glm_spec* spec,
                                        const float* theta, int64_t n_chains, int64_t P,
                                        const float* X, const float* y, const int32_t* idx,
                                        int64_t batch_size, int64_t observation_count,
                                        float* potential, float* variance, float* grad,
                                        void* workspace, size_t workspace_bytes, int path,
                                        void* nccl_comm, int rank, int n_ranks, float* scratch) {
  SGMC_REQUIRE(spec && theta && X && y && potential && grad && scratch, "null argument");
  SGMC_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad rank");
  SGMC_REQUIRE(batch_size % n_ranks == 0 && observation_count % n_ranks == 0,
               "batch size and observation count must be multiples of the rank count");
  const int64_t C = n_chains, n_r = batch_size / n_ranks, row0 = (int64_t)rank * n_r;
  float* ell = scratch;                       // f32[C][n_r]
  float* extras = scratch + (size_t)C * n_r;  // f32[C][3]: U_r, sum ell, sum ell^2
  sgmc_glm_spec local = *spec;
  if (rank != 0) { local.prior = kPriorFlat; local.prior_size = 0; }
  FusedSgld none{};
  if (int e = glm_dispatch(stream, &local, theta, C, P, idx ? X : X + row0 * spec->d,
                           idx ? y : y + row0, idx ? idx + row0 : nullptr, nullptr, n_r,
                           observation_count / n_ranks, extras + (size_t)3 * C, nullptr, grad, ell,
                           workspace, workspace_bytes, path, none))
    return e;
  // (extras + 3C is a C-float slot behind the table: the local potential lands there first)
  launch_pdl(k_row_shard_stats, dim3((unsigned)C), dim3(256), 0, (cudaStream_t)stream,
             (const float*)ell, (int)n_r, (const float*)(extras + (size_t)3 * C), extras);
  if (post_launch("k_row_shard_stats")) return 1;
  if (nccl_comm == nullptr) return 0;         // partials only (single-process emulation)
  if (int e = sgmc_nccl_allreduce_sum_f32(nccl_comm, stream, grad, grad, (size_t)C * P)) return e;
  if (int e = sgmc_nccl_allreduce_sum_f32(nccl_comm, stream, extras, extras, (size_t)3 * C))
    return e;
  return sgmc_glm_row_shard_finalize(stream, extras, batch_size, potential, variance, C);
}

int sgmc_glm_prepare_minibatch(void* stream, const sgmc_glm_spec* spec, int64_t n_chains,
                               const float* X, const int32_t* idx, int64_t batch_size,
                               void* workspace, size_t workspace_bytes, int path, int slot) {
  SGMC_REQUIRE(path == 1 || path == 2, "staging is a tensor-core path operation");
  SGMC_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
  FusedSgld none{};
  XStage xs;
  xs.slot = slot;
  xs.only_x = true;
  const int64_t P = spec ? spec->d : 0;
  return glm_dispatch(stream, spec, nullptr, n_chains, P, X, nullptr, idx, nullptr, batch_size, 1,
                      nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes, path, none, 0,
                      nullptr, xs);
}

int sgmc_glm_sgld_step(void* stream, const sgmc_glm_spec* spec, float* theta, float* v,
                       int64_t n_chains, int64_t P, const float* X, const float* y,
                       const int32_t* idx, const float* mask, int64_t batch_size,
                       int64_t observation_count, float* potential, float* variance,
                       float* grad, const uint32_t* keys_in, uint32_t* keys_out,
                       float step_size, float temperature, float alpha, float lmbd,
                       void* workspace, size_t workspace_bytes, int path, int prng_layout,
                       int write_grad, const float* temp_per_chain, void* wait_event,
                       const int64_t* leaf_sizes, int n_leaves, int flags) {
  SGMC_REQUIRE(grad && keys_in && keys_out, "null argument");
  const int64_t whole = P;
  if (leaf_sizes == nullptr) {       // the sample is one leaf
    leaf_sizes = &whole;
    n_leaves = 1;
  }
  SGMC_REQUIRE(keys_in != keys_out, "keys_out must not alias keys_in");
  bool applied = false;
  FusedSgld fu{};
  // per-chain temperatures / an event to wait for before the update: only the
  // plain kernel sequence supports them
  fu.requested = temp_per_chain == nullptr && wait_event == nullptr && n_leaves == 1;
  fu.theta_rw = theta; fu.v = v; fu.keys_in = keys_in; fu.keys_out = keys_out;
  fu.step_size = step_size; fu.temperature = temperature; fu.alpha = alpha; fu.lmbd = lmbd;
  fu.layout = prng_layout; fu.applied = &applied; fu.write_grad = write_grad != 0;
  prof_mark((cudaStream_t)stream, 0);
  XStage xs;
  if ((flags & SGMC_STEP_X_STAGED) && path != 0) {
    xs.prepared = true;
    xs.slot = (flags & SGMC_STEP_X_SLOT1) ? 1 : 0;
  }
  CarryCtx cc{};
  if ((flags & (SGMC_STEP_CARRY_INIT | SGMC_STEP_CARRY)) && path != 0 && n_leaves == 1) {
    cc.mode = (flags & SGMC_STEP_CARRY_INIT) ? 1 : 2;
    cc.keys_in = keys_in; cc.keys_out = keys_out; cc.prng_layout = prng_layout;
  }
  if (int e = glm_dispatch(stream, spec, theta, n_chains, P, X, y, idx, mask, batch_size,
                           observation_count, potential, variance, grad, nullptr, workspace,
                           workspace_bytes, path, fu, 0, cc.mode ? &cc : nullptr, xs))
    return e;
  if (applied) return 0;
  if (wait_event != nullptr &&
      check_cuda(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)wait_event, 0),
                 "cudaStreamWaitEvent"))
    return 1;
  prof_mark((cudaStream_t)stream, 2);
  if (cc.active && write_grad) cc.out.grad_rw = grad;   // hand the completed gradient out
  if (cc.active) {  // the update also writes Theta's operand form for the next step
    const int e = sgld_update_split((cudaStream_t)stream, theta, v, grad, keys_in, keys_out,
                                    n_chains, P, step_size, temperature, temp_per_chain, alpha,
                                    lmbd, prng_layout, cc.out);
    prof_mark((cudaStream_t)stream, 3);
    return e;
  }
  // the stand-alone fused noise + update kernel
  if (v)
    return sgmc_sgld_rms_update(stream, theta, v, grad, keys_in, keys_out, n_chains, leaf_sizes,
                                n_leaves, step_size, temperature, temp_per_chain, alpha, lmbd,
                                prng_layout);
  return sgmc_sgld_update(stream, theta, grad, keys_in, keys_out, n_chains, leaf_sizes, n_leaves,
                          step_size, temperature, temp_per_chain, prng_layout);
}

// K Langevin steps over minibatches that live in HOST memory: the inner loop of
// solver.mcmc (solver.py:152-160) with the reference's host data cache
// (data/core.py:664-791) in native code.  Two sources:
//   staged  (sgmc_glm_sgld_scan_host): step k reads host batch k mod host_batch_count =
//           host_batches + that * stride floats of page-locked memory, laid out [n][d]
//           rows followed by [n] labels (gathered by sgmc_host_gather_batches), copied by
//           the DMA engine;
//   pulled  (sgmc_glm_sgld_scan_pull): the data set itself is mapped (sgmc_host_register)
//           and a kernel reads the rows idx_all[k][:] of step k over the host link
//           (sgmc_pull_rows) -- no host gather, no staging.
// A ring of n_slots device buffers is filled on a transfer stream n_slots - 1 batches
// ahead of the sampling stream, the operand staging of batch k + 1 runs on the copy
// stream while step k samples (events order reuse); every step's (U, var) rows are read
// back to host_results[k] on a third stream.  Nothing synchronises: the caller waits on
// the two streams it passed when it needs the results.
namespace {
struct HostSource {
  const float* host_batches = nullptr;   // staged
  int64_t host_batch_count = 0;
  const float* X_mapped = nullptr;       // pulled
  const float* y_mapped = nullptr;
  const int32_t* idx_all = nullptr;      // device int32[n_steps][n]
  int pull_ctas = 0;
  // hybrid: the first rows_dma rows of a rank's slice come from host_batches (DMA engine),
  // the rest is pulled (both sources set); -1 = all rows from the one source that is set
  int64_t rows_dma = -1;
};

int scan_host_impl(void* stream, void* copy_stream, const sgmc_glm_spec* spec,
                   float* theta, float* v, int64_t n_chains, int64_t P, const HostSource& src,
                   int64_t n_steps, int64_t batch_size,
                   int64_t observation_count, float* device_slots, int n_slots,
                   float* potential_variance, float* host_results, float* grad,
                   uint32_t* keys_a, uint32_t* keys_b, const float* step_sizes,
                   float temperature, float alpha, float lmbd, void* workspace,
                   size_t workspace_bytes, int path, int prng_layout,
                   void* nccl_comm, int rank, int n_ranks, const uint8_t* keep,
                   float* samples_out, float* scalars_out, int64_t capacity,
                   int64_t* kept) {
  const bool hybrid = src.X_mapped != nullptr && src.host_batches != nullptr;
  const bool pull = src.X_mapped != nullptr && !hybrid;
  SGMC_REQUIRE(spec && theta && device_slots && potential_variance && grad &&
               keys_a && keys_b && step_sizes, "null argument");
  SGMC_REQUIRE(src.X_mapped ? (src.idx_all && (hybrid || src.y_mapped)) : src.host_batches != nullptr,
               "no minibatch source");
  SGMC_REQUIRE(src.host_batches == nullptr || src.host_batch_count >= 1, "no host batches");
  SGMC_REQUIRE(keep == nullptr || samples_out == nullptr || (scalars_out && kept),
               "sample collection needs scalars_out and kept");
  SGMC_REQUIRE(n_slots >= 2 && n_slots <= 8 && n_steps >= 0, "2..8 slots");
  const bool sharded = nccl_comm != nullptr && n_ranks > 1;
  SGMC_REQUIRE(!sharded || (batch_size % n_ranks == 0 && rank >= 0 && rank < n_ranks),
               "sharded upload: batch_size must be a multiple of the rank count");
  cudaStream_t ms = (cudaStream_t)stream, cs = (cudaStream_t)copy_stream;
  const int64_t n = batch_size, d = spec->d, C = n_chains;
  const size_t stride = (size_t)n * d + n;                 // floats per device batch
  // Every rank of a chain-sharded job consumes the SAME minibatch.  Sharded upload: a
  // rank moves only its n / R rows (then all n labels) over its host link; the rows are
  // all-gathered over NVLink into the device slot, so the host link carries 1 / R of
  // the batch per rank instead of R identical copies.
  const int64_t rows_local = sharded ? n / n_ranks : n;
  const int64_t rows_dma = hybrid ? src.rows_dma : rows_local;      // rows of the slice staged on the host
  SGMC_REQUIRE(rows_dma >= 0 && rows_dma <= rows_local, "bad DMA / pull split");
  const size_t host_stride = (size_t)rows_dma * d + n;
  cudaStream_t rs = nullptr;                               // read-back stream (D2H runs
  cudaStream_t hs = nullptr;                               // beside H2D); transfer stream
  cudaStream_t ps = nullptr;                               // hybrid: the pull beside the DMA
  if (check_cuda(cudaStreamCreateWithFlags(&rs, cudaStreamNonBlocking), "stream") ||
      check_cuda(cudaStreamCreateWithFlags(&hs, cudaStreamNonBlocking), "stream") ||
      (hybrid && check_cuda(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking), "stream")))
    return 1;
  cudaEvent_t copied[8], consumed[8], computed[2], read_back[2], begin, pulled[8];
  if (hybrid)
    for (int i = 0; i < n_slots; ++i)
      if (check_cuda(cudaEventCreateWithFlags(&pulled[i], cudaEventDisableTiming), "event")) return 1;
  if (check_cuda(cudaEventCreateWithFlags(&begin, cudaEventDisableTiming), "event")) return 1;
  cudaEventRecord(begin, cs);                              // what the caller queued on the copy
  cudaStreamWaitEvent(hs, begin, 0);                       // stream (index upload) comes first
  if (hybrid) cudaStreamWaitEvent(ps, begin, 0);
  for (int i = 0; i < n_slots; ++i) {
    if (check_cuda(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming), "event") ||
        check_cuda(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming), "event"))
      return 1;
    cudaEventRecord(consumed[i], ms);                      // all slots start free
  }
  for (int i = 0; i < 2; ++i) {
    if (check_cuda(cudaEventCreateWithFlags(&computed[i], cudaEventDisableTiming), "event") ||
        check_cuda(cudaEventCreateWithFlags(&read_back[i], cudaEventDisableTiming), "event"))
      return 1;
    cudaEventRecord(read_back[i], rs);
  }
  int rc = 0;
  auto prefetch = [&](int64_t k) {
    const int sl = (int)(k % n_slots);
    float* dst = device_slots + sl * stride;
    cudaStreamWaitEvent(hs, consumed[sl], 0);
    if (pull) {
      if (rc == 0)
        rc = sgmc_pull_rows(hs, src.X_mapped, src.y_mapped, src.idx_all + k * n, n,
                            (int64_t)rank * (sharded ? rows_local : 0), rows_local, d, dst,
                            dst + (size_t)n * d, src.pull_ctas);
    } else {
      const float* hb = src.host_batches + (k % src.host_batch_count) * host_stride;
      const int64_t row0 = (int64_t)rank * (sharded ? rows_local : 0);
      if (hybrid && rows_dma < rows_local) {
        // the tail of the slice is pulled by the GPU while the DMA engine copies its head
        cudaStreamWaitEvent(ps, consumed[sl], 0);
        if (rc == 0)
          rc = sgmc_pull_rows(ps, src.X_mapped, nullptr, src.idx_all + k * n, n, row0 + rows_dma,
                              rows_local - rows_dma, d, dst, nullptr, src.pull_ctas);
        cudaEventRecord(pulled[sl], ps);
      }
      if (rows_dma > 0)
        cudaMemcpyAsync(dst + (size_t)row0 * d, hb, (size_t)rows_dma * d * 4,
                        cudaMemcpyHostToDevice, hs);
      cudaMemcpyAsync(dst + (size_t)n * d, hb + (size_t)rows_dma * d, (size_t)n * 4,
                      cudaMemcpyHostToDevice, hs);
      if (hybrid && rows_dma < rows_local) cudaStreamWaitEvent(hs, pulled[sl], 0);
    }
    if (sharded && rc == 0)
      rc = sgmc_nccl_allgather(nccl_comm, hs, dst + (size_t)rank * rows_local * d, dst,
                               (size_t)rows_local * d * 4);
    cudaEventRecord(copied[sl], hs);
  };
  // Tensor-core paths: the operand staging of batch k+1 runs on the copy stream (as soon
  // as its rows have arrived) while step k samples; two copies of the operands.
  const bool piped = path != 0 && n_steps > 1 && !option(SGMC_OPT_NO_PIPELINE);
  cudaEvent_t staged[2] = {nullptr, nullptr}, xfree[2] = {nullptr, nullptr};
  if (piped)
    for (int i = 0; i < 2; ++i)
      if (check_cuda(cudaEventCreateWithFlags(&staged[i], cudaEventDisableTiming), "event") ||
          check_cuda(cudaEventCreateWithFlags(&xfree[i], cudaEventDisableTiming), "event"))
        return 1;
  auto stage = [&](int64_t k) -> int {
    const int sl = (int)(k % n_slots), xsl = (int)(k & 1);
    cudaStreamWaitEvent(cs, copied[sl], 0);                  // the rows of batch k are there
    if (k >= 2) cudaStreamWaitEvent(cs, xfree[xsl], 0);      // step k-2 is done with this copy
    if (int e = sgmc_glm_prepare_minibatch(cs, spec, C, device_slots + sl * stride, nullptr, n,
                                           workspace, workspace_bytes, path, xsl))
      return e;
    return check_cuda(cudaEventRecord(staged[xsl], cs), "event record");
  };
  for (int64_t k = 0; k < n_steps && k < n_slots - 1; ++k) prefetch(k);
  if (piped && rc == 0 && n_steps > 0) rc = stage(0);
  for (int64_t k = 0; k < n_steps && rc == 0; ++k) {
    if (k + n_slots - 1 < n_steps) prefetch(k + n_slots - 1);
    if (piped && k + 1 < n_steps && rc == 0 && (rc = stage(k + 1)) != 0) break;
    if (rc) break;
    const int sl = (int)(k % n_slots);
    float* Xb = device_slots + sl * stride;
    float* uv = potential_variance + (k & 1) * 2 * C;      // (U, var) double-buffered
    int flags = k == 0 ? SGMC_STEP_CARRY_INIT : SGMC_STEP_CARRY;
    if (piped) {
      cudaStreamWaitEvent(ms, staged[k & 1], 0);           // implies the transfer
      flags |= SGMC_STEP_X_STAGED | ((k & 1) ? SGMC_STEP_X_SLOT1 : 0);
    } else {
      cudaStreamWaitEvent(ms, copied[sl], 0);
    }
    cudaStreamWaitEvent(ms, read_back[k & 1], 0);          // its previous contents are on the host
    rc = sgmc_glm_sgld_step(ms, spec, theta, v, C, P, Xb, Xb + n * d, nullptr, nullptr, n,
                            observation_count, uv, uv + C, grad, (k & 1) ? keys_b : keys_a,
                            (k & 1) ? keys_a : keys_b, step_sizes[k], temperature, alpha, lmbd,
                            workspace, workspace_bytes, path, prng_layout, 0, nullptr, nullptr,
                            nullptr, 0, flags);
    cudaEventRecord(consumed[sl], ms);
    if (piped) cudaEventRecord(xfree[k & 1], ms);
    if (rc == 0 && keep && keep[k] && samples_out && *kept < capacity) {   // io.py:696-713
      if (check_cuda(cudaMemcpyAsync(samples_out + *kept * C * P, theta, (size_t)C * P * 4,
                                     cudaMemcpyDeviceToDevice, ms), "collect") ||
          check_cuda(cudaMemcpyAsync(scalars_out + *kept * C, uv, (size_t)C * 4,
                                     cudaMemcpyDeviceToDevice, ms), "collect"))
        rc = 1;
      else
        ++*kept;
    }
    if (host_results != nullptr) {
      cudaEventRecord(computed[k & 1], ms);
      cudaStreamWaitEvent(rs, computed[k & 1], 0);
      cudaMemcpyAsync(host_results + k * 2 * C, uv, (size_t)2 * C * 4, cudaMemcpyDeviceToHost, rs);
      cudaEventRecord(read_back[k & 1], rs);
    }
  }
  // the caller waits on `stream` and `copy_stream`: order the read-back and transfer
  // streams before the end of the copy stream
  cudaEventRecord(computed[0], rs);
  cudaStreamWaitEvent(cs, computed[0], 0);
  cudaEventRecord(begin, hs);
  cudaStreamWaitEvent(cs, begin, 0);
  if (piped)
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(staged[i]); cudaEventDestroy(xfree[i]); }
  // events can go (destruction is deferred by the runtime until they have completed)
  for (int i = 0; i < n_slots; ++i) { cudaEventDestroy(copied[i]); cudaEventDestroy(consumed[i]); }
  for (int i = 0; i < 2; ++i) { cudaEventDestroy(computed[i]); cudaEventDestroy(read_back[i]); }
  if (hybrid) {
    cudaEventRecord(begin, ps);
    cudaStreamWaitEvent(cs, begin, 0);
    for (int i = 0; i < n_slots; ++i) cudaEventDestroy(pulled[i]);
    cudaStreamDestroy(ps);
  }
  cudaEventDestroy(begin);
  cudaStreamDestroy(rs);
  cudaStreamDestroy(hs);
  if (rc) return rc;
  return check_cuda(cudaGetLastError(), "sgmc_glm_sgld_scan_host");
}
}  // namespace

int sgmc_glm_sgld_scan_host(void* stream, void* copy_stream, const sgmc_glm_spec* spec,
                            float* theta, float* v, int64_t n_chains, int64_t P,
                            const float* host_batches, int64_t host_batch_count,
                            int64_t n_steps, int64_t batch_size,
                            int64_t observation_count, float* device_slots, int n_slots,
                            float* potential_variance, float* host_results, float* grad,
                            uint32_t* keys_a, uint32_t* keys_b, const float* step_sizes,
                            float temperature, float alpha, float lmbd, void* workspace,
                            size_t workspace_bytes, int path, int prng_layout,
                            void* nccl_comm, int rank, int n_ranks, const uint8_t* keep,
                            float* samples_out, float* scalars_out, int64_t capacity,
                            int64_t* kept) {
  HostSource src;
  src.host_batches = host_batches;
  src.host_batch_count = host_batch_count;
  return scan_host_impl(stream, copy_stream, spec, theta, v, n_chains, P, src, n_steps,
                        batch_size, observation_count, device_slots, n_slots,
                        potential_variance, host_results, grad, keys_a, keys_b, step_sizes,
                        temperature, alpha, lmbd, workspace, workspace_bytes, path, prng_layout,
                        nccl_comm, rank, n_ranks, keep, samples_out, scalars_out, capacity, kept);
}

int sgmc_glm_sgld_scan_hybrid(void* stream, void* copy_stream, const sgmc_glm_spec* spec,
                              float* theta, float* v, int64_t n_chains, int64_t P,
                              const float* host_batches, int64_t host_batch_count,
                              int64_t rows_dma, const float* X_mapped, const int32_t* idx_all,
                              int pull_ctas, int64_t n_steps, int64_t batch_size,
                              int64_t observation_count, float* device_slots, int n_slots,
                              float* potential_variance, float* host_results, float* grad,
                              uint32_t* keys_a, uint32_t* keys_b, const float* step_sizes,
                              float temperature, float alpha, float lmbd, void* workspace,
                              size_t workspace_bytes, int path, int prng_layout,
                              void* nccl_comm, int rank, int n_ranks, const uint8_t* keep,
                              float* samples_out, float* scalars_out, int64_t capacity,
                              int64_t* kept) {
  HostSource src;
  src.host_batches = host_batches;
  src.host_batch_count = host_batch_count;
  src.rows_dma = rows_dma;
  src.X_mapped = X_mapped;
  src.idx_all = idx_all;
  src.pull_ctas = pull_ctas;
  return scan_host_impl(stream, copy_stream, spec, theta, v, n_chains, P, src, n_steps,
                        batch_size, observation_count, device_slots, n_slots,
                        potential_variance, host_results, grad, keys_a, keys_b, step_sizes,
                        temperature, alpha, lmbd, workspace, workspace_bytes, path, prng_layout,
                        nccl_comm, rank, n_ranks, keep, samples_out, scalars_out, capacity, kept);
}

int sgmc_glm_sgld_scan_pull(void* stream, void* copy_stream, const sgmc_glm_spec* spec,
                            float* theta, float* v, int64_t n_chains, int64_t P,
                            const float* X_mapped, const float* y_mapped,
                            const int32_t* idx_all, int pull_ctas,
                            int64_t n_steps, int64_t batch_size,
                            int64_t observation_count, float* device_slots, int n_slots,
                            float* potential_variance, float* host_results, float* grad,
                            uint32_t* keys_a, uint32_t* keys_b, const float* step_sizes,
                            float temperature, float alpha, float lmbd, void* workspace,
                            size_t workspace_bytes, int path, int prng_layout,
                            void* nccl_comm, int rank, int n_ranks, const uint8_t* keep,
                            float* samples_out, float* scalars_out, int64_t capacity,
                            int64_t* kept) {
  HostSource src;
  src.X_mapped = X_mapped;
  src.y_mapped = y_mapped;
  src.idx_all = idx_all;
  src.pull_ctas = pull_ctas;
  return scan_host_impl(stream, copy_stream, spec, theta, v, n_chains, P, src, n_steps,
                        batch_size, observation_count, device_slots, n_slots,
                        potential_variance, host_results, grad, keys_a, keys_b, step_sizes,
                        temperature, alpha, lmbd, workspace, workspace_bytes, path, prng_layout,
                        nccl_comm, rank, n_ranks, keep, samples_out, scalars_out, capacity, kept);
}

// K Langevin steps over a data set resident in HBM, one C call: the lax.scan of
// solver.mcmc (solver.py:152-160) with the draw of data/numpy_loader.py:128-141
// (idx_all == NULL) or pre-drawn index rows (idx_all = device int32[K][n], the
// host loader's cached index blocks), the step of integrator.py:860-922 and the
// collection of io.py:696-713: after step k with keep[k] != 0 the sample and its
// potential are copied to slot *kept of samples_out / scalars_out.
int sgmc_glm_sgld_scan_device(void* stream, const sgmc_glm_spec* spec, float* theta, float* v,
                              int64_t n_chains, int64_t P, const float* X, const float* y,
                              int64_t observation_count, int64_t batch_size,
                              uint32_t* data_key_a, uint32_t* data_key_b, int32_t* idx_buf,
                              const int32_t* idx_all, float* potential, float* variance,
                              float* grad, uint32_t* keys_a, uint32_t* keys_b,
                              const int64_t* leaf_sizes, int n_leaves,
                              const float* step_sizes, const float* temperatures,
                              const uint8_t* keep, int64_t n_steps, float* samples_out,
                              float* scalars_out, int64_t capacity, int64_t* kept,
                              float alpha, float lmbd, void* workspace,
                              size_t workspace_bytes, int path, int prng_layout) {
  SGMC_REQUIRE(spec && theta && X && y && potential && grad && keys_a && keys_b && step_sizes &&
               temperatures && kept, "null argument");
  SGMC_REQUIRE(idx_all != nullptr || (data_key_a && data_key_b && idx_buf),
               "device draws need the data keys and an index buffer");
  cudaStream_t ms = (cudaStream_t)stream;
  const int64_t C = n_chains, n = batch_size;
  // Tensor-core paths: the minibatch pipeline (index draw + operand staging) of step k+1
  // runs on a side stream while step k samples; two copies of the operands / indices.
  const bool piped = path != 0 && n_steps > 1 && !option(SGMC_OPT_NO_PIPELINE);
  cudaStream_t xs = nullptr;
  cudaEvent_t x_ready[2] = {nullptr, nullptr}, slot_free[2] = {nullptr, nullptr};
  int32_t* idx_pp[2] = {idx_buf, idx_buf};
  if (piped) {
    if (check_cuda(cudaStreamCreateWithFlags(&xs, cudaStreamNonBlocking), "side stream")) return 1;
    for (int i = 0; i < 2; ++i)
      if (check_cuda(cudaEventCreateWithFlags(&x_ready[i], cudaEventDisableTiming), "event") ||
          check_cuda(cudaEventCreateWithFlags(&slot_free[i], cudaEventDisableTiming), "event"))
        return 1;
    if (!idx_all) {
      float* ws = reinterpret_cast<float*>(
          (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
      idx_pp[1] = glm_tc_spare_idx(ws + 2 * (size_t)C * n, C, n, spec->d);
    }
    // the side stream starts after everything already queued on the sampling stream
    cudaEventRecord(slot_free[0], ms);
    cudaStreamWaitEvent(xs, slot_free[0], 0);
  }
  auto stage = [&](int64_t k) -> int {     // draw + operand staging of step k on the side stream
    const int sl = (int)(k & 1);
    const int32_t* idx = idx_all ? idx_all + k * n : idx_pp[sl];
    if (k >= 2) cudaStreamWaitEvent(xs, slot_free[sl], 0);   // step k-2 is done with this copy
    if (!idx_all) {
      if (int e = sgmc_minibatch_draw(xs, (k & 1) ? data_key_b : data_key_a,
                                      (k & 1) ? data_key_a : data_key_b, idx_pp[sl], n,
                                      observation_count, prng_layout))
        return e;
    }
    if (int e = sgmc_glm_prepare_minibatch(xs, spec, C, X, idx, n, workspace, workspace_bytes, path,
                                           sl))
      return e;
    return check_cuda(cudaEventRecord(x_ready[sl], xs), "event record");
  };
  int rc = 0;
  if (piped) rc = stage(0);
  for (int64_t k = 0; k < n_steps && rc == 0; ++k) {
    const int sl = (int)(k & 1);
    const int32_t* idx = idx_all ? idx_all + k * n : (piped ? idx_pp[sl] : idx_buf);
    int flags = k == 0 ? SGMC_STEP_CARRY_INIT : SGMC_STEP_CARRY;
    if (piped) {
      if (k + 1 < n_steps && (rc = stage(k + 1)) != 0) break;
      cudaStreamWaitEvent(ms, x_ready[sl], 0);
      flags |= SGMC_STEP_X_STAGED | (sl ? SGMC_STEP_X_SLOT1 : 0);
    } else if (!idx_all) {
      if ((rc = sgmc_minibatch_draw(stream, (k & 1) ? data_key_b : data_key_a,
                                    (k & 1) ? data_key_a : data_key_b, idx_buf, n,
                                    observation_count, prng_layout)) != 0)
        break;
    }
    if ((rc = sgmc_glm_sgld_step(stream, spec, theta, v, C, P, X, y, idx, nullptr, n,
                                 observation_count, potential, variance, grad,
                                 (k & 1) ? keys_b : keys_a, (k & 1) ? keys_a : keys_b,
                                 step_sizes[k], temperatures[k], alpha, lmbd, workspace,
                                 workspace_bytes, path, prng_layout, 0, nullptr, nullptr,
                                 leaf_sizes, n_leaves, flags)) != 0)
      break;
    if (piped) cudaEventRecord(slot_free[sl], ms);
    if (keep && keep[k] && samples_out && *kept < capacity) {
      if (check_cuda(cudaMemcpyAsync(samples_out + *kept * C * P, theta, (size_t)C * P * 4,
                                     cudaMemcpyDeviceToDevice, ms), "collect") ||
          check_cuda(cudaMemcpyAsync(scalars_out + *kept * C, potential, (size_t)C * 4,
                                     cudaMemcpyDeviceToDevice, ms), "collect")) {
        rc = 1;
        break;
      }
      ++*kept;
    }
  }
  if (piped) {
    // the data keys are advanced on the side stream: order the caller's next use after it
    cudaEventRecord(x_ready[0], xs);
    cudaStreamWaitEvent(ms, x_ready[0], 0);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(x_ready[i]); cudaEventDestroy(slot_free[i]); }
    cudaStreamDestroy(xs);
  }
  return rc;
}

}  // extern "C"
