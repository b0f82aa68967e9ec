// Bayesian MLP classifier potential + gradient, batched over chains (BASELINE.json
// configs[2]: 784-512-512-10, tanh, 256 chains).
//
// Replaces, for the dense-network likelihood, what the reference obtains from
// jax.value_and_grad(potential_fn) (jax_sgmc/integrator.py:593, :166, :792) over
// potential.minibatch_potential (jax_sgmc/potential.py:159-214) with the per-observation
// likelihood of examples/cifar.md:196-204 evaluated under vmap (potential.py:141-156):
// the forward pass, the softmax cross entropy, and the hand-derived reverse pass.
// Every chain has its own weights, so each layer is a chain-batched GEMM (blockIdx.z =
// chain); the minibatch is shared by all chains (gathered through `idx` inside the
// operand loads).  Launches per evaluation for L layers: 1 (prior) + L (forward) + 1
// (head) + per layer 2..3 (bias gradient, weight gradient, input gradient).
//
// This file is the fp32 path: 128 x 128 x 16 shared-memory tiles, 8 x 8 register
// micro-tiles, FFMA; operand loads are scalar and bounds-checked so any offset /
// alignment of the layers inside the raveled sample works (the reference's pytrees put
// biases of 10 floats between the weight matrices).
#include <algorithm>

#include "common.cuh"
#include "glm_math.cuh"

namespace sgmc {

constexpr int kBM = 128, kBN = 128, kBK = 16, kBThreads = 256, kBPad = 4;
constexpr int kSumsqParts = 32;

enum : int { kEpiStore = 0, kEpiBiasTanh = 1, kEpiBias = 2, kEpiDtanh = 3, kEpiPrior = 4 };

struct BgemmArgs {
  // C[b] (M x N) = opA(A[b]) (M x K) . opB(B[b]) (K x N)
  const float* A; int64_t a_batch, lda; const int32_t* a_idx;   // a_idx: gather on A's strided dim
  const float* B; int64_t b_batch, ldb;
  float* C; int64_t c_batch, ldc;
  int M, N, K;
  const float* bias; int64_t bias_batch;                        // kEpiBias*: + bias[n]
  const float* aux; int64_t aux_batch, ldaux;                   // kEpiDtanh: * (1 - aux^2); kEpiPrior: + coef * aux
  float coef;
  int64_t p0, prior_lo, prior_hi;                               // kEpiPrior: flat index of C(0,0), prior range
  // split-K (reductions over hundreds of thousands of rows with a tiny output: the conv
  // weight gradients): blockIdx.z = batch * n_split + split, a split covers k_chunk of K and
  // writes its raw accumulators to C + batch * c_batch + split * c_split (k_splitk_reduce adds
  // them in fixed order).  n_split <= 1: plain GEMM.
  int n_split, k_chunk; int64_t c_split;
};

// One operand tile -> registers.  MN_CONTIG: element (k, mn) at base + row(k) * ld + mn
// (the tile's M / N index is contiguous in memory); otherwise element (mn, k) at
// base + row(mn) * ld + k (K contiguous).  `row` applies the optional gather.
// `vec` (uniform per CTA, see operand_vec): two 16-byte loads per thread along the contiguous
// index instead of eight 4-byte ones; the register / shared-memory mapping differs, the
// values that land in As / Bs do not.
template <bool MN_CONTIG>
__device__ __forceinline__ void load_tile(float (&r)[8], const float* __restrict__ base, int64_t ld,
                                          const int32_t* __restrict__ idx, int mn0, int k0, int MN,
                                          int K, int t, bool vec) {
  if (vec) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int mn = MN_CONTIG ? mn0 + (t & 31) * 4 : mn0 + (t >> 2) + 64 * h;
      const int k = MN_CONTIG ? k0 + (t >> 5) + 8 * h : k0 + (t & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mn < MN && k < K) {
        const int64_t row = MN_CONTIG ? (idx ? idx[k] : k) : (idx ? idx[mn] : mn);
        v = __ldg(reinterpret_cast<const float4*>(base + row * ld + (MN_CONTIG ? mn : k)));
      }
      r[4 * h] = v.x; r[4 * h + 1] = v.y; r[4 * h + 2] = v.z; r[4 * h + 3] = v.w;
    }
  } else if (MN_CONTIG) {
    const int mn = mn0 + (t & 127);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + (t >> 7) + 2 * i;
      float v = 0.f;
      if (mn < MN && k < K) v = __ldg(base + (int64_t)(idx ? idx[k] : k) * ld + mn);
      r[i] = v;
    }
  } else {
    const int k = k0 + (t & 15);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int mn = mn0 + (t >> 4) + 16 * i;
      float v = 0.f;
      if (mn < MN && k < K) v = __ldg(base + (int64_t)(idx ? idx[mn] : mn) * ld + k);
      r[i] = v;
    }
  }
}

template <bool MN_CONTIG>
__device__ __forceinline__ void store_tile(float (*s)[kBM + kBPad], const float (&r)[8], int t,
                                           bool vec) {
  if (vec && MN_CONTIG) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      *reinterpret_cast<float4*>(&s[(t >> 5) + 8 * h][(t & 31) * 4]) =
          make_float4(r[4 * h], r[4 * h + 1], r[4 * h + 2], r[4 * h + 3]);
  } else if (vec) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[(t & 3) * 4 + j][(t >> 2) + 64 * h] = r[4 * h + j];
  } else if (MN_CONTIG) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s[(t >> 7) + 2 * i][t & 127] = r[i];
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) s[t & 15][(t >> 4) + 16 * i] = r[i];
  }
}

// 16-byte loads are legal for an operand when its (batch- and split-advanced) base and its
// leading dimension are 16-byte multiples and the extent along the contiguous index is a
// multiple of four (a float4 is then entirely inside or entirely outside the matrix).
__device__ __forceinline__ bool operand_vec(const float* base, int64_t ld, int contig_extent) {
  return ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (ld & 3) == 0 && (contig_extent & 3) == 0;
}

// A_T: A is stored K x M (M contiguous), else M x K (K contiguous).
// B_T: B is stored N x K (K contiguous), else K x N (N contiguous).
// BN: tile width, 128 or 64 (outputs of at most 64 columns -- the convolutions' channels, the
// class logits: the thread tile drops its second column half instead of multiplying padding).
template <bool A_T, bool B_T, int EPI, int BN = kBN>
__global__ void __launch_bounds__(kBThreads, BN == 64 ? 3 : 2) k_bgemm(const BgemmArgs a) {
  static_assert(BN == 128 || BN == 64, "tile width");
  constexpr int NH = BN / 64;          // column halves per thread
  __shared__ __align__(16) float As[2][kBK][kBM + kBPad];
  __shared__ __align__(16) float Bs[2][kBK][kBN + kBPad];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int n_end = min(a.N, n0 + BN);   // operand B: columns of this tile only
  const int ns = a.n_split > 1 ? a.n_split : 1;
  const int64_t b = blockIdx.z / ns;
  const int sp = (int)(blockIdx.z - b * ns);
  const int k_begin = ns > 1 ? sp * a.k_chunk : 0;
  const int K = ns > 1 ? min(a.k_chunk, a.K - k_begin) : a.K;
  // advance both operands along K to the split's range
  const float* A = a.A + b * a.a_batch + (a.a_idx ? 0 : (A_T ? (int64_t)k_begin * a.lda : k_begin));
  const float* B = a.B + b * a.b_batch + (B_T ? k_begin : (int64_t)k_begin * a.ldb);
  const int32_t* a_idx = a.a_idx ? a.a_idx + (A_T ? k_begin : 0) : nullptr;
  pdl_launch_dependents();
  pdl_wait();

  // accumulators as column pairs: one packed fma.rn.f32x2 per pair (same per-element
  // rounding as 64 scalar FMAs, half the issue slots)
  float2 acc[8][2 * NH];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 2 * NH; ++j) acc[i][j] = make_float2(0.f, 0.f);

  float ra[8], rb[8];
  const int kt = (K + kBK - 1) / kBK;
  const bool va = operand_vec(A, a.lda, A_T ? a.M : K);
  const bool vb = operand_vec(B, a.ldb, B_T ? K : a.N);
  load_tile<A_T>(ra, A, a.lda, a_idx, m0, 0, a.M, K, t, va);
  load_tile<!B_T>(rb, B, a.ldb, nullptr, n0, 0, n_end, K, t, vb);
  store_tile<A_T>(As[0], ra, t, va);
  store_tile<!B_T>(Bs[0], rb, t, vb);
  __syncthreads();
  for (int it = 0; it < kt; ++it) {
    const int cur = it & 1;
    if (it + 1 < kt) {
      load_tile<A_T>(ra, A, a.lda, a_idx, m0, (it + 1) * kBK, a.M, K, t, va);
      load_tile<!B_T>(rb, B, a.ldb, nullptr, n0, (it + 1) * kBK, n_end, K, t, vb);
    }
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float2 bv[2 * NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float4 q = *reinterpret_cast<const float4*>(&Bs[cur][kk][64 * h + tx * 4]);
        bv[2 * h] = make_float2(q.x, q.y);
        bv[2 * h + 1] = make_float2(q.z, q.w);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 ai = make_float2(av[i], av[i]);
#pragma unroll
        for (int j = 0; j < 2 * NH; ++j) acc[i][j] = __ffma2_rn(ai, bv[j], acc[i][j]);
      }
    }
    if (it + 1 < kt) {
      store_tile<A_T>(As[cur ^ 1], ra, t, va);
      store_tile<!B_T>(Bs[cur ^ 1], rb, t, vb);
    }
    __syncthreads();
  }

  float* Cb = a.C + b * a.c_batch + (ns > 1 ? sp * a.c_split : 0);
  const float* bias = (EPI == kEpiBiasTanh || EPI == kEpiBias) ? a.bias + b * a.bias_batch : nullptr;
  const float* aux = (EPI == kEpiDtanh || EPI == kEpiPrior) ? a.aux + b * a.aux_batch : nullptr;
  // four consecutive columns per (row, half): one 16-byte store (and aux load) when the
  // output / aux rows allow it
  const bool vc = operand_vec(Cb, a.ldc, a.N);
  const bool vx = aux != nullptr && operand_vec(aux, a.ldaux, a.N);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= a.M) continue;
#pragma unroll
    for (int jh = 0; jh < NH; ++jh) {
      const int nb = n0 + 64 * jh + tx * 4;
      if (nb >= a.N) continue;
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (EPI == kEpiDtanh || EPI == kEpiPrior) {
        if (vx) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(aux + (int64_t)m * a.ldaux + nb));
          x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (nb + j < a.N) x[j] = __ldg(aux + (int64_t)m * a.ldaux + nb + j);
        }
      }
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = nb + j;
        v[j] = (j & 1) ? acc[i][2 * jh + (j >> 1)].y : acc[i][2 * jh + (j >> 1)].x;
        if (n >= a.N) continue;
        if (EPI == kEpiBiasTanh) v[j] = tanhf(v[j] + __ldg(bias + n));
        if (EPI == kEpiBias) v[j] = v[j] + __ldg(bias + n);
        if (EPI == kEpiDtanh) v[j] = v[j] * (1.0f - x[j] * x[j]);
        if (EPI == kEpiPrior) {
          const int64_t p = a.p0 + (int64_t)m * a.ldc + n;
          if (p >= a.prior_lo && p < a.prior_hi) v[j] = fmaf(a.coef, x[j], v[j]);
        }
      }
      float* dst = Cb + (int64_t)m * a.ldc + nb;
      if (vc) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + j < a.N) dst[j] = v[j];
      }
    }
  }
}

template <bool A_T, bool B_T, int EPI>
static int launch_bgemm(cudaStream_t s, const BgemmArgs& a, int64_t batch, const char* name) {
  SGMC_REQUIRE(batch <= 65535, "%s: more than 65535 chains per call", name);
  const int64_t z = batch * (a.n_split > 1 ? a.n_split : 1);
  SGMC_REQUIRE(z <= 65535, "%s: too many chains x K splits", name);
  const unsigned gy = (unsigned)((a.M + kBM - 1) / kBM);
  if (a.N <= 64)
    launch_pdl(k_bgemm<A_T, B_T, EPI, 64>, dim3(1, gy, (unsigned)z), dim3(kBThreads), 0, s, a);
  else
    launch_pdl(k_bgemm<A_T, B_T, EPI, kBN>, dim3((unsigned)((a.N + kBN - 1) / kBN), gy, (unsigned)z),
               dim3(kBThreads), 0, s, a);
  return post_launch(name);
}

// ---- prior: per-chain partial sums of theta^2 over [lo, hi) ------------------------
__global__ void __launch_bounds__(256) k_mlp_sumsq(const float* __restrict__ theta, int64_t P,
                                                   int64_t lo, int64_t hi,
                                                   float* __restrict__ parts) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t c = blockIdx.y;
  const int64_t span = hi - lo, per = (span + kSumsqParts - 1) / kSumsqParts;
  const int64_t b0 = lo + blockIdx.x * per, b1 = min(hi, b0 + per);
  const float* row = theta + c * P;
  float s = 0.f;
  for (int64_t p = b0 + threadIdx.x; p < b1; p += 256) {
    const float v = row[p];
    s = fmaf(v, v, s);
  }
  __shared__ float red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) parts[c * kSumsqParts + blockIdx.x] = red[0];
}

// ---- head: softmax cross entropy, statistics, d logits ------------------------------
// One CTA per chain.  ell_i = logits[label_i] - logsumexp(logits_i) (the max-shifted form
// of jax.nn.log_softmax); U = (-N mean(ell) - prior) / T or the masked form
// (potential.py:183-185, :210); var(ell) (integrator.py:880);
// dlogits_i = cot_i (onehot - softmax), cot_i = (-N / n) / T (* mask_i).
struct HeadArgs {
  const float* logits;      // [C][n][K]
  float* dlogits;           // [C][n][K] or null
  float* ell_ws;            // [C][n]
  float* ell_out;           // [C][n] or null
  const float* y; const int32_t* idx; const float* mask;
  int n, K;
  float n_obs, inv_temperature, prior_half_inv;
  const float* sumsq_parts; // [C][kSumsqParts] or null (flat prior)
  float* potential; float* variance;
};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  __syncthreads();
  red[threadIdx.x] = v;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  return red[0];
}

__global__ void __launch_bounds__(256) k_mlp_head(const HeadArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[256];
  const int64_t c = blockIdx.x;
  const float cot = (-a.n_obs / (float)a.n) * a.inv_temperature;
  float s_ell = 0.f, s_mask = 0.f;
  for (int i = threadIdx.x; i < a.n; i += 256) {
    const float* lg = a.logits + (c * a.n + i) * a.K;
    const int label = (int)a.y[a.idx ? a.idx[i] : i];
    float m = -INFINITY;
    for (int k = 0; k < a.K; ++k) m = fmaxf(m, lg[k]);
    float se = 0.f;
    for (int k = 0; k < a.K; ++k) se += expf(lg[k] - m);
    const float lse = logf(se);
    const float mk = a.mask ? a.mask[i] : 1.0f;
    const float l = (lg[label] - m) - lse;
    a.ell_ws[c * a.n + i] = l;
    if (a.ell_out) a.ell_out[c * a.n + i] = l;
    s_ell += l;
    s_mask = fmaf(l, mk, s_mask);
    if (a.dlogits) {
      float* dl = a.dlogits + (c * a.n + i) * a.K;
      const float ci = cot * mk;
      for (int k = 0; k < a.K; ++k) {
        const float soft = expf(lg[k] - m) / se;
        dl[k] = ci * ((k == label ? 1.0f : 0.0f) - soft);
      }
    }
  }
  const float sum = block_sum_256(s_ell, red);
  const float msum = block_sum_256(s_mask, red);
  const float mean = sum / (float)a.n;
  float dev = 0.f;
  for (int i = threadIdx.x; i < a.n; i += 256) {
    const float dlt = a.ell_ws[c * a.n + i] - mean;
    dev = fmaf(dlt, dlt, dev);
  }
  const float m2 = block_sum_256(dev, red);
  if (threadIdx.x == 0) {
    float prior = 0.f;
    if (a.sumsq_parts) {
      float ss = 0.f;
      for (int q = 0; q < kSumsqParts; ++q) ss += a.sumsq_parts[c * kSumsqParts + q];
      prior = -a.prior_half_inv * ss;
    }
    const float L = a.mask ? (-a.n_obs / (float)a.n) * msum : -a.n_obs * mean;
    a.potential[c] = (L - prior) * a.inv_temperature;
    if (a.variance) a.variance[c] = m2 / (float)a.n;
  }
}

// ---- bias gradient: db[c][o] = sum_i dA[c][i][o] (+ prior term) -----------------------
__global__ void __launch_bounds__(256) k_mlp_bias_grad(const float* __restrict__ dA, int n, int out,
                                                       const float* __restrict__ theta,
                                                       float* __restrict__ grad, int64_t P,
                                                       int64_t b_off, float coef, int64_t prior_lo,
                                                       int64_t prior_hi) {
  pdl_launch_dependents();
  pdl_wait();
  const int o = blockIdx.x * 256 + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (o >= out) return;
  const float* src = dA + c * (int64_t)n * out + o;
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += src[(int64_t)i * out];
  const int64_t p = b_off + o;
  if (p >= prior_lo && p < prior_hi) s = fmaf(coef, theta[c * P + p], s);
  grad[c * P + p] = s;
}

struct MlpWs {
  float* H[SGMC_MLP_MAX_LAYERS + 1];   // activations h_1 .. h_{L-1}, logits = H[L]
  float* dA[SGMC_MLP_MAX_LAYERS + 1];  // dA_1 .. dA_L
  float* ell; float* sumsq;
};

static size_t mlp_carve(MlpWs* w, uint8_t* base, const sgmc_mlp_spec& s, int64_t C, int64_t n) {
  size_t off = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += (floats * 4 + 255) & ~(size_t)255;
    return p;
  };
  for (int l = 1; l <= s.n_layers; ++l) {
    float* h = take((size_t)C * n * s.sizes[l]);
    float* d = take((size_t)C * n * s.sizes[l]);
    if (w) { w->H[l] = h; w->dA[l] = d; }
  }
  float* ell = take((size_t)C * n);
  float* sq = take((size_t)C * kSumsqParts);
  if (w) { w->ell = ell; w->sumsq = sq; }
  return off;
}

static int mlp_check(const sgmc_mlp_spec* s, int64_t P) {
  SGMC_REQUIRE(s != nullptr, "null spec");
  SGMC_REQUIRE(s->n_layers >= 1 && s->n_layers <= SGMC_MLP_MAX_LAYERS, "1..%d layers",
               SGMC_MLP_MAX_LAYERS);
  SGMC_REQUIRE(s->activation == 0, "activation: 0 (tanh)");
  SGMC_REQUIRE(s->prior == kPriorFlat || s->prior == kPriorGaussian, "prior: flat or gaussian");
  for (int l = 0; l < s->n_layers; ++l) {
    SGMC_REQUIRE(s->sizes[l] > 0 && s->sizes[l + 1] > 0, "layer widths must be positive");
    SGMC_REQUIRE(s->w_off[l] >= 0 && s->w_off[l] + (int64_t)s->sizes[l] * s->sizes[l + 1] <= P &&
                 s->b_off[l] >= 0 && s->b_off[l] + s->sizes[l + 1] <= P,
                 "layer %d lies outside the sample", l);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// Convolutional classifier (BASELINE.json configs[4]: SGGMC / AMAGOLD on a CIFAR-10-shape
// CNN): 3x3 convolutions with zero padding 1 and stride 1 or 2, tanh, then one dense layer
// on the flattened NHWC feature map and the softmax cross entropy of the head above.  Every
// chain has its own filters, the minibatch is shared: a convolution is im2col (patch index
// (kh*3 + kw)*Cin + cin, the row-major order of the HWIO filter tensor) followed by the
// chain-batched GEMM of the dense layers; its reverse pass is the same two GEMMs as a dense
// layer (dW = patches^T . dZ, dPatches = dZ . W^T) followed by col2im.
struct ConvGeom {
  int n, Hi, Wi, Cin, Ho, Wo, stride;
};

// dst[b][(i, ho, wo)][(kh*3 + kw)*Cin + c] = src[b][i'][ho*s + kh - 1][wo*s + kw - 1][c]
// (zero outside); i' = idx[i] when the source is the data set itself (first layer).
__global__ void __launch_bounds__(256) k_im2col(const float* __restrict__ src, int64_t src_batch,
                                                const int32_t* __restrict__ idx, float* __restrict__ dst,
                                                const ConvGeom g, int64_t total_per_batch) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t b = blockIdx.y;
  const int K = 9 * g.Cin;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total_per_batch;
       e += (int64_t)gridDim.x * 256) {
    const int k = (int)(e % K);
    const int64_t row = e / K;
    const int c = k % g.Cin, kk = k / g.Cin, kh = kk / 3, kw = kk - kh * 3;
    const int wo = (int)(row % g.Wo);
    const int64_t r2 = row / g.Wo;
    const int ho = (int)(r2 % g.Ho);
    const int i = (int)(r2 / g.Ho);
    const int hi = ho * g.stride + kh - 1, wi = wo * g.stride + kw - 1;
    float v = 0.f;
    if (hi >= 0 && hi < g.Hi && wi >= 0 && wi < g.Wi) {
      const int64_t obs = idx ? idx[i] : i;
      v = __ldg(src + b * src_batch + ((obs * g.Hi + hi) * g.Wi + wi) * g.Cin + c);
    }
    dst[b * total_per_batch + e] = v;
  }
}

// dZ[b][(i, hi, wi)][c] = (sum over the patches that contain input pixel (hi, wi)) dP * (1 - h^2)
__global__ void __launch_bounds__(256) k_col2im_dtanh(const float* __restrict__ dP,
                                                      const float* __restrict__ h,
                                                      float* __restrict__ dZ, const ConvGeom g,
                                                      int64_t in_per_batch) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t b = blockIdx.y;
  const int K = 9 * g.Cin;
  const int64_t rows = (int64_t)g.n * g.Ho * g.Wo;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < in_per_batch;
       e += (int64_t)gridDim.x * 256) {
    const int c = (int)(e % g.Cin);
    const int64_t r1 = e / g.Cin;
    const int wi = (int)(r1 % g.Wi);
    const int64_t r2 = r1 / g.Wi;
    const int hi = (int)(r2 % g.Hi);
    const int i = (int)(r2 / g.Hi);
    float s = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int th = hi + 1 - kh;
      if (th < 0 || th % g.stride) continue;
      const int ho = th / g.stride;
      if (ho >= g.Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tw = wi + 1 - kw;
        if (tw < 0 || tw % g.stride) continue;
        const int wo = tw / g.stride;
        if (wo >= g.Wo) continue;
        const int64_t row = ((int64_t)i * g.Ho + ho) * g.Wo + wo;
        s += dP[(b * rows + row) * K + (kh * 3 + kw) * g.Cin + c];
      }
    }
    const float hv = h[b * in_per_batch + e];
    dZ[b * in_per_batch + e] = s * (1.0f - hv * hv);
  }
}

// The same two kernels for channel counts that are multiples of four (every layer after the
// first on RGB input): four channels per thread as one 16-byte access, 32-bit index
// arithmetic (the per-chain extent is checked against 2^31 at the call site).  The sums of
// col2im run in the same (kh, kw) order, so the results are bit-identical to the scalar form.
__global__ void __launch_bounds__(256) k_im2col_v4(const float* __restrict__ src, int64_t src_batch,
                                                   const int32_t* __restrict__ idx, float* __restrict__ dst,
                                                   const ConvGeom g, uint32_t quads_per_batch) {
  pdl_launch_dependents();
  pdl_wait();
  const float* sb = src + blockIdx.y * src_batch;
  float4* db = reinterpret_cast<float4*>(dst) + (int64_t)blockIdx.y * quads_per_batch;
  const uint32_t C4 = (uint32_t)g.Cin >> 2, K4 = 9u * C4;
  for (uint32_t e = blockIdx.x * 256u + threadIdx.x; e < quads_per_batch; e += gridDim.x * 256u) {
    const uint32_t k4 = e % K4, row = e / K4;
    const uint32_t c4 = k4 % C4, kk = k4 / C4, kh = kk / 3u, kw = kk - kh * 3u;
    const uint32_t wo = row % (uint32_t)g.Wo, r2 = row / (uint32_t)g.Wo;
    const uint32_t ho = r2 % (uint32_t)g.Ho, i = r2 / (uint32_t)g.Ho;
    const int hi = (int)(ho * g.stride + kh) - 1, wi = (int)(wo * g.stride + kw) - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hi >= 0 && hi < g.Hi && wi >= 0 && wi < g.Wi) {
      const int64_t obs = idx ? idx[i] : (int64_t)i;
      v = __ldg(reinterpret_cast<const float4*>(sb + ((obs * g.Hi + hi) * g.Wi + wi) * g.Cin) + c4);
    }
    db[e] = v;
  }
}

__global__ void __launch_bounds__(256) k_col2im_dtanh_v4(const float* __restrict__ dP,
                                                         const float* __restrict__ h,
                                                         float* __restrict__ dZ, const ConvGeom g,
                                                         uint32_t quads_per_batch) {
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t C4 = (uint32_t)g.Cin >> 2, K4 = 9u * C4;
  const int64_t rows = (int64_t)g.n * g.Ho * g.Wo;
  const float4* pb = reinterpret_cast<const float4*>(dP) + (int64_t)blockIdx.y * rows * K4;
  const float4* hb = reinterpret_cast<const float4*>(h) + (int64_t)blockIdx.y * quads_per_batch;
  float4* zb = reinterpret_cast<float4*>(dZ) + (int64_t)blockIdx.y * quads_per_batch;
  for (uint32_t e = blockIdx.x * 256u + threadIdx.x; e < quads_per_batch; e += gridDim.x * 256u) {
    const uint32_t c4 = e % C4, r1 = e / C4;
    const int wi = (int)(r1 % (uint32_t)g.Wi);
    const uint32_t r2 = r1 / (uint32_t)g.Wi;
    const int hi = (int)(r2 % (uint32_t)g.Hi);
    const uint32_t i = r2 / (uint32_t)g.Hi;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int th = hi + 1 - kh;
      if (th < 0 || th % g.stride) continue;
      const int ho = th / g.stride;
      if (ho >= g.Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tw = wi + 1 - kw;
        if (tw < 0 || tw % g.stride) continue;
        const int wo = tw / g.stride;
        if (wo >= g.Wo) continue;
        const int64_t row = ((int64_t)i * g.Ho + ho) * g.Wo + wo;
        const float4 q = pb[row * K4 + (uint32_t)(kh * 3 + kw) * C4 + c4];
        s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
      }
    }
    const float4 hv = hb[e];
    zb[e] = make_float4(s.x * (1.0f - hv.x * hv.x), s.y * (1.0f - hv.y * hv.y),
                        s.z * (1.0f - hv.z * hv.z), s.w * (1.0f - hv.w * hv.w));
  }
}

// out[b][m][n] = sum_s part[b][s][m][n] in split order (+ coef * aux on the prior range):
// the second stage of the split-K GEMMs and of the split column sums
__global__ void __launch_bounds__(256) k_splitk_reduce(const float* __restrict__ part, int n_split,
                                                       int64_t MN, int N, float* __restrict__ out,
                                                       int64_t out_batch, int64_t ldc,
                                                       const float* __restrict__ aux,
                                                       int64_t aux_batch, float coef, int64_t p0,
                                                       int64_t prior_lo, int64_t prior_hi,
                                                       const float* __restrict__ bias,
                                                       int64_t bias_batch) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t b = blockIdx.y;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < MN; e += (int64_t)gridDim.x * 256) {
    const float* src = part + b * n_split * MN + e;
    float s = 0.f;
    for (int q = 0; q < n_split; ++q) s += src[(int64_t)q * MN];
    const int64_t m = e / N, n = e - m * N;
    const int64_t o = m * ldc + n;
    const int64_t p = p0 + o;
    if (aux != nullptr && p >= prior_lo && p < prior_hi) s = fmaf(coef, aux[b * aux_batch + o], s);
    if (bias != nullptr) s += bias[b * bias_batch + n];
    out[b * out_batch + o] = s;
  }
}

// Tall-skinny weight gradient of a convolution with few channels (M N <= 1024 outputs over
// hundreds of thousands of rows, e.g. 27 x 32 for the first layer on RGB images), where a
// 128 x 128 GEMM tile would be 95 % padding: part[c][s][m][n] = sum over the split's rows
// of P[row][m] dZ[c][row][n].  A CTA stages 64 rows of both operands in shared memory; a
// thread owns up to four outputs -- four neighbouring columns of one output row when N is a
// multiple of four (one 16-byte and one 4-byte shared-memory load per four FMAs), any four
// otherwise.  Every output adds its rows in row order either way.
__global__ void __launch_bounds__(256) k_dw_tallskinny(const float* __restrict__ Pm, int64_t p_batch,
                                                       const float* __restrict__ dZ, int64_t rows,
                                                       int M, int N, int64_t row_chunk,
                                                       float* __restrict__ part) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float ts_smem[];
  float* sP = ts_smem;                 // [64][M]
  float* sZ = ts_smem + 64 * M;        // [64][N]
  const int64_t c = blockIdx.y, sp = blockIdx.x;
  const int64_t r0 = sp * row_chunk, r1 = min(rows, r0 + row_chunk);
  const float* P = Pm + c * p_batch;
  const float* Z = dZ + c * rows * N;
  const int MN = M * N;
  const bool quads = (N & 3) == 0;     // M N / 4 <= 256 quads: one per thread
  int om[4], on[4];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int o = quads ? threadIdx.x * 4 + q : threadIdx.x + q * 256;
    om[q] = o < MN ? o / N : 0;
    on[q] = o < MN ? o - om[q] * N : 0;
  }
  for (int64_t r = r0; r < r1; r += 64) {
    const int nr = (int)min((int64_t)64, r1 - r);
    __syncthreads();
    for (int e = threadIdx.x; e < nr * M; e += 256) sP[e] = P[r * M + e];
    for (int e = threadIdx.x; e < nr * N; e += 256) sZ[e] = Z[r * N + e];
    __syncthreads();
    if (quads) {
      if (threadIdx.x * 4 < MN) {
        const float* p = sP + om[0];
        const float* z = sZ + on[0];
#pragma unroll 8
        for (int i = 0; i < nr; ++i) {
          const float a = p[i * M];
          const float4 q = *reinterpret_cast<const float4*>(z + i * N);
          acc[0] = fmaf(a, q.x, acc[0]);
          acc[1] = fmaf(a, q.y, acc[1]);
          acc[2] = fmaf(a, q.z, acc[2]);
          acc[3] = fmaf(a, q.w, acc[3]);
        }
      }
      continue;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (threadIdx.x + q * 256 >= MN) break;
      float a = acc[q];
      for (int i = 0; i < nr; ++i) a = fmaf(sP[i * M + om[q]], sZ[i * N + on[q]], a);
      acc[q] = a;
    }
  }
  float* out = part + (c * gridDim.x + sp) * MN;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int o = quads ? threadIdx.x * 4 + q : threadIdx.x + q * 256;
    if (o < MN) out[o] = acc[q];
  }
}

// part[c][s][o] = sum over the split's rows of dZ[c][row][o]: 32 columns x 8 row lanes
__global__ void __launch_bounds__(256) k_colsum_part(const float* __restrict__ dZ, int64_t rows,
                                                     int out, int64_t row_chunk,
                                                     float* __restrict__ part) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + tx;
  const int64_t c = blockIdx.y, sp = blockIdx.z;
  const int64_t r0 = sp * row_chunk, r1 = min(rows, r0 + row_chunk);
  float s = 0.f;
  if (o < out) {
    const float* src = dZ + c * rows * out + o;
    for (int64_t r = r0 + ty; r < r1; r += 8) s += src[r * out];
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && o < out) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][tx];
    part[(c * gridDim.z + sp) * out + o] = t;
  }
}

// db[c][o] = sum over `rows` rows of dZ[c][row][o] (+ prior term): 32 columns x 8 row lanes
__global__ void __launch_bounds__(256) k_colsum_grad(const float* __restrict__ dZ, int64_t rows,
                                                     int out, const float* __restrict__ theta,
                                                     float* __restrict__ grad, int64_t P,
                                                     int64_t b_off, float coef, int64_t prior_lo,
                                                     int64_t prior_hi) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + tx;
  const int64_t c = blockIdx.y;
  float s = 0.f;
  if (o < out) {
    const float* src = dZ + c * rows * out + o;
    for (int64_t r = ty; r < rows; r += 8) s += src[r * out];
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && o < out) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][tx];
    const int64_t p = b_off + o;
    if (p >= prior_lo && p < prior_hi) t = fmaf(coef, theta[c * P + p], t);
    grad[c * P + p] = t;
  }
}

struct CnnLayer {
  ConvGeom g;
  int64_t rows, K;       // rows = n * Ho * Wo, K = 9 * Cin
  int Cout;
};

struct CnnWs {
  float* P[SGMC_CNN_MAX_CONV];       // patches (layer 0: shared by the chains)
  float* H[SGMC_CNN_MAX_CONV + 1];   // H[l + 1] = tanh output of conv layer l
  float* dZ[SGMC_CNN_MAX_CONV + 1];  // dZ[l + 1] = gradient w.r.t. the pre-activation of layer l
  float* dP;                         // largest patch-gradient buffer (layers >= 1)
  float* part;                       // split-K partials of the conv weight / bias gradients
  float* logits; float* dlogits; float* ell; float* sumsq;
};

constexpr int kSplitChunk = 2048;    // rows of K per split of a conv weight gradient
static int cnn_splits(int64_t rows) { return (int)std::min<int64_t>((rows + kSplitChunk - 1) / kSplitChunk, 512); }

static int cnn_layers(const sgmc_cnn_spec& s, int64_t n, CnnLayer* L) {
  int Hi = s.height, Wi = s.width;
  for (int l = 0; l < s.n_conv; ++l) {
    const int st = s.stride[l];
    CnnLayer& q = L[l];
    q.g.n = (int)n; q.g.Hi = Hi; q.g.Wi = Wi; q.g.Cin = s.channels[l]; q.g.stride = st;
    q.g.Ho = (Hi + 2 - 3) / st + 1; q.g.Wo = (Wi + 2 - 3) / st + 1;
    q.rows = n * (int64_t)q.g.Ho * q.g.Wo; q.K = 9 * (int64_t)s.channels[l];
    q.Cout = s.channels[l + 1];
    Hi = q.g.Ho; Wi = q.g.Wo;
  }
  return Hi * Wi * s.channels[s.n_conv];      // features of the dense head
}

// The dense head has n x n_classes outputs per chain over F features: with few chains that
// is a handful of tiles with a long serial K loop, so K is split like the weight gradients
// (512 features per split at least, 16 splits at most) whenever the plain grid would not
// fill the SMs.
static int cnn_head_splits(int64_t F, int64_t n, int64_t C) {
  if (((n + kBM - 1) / kBM) * C >= 2 * 148 || F < 1024) return 1;
  return (int)std::min<int64_t>(std::min<int64_t>(16, F / 512), 65535 / C);
}

static size_t cnn_carve(CnnWs* w, uint8_t* base, const sgmc_cnn_spec& s, int64_t C, int64_t n) {
  CnnLayer L[SGMC_CNN_MAX_CONV];
  const int64_t F_head = cnn_layers(s, n, L);
  size_t off = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += (floats * 4 + 255) & ~(size_t)255;
    return p;
  };
  size_t dp_max = 0, part_max = 1;
  for (int l = 0; l < s.n_conv; ++l) {
    float* p = take((size_t)(l == 0 ? 1 : C) * L[l].rows * L[l].K);
    float* h = take((size_t)C * L[l].rows * L[l].Cout);
    float* d = take((size_t)C * L[l].rows * L[l].Cout);
    if (l > 0) dp_max = std::max(dp_max, (size_t)C * L[l].rows * L[l].K);
    part_max = std::max(part_max, (size_t)C * std::max(cnn_splits(L[l].rows), 2) * L[l].K * L[l].Cout);
    if (w) { w->P[l] = p; w->H[l + 1] = h; w->dZ[l + 1] = d; }
  }
  part_max = std::max(part_max, (size_t)C * cnn_head_splits(F_head, n, C) * n * s.n_classes);
  float* dp = take(dp_max ? dp_max : 1);
  float* part = take(part_max);
  if (w) w->part = part;
  float* lg = take((size_t)C * n * s.n_classes);
  float* dl = take((size_t)C * n * s.n_classes);
  float* ell = take((size_t)C * n);
  float* sq = take((size_t)C * kSumsqParts);
  if (w) { w->dP = dp; w->logits = lg; w->dlogits = dl; w->ell = ell; w->sumsq = sq; }
  return off;
}

static int cnn_check(const sgmc_cnn_spec* s, int64_t P) {
  SGMC_REQUIRE(s != nullptr, "null spec");
  SGMC_REQUIRE(s->n_conv >= 1 && s->n_conv <= SGMC_CNN_MAX_CONV, "1..%d conv layers",
               SGMC_CNN_MAX_CONV);
  SGMC_REQUIRE(s->height > 0 && s->width > 0 && s->n_classes > 0, "bad geometry");
  SGMC_REQUIRE(s->prior == kPriorFlat || s->prior == kPriorGaussian, "prior: flat or gaussian");
  int Hi = s->height, Wi = s->width;
  for (int l = 0; l < s->n_conv; ++l) {
    SGMC_REQUIRE(s->channels[l] > 0 && s->channels[l + 1] > 0, "channels must be positive");
    SGMC_REQUIRE(s->stride[l] == 1 || s->stride[l] == 2, "stride 1 or 2");
    SGMC_REQUIRE(s->w_off[l] >= 0 && s->w_off[l] + 9ll * s->channels[l] * s->channels[l + 1] <= P &&
                 s->b_off[l] >= 0 && s->b_off[l] + s->channels[l + 1] <= P,
                 "conv layer %d lies outside the sample", l);
    Hi = (Hi - 1) / s->stride[l] + 1; Wi = (Wi - 1) / s->stride[l] + 1;
  }
  const int64_t F = (int64_t)Hi * Wi * s->channels[s->n_conv];
  SGMC_REQUIRE(s->w_off[s->n_conv] >= 0 && s->w_off[s->n_conv] + F * s->n_classes <= P &&
               s->b_off[s->n_conv] >= 0 && s->b_off[s->n_conv] + s->n_classes <= P,
               "the dense head lies outside the sample");
  return 0;
}

}  // namespace sgmc

using namespace sgmc;

extern "C" {

size_t sgmc_cnn_workspace_bytes(const sgmc_cnn_spec* spec, int64_t n_chains, int64_t batch_size) {
  if (!spec || spec->n_conv < 1 || spec->n_conv > SGMC_CNN_MAX_CONV) return 0;
  return cnn_carve(nullptr, nullptr, *spec, n_chains, batch_size) + 256;
}

int sgmc_cnn_potential_grad(void* stream, const sgmc_cnn_spec* spec, const float* theta,
                            int64_t n_chains, int64_t P, const float* X, const float* y,
                            const int32_t* idx, const float* mask, int64_t batch_size,
                            int64_t observation_count, float* potential, float* variance,
                            float* grad, float* ell, void* workspace, size_t workspace_bytes) {
  if (int e = cnn_check(spec, P)) return e;
  SGMC_REQUIRE(theta && X && y && potential && workspace, "null argument");
  const int64_t C = n_chains, n = batch_size;
  SGMC_REQUIRE(C > 0 && C <= 65535 && n > 0 && n <= (1 << 20), "bad sizes");
  SGMC_REQUIRE(workspace_bytes >= sgmc_cnn_workspace_bytes(spec, C, n), "workspace too small");
  const sgmc_cnn_spec& sp = *spec;
  const int NL = sp.n_conv;
  cudaStream_t s = (cudaStream_t)stream;
  CnnWs w{};
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) &
                                             ~(uintptr_t)255);
  cnn_carve(&w, base, sp, C, n);
  CnnLayer L[SGMC_CNN_MAX_CONV];
  const int F = cnn_layers(sp, n, L);
  for (int l = 0; l < NL; ++l)
    SGMC_REQUIRE(L[l].rows < (1ll << 31) && L[l].rows * L[l].K < (1ll << 40), "layer %d too large", l);
  const bool gauss = sp.prior == kPriorGaussian;
  const int64_t prior_lo = gauss ? sp.prior_off : 0, prior_hi = gauss ? sp.prior_off + sp.prior_size : 0;
  const float inv_T = 1.0f / sp.temperature;
  const float coef = gauss ? (1.0f / (sp.prior_scale * sp.prior_scale)) / sp.temperature : 0.f;
  auto blocks = [](int64_t total) { return (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 32); };

  if (gauss && prior_hi > prior_lo) {
    launch_pdl(k_mlp_sumsq, dim3(kSumsqParts, (unsigned)C), dim3(256), 0, s, theta, P, prior_lo,
               prior_hi, w.sumsq);
    if (post_launch("k_mlp_sumsq")) return 1;
  }
  // ---- forward ---------------------------------------------------------------------------
  for (int l = 0; l < NL; ++l) {
    const int64_t per = L[l].rows * L[l].K;
    const int64_t in_per = (int64_t)n * L[l].g.Hi * L[l].g.Wi * L[l].g.Cin;
    const float* im_src = l == 0 ? X : (const float*)w.H[l];
    const bool v4 = L[l].g.Cin % 4 == 0 && per < (1ll << 31) &&
                    (reinterpret_cast<uintptr_t>(im_src) & 15) == 0;
    if (v4)
      launch_pdl(k_im2col_v4, dim3(blocks(per / 4), l == 0 ? 1u : (unsigned)C), dim3(256), 0, s, im_src,
                 l == 0 ? (int64_t)0 : in_per, l == 0 ? idx : (const int32_t*)nullptr, w.P[l], L[l].g,
                 (uint32_t)(per / 4));
    else if (l == 0)
      launch_pdl(k_im2col, dim3(blocks(per), 1), dim3(256), 0, s, X, (int64_t)0, idx, w.P[0], L[0].g, per);
    else
      launch_pdl(k_im2col, dim3(blocks(per), (unsigned)C), dim3(256), 0, s, (const float*)w.H[l],
                 in_per, (const int32_t*)nullptr, w.P[l], L[l].g, per);
    if (post_launch("k_im2col")) return 1;
    BgemmArgs g{};
    g.A = w.P[l]; g.a_batch = l == 0 ? 0 : per; g.lda = L[l].K;
    g.B = theta + sp.w_off[l]; g.b_batch = P; g.ldb = L[l].Cout;
    g.C = w.H[l + 1]; g.c_batch = L[l].rows * L[l].Cout; g.ldc = L[l].Cout;
    g.M = (int)L[l].rows; g.N = L[l].Cout; g.K = (int)L[l].K;
    g.bias = theta + sp.b_off[l]; g.bias_batch = P;
    if (launch_bgemm<false, false, kEpiBiasTanh>(s, g, C, "k_bgemm<conv,tanh>")) return 1;
  }
  {   // dense head on the flattened NHWC map: logits [n, K] = H_L [n, F] . W [F, K] + b
    BgemmArgs g{};
    g.A = w.H[NL]; g.a_batch = n * (int64_t)F; g.lda = F;
    g.B = theta + sp.w_off[NL]; g.b_batch = P; g.ldb = sp.n_classes;
    g.C = w.logits; g.c_batch = n * (int64_t)sp.n_classes; g.ldc = sp.n_classes;
    g.M = (int)n; g.N = sp.n_classes; g.K = F;
    g.bias = theta + sp.b_off[NL]; g.bias_batch = P;
    const int Sh = cnn_head_splits(F, n, C);
    if (Sh > 1) {
      const int64_t MN = n * (int64_t)sp.n_classes;
      g.C = w.part; g.c_batch = Sh * MN; g.c_split = MN;
      g.n_split = Sh; g.k_chunk = (int)((((int64_t)F + Sh - 1) / Sh + kBK - 1) / kBK * kBK);
      if (launch_bgemm<false, false, kEpiStore>(s, g, C, "k_bgemm<head, split-K>")) return 1;
      launch_pdl(k_splitk_reduce, dim3((unsigned)std::min<int64_t>((MN + 255) / 256, 1024), (unsigned)C),
                 dim3(256), 0, s, (const float*)w.part, Sh, MN, sp.n_classes, w.logits, MN,
                 (int64_t)sp.n_classes, (const float*)nullptr, (int64_t)0, 0.f, (int64_t)0,
                 (int64_t)0, (int64_t)0, theta + sp.b_off[NL], P);
      if (post_launch("k_splitk_reduce<head>")) return 1;
    } else if (launch_bgemm<false, false, kEpiBias>(s, g, C, "k_bgemm<head>")) {
      return 1;
    }
  }
  HeadArgs h{};
  h.logits = w.logits; h.dlogits = grad ? w.dlogits : nullptr; h.ell_ws = w.ell; h.ell_out = ell;
  h.y = y; h.idx = idx; h.mask = mask; h.n = (int)n; h.K = sp.n_classes;
  h.n_obs = (float)observation_count; h.inv_temperature = inv_T;
  h.prior_half_inv = gauss ? 0.5f / (sp.prior_scale * sp.prior_scale) : 0.f;
  h.sumsq_parts = gauss && prior_hi > prior_lo ? w.sumsq : nullptr;
  h.potential = potential; h.variance = variance;
  launch_pdl(k_mlp_head, dim3((unsigned)C), dim3(256), 0, s, h);
  if (post_launch("k_mlp_head")) return 1;
  if (!grad) return 0;
  // ---- reverse pass: head ------------------------------------------------------------------
  launch_pdl(k_colsum_grad, dim3((unsigned)((sp.n_classes + 31) / 32), (unsigned)C), dim3(256), 0, s,
             (const float*)w.dlogits, n, sp.n_classes, theta, grad, P, (int64_t)sp.b_off[NL], coef,
             prior_lo, prior_hi);
  if (post_launch("k_colsum_grad")) return 1;
  {   // dW_head [F, K] = H_L^T [F, n] . dlogits [n, K]
    BgemmArgs g{};
    g.A = w.H[NL]; g.a_batch = n * (int64_t)F; g.lda = F;
    g.B = w.dlogits; g.b_batch = n * (int64_t)sp.n_classes; g.ldb = sp.n_classes;
    g.C = grad + sp.w_off[NL]; g.c_batch = P; g.ldc = sp.n_classes;
    g.M = F; g.N = sp.n_classes; g.K = (int)n;
    g.aux = theta + sp.w_off[NL]; g.aux_batch = P; g.ldaux = sp.n_classes;
    g.coef = coef; g.p0 = sp.w_off[NL]; g.prior_lo = prior_lo; g.prior_hi = prior_hi;
    if (launch_bgemm<true, false, kEpiPrior>(s, g, C, "k_bgemm<dW head>")) return 1;
  }
  {   // dZ_L [n, F] = (dlogits [n, K] . W^T [K, F]) * (1 - H_L^2)
    BgemmArgs g{};
    g.A = w.dlogits; g.a_batch = n * (int64_t)sp.n_classes; g.lda = sp.n_classes;
    g.B = theta + sp.w_off[NL]; g.b_batch = P; g.ldb = sp.n_classes;      // stored [F][K] = N x K
    g.C = w.dZ[NL]; g.c_batch = n * (int64_t)F; g.ldc = F;
    g.M = (int)n; g.N = F; g.K = sp.n_classes;
    g.aux = w.H[NL]; g.aux_batch = n * (int64_t)F; g.ldaux = F;
    if (launch_bgemm<false, true, kEpiDtanh>(s, g, C, "k_bgemm<dH head>")) return 1;
  }
  // ---- reverse pass: convolutions ------------------------------------------------------------
  for (int l = NL - 1; l >= 0; --l) {
    const int64_t per = L[l].rows * L[l].K;
    const int S = cnn_splits(L[l].rows);
    const int64_t chunk = (L[l].rows + S - 1) / S;
    SGMC_REQUIRE(C * (int64_t)S <= 65535, "too many chains x K splits");
    {   // db_l [Cout] = column sums of dZ over all rows, in S row chunks then in order
      launch_pdl(k_colsum_part, dim3((unsigned)((L[l].Cout + 31) / 32), (unsigned)C, (unsigned)S),
                 dim3(256), 0, s, (const float*)w.dZ[l + 1], L[l].rows, L[l].Cout, chunk, w.part);
      if (post_launch("k_colsum_part")) return 1;
      launch_pdl(k_splitk_reduce, dim3(1, (unsigned)C), dim3(256), 0, s, (const float*)w.part, S,
                 (int64_t)L[l].Cout, L[l].Cout, grad + sp.b_off[l], P, (int64_t)L[l].Cout,
                 theta + sp.b_off[l], P, coef, (int64_t)sp.b_off[l], prior_lo, prior_hi,
                 (const float*)nullptr, (int64_t)0);
      if (post_launch("k_splitk_reduce")) return 1;
    }
    {   // dW_l [9 Cin, Cout] = patches^T [9 Cin, rows] . dZ [rows, Cout]: split over the rows
      BgemmArgs g{};
      const int64_t MN = L[l].K * L[l].Cout;
      g.A = w.P[l]; g.a_batch = l == 0 ? 0 : per; g.lda = L[l].K;
      g.B = w.dZ[l + 1]; g.b_batch = L[l].rows * L[l].Cout; g.ldb = L[l].Cout;
      g.C = w.part; g.c_batch = (int64_t)S * MN; g.ldc = L[l].Cout;
      g.M = (int)L[l].K; g.N = L[l].Cout; g.K = (int)L[l].rows;
      g.n_split = S; g.k_chunk = (int)chunk; g.c_split = MN;
      if (S == 1) { g.n_split = 2; g.k_chunk = (int)L[l].rows; }      // (one split: same code path)
      const int S_eff = S == 1 ? 2 : S;
      g.c_batch = (int64_t)S_eff * MN;
      if (MN <= 1024 && (64 * (L[l].K + L[l].Cout)) * 4 <= 48 * 1024) {
        // few channels: the tall-skinny kernel instead of 128 x 128 tiles of padding
        launch_pdl(k_dw_tallskinny, dim3((unsigned)S_eff, (unsigned)C), dim3(256),
                   (size_t)(64 * (L[l].K + L[l].Cout)) * 4, s, (const float*)w.P[l],
                   l == 0 ? (int64_t)0 : per, (const float*)w.dZ[l + 1], L[l].rows, (int)L[l].K,
                   L[l].Cout, S == 1 ? L[l].rows : chunk, w.part);
        if (post_launch("k_dw_tallskinny")) return 1;
      } else if (launch_bgemm<true, false, kEpiStore>(s, g, C, "k_bgemm<dW conv, split-K>")) {
        return 1;
      }
      launch_pdl(k_splitk_reduce, dim3((unsigned)std::min<int64_t>((MN + 255) / 256, 1024), (unsigned)C),
                 dim3(256), 0, s, (const float*)w.part, S_eff, MN, L[l].Cout, grad + sp.w_off[l], P,
                 (int64_t)L[l].Cout, theta + sp.w_off[l], P, coef, (int64_t)sp.w_off[l], prior_lo,
                 prior_hi, (const float*)nullptr, (int64_t)0);
      if (post_launch("k_splitk_reduce")) return 1;
    }
    if (l > 0) {   // dPatches [rows, 9 Cin] = dZ [rows, Cout] . W^T, then col2im and tanh'
      BgemmArgs g{};
      g.A = w.dZ[l + 1]; g.a_batch = L[l].rows * L[l].Cout; g.lda = L[l].Cout;
      g.B = theta + sp.w_off[l]; g.b_batch = P; g.ldb = L[l].Cout;     // stored [9 Cin][Cout] = N x K
      g.C = w.dP; g.c_batch = per; g.ldc = L[l].K;
      g.M = (int)L[l].rows; g.N = (int)L[l].K; g.K = L[l].Cout;
      if (launch_bgemm<false, true, kEpiStore>(s, g, C, "k_bgemm<dPatches>")) return 1;
      const int64_t in_per = (int64_t)n * L[l].g.Hi * L[l].g.Wi * L[l].g.Cin;
      if (L[l].g.Cin % 4 == 0 && in_per < (1ll << 31) && per < (1ll << 31))
        launch_pdl(k_col2im_dtanh_v4, dim3(blocks(in_per / 4), (unsigned)C), dim3(256), 0, s,
                   (const float*)w.dP, (const float*)w.H[l], w.dZ[l], L[l].g, (uint32_t)(in_per / 4));
      else
        launch_pdl(k_col2im_dtanh, dim3(blocks(in_per), (unsigned)C), dim3(256), 0, s,
                   (const float*)w.dP, (const float*)w.H[l], w.dZ[l], L[l].g, in_per);
      if (post_launch("k_col2im_dtanh")) return 1;
    }
  }
  return 0;
}

size_t sgmc_mlp_workspace_bytes(const sgmc_mlp_spec* spec, int64_t n_chains, int64_t batch_size) {
  if (!spec || spec->n_layers < 1 || spec->n_layers > SGMC_MLP_MAX_LAYERS) return 0;
  return mlp_carve(nullptr, nullptr, *spec, n_chains, batch_size) + 256;
}

int sgmc_mlp_potential_grad(void* stream, const sgmc_mlp_spec* spec, const float* theta,
                            int64_t n_chains, int64_t P, const float* X, const float* y,
                            const int32_t* idx, const float* mask, int64_t batch_size,
                            int64_t observation_count, float* potential, float* variance,
                            float* grad, float* ell, void* workspace, size_t workspace_bytes) {
  if (int e = mlp_check(spec, P)) return e;
  SGMC_REQUIRE(theta && X && y && potential && workspace, "null argument");
  const int64_t C = n_chains, n = batch_size;
  SGMC_REQUIRE(C > 0 && n > 0 && n <= (1 << 24), "bad sizes");
  SGMC_REQUIRE(workspace_bytes >= sgmc_mlp_workspace_bytes(spec, C, n), "workspace too small");
  const sgmc_mlp_spec& sp = *spec;
  const int L = sp.n_layers;
  cudaStream_t s = (cudaStream_t)stream;
  MlpWs w;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) &
                                             ~(uintptr_t)255);
  mlp_carve(&w, base, sp, C, n);
  const bool gauss = sp.prior == kPriorGaussian;
  const int64_t prior_lo = gauss ? sp.prior_off : 0, prior_hi = gauss ? sp.prior_off + sp.prior_size : 0;
  const float inv_T = 1.0f / sp.temperature;
  // -(d prior / d theta) / T = theta / scale^2 / T
  const float coef = gauss ? (1.0f / (sp.prior_scale * sp.prior_scale)) / sp.temperature : 0.f;

  if (gauss && prior_hi > prior_lo) {
    launch_pdl(k_mlp_sumsq, dim3(kSumsqParts, (unsigned)C), dim3(256), 0, s, theta, P, prior_lo,
               prior_hi, w.sumsq);
    if (post_launch("k_mlp_sumsq")) return 1;
  }
  // ---- forward: h_l = tanh(h_{l-1} W_l + b_l), logits = h_{L-1} W_L + b_L -------------
  for (int l = 0; l < L; ++l) {
    BgemmArgs g{};
    const int in = sp.sizes[l], out = sp.sizes[l + 1];
    if (l == 0) { g.A = X; g.a_batch = 0; g.lda = in; g.a_idx = idx; }
    else { g.A = w.H[l]; g.a_batch = n * (int64_t)in; g.lda = in; }
    g.B = theta + sp.w_off[l]; g.b_batch = P; g.ldb = out;
    g.C = w.H[l + 1]; g.c_batch = n * (int64_t)out; g.ldc = out;
    g.M = (int)n; g.N = out; g.K = in;
    g.bias = theta + sp.b_off[l]; g.bias_batch = P;
    if (l < L - 1) {
      if (launch_bgemm<false, false, kEpiBiasTanh>(s, g, C, "k_bgemm<fwd,tanh>")) return 1;
    } else {
      if (launch_bgemm<false, false, kEpiBias>(s, g, C, "k_bgemm<fwd,logits>")) return 1;
    }
  }
  // ---- head ------------------------------------------------------------------------------
  HeadArgs h{};
  h.logits = w.H[L]; h.dlogits = grad ? w.dA[L] : nullptr; h.ell_ws = w.ell; h.ell_out = ell;
  h.y = y; h.idx = idx; h.mask = mask; h.n = (int)n; h.K = sp.sizes[L];
  h.n_obs = (float)observation_count; h.inv_temperature = inv_T;
  h.prior_half_inv = gauss ? 0.5f / (sp.prior_scale * sp.prior_scale) : 0.f;
  h.sumsq_parts = gauss && prior_hi > prior_lo ? w.sumsq : nullptr;
  h.potential = potential; h.variance = variance;
  launch_pdl(k_mlp_head, dim3((unsigned)C), dim3(256), 0, s, h);
  if (post_launch("k_mlp_head")) return 1;
  if (!grad) return 0;
  // ---- reverse pass ------------------------------------------------------------------------
  for (int l = L - 1; l >= 0; --l) {
    const int in = sp.sizes[l], out = sp.sizes[l + 1];
    launch_pdl(k_mlp_bias_grad, dim3((unsigned)((out + 255) / 256), (unsigned)C), dim3(256), 0, s,
               (const float*)w.dA[l + 1], (int)n, out, theta, grad, P, (int64_t)sp.b_off[l], coef,
               prior_lo, prior_hi);
    if (post_launch("k_mlp_bias_grad")) return 1;
    {   // dW_l [in, out] = h_{l-1}^T [in, n] . dA_l [n, out]  (+ prior term) -> grad
      BgemmArgs g{};
      if (l == 0) { g.A = X; g.a_batch = 0; g.lda = in; g.a_idx = idx; }
      else { g.A = w.H[l]; g.a_batch = n * (int64_t)in; g.lda = in; }
      g.B = w.dA[l + 1]; g.b_batch = n * (int64_t)out; g.ldb = out;
      g.C = grad + sp.w_off[l]; g.c_batch = P; g.ldc = out;
      g.M = in; g.N = out; g.K = (int)n;
      g.aux = theta + sp.w_off[l]; g.aux_batch = P; g.ldaux = out;
      g.coef = coef; g.p0 = sp.w_off[l]; g.prior_lo = prior_lo; g.prior_hi = prior_hi;
      if (launch_bgemm<true, false, kEpiPrior>(s, g, C, "k_bgemm<dW>")) return 1;
    }
    if (l > 0) {   // dA_{l-1} [n, in] = (dA_l [n, out] . W_l^T [out, in]) * (1 - h_{l-1}^2)
      BgemmArgs g{};
      g.A = w.dA[l + 1]; g.a_batch = n * (int64_t)out; g.lda = out;
      g.B = theta + sp.w_off[l]; g.b_batch = P; g.ldb = out;      // stored [in][out] = N x K
      g.C = w.dA[l]; g.c_batch = n * (int64_t)in; g.ldc = in;
      g.M = (int)n; g.N = in; g.K = out;
      g.aux = w.H[l]; g.aux_batch = n * (int64_t)in; g.ldaux = in;
      if (launch_bgemm<false, true, kEpiDtanh>(s, g, C, "k_bgemm<dH>")) return 1;
    }
  }
  return 0;
}

}  // extern "C"
