// Stand-alone jax.random kernels (split, bits, uniform, normal, randint) and
// the DeviceNumpyDataLoader minibatch draw (jax_sgmc/data/numpy_loader.py:
// 128-141) + row gather (jax_sgmc/data/core.py:642-660).
//
// These are the parity surface for the PRNG (bit-exact against oracle/prng.py
// and the public jax.random vectors); the hot path itself generates its noise
// inside the fused update kernels (update_kernels.cu).
#include "common.cuh"
#include "rng.cuh"

namespace sgmc {

__global__ void k_split(const uint32_t* __restrict__ keys_in,
                        uint32_t* __restrict__ keys_out, int64_t n_keys,
                        int num, int layout) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys * num) return;
  const int64_t c = i / num;
  const uint32_t j = (uint32_t)(i - c * num);
  Key k{keys_in[2 * c], keys_in[2 * c + 1]};
  const Key o = split_key(k, j, (uint32_t)num, layout);
  keys_out[2 * i] = o.k0;
  keys_out[2 * i + 1] = o.k1;
}

enum { kBits = 0, kUniform = 1, kNormal = 2 };

// one thread per threefry block of one key (two outputs in the original layout)
template <int KIND>
__global__ void k_random(const uint32_t* __restrict__ keys, void* __restrict__ out,
                         int64_t n_keys, int64_t n, float lo, float scale,
                         int layout) {
  const int64_t half = layout == 0 ? (n + 1) / 2 : n;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys * half) return;
  const int64_t c = i / half, j = i - c * half;
  Key k{keys[2 * c], keys[2 * c + 1]};
  uint32_t w[2];
  int64_t pos[2];
  int cnt;
  if (layout == 0) {
    uint32_t x0 = (uint32_t)j, x1 = (j + half < n) ? (uint32_t)(j + half) : 0u;
    threefry2x32(k, x0, x1);
    w[0] = x0; w[1] = x1;
    pos[0] = j; pos[1] = j + half;
    cnt = (j + half < n) ? 2 : 1;
  } else {
    w[0] = random_word(k, (uint64_t)j, (uint64_t)n, 1);
    pos[0] = j;
    cnt = 1;
  }
  for (int q = 0; q < cnt; ++q) {
    const int64_t o = c * n + pos[q];
    if (KIND == kBits) ((uint32_t*)out)[o] = w[q];
    if (KIND == kUniform) ((float*)out)[o] = bits_to_uniform(w[q], lo, scale);
    if (KIND == kNormal) ((float*)out)[o] = bits_to_normal(w[q]);
  }
}

// jax._src.random._randint for int32: offset = ((hi % span) * mult + lo % span)
// % span with mult = (2^16 % span)^2 % span (uint32 wrap-around arithmetic).
__device__ __forceinline__ int32_t randint_one(Key k1, Key k2, uint64_t i,
                                               uint64_t n, int32_t minval,
                                               uint32_t span, uint32_t mult,
                                               int layout) {
  const uint32_t hi = random_word(k1, i, n, layout);
  const uint32_t lo = random_word(k2, i, n, layout);
  uint32_t off = (hi % span) * mult + (lo % span);
  off %= span;
  return (int32_t)((uint32_t)minval + off);
}

__global__ void k_randint(const uint32_t* __restrict__ key_in,
                          uint32_t* __restrict__ key_out,
                          int32_t* __restrict__ out, int64_t n, int32_t minval,
                          uint32_t span, uint32_t mult, int draw_mode,
                          int layout) {
  // draw_mode 0: randint(key);  1: key', sub = split(key); randint(sub)
  pdl_launch_dependents();
  pdl_wait();
  Key k{key_in[0], key_in[1]};
  if (draw_mode == 1) {
    Key nk, sub;
    split2(k, layout, nk, sub);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      key_out[0] = nk.k0;
      key_out[1] = nk.k1;
    }
    k = sub;
  }
  Key k1, k2;
  split2(k, layout, k1, k2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = randint_one(k1, k2, (uint64_t)i, (uint64_t)n, minval, span, mult,
                         layout);
}

__global__ void k_gather_rows(const float* __restrict__ src,
                              const int32_t* __restrict__ idx,
                              float* __restrict__ out, int64_t n,
                              int64_t row_elems) {
  // one warp per row chunk; float4 when the row is 16B-tileable
  const int64_t row = blockIdx.x;
  if (row >= n) return;
  const float* s = src + (int64_t)idx[row] * row_elems;
  float* o = out + row * row_elems;
  if ((row_elems & 3) == 0 &&
      ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* o4 = reinterpret_cast<float4*>(o);
    for (int64_t j = threadIdx.x; j < row_elems / 4; j += blockDim.x)
      o4[j] = __ldg(s4 + j);
  } else {
    for (int64_t j = threadIdx.x; j < row_elems; j += blockDim.x) o[j] = s[j];
  }
}

static int randint_params(int32_t minval, int32_t maxval, uint32_t* span,
                          uint32_t* mult) {
  uint32_t s = (uint32_t)((int64_t)maxval - (int64_t)minval);
  if (maxval <= minval) s = 1u;
  uint32_t m = (1u << 16) % s;
  m = (uint32_t)(((uint64_t)m * m) & 0xFFFFFFFFull) % s;   // uint32 wrap as in XLA
  *span = s;
  *mult = m;
  return 0;
}

}  // namespace sgmc

using namespace sgmc;

template <int KIND>
static int random_common(void* stream, const uint32_t* keys, void* out,
                         int64_t n_keys, int64_t n, float lo, float scale,
                         int layout, const char* name) {
  SGMC_REQUIRE(n_keys >= 0 && n >= 0, "bad sizes");
  const int64_t half = layout == 0 ? (n + 1) / 2 : n;
  const int64_t total = n_keys * half;
  if (total == 0) return 0;
  k_random<KIND><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      keys, out, n_keys, n, lo, scale, layout);
  return post_launch(name);
}

extern "C" {

int sgmc_prng_split(void* stream, const uint32_t* keys_in, uint32_t* keys_out,
                    int64_t n_keys, int num, int prng_layout) {
  SGMC_REQUIRE(n_keys >= 0 && num >= 1, "bad split sizes");
  const int64_t total = n_keys * num;
  if (total == 0) return 0;
  k_split<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      keys_in, keys_out, n_keys, num, prng_layout);
  return post_launch("sgmc_prng_split");
}

int sgmc_random_bits(void* stream, const uint32_t* keys, uint32_t* out,
                     int64_t n_keys, int64_t n, int prng_layout) {
  return random_common<kBits>(stream, keys, out, n_keys, n, 0.f, 1.f,
                              prng_layout, "sgmc_random_bits");
}
int sgmc_uniform(void* stream, const uint32_t* keys, float* out, int64_t n_keys,
                 int64_t n, float minval, float maxval, int prng_layout) {
  return random_common<kUniform>(stream, keys, out, n_keys, n, minval,
                                 maxval - minval, prng_layout, "sgmc_uniform");
}
int sgmc_normal(void* stream, const uint32_t* keys, float* out, int64_t n_keys,
                int64_t n, int prng_layout) {
  return random_common<kNormal>(stream, keys, out, n_keys, n, 0.f, 1.f,
                                prng_layout, "sgmc_normal");
}

int sgmc_randint(void* stream, const uint32_t* key, int32_t* out, int64_t n,
                 int32_t minval, int32_t maxval, int prng_layout) {
  if (n == 0) return 0;
  uint32_t span, mult;
  randint_params(minval, maxval, &span, &mult);
  const unsigned grid = (unsigned)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256);
  k_randint<<<grid, 256, 0, (cudaStream_t)stream>>>(key, nullptr, out, n, minval,
                                                    span, mult, 0, prng_layout);
  return post_launch("sgmc_randint");
}

int sgmc_minibatch_draw(void* stream, const uint32_t* key_in, uint32_t* key_out,
                        int32_t* idx, int64_t batch_size,
                        int64_t observation_count, int prng_layout) {
  SGMC_REQUIRE(key_in != key_out, "key_out must not alias key_in");
  SGMC_REQUIRE(batch_size > 0 && observation_count > 0 &&
               observation_count < (1ll << 31), "bad minibatch sizes");
  uint32_t span, mult;
  randint_params(0, (int32_t)observation_count, &span, &mult);
  const unsigned grid =
      (unsigned)((batch_size + 255) / 256 > 1184 ? 1184 : (batch_size + 255) / 256);
  launch_pdl(k_randint, dim3(grid), dim3(256), 0, (cudaStream_t)stream, key_in, key_out, idx,
             batch_size, (int32_t)0, span, mult, 1, prng_layout);
  return post_launch("sgmc_minibatch_draw");
}

int sgmc_gather_rows(void* stream, const float* src, const int32_t* idx,
                     float* out, int64_t n, int64_t row_elems) {
  if (n == 0 || row_elems == 0) return 0;
  k_gather_rows<<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(src, idx, out, n,
                                                              row_elems);
  return post_launch("sgmc_gather_rows");
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Synthetic Bayesian-logistic-regression data set generated in HBM
// (SURVEY.md section 8d, config C2): kx, kw, ky = split(key, 3);
// X = normal(kx, (N, d)) / sqrt(d); w = normal(kw, (d,));
// y_i = uniform(ky)_i < sigmoid(x_i . w).  Benchmark / test support: the data
// never exists on the host.
namespace sgmc {

__global__ void k_synth_x(Key kx, float* __restrict__ X, uint64_t total,
                          float inv_sqrt_d, int layout) {
  const uint64_t half = layout == 0 ? (total + 1) / 2 : total;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < half;
       j += (uint64_t)gridDim.x * blockDim.x) {
    if (layout == 0) {
      uint32_t x0 = (uint32_t)j, x1 = (j + half < total) ? (uint32_t)(j + half) : 0u;
      threefry2x32(kx, x0, x1);
      X[j] = __fmul_rn(bits_to_normal(x0), inv_sqrt_d);
      if (j + half < total) X[j + half] = __fmul_rn(bits_to_normal(x1), inv_sqrt_d);
    } else {
      X[j] = __fmul_rn(bits_to_normal(random_word(kx, j, total, 1)), inv_sqrt_d);
    }
  }
}

__global__ void k_synth_y(Key ky, const float* __restrict__ X,
                          const float* __restrict__ w, float* __restrict__ y,
                          int64_t N, int d, int layout) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= N) return;
  float acc = 0.f;
  for (int j = lane; j < d; j += 32) acc = fmaf(X[row * d + j], w[j], acc);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    const float u = bits_to_uniform(random_word(ky, (uint64_t)row, (uint64_t)N, layout),
                                    0.0f, 1.0f);
    const float p = 1.0f / (1.0f + expf(-acc));
    y[row] = u < p ? 1.0f : 0.0f;
  }
}

}  // namespace sgmc

extern "C" int sgmc_synth_logistic_data(void* stream, const uint32_t* key_host,
                                        float* X, float* y, float* w,
                                        int64_t N, int64_t d, int prng_layout) {
  using namespace sgmc;
  SGMC_REQUIRE(N > 0 && d > 0 && N * d < (1ll << 32), "synthetic set too large");
  // key schedule on the host side of the launch (3-way split) is done by a
  // one-thread kernel-free path: derive on device through k_split semantics.
  uint32_t* dkeys = nullptr;
  if (check_cuda(cudaMalloc(&dkeys, 8 * sizeof(uint32_t)), "cudaMalloc")) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemcpyAsync(dkeys, key_host, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s);
  k_split<<<1, 32, 0, s>>>(dkeys, dkeys + 2, 1, 3, prng_layout);
  if (post_launch("k_split")) return 1;
  uint32_t hk[6];
  cudaMemcpyAsync(hk, dkeys + 2, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
  if (check_cuda(cudaStreamSynchronize(s), "sync")) return 1;
  cudaFree(dkeys);
  const Key kx{hk[0], hk[1]}, kw{hk[2], hk[3]}, ky{hk[4], hk[5]};
  const float inv_sqrt_d = 1.0f / sqrtf((float)d);
  k_synth_x<<<sm_count() * 8, 256, 0, s>>>(kx, X, (uint64_t)(N * d), inv_sqrt_d,
                                          prng_layout);
  if (post_launch("k_synth_x")) return 1;
  k_synth_x<<<1, 256, 0, s>>>(kw, w, (uint64_t)d, 1.0f, prng_layout);
  if (post_launch("k_synth_w")) return 1;
  k_synth_y<<<(unsigned)((N * 32 + 255) / 256), 256, 0, s>>>(ky, X, w, y, N, (int)d,
                                                            prng_layout);
  return post_launch("k_synth_y");
}
