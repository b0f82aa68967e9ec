// Generic "one elementwise pass with in-kernel jax.random noise" skeleton.
//
// Every fused integrator update (SGLD, pSGLD, SGHMC, OBABO) is an elementwise
// map over the chain-batched flat sample f32[C][P] that consumes one standard
// normal per element, drawn exactly as integrator.random_tree does
// (jax_sgmc/integrator.py:119-135): per chain `splits = split(key, n_leaves)`,
// leaf l gets normal(splits[l], leaf.shape).  In the original threefry layout
// element j of a leaf shares its threefry block with element j + ceil(n/2), so
// a thread owns a *group* of 4 such pairs = 8 elements = two float4 vectors
// (coalesced 128-bit loads/stores on both halves of the leaf) and spends
// exactly one threefry block per two normals.
//
// Work decomposition: a warp-tile is 32 groups of one chain (tiles never cross
// a chain, so per-chain reductions are plain warp reductions).  Tiles are dealt
// to a persistent grid (a multiple of the SM count) as contiguous ranges; each
// CTA first derives the per-(chain, leaf) noise keys for the chains in its
// range into shared memory (lanes work on different chains in parallel), then
// its warps stream the tiles.  HBM-bound: no shared-memory staging of the data
// (no reuse), 128-bit accesses, loads issued ahead of the ALU-heavy noise
// generation.
#pragma once
#include "common.cuh"
#include "rng.cuh"

#include <type_traits>

namespace sgmc {

constexpr int kNoiseThreads = 512;
constexpr int kNoiseWarps = kNoiseThreads / 32;

struct LeafTable {
  int32_t n_leaves;
  uint32_t P;                       // elements per chain
  uint32_t groups;                  // groups per chain
  uint32_t tiles_per_chain;         // ceil(groups / 32)
  uint32_t off[SGMC_MAX_LEAVES];    // leaf offset in the flat sample
  uint32_t size[SGMC_MAX_LEAVES];   // leaf size
  uint32_t gstart[SGMC_MAX_LEAVES + 1];
  uint8_t vec_ok[SGMC_MAX_LEAVES];  // float4 path legal for this leaf
  // The leaf could take the float4 path but its first element is not 16-byte aligned in every
  // chain (P or the leaf offset is not a multiple of four: the 669 706-parameter classifier of
  // C3).  Such a leaf gets ONE EXTRA group: group 0 covers the first s = (4 - misalignment) & 3
  // pairs of the chain's copy one by one, groups k >= 1 the pairs [s + 4(k-1), s + 4k), which
  // start 16-byte aligned in that chain.  The noise belongs to the pair index, not to the
  // group, so the draws are unchanged.
  uint8_t shifted[SGMC_MAX_LEAVES];
};

// How the noise key of this pass derives from the chain key (see the op
// descriptions in include/sgmc_b200.h).
enum KeyMode : int {
  kKeyDirect = 0,   // noise key = key (random_tree(key, ...)); nothing written
  kKeySplit2 = 1,   // key', sub = split(key); noise = sub; keys_out = key'
  kKeySplit3A = 2,  // key', s1, s2 = split(key,3); noise = s1; keys_out = key'
  kKeySplit3B = 3,  // noise = s2; nothing written
  kKeyCached = 4,   // keys_in = u32[C][L][2] per-(chain, leaf) NOISE keys derived by an
                    // earlier kernel of the step (k_prepare_all); nothing written
};

int build_leaf_table(LeafTable* t, const int64_t* leaf_sizes, int n_leaves,
                     int64_t n_chains, bool ptrs_aligned16);

struct NoiseLaunch {
  int grid;
  int threads;              // CTA size: a multiple of 32 in [kNoiseThreads / 2, kNoiseThreads]
  size_t smem;
  int64_t tiles_total;
  int max_chains_per_cta;
};
int plan_noise_launch(const LeafTable& t, int64_t n_chains, const void* kernel,
                      NoiseLaunch* out);

// Noise for one group: pairs j0..j0+3 of a leaf with `size` elements.
// nA[q] belongs to element j0+q (valid if j0+q < half), nB[q] to element
// half + j0 + q (valid if that is < size).
template <int LAYOUT, bool FULL = false>
__device__ __forceinline__ void pair_bits(Key lk, uint32_t j, uint32_t half,
                                          uint32_t size, uint32_t& wa,
                                          uint32_t& wb) {
  if (LAYOUT == 0) {
    wa = j;
    wb = (FULL || j + half < size) ? j + half : 0u;   // FULL: partner known valid
    threefry2x32(lk, wa, wb);
  } else {
    uint32_t a0 = 0u, a1 = j;
    threefry2x32(lk, a0, a1);
    wa = a0 ^ a1;
    uint32_t b0 = 0u, b1 = j + half;
    threefry2x32(lk, b0, b1);
    wb = b0 ^ b1;
  }
}

template <int LAYOUT, bool FULL = false>
__device__ __forceinline__ void group_noise(Key lk, uint32_t j0, uint32_t half,
                                            uint32_t size, float nA[4],
                                            float nB[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    NormalPartial pa, pb;
    uint32_t wa, wb;
    pair_bits<LAYOUT, FULL>(lk, j0 + q, half, size, wa, wb);
    nA[q] = normal_main(wa, pa);
    nB[q] = normal_main(wb, pb);
    // 0.34 % of the draws need erf_inv's w >= 5 polynomial
    if (normal_is_tail(pa) | normal_is_tail(pb)) {
      if (normal_is_tail(pa)) nA[q] = normal_tail(pa);
      if (normal_is_tail(pb)) nB[q] = normal_tail(pb);
    }
  }
}

// Op concept:
//   struct Op {
//     struct Regs;                                    // staged inputs
//     void load_vec(Regs&, int64_t iA, int64_t iB) const;      // float4 x2
//     float apply_vec(Regs&, const float nA[4], const float nB[4],
//                     int64_t iA, int64_t iB, int64_t chain, uint32_t eA,
//                     uint32_t eB) const;              // returns a partial sum
//     float apply_one(int64_t i, float noise, int64_t chain, uint32_t e) const;
//     void reduce(int64_t chain, uint32_t tile_in_chain, uint32_t tiles_per_chain,
//                 float warp_sum) const;                        // lane 0 only, if kReduce
//     static constexpr bool kReduce;
//     static constexpr bool kSplit;   // optional: apply_vec also returns |max| and
//                                     // reduce2(chain, tile_in_chain, sum, max) is called
//   };
// eA/eB/e are element offsets inside the chain (for per-parameter vectors such
// as mass or friction).
template <class Op, class = void>
struct OpSplits { static constexpr bool value = false; };
template <class Op>
struct OpSplits<Op, std::enable_if_t<Op::kSplit>> { static constexpr bool value = true; };

template <int LAYOUT, class Op>
__global__ void __launch_bounds__(kNoiseThreads, 2)
k_noise_pass(const __grid_constant__ LeafTable tab,
             const uint32_t* __restrict__ keys_in,
             uint32_t* __restrict__ keys_out, int64_t n_chains,
             int64_t tiles_total, int key_mode, const Op op) {
  extern __shared__ Key s_keys[];
  pdl_launch_dependents();   // the next kernel may start ramping up behind this grid
  pdl_wait();                // ... and this one waits here for the grids before it
  const int L = tab.n_leaves;
  const int64_t t0 = tiles_total * (int64_t)blockIdx.x / gridDim.x;
  const int64_t t1 = tiles_total * ((int64_t)blockIdx.x + 1) / gridDim.x;
  if (t0 >= t1) return;
  const int64_t c_lo = t0 / tab.tiles_per_chain;
  const int64_t c_hi = (t1 - 1) / tab.tiles_per_chain;
  const int n_ch = (int)(c_hi - c_lo + 1);

  // ---- prologue: per-(chain, leaf) noise keys --------------------------
  const int n_threads = blockDim.x, n_warps = n_threads >> 5;
  if (key_mode == kKeyCached) {
    // the noise keys were derived by an earlier kernel of the step: one coalesced load
    const uint2* ck = reinterpret_cast<const uint2*>(keys_in) + c_lo * L;
    for (int idx = threadIdx.x; idx < n_ch * L; idx += n_threads) {
      const uint2 k = ck[idx];
      s_keys[idx] = Key{k.x, k.y};
    }
  } else
  for (int idx = threadIdx.x; idx < n_ch * L; idx += n_threads) {
    const int ci = idx / L, l = idx - ci * L;
    const int64_t c = c_lo + ci;
    Key k;
    k.k0 = keys_in[2 * c];
    k.k1 = keys_in[2 * c + 1];
    Key nk = k, newk = k;
    bool write = false;
    if (key_mode == kKeySplit2) {
      split2(k, LAYOUT, newk, nk);
      write = true;
    } else if (key_mode == kKeySplit3A || key_mode == kKeySplit3B) {
      Key s1, s2;
      split3(k, LAYOUT, newk, s1, s2);
      nk = (key_mode == kKeySplit3A) ? s1 : s2;
      write = (key_mode == kKeySplit3A);
    }
    s_keys[idx] = key_mode == 99 ? k /* debug: no derivation */
                                 : split_key(nk, (uint32_t)l, (uint32_t)L, LAYOUT);
    // the CTA that owns the chain's first tile publishes the new chain key
    if (write && l == 0 && c * tab.tiles_per_chain >= t0) {
      keys_out[2 * c] = newk.k0;
      keys_out[2 * c + 1] = newk.k1;
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rotate the first warp per CTA so the odd tiles of the CTAs spread evenly
  // over the four SM sub-partitions (warp w runs on sub-partition w % 4)
  const int wrot = (warp + blockIdx.x) % n_warps;
  // (chain, tile-in-chain) advance incrementally: no division in the loop
  const uint32_t tpc = tab.tiles_per_chain;
  int64_t c = (t0 + wrot) / tpc;
  uint32_t lt = (uint32_t)((t0 + wrot) - c * tpc);
  for (int64_t tile = t0 + wrot; tile < t1; tile += n_warps, lt += n_warps) {
    while (lt >= tpc) { lt -= tpc; ++c; }
    const uint32_t g = lt * 32u + lane;
    float partial = 0.0f;
    float amax = 0.0f;
    if (g < tab.groups) {
      int l = 0;
      while (l + 1 < L && g >= tab.gstart[l + 1]) ++l;
      const uint32_t size = tab.size[l];
      const uint32_t half = (size + 1u) >> 1;
      uint32_t j0 = (g - tab.gstart[l]) * 4u, cnt = 4u;
      const Key lk = s_keys[(int)(c - c_lo) * L + l];
      const int64_t base = c * (int64_t)tab.P + tab.off[l];
      if (tab.shifted[l]) {
        const uint32_t s = (4u - ((uint32_t)base & 3u)) & 3u;
        if (j0 == 0u) cnt = s;
        else j0 = j0 - 4u + s;
      }
      const uint32_t eA = tab.off[l] + j0, eB = eA + half;
      float nA[4], nB[4];
      if (tab.vec_ok[l] && cnt == 4u && j0 + 4u <= half) {
        typename Op::Regs r;
        op.load_vec(r, base + j0, base + half + j0);   // loads first ...
        group_noise<LAYOUT, true>(lk, j0, half, size, nA, nB);  // ... then ALU work
        if constexpr (OpSplits<Op>::value)
          partial = op.apply_vec2(r, nA, nB, base + j0, base + half + j0, c, eA, eB, amax);
        else
          partial = op.apply_vec(r, nA, nB, base + j0, base + half + j0, c, eA, eB);
      } else {
        group_noise<LAYOUT>(lk, j0, half, size, nA, nB);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if ((uint32_t)q < cnt && j0 + q < half) {
            partial += op.apply_one(base + j0 + q, nA[q], c, eA + q);
            if (j0 + q + half < size)
              partial += op.apply_one(base + half + j0 + q, nB[q], c, eB + q);
          }
        }
      }
    }
    if constexpr (OpSplits<Op>::value) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        partial += __shfl_xor_sync(0xffffffffu, partial, s);
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, s));
      }
      if (lane == 0) op.reduce2(c, lt, partial, amax);
    } else if constexpr (Op::kReduce) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1)
        partial += __shfl_xor_sync(0xffffffffu, partial, s);
      if (lane == 0) op.reduce(c, lt, tpc, partial);
    }
  }
}

template <class Op>
int launch_noise_pass(cudaStream_t stream, const LeafTable& tab,
                      const uint32_t* keys_in, uint32_t* keys_out,
                      int64_t n_chains, int key_mode, int layout, const Op& op,
                      const char* name) {
  NoiseLaunch nl;
  const void* fn = layout == 0 ? (const void*)k_noise_pass<0, Op>
                               : (const void*)k_noise_pass<1, Op>;
  if (plan_noise_launch(tab, n_chains, fn, &nl)) return 1;
  if (nl.tiles_total == 0) return 0;
  if (layout == 0)
    launch_pdl(k_noise_pass<0, Op>, dim3(nl.grid), dim3(nl.threads), nl.smem, stream, tab,
               keys_in, keys_out, n_chains, nl.tiles_total, key_mode, op);
  else
    launch_pdl(k_noise_pass<1, Op>, dim3(nl.grid), dim3(nl.threads), nl.smem, stream, tab,
               keys_in, keys_out, n_chains, nl.tiles_total, key_mode, op);
  return post_launch(name);
}

}  // namespace sgmc
