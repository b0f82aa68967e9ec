// Fused integrator / preconditioner updates with in-kernel jax.random noise.
//
//   sgmc_sgld_update, sgmc_sgld_rms_update   integrator.py:860-922,
//                                            adaption.py:254-291
//   sgmc_sghmc_begin, sgmc_sghmc_step        integrator.py:603-666, :716-757
//   sgmc_obabo_pass_a, sgmc_obabo_pass_b     integrator.py:177-273
//   sgmc_normal_like                         integrator.py:119-135
//
// All are instances of k_noise_pass (noise_pass.cuh).  The elementwise
// arithmetic follows SURVEY.md Appendix A operation by operation with explicit
// round-to-nearest intrinsics (no FMA contraction), so that given the same
// gradient the result equals the NumPy oracle bit for bit.
// Roofline: HBM.  Algorithmic bytes per parameter: SGLD 12, pSGLD 20,
// SGHMC step 20, OBABO pass A 20 + pass B 12.
#include "noise_pass.cuh"
#include "sgld_math.cuh"
#include "sgld_apply_tile.cuh"
#include "sgld_split.cuh"
#include "tc_ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cmath>
#include <map>
#include <mutex>
#include <utility>

namespace sgmc {

int build_leaf_table(LeafTable* t, const int64_t* leaf_sizes, int n_leaves,
                     int64_t n_chains, bool ptrs_aligned16) {
  SGMC_REQUIRE(n_leaves >= 1 && n_leaves <= SGMC_MAX_LEAVES,
               "n_leaves=%d outside [1,%d]", n_leaves, SGMC_MAX_LEAVES);
  SGMC_REQUIRE(n_chains >= 0, "n_chains < 0");
  int64_t off = 0, g = 0;
  for (int l = 0; l < n_leaves; ++l) {
    SGMC_REQUIRE(leaf_sizes[l] >= 0 && leaf_sizes[l] < (1ll << 31),
                 "leaf %d size %lld unsupported", l, (long long)leaf_sizes[l]);
    off += leaf_sizes[l];
  }
  SGMC_REQUIRE(off < (1ll << 31), "flat sample too large (%lld)", (long long)off);
  const int64_t P = off;
  off = 0;
  for (int l = 0; l < n_leaves; ++l) {
    const int64_t size = leaf_sizes[l], half = (size + 1) / 2;
    t->off[l] = (uint32_t)off;
    t->size[l] = (uint32_t)size;
    t->gstart[l] = (uint32_t)g;
    // float4 path: both halves of the leaf start 16-byte aligned -- in every chain when P and
    // the offset are multiples of four, else after a per-chain shift (see LeafTable::shifted)
    const bool capable = ptrs_aligned16 && half % 4 == 0 && size % 2 == 0 && size > 0;
    const bool aligned_everywhere = P % 4 == 0 && off % 4 == 0;
    t->vec_ok[l] = capable ? 1 : 0;
    t->shifted[l] = (capable && !aligned_everywhere) ? 1 : 0;
    g += (half + 3) / 4 + t->shifted[l];
    off += size;
  }
  t->gstart[n_leaves] = (uint32_t)g;
  t->n_leaves = n_leaves;
  t->P = (uint32_t)P;
  t->groups = (uint32_t)g;
  t->tiles_per_chain = (uint32_t)((g + 31) / 32);
  return 0;
}

int plan_noise_launch(const LeafTable& t, int64_t n_chains, const void* kernel,
                      NoiseLaunch* out) {
  out->tiles_total = n_chains * (int64_t)t.tiles_per_chain;
  out->grid = 0;
  out->smem = 0;
  if (out->tiles_total == 0) return 0;
  // persistent grid: a multiple of the SM count, bounded by the work
  const int sms = sm_count();
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kNoiseThreads,
                                                    2048) != cudaSuccess ||
      per_sm < 1) {
    cudaGetLastError();
    per_sm = 4;
  }
  // one resident wave exactly: every CTA is live from the start, tiles are
  // dealt evenly, nothing is left for a ragged second wave
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > out->tiles_total) grid = out->tiles_total;
  for (;;) {
    const int64_t tiles_per_cta = (out->tiles_total + grid - 1) / grid;
    const int64_t chains = tiles_per_cta / t.tiles_per_chain + 2;
    const size_t smem = (size_t)chains * t.n_leaves * sizeof(Key);
    if (smem <= 40 * 1024 || grid >= out->tiles_total) {
      SGMC_REQUIRE(smem <= 40 * 1024, "key cache too large (%zu B)", smem);
      out->smem = smem;
      out->max_chains_per_cta = (int)chains;
      break;
    }
    grid = grid * 2 > out->tiles_total ? out->tiles_total : grid * 2;
  }
  out->grid = (int)grid;
  // Warps per CTA: a warp walks whole tiles, so with only a few tiles per warp the
  // slowest warp sets the pace (C2: 55.4 tiles per CTA over 16 warps = 3.46 -> 4
  // rounds, 86 % busy; over 14 warps = 3.95 -> 4 rounds, 99 % busy).  Pick the
  // block size that wastes the fewest tile slots (ties: more warps).
  const int64_t tiles_per_cta = (out->tiles_total + grid - 1) / grid;
  int best_w = kNoiseWarps;
  double best_eff = -1.0;
  for (int w = kNoiseWarps; w >= kNoiseWarps / 2; --w) {
    const int64_t rounds = (tiles_per_cta + w - 1) / w;
    const double eff = (double)tiles_per_cta / (double)(rounds * w);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best_w = w;
    }
  }
  out->threads = best_w * 32;
  return 0;
}

// ---------------------------------------------------------------------------
#define SGMC_F4_MAP(dst, expr)                          \
  {                                                     \
    float4 _o;                                          \
    { const int k = 0; _o.x = (expr); }                 \
    { const int k = 1; _o.y = (expr); }                 \
    { const int k = 2; _o.z = (expr); }                 \
    { const int k = 3; _o.w = (expr); }                 \
    dst = _o;                                           \
  }
__device__ __forceinline__ float f4get(const float4& v, int k) {
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

// ---- random_tree only -----------------------------------------------------
struct NormalLikeOp {
  float* out;
  static constexpr bool kReduce = false;
  struct Regs {};
  __device__ void load_vec(Regs&, int64_t, int64_t) const {}
  __device__ float apply_vec(Regs&, const float nA[4], const float nB[4],
                             int64_t iA, int64_t iB, int64_t, uint32_t,
                             uint32_t) const {
    st4(out, iA, make_float4(nA[0], nA[1], nA[2], nA[3]));
    st4(out, iB, make_float4(nB[0], nB[1], nB[2], nB[3]));
    return 0.f;
  }
  __device__ float apply_one(int64_t i, float n, int64_t, uint32_t) const {
    out[i] = n;
    return 0.f;
  }
};

// ---- SGLD / pSGLD -----------------------------------------------------------
// integrator.py:882-912; adaption.py:270-272, :289-291.
// FAST = false: IEEE sqrt / reciprocal and unfused mul/add, bit-identical to the
// oracle given the same gradient.  FAST = true (default at run time, see
// sgmc_set_option): MUFU rsqrt / rcp / sqrt approximations (<= 2 ulp) and FMA
// contraction in the preconditioner arithmetic only -- the noise stays
// bit-exact; trajectories stay within the 1e-5 parity tolerance.
template <bool RMS, bool FAST>
struct SgldOp {
  float* theta;
  float* v;
  const float* grad;
  const float* temp_per_chain;
  float eps, neg_eps, noise_scale, alpha, one_m_alpha, lmbd;
  static constexpr bool kReduce = false;
  struct Regs {
    float4 tA, tB, gA, gB, vA, vB;
  };
  __device__ __forceinline__ float scale_for(int64_t c) const {
    if (temp_per_chain == nullptr) return noise_scale;
    return __fsqrt_rn(__fmul_rn(__fmul_rn(2.0f, temp_per_chain[c]), eps));
  }
  __device__ __forceinline__ float one(float t, float g, float& vv, float xi,
                                       float ns) const {
    return sgld_one<RMS, FAST>(t, g, vv, xi, ns, neg_eps, alpha, one_m_alpha, lmbd);
  }
  __device__ void load_vec(Regs& r, int64_t iA, int64_t iB) const {
    r.tA = ld4(theta, iA);
    r.tB = ld4(theta, iB);
    r.gA = ld4(grad, iA);
    r.gB = ld4(grad, iB);
    if (RMS) {
      r.vA = ld4(v, iA);
      r.vB = ld4(v, iB);
    }
  }
  __device__ float apply_vec(Regs& r, const float nA[4], const float nB[4],
                             int64_t iA, int64_t iB, int64_t c, uint32_t,
                             uint32_t) const {
    const float ns = scale_for(c);
    float va[4] = {r.vA.x, r.vA.y, r.vA.z, r.vA.w};
    float vb[4] = {r.vB.x, r.vB.y, r.vB.z, r.vB.w};
    float4 oA, oB;
    SGMC_F4_MAP(oA, one(f4get(r.tA, k), f4get(r.gA, k), va[k], nA[k], ns));
    SGMC_F4_MAP(oB, one(f4get(r.tB, k), f4get(r.gB, k), vb[k], nB[k], ns));
    st4(theta, iA, oA);
    st4(theta, iB, oB);
    if (RMS) {
      st4(v, iA, make_float4(va[0], va[1], va[2], va[3]));
      st4(v, iB, make_float4(vb[0], vb[1], vb[2], vb[3]));
    }
    return 0.f;
  }
  __device__ float apply_one(int64_t i, float n, int64_t c, uint32_t) const {
    float vv = RMS ? v[i] : 0.f;
    theta[i] = one(theta[i], grad[i], vv, n, scale_for(c));
    if (RMS) v[i] = vv;
    return 0.f;
  }
};

// ---- SGLD / pSGLD that also emits the tensor-core operand of the NEXT potential ----
// sgmc_glm_sgld_step in carried mode: the GLM potential's GEMM1 consumes Theta as
// fp16 hi/lo (or bf16) K-major rows.  The update has theta' in registers, so it
// writes that split itself (scale[c] = a power of two chosen from the row's
// previous |max|; k_prepare_all validates it against the new |max| and only
// re-splits rows that left the safe window) together with the row statistics the
// potential needs: |max| (atomicMax on the bit pattern) and the per-warp-tile
// sum of squares over the gaussian-prior range (combined in fixed order by
// k_prepare_all, so the potential stays run-to-run deterministic).  This removes
// the Theta pass of k_prepare_all (16.8 MB read + 16.8 MB written at C2) from
// every step.
template <bool RMS, bool FAST, int FMT>
struct SgldSplitOp : SgldOp<RMS, FAST> {
  using Base = SgldOp<RMS, FAST>;
  using Regs = typename Base::Regs;
  static constexpr bool kSplit = true;
  void* th_hi;                 // fp16 (FMT 1) or bf16 (FMT 2) [C][P]
  __half* th_lo;               // fp16 [C][P] (FMT 1)
  const float* scale;          // f32[C]
  uint32_t* amax_bits;         // u32[C]
  float* sumsq_part;           // f32[C][tiles_per_chain]
  uint32_t tiles_per_chain;
  uint32_t prior_lo, prior_hi; // element range of the gaussian prior (empty: lo == hi)
  float prior_coef;            // != 0: `grad` lacks the prior term theta * coef (added here)
  float* grad_rw;              // completed gradient written back (or null)

  // g + theta * coef on the prior range: the same fmaf the gradient epilogue of the
  // potential kernel applies when it adds the prior itself
  __device__ __forceinline__ float4 with_prior(const float4& g, const float4& t, uint32_t e) const {
    float4 o = g;
    if (e >= prior_lo && e + 4u <= prior_hi) {
      o.x = fmaf(t.x, prior_coef, g.x); o.y = fmaf(t.y, prior_coef, g.y);
      o.z = fmaf(t.z, prior_coef, g.z); o.w = fmaf(t.w, prior_coef, g.w);
    } else {
      if (e >= prior_lo && e < prior_hi) o.x = fmaf(t.x, prior_coef, g.x);
      if (e + 1u >= prior_lo && e + 1u < prior_hi) o.y = fmaf(t.y, prior_coef, g.y);
      if (e + 2u >= prior_lo && e + 2u < prior_hi) o.z = fmaf(t.z, prior_coef, g.z);
      if (e + 3u >= prior_lo && e + 3u < prior_hi) o.w = fmaf(t.w, prior_coef, g.w);
    }
    return o;
  }
  template <bool HINT = false>
  __device__ __forceinline__ void emit4(int64_t i, const float4& t, float s, uint64_t pol = 0) const {
    const float v0 = t.x * s, v1 = t.y * s, v2 = t.z * s, v3 = t.w * s;
    if (FMT == 1) {
      const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y);
      const __half2 l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
      uint2 ph, pl;
      ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
      pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
      if (HINT) {
        st2u_hint(reinterpret_cast<__half*>(th_hi) + i, ph, pol);
        st2u_hint(th_lo + i, pl, pol);
      } else {
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(th_hi) + i) = ph;
        *reinterpret_cast<uint2*>(th_lo + i) = pl;
      }
    } else {
      const __nv_bfloat162 b01 = __floats2bfloat162_rn(v0, v1), b23 = __floats2bfloat162_rn(v2, v3);
      uint2 pb;
      pb.x = *reinterpret_cast<const uint32_t*>(&b01); pb.y = *reinterpret_cast<const uint32_t*>(&b23);
      if (HINT) st2u_hint(reinterpret_cast<__nv_bfloat16*>(th_hi) + i, pb, pol);
      else *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(th_hi) + i) = pb;
    }
  }
  __device__ __forceinline__ float sq4(const float4& t, uint32_t e, float& amax) const {
    amax = fmaxf(fmaxf(amax, fmaxf(fabsf(t.x), fabsf(t.y))), fmaxf(fabsf(t.z), fabsf(t.w)));
    if (e >= prior_lo && e + 4u <= prior_hi)
      return (t.x * t.x + t.y * t.y) + (t.z * t.z + t.w * t.w);
    float q = 0.f;
    if (e >= prior_lo && e < prior_hi) q = fmaf(t.x, t.x, q);
    if (e + 1u >= prior_lo && e + 1u < prior_hi) q = fmaf(t.y, t.y, q);
    if (e + 2u >= prior_lo && e + 2u < prior_hi) q = fmaf(t.z, t.z, q);
    if (e + 3u >= prior_lo && e + 3u < prior_hi) q = fmaf(t.w, t.w, q);
    return q;
  }
  __device__ float apply_vec2(Regs& r, const float nA[4], const float nB[4], int64_t iA,
                              int64_t iB, int64_t c, uint32_t eA, uint32_t eB,
                              float& amax) const {
    const float ns = Base::scale_for(c);
    float va[4] = {r.vA.x, r.vA.y, r.vA.z, r.vA.w};
    float vb[4] = {r.vB.x, r.vB.y, r.vB.z, r.vB.w};
    if (prior_coef != 0.f) {
      r.gA = with_prior(r.gA, r.tA, eA);
      r.gB = with_prior(r.gB, r.tB, eB);
      if (grad_rw) {
        st4(grad_rw, iA, r.gA);
        st4(grad_rw, iB, r.gB);
      }
    }
    float4 oA, oB;
    SGMC_F4_MAP(oA, Base::one(f4get(r.tA, k), f4get(r.gA, k), va[k], nA[k], ns));
    SGMC_F4_MAP(oB, Base::one(f4get(r.tB, k), f4get(r.gB, k), vb[k], nB[k], ns));
    st4(Base::theta, iA, oA);
    st4(Base::theta, iB, oB);
    if (RMS) {
      st4(Base::v, iA, make_float4(va[0], va[1], va[2], va[3]));
      st4(Base::v, iB, make_float4(vb[0], vb[1], vb[2], vb[3]));
    }
    const float s = FMT == 1 ? __ldg(scale + c) : 1.0f;
    emit4(iA, oA, s);
    emit4(iB, oB, s);
    return sq4(oA, eA, amax) + sq4(oB, eB, amax);
  }
  __device__ void reduce2(int64_t c, uint32_t tile_in_chain, float sum, float amax) const {
    sumsq_part[c * tiles_per_chain + tile_in_chain] = sum;
    atomicMax(amax_bits + c, __float_as_uint(amax));   // amax >= 0: uint order == float order
  }
};

// ---- SGHMC ------------------------------------------------------------------
// begin: p = sqrt(m) * xi (integrator.py:737-738); theta += eps*(inv_m*p) (:610-612)
struct SghmcBeginOp {
  float* theta;
  float* mom;
  const float* mass;   // f32[P] or null
  float eps;
  static constexpr bool kReduce = false;
  struct Regs {
    float4 tA, tB;
  };
  __device__ __forceinline__ float one(float t, float xi, uint32_t e, float& p) const {
    float inv_m = 1.0f, sqrt_m = 1.0f;
    if (mass) {
      const float m = mass[e];
      inv_m = __frcp_rn(m);
      sqrt_m = __fsqrt_rn(m);
    }
    p = __fmul_rn(sqrt_m, xi);
    return __fadd_rn(t, __fmul_rn(eps, __fmul_rn(inv_m, p)));
  }
  __device__ void load_vec(Regs& r, int64_t iA, int64_t iB) const {
    r.tA = ld4(theta, iA);
    r.tB = ld4(theta, iB);
  }
  __device__ float apply_vec(Regs& r, const float nA[4], const float nB[4],
                             int64_t iA, int64_t iB, int64_t, uint32_t eA,
                             uint32_t eB) const {
    float pa[4], pb[4];
    float4 oA, oB;
    SGMC_F4_MAP(oA, one(f4get(r.tA, k), nA[k], eA + k, pa[k]));
    SGMC_F4_MAP(oB, one(f4get(r.tB, k), nB[k], eB + k, pb[k]));
    st4(theta, iA, oA);
    st4(theta, iB, oB);
    st4(mom, iA, make_float4(pa[0], pa[1], pa[2], pa[3]));
    st4(mom, iB, make_float4(pb[0], pb[1], pb[2], pb[3]));
    return 0.f;
  }
  __device__ float apply_one(int64_t i, float n, int64_t, uint32_t e) const {
    float p;
    theta[i] = one(theta[i], n, e, p);
    mom[i] = p;
    return 0.f;
  }
};

// step: integrator.py:616-655 (+ next position update :610-612 unless last)
struct SghmcStepOp {
  float* theta;
  float* mom;
  const float* grad;
  const float* friction;  // f32[P] or null -> friction_scalar
  const float* mass;      // f32[P] or null
  float eps, neg_eps, noise_scale, friction_scalar;
  int last;
  // Fisher noise model (adaption.fisher_information): cb_diff_sqrt as f32[C][P], or null.
  // The noise becomes sqrt(2 eps) * (cb_diff_sqrt o xi) and is NOT multiplied by the
  // friction (integrator.py:632-650).
  const float* noise_mul = nullptr;
  static constexpr bool kReduce = false;
  struct Regs {
    float4 tA, tB, pA, pB, gA, gB;
  };
  __device__ __forceinline__ float one(float t, float& p, float g, float xi,
                                       uint32_t e, int64_t gi) const {
    const float C = friction ? friction[e] : friction_scalar;
    const float inv_m = mass ? __frcp_rn(mass[e]) : 1.0f;
    const float m = __fmul_rn(inv_m, p);
    const float p1 = __fadd_rn(p, __fmul_rn(__fmul_rn(neg_eps, C), m));
    const float p2 = __fadd_rn(p1, __fmul_rn(neg_eps, g));
    const float p3 = noise_mul
        ? __fadd_rn(p2, __fmul_rn(noise_scale, __fmul_rn(noise_mul[gi], xi)))
        : __fadd_rn(p2, __fmul_rn(C, __fmul_rn(noise_scale, xi)));
    p = p3;
    if (last) return t;
    return __fadd_rn(t, __fmul_rn(eps, __fmul_rn(inv_m, p3)));
  }
  __device__ void load_vec(Regs& r, int64_t iA, int64_t iB) const {
    if (!last) {
      r.tA = ld4(theta, iA);
      r.tB = ld4(theta, iB);
    }
    r.pA = ld4(mom, iA);
    r.pB = ld4(mom, iB);
    r.gA = ld4(grad, iA);
    r.gB = ld4(grad, iB);
  }
  __device__ float apply_vec(Regs& r, const float nA[4], const float nB[4],
                             int64_t iA, int64_t iB, int64_t, uint32_t eA,
                             uint32_t eB) const {
    float pa[4] = {r.pA.x, r.pA.y, r.pA.z, r.pA.w};
    float pb[4] = {r.pB.x, r.pB.y, r.pB.z, r.pB.w};
    float4 oA, oB;
    SGMC_F4_MAP(oA, one(f4get(r.tA, k), pa[k], f4get(r.gA, k), nA[k], eA + k, iA + k));
    SGMC_F4_MAP(oB, one(f4get(r.tB, k), pb[k], f4get(r.gB, k), nB[k], eB + k, iB + k));
    if (!last) {
      st4(theta, iA, oA);
      st4(theta, iB, oB);
    }
    st4(mom, iA, make_float4(pa[0], pa[1], pa[2], pa[3]));
    st4(mom, iB, make_float4(pb[0], pb[1], pb[2], pb[3]));
    return 0.f;
  }
  __device__ float apply_one(int64_t i, float n, int64_t, uint32_t e) const {
    float p = mom[i];
    const float t = one(last ? 0.f : theta[i], p, grad[i], n, e, i);
    if (!last) theta[i] = t;
    mom[i] = p;
    return 0.f;
  }
};

// ---- OBABO --------------------------------------------------------------------
// O(p, xi) = (sqrt(a) p) + (sqrt((1-a)T) * (sqrt(m) xi))   integrator.py:192-200
struct ObaboAOp {   // integrator.py:210-240
  float* theta;
  float* mom;
  const float* grad;
  float* ke;           // f32[C] accumulator (kinetic_energy_start)
  const float* mass;
  float eps, sqrt_a, o_noise, neg_half_eps;
  // adapted per-chain mass matrix (adaption.mass_matrix): M^-1 and M^1/2 as f32[C][P], or null
  const float* mass_inv = nullptr;
  const float* mass_sqrt = nullptr;
  static constexpr bool kReduce = true;
  struct Regs {
    float4 tA, tB, pA, pB, gA, gB;
  };
  __device__ __forceinline__ float one(float& t, float& p, float g, float xi,
                                       uint32_t e, int64_t gi) const {
    float inv_m = 1.0f, sqrt_m = 1.0f;
    if (mass_inv) {
      inv_m = mass_inv[gi];
      sqrt_m = mass_sqrt[gi];
    } else if (mass) {
      const float m = mass[e];
      inv_m = __frcp_rn(m);
      sqrt_m = __fsqrt_rn(m);
    }
    const float p1 = __fadd_rn(__fmul_rn(sqrt_a, p),
                               __fmul_rn(o_noise, __fmul_rn(sqrt_m, xi)));
    const float ke1 = __fmul_rn(p1, __fmul_rn(inv_m, p1));
    const float p2 = __fadd_rn(__fmul_rn(neg_half_eps, g), p1);
    t = __fadd_rn(t, __fmul_rn(eps, __fmul_rn(inv_m, p2)));
    p = p2;
    return ke1;
  }
  __device__ void load_vec(Regs& r, int64_t iA, int64_t iB) const {
    r.tA = ld4(theta, iA); r.tB = ld4(theta, iB);
    r.pA = ld4(mom, iA);   r.pB = ld4(mom, iB);
    r.gA = ld4(grad, iA);  r.gB = ld4(grad, iB);
  }
  __device__ float apply_vec(Regs& r, const float nA[4], const float nB[4],
                             int64_t iA, int64_t iB, int64_t, uint32_t eA,
                             uint32_t eB) const {
    float ta[4] = {r.tA.x, r.tA.y, r.tA.z, r.tA.w};
    float tb[4] = {r.tB.x, r.tB.y, r.tB.z, r.tB.w};
    float pa[4] = {r.pA.x, r.pA.y, r.pA.z, r.pA.w};
    float pb[4] = {r.pB.x, r.pB.y, r.pB.z, r.pB.w};
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s += one(ta[k], pa[k], f4get(r.gA, k), nA[k], eA + k, iA + k);
      s += one(tb[k], pb[k], f4get(r.gB, k), nB[k], eB + k, iB + k);
    }
    st4(theta, iA, make_float4(ta[0], ta[1], ta[2], ta[3]));
    st4(theta, iB, make_float4(tb[0], tb[1], tb[2], tb[3]));
    st4(mom, iA, make_float4(pa[0], pa[1], pa[2], pa[3]));
    st4(mom, iB, make_float4(pb[0], pb[1], pb[2], pb[3]));
    return s;
  }
  __device__ float apply_one(int64_t i, float n, int64_t, uint32_t e) const {
    float t = theta[i], p = mom[i];
    const float s = one(t, p, grad[i], n, e, i);
    theta[i] = t;
    mom[i] = p;
    return s;
  }
  __device__ void reduce(int64_t c, uint32_t lt, uint32_t tpc, float s) const {
    ke[c * tpc + lt] = s;         // per-tile partial; k_energy_finish adds them in tile order
  }
};

struct ObaboBOp {   // integrator.py:248-261
  float* mom;
  const float* grad;
  float* ke;           // kinetic_energy_end
  const float* mass;
  float sqrt_a, o_noise, neg_half_eps;
  const float* mass_inv = nullptr;    // adapted per-chain mass matrix, f32[C][P] each
  const float* mass_sqrt = nullptr;
  static constexpr bool kReduce = true;
  struct Regs {
    float4 pA, pB, gA, gB;
  };
  __device__ __forceinline__ float one(float& p, float g, float xi, uint32_t e,
                                       int64_t gi) const {
    float inv_m = 1.0f, sqrt_m = 1.0f;
    if (mass_inv) {
      inv_m = mass_inv[gi];
      sqrt_m = mass_sqrt[gi];
    } else if (mass) {
      const float m = mass[e];
      inv_m = __frcp_rn(m);
      sqrt_m = __fsqrt_rn(m);
    }
    const float p3 = __fadd_rn(__fmul_rn(neg_half_eps, g), p);
    const float ke3 = __fmul_rn(p3, __fmul_rn(inv_m, p3));
    p = __fadd_rn(__fmul_rn(sqrt_a, p3),
                  __fmul_rn(o_noise, __fmul_rn(sqrt_m, xi)));
    return ke3;
  }
  __device__ void load_vec(Regs& r, int64_t iA, int64_t iB) const {
    r.pA = ld4(mom, iA);  r.pB = ld4(mom, iB);
    r.gA = ld4(grad, iA); r.gB = ld4(grad, iB);
  }
  __device__ float apply_vec(Regs& r, const float nA[4], const float nB[4],
                             int64_t iA, int64_t iB, int64_t, uint32_t eA,
                             uint32_t eB) const {
    float pa[4] = {r.pA.x, r.pA.y, r.pA.z, r.pA.w};
    float pb[4] = {r.pB.x, r.pB.y, r.pB.z, r.pB.w};
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s += one(pa[k], f4get(r.gA, k), nA[k], eA + k, iA + k);
      s += one(pb[k], f4get(r.gB, k), nB[k], eB + k, iB + k);
    }
    st4(mom, iA, make_float4(pa[0], pa[1], pa[2], pa[3]));
    st4(mom, iB, make_float4(pb[0], pb[1], pb[2], pb[3]));
    return s;
  }
  __device__ float apply_one(int64_t i, float n, int64_t, uint32_t e) const {
    float p = mom[i];
    const float s = one(p, grad[i], n, e, i);
    mom[i] = p;
    return s;
  }
  __device__ void reduce(int64_t c, uint32_t lt, uint32_t tpc, float s) const {
    ke[c * tpc + lt] = s;         // per-tile partial; k_energy_finish adds them in tile order
  }
};

// ---- reversible leapfrog (AMAGOLD) -------------------------------------------------
// One _body_fun of integrator.reversible_leapfrog (integrator.py:395-466) after
// the gradient, with the NEXT position update folded in:
//   p' = (((1-eps*f) p) + (-eps g) + sqrt(4 f eps) (sqrt(m) xi)) * (1/(1+eps*f))
//   energy[c] += (0.5 eps) * sum((p + p') * (inv_m g))                 (:395-400)
//   theta  += pos_scale * (inv_m p')    pos_scale = eps (next body's :411-418) or
//                                       0.5 eps (closing half step :545-546)
struct RevLeapfrogOp {
  float* theta;
  float* mom;
  const float* grad;
  float* energy;       // f32[C] accumulator (LeapfrogState.potential)
  const float* mass;   // f32[P] or null
  float decay, neg_eps, noise_scale, inv_norm, half_eps, pos_scale;
  const float* mass_inv = nullptr;    // adapted per-chain mass matrix, f32[C][P] each
  const float* mass_sqrt = nullptr;
  static constexpr bool kReduce = true;
  struct Regs {
    float4 tA, tB, pA, pB, gA, gB;
  };
  __device__ __forceinline__ float one(float& t, float& p, float g, float xi,
                                       uint32_t e, int64_t gi) const {
    float inv_m = 1.0f, sqrt_m = 1.0f;
    if (mass_inv) {
      inv_m = mass_inv[gi];
      sqrt_m = mass_sqrt[gi];
    } else if (mass) {
      const float m = mass[e];
      inv_m = __frcp_rn(m);
      sqrt_m = __fsqrt_rn(m);
    }
    const float noise = __fmul_rn(noise_scale, __fmul_rn(sqrt_m, xi));
    const float un = __fadd_rn(__fadd_rn(__fmul_rn(decay, p), __fmul_rn(neg_eps, g)), noise);
    const float pn = __fmul_rn(inv_norm, un);
    const float en = __fmul_rn(__fadd_rn(p, pn), __fmul_rn(inv_m, g));
    t = __fadd_rn(t, __fmul_rn(pos_scale, __fmul_rn(inv_m, pn)));
    p = pn;
    return en;
  }
  __device__ void load_vec(Regs& r, int64_t iA, int64_t iB) const {
    r.tA = ld4(theta, iA); r.tB = ld4(theta, iB);
    r.pA = ld4(mom, iA);   r.pB = ld4(mom, iB);
    r.gA = ld4(grad, iA);  r.gB = ld4(grad, iB);
  }
  __device__ float apply_vec(Regs& r, const float nA[4], const float nB[4],
                             int64_t iA, int64_t iB, int64_t, uint32_t eA,
                             uint32_t eB) const {
    float ta[4] = {r.tA.x, r.tA.y, r.tA.z, r.tA.w};
    float tb[4] = {r.tB.x, r.tB.y, r.tB.z, r.tB.w};
    float pa[4] = {r.pA.x, r.pA.y, r.pA.z, r.pA.w};
    float pb[4] = {r.pB.x, r.pB.y, r.pB.z, r.pB.w};
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s += one(ta[k], pa[k], f4get(r.gA, k), nA[k], eA + k, iA + k);
      s += one(tb[k], pb[k], f4get(r.gB, k), nB[k], eB + k, iB + k);
    }
    st4(theta, iA, make_float4(ta[0], ta[1], ta[2], ta[3]));
    st4(theta, iB, make_float4(tb[0], tb[1], tb[2], tb[3]));
    st4(mom, iA, make_float4(pa[0], pa[1], pa[2], pa[3]));
    st4(mom, iB, make_float4(pb[0], pb[1], pb[2], pb[3]));
    return s;
  }
  __device__ float apply_one(int64_t i, float n, int64_t, uint32_t e) const {
    float t = theta[i], p = mom[i];
    const float s = one(t, p, grad[i], n, e, i);
    theta[i] = t;
    mom[i] = p;
    return s;
  }
  __device__ void reduce(int64_t c, uint32_t lt, uint32_t tpc, float s) const {
    energy[c * tpc + lt] = s;     // per-tile partial; k_energy_finish adds them in tile order
  }
};

// Adapted mass matrix of the *_adapted entry points: set for the duration of one call on
// the calling thread, picked up where the constant-mass entry points build their ops.
struct AdaptedMass { const float* inv = nullptr; const float* sqrt = nullptr; };
static thread_local AdaptedMass g_adapted;
static thread_local const float* g_noise_mul = nullptr;   // sgmc_sghmc_step_noise_model
struct AdaptedMassScope {
  AdaptedMassScope(const float* inv, const float* sqrt) { g_adapted.inv = inv; g_adapted.sqrt = sqrt; }
  ~AdaptedMassScope() { g_adapted = AdaptedMass{}; }
};

// ---- per-chain energies in a fixed order ------------------------------------------------
// The OBABO and reversible-leapfrog passes reduce <p, M^-1 p> / <p + p', M^-1 g> per chain.
// A chain's warp-tiles run on different CTAs in no particular order, so the pass stores one
// partial per (chain, tile) and k_energy_finish adds them in tile order: the energy -- and
// with it every Metropolis-Hastings decision -- is reproducible run to run.  The partials
// live in a scratch buffer cached per (device, stream): calls on one stream are ordered, so
// they can share it; it only ever grows.
__global__ void __launch_bounds__(256) k_energy_finish(const float* __restrict__ part, uint32_t tpc,
                                                       float* __restrict__ out, float scale,
                                                       int64_t n_chains) {
  pdl_wait();
  const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= n_chains) return;
  const float* p = part + c * tpc;
  float s = 0.f;
  for (uint32_t t = lane; t < tpc; t += 32) s += p[t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[c] += scale * s;
}

static float* energy_scratch(cudaStream_t stream, size_t floats) {
  struct Buf { float* p = nullptr; size_t cap = 0; };
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, Buf> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  Buf& b = cache[{dev, stream}];
  if (b.cap < floats) {
    if (b.p) {   // earlier passes on this stream may still read the old buffer
      if (cudaStreamSynchronize(stream) != cudaSuccess) return nullptr;
      cudaFree(b.p);
      b = Buf{};
    }
    const size_t cap = floats + floats / 4 + 1024;
    if (cudaMalloc(&b.p, cap * sizeof(float)) != cudaSuccess) { b.p = nullptr; return nullptr; }
    b.cap = cap;
  }
  return b.p;
}

// Runs `op` (whose energy pointer the caller has set to the scratch) and then the ordered sum
// out[c] += scale * sum_t part[c][t].
template <class Op>
static int launch_energy_pass(cudaStream_t stream, const LeafTable& tab, const uint32_t* keys_in,
                              uint32_t* keys_out, int64_t n_chains, int key_mode, int layout,
                              const Op& op, const float* part, float* out, float scale,
                              const char* name) {
  if (int e = launch_noise_pass(stream, tab, keys_in, keys_out, n_chains, key_mode, layout, op, name))
    return e;
  if (n_chains == 0 || tab.tiles_per_chain == 0) return 0;
  launch_pdl(k_energy_finish, dim3((unsigned)((n_chains + 7) / 8)), dim3(256), 0, stream, part,
             tab.tiles_per_chain, out, scale, n_chains);
  return post_launch(name);
}

static bool aligned16(std::initializer_list<const void*> ps) {
  for (const void* p : ps)
    if (p && (reinterpret_cast<uintptr_t>(p) & 15u)) return false;
  return true;
}

}  // namespace sgmc

using namespace sgmc;

extern "C" {

int sgmc_normal_like(void* stream, const uint32_t* keys, float* noise,
                     int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                     int prng_layout) {
  LeafTable tab;
  if (int e = build_leaf_table(&tab, leaf_sizes, n_leaves, n_chains,
                               aligned16({noise}))) return e;
  NormalLikeOp op{noise};
  return launch_noise_pass((cudaStream_t)stream, tab, keys, nullptr, n_chains,
                           option(1) == 99 ? 99 : kKeyDirect, prng_layout, op,
                           "sgmc_normal_like");
}

static int sgld_common(void* stream, float* theta, float* v, const float* grad,
                       const uint32_t* keys_in, uint32_t* keys_out,
                       int64_t n_chains, const int64_t* leaf_sizes,
                       int n_leaves, float step_size, float temperature,
                       const float* temp_per_chain, float alpha, float lmbd,
                       int prng_layout, bool rms) {
  SGMC_REQUIRE(keys_in != keys_out, "keys_out must not alias keys_in");
  LeafTable tab;
  if (int e = build_leaf_table(&tab, leaf_sizes, n_leaves, n_chains,
                               aligned16({theta, v, grad}))) return e;
  // integrator.py:882-884: (-eps), sqrt(2*T*eps) in f32
  const float eps = step_size;
  const float ns = sqrtf((2.0f * temperature) * eps);
  if (rms && option(SGMC_OPT_EXACT_UPDATE_MATH)) {
    SgldOp<true, false> op{theta, v, grad, temp_per_chain, eps, -eps, ns,
                           alpha, 1.0f - alpha, lmbd};
    return launch_noise_pass((cudaStream_t)stream, tab, keys_in, keys_out,
                             n_chains, kKeySplit2, prng_layout, op,
                             "sgmc_sgld_rms_update");
  }
  if (rms) {
    SgldOp<true, true> op{theta, v, grad, temp_per_chain, eps, -eps, ns,
                          alpha, 1.0f - alpha, lmbd};
    return launch_noise_pass((cudaStream_t)stream, tab, keys_in, keys_out,
                             n_chains, kKeySplit2, prng_layout, op,
                             "sgmc_sgld_rms_update");
  }
  SgldOp<false, false> op{theta, nullptr, grad, temp_per_chain, eps, -eps, ns,
                          0.f, 0.f, 0.f};
  return launch_noise_pass((cudaStream_t)stream, tab, keys_in, keys_out,
                           n_chains, kKeySplit2, prng_layout, op,
                           "sgmc_sgld_update");
}

}  // extern "C"

namespace sgmc {

// The carried step's update when its noise was already generated (xi, by the potential
// kernel's shadow job): one elementwise pass, one warp per 256 consecutive elements of a
// chain (two float4 per lane), same arithmetic and outputs as SgldSplitOp::apply_vec2.
// HBM-bound: reads theta, grad, v, xi (grad and xi were just written: L2 hits), writes
// theta, v and the fp16 hi/lo (or bf16) operand form.
template <bool RMS, bool FAST, int FMT>
__global__ void __launch_bounds__(256)
k_sgld_apply_split(const SgldSplitOp<RMS, FAST, FMT> op, const float* __restrict__ xi, int64_t P,
                   int64_t n_tiles) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tile >= n_tiles) return;
  const uint32_t tpc = op.tiles_per_chain;
  const int64_t c = tile / tpc;
  const uint32_t t = (uint32_t)(tile - c * tpc);
  const float ns = op.scale_for(c);
  const float s = FMT == 1 ? __ldg(op.scale + c) : 1.0f;
  const uint64_t keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
  float sum = 0.f, amax = 0.f;
  const uint32_t e0 = t * 256u;
  if (e0 + 256u <= (uint32_t)P && e0 >= op.prior_lo && e0 + 256u <= op.prior_hi &&
      op.prior_coef != 0.f && op.grad_rw == nullptr) {
    // the common tile: whole, inside the prior range (sgld_apply_tile.cuh)
    ApplyTileArgs ta;
    ta.theta = op.theta; ta.v = op.v; ta.grad = op.grad; ta.xi = xi;
    ta.th_hi = op.th_hi; ta.th_lo = op.th_lo; ta.scale = op.scale;
    ta.amax_bits = op.amax_bits; ta.sumsq_part = op.sumsq_part; ta.P = P;
    ta.tiles_per_chain = tpc; ta.prior_on = 1; ta.prior_coef = op.prior_coef;
    ta.noise_scale = ns; ta.neg_eps = op.neg_eps; ta.alpha = op.alpha;
    ta.one_m_alpha = op.one_m_alpha; ta.lmbd = op.lmbd;
    apply_tile_fast<RMS, FAST, FMT>(ta, c, t, lane, keep, drop);
    return;
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t e = t * 256u + (uint32_t)h * 128u + (uint32_t)lane * 4u;
      if (e < (uint32_t)P) {                       // P % 8 == 0: whole float4 or nothing
        const int64_t i = c * P + e;
        const float4 th = ld4_hint(op.theta, i, keep);
        float4 g = ld4_hint(op.grad, i, drop);
        const float4 x = ld4_hint(xi, i, drop);
        float4 vv = RMS ? ld4_hint(op.v, i, keep) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (op.prior_coef != 0.f) {
          g = op.with_prior(g, th, e);
          if (op.grad_rw) st4(op.grad_rw, i, g);
        }
        float4 o;
        o.x = op.one(th.x, g.x, vv.x, x.x, ns);
        o.y = op.one(th.y, g.y, vv.y, x.y, ns);
        o.z = op.one(th.z, g.z, vv.z, x.z, ns);
        o.w = op.one(th.w, g.w, vv.w, x.w, ns);
        st4_hint(op.theta, i, o, keep);
        if (RMS) st4_hint(op.v, i, vv, keep);
        op.template emit4<true>(i, o, s, keep);
        sum += op.sq4(o, e, amax);
      }
    }
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, k);
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, k));
  }
  if (lane == 0) op.reduce2(c, t, sum, amax);
}

// The same update as a STREAMING kernel: two persistent CTAs per SM, a producer thread keeps
// kStreamStages units (2048 consecutive parameters: theta, grad, xi, v = 32 KB) in flight as
// bulk async copies into shared memory, eight compute warps work through the units (a warp
// owns the same 256 elements, in the same order, as in k_sgld_apply_split: identical bits),
// and the four outputs leave through double-buffered shared memory as bulk stores.  The bytes
// in flight per SM (128 KB) no longer depend on registers or occupancy.  MEASURED (C2 step,
// B200): 21.4 us against 20.5 us of the one-shot kernel above (one CTA per SM with three
// stages: 26.7 us -- eight compute warps cannot hide their own instruction latencies), i.e.
// no gain: in steady state the update moves 117 MB (theta, grad, xi, v in; theta, v and the
// fp16 hi / lo operand form out), 75 MB of them through DRAM (grad and xi hit the L2), at
// 5.7 TB/s = 0.87 of the measured copy bandwidth -- the kernel is at the roof of the bytes it
// moves, not of the 83.9 MB the algorithm needs.  Kept as an opt-in A/B
// (SGMC_OPT_STREAM_UPDATE); needs P % 256 == 0, the prior on the whole sample (or none) and
// no gradient write-back.
constexpr int kStreamUnit = 2048;                  // parameters per unit: 8 warps x 256
constexpr int kStreamStages = 2;                  // per CTA; two CTAs share an SM
constexpr int kStreamInBytes = 4 * kStreamUnit * 4;            // theta, grad, xi, v
constexpr int kStreamOutBytes = 2 * kStreamUnit * 4 + 2 * kStreamUnit * 2;   // theta, v, hi, lo
constexpr int kStreamSmem = kStreamStages * kStreamInBytes + 2 * kStreamOutBytes + 128;
constexpr int kStreamThreads = 256 + 32;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes,
                                          uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;"
      :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes,
                                           uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_wait_read_1() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

template <bool RMS, bool FAST, int FMT>
__global__ void __launch_bounds__(kStreamThreads, 2)
k_sgld_apply_stream(const SgldSplitOp<RMS, FAST, FMT> op, const float* __restrict__ xi, int64_t P,
                    int64_t n_units, int64_t n_params) {
  extern __shared__ uint8_t stream_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(stream_smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* in_buf = smem;
  uint8_t* out_buf = smem + kStreamStages * kStreamInBytes;
  __shared__ uint64_t full_bar[kStreamStages], empty_bar[kStreamStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int st = 0; st < kStreamStages; ++st) {
      mbar_init(&full_bar[st], 1);
      mbar_init(&empty_bar[st], 8);
    }
    fence_barrier_init();
  }
  pdl_launch_dependents();
  pdl_wait();
  __syncthreads();
  const uint64_t keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
  const int64_t first = blockIdx.x, stride = gridDim.x;
  if (warp == 8) {
    // ===== producer: bulk loads of the next units =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t u = first; u < n_units; u += stride, ++it) {
        const int st = it % kStreamStages;
        if (it >= (uint32_t)kStreamStages) mbar_wait(&empty_bar[st], ((it / kStreamStages) - 1) & 1);
        const int64_t e = u * kStreamUnit;
        const int64_t left = n_params - e;
        const uint32_t bytes = (uint32_t)((left < kStreamUnit ? left : kStreamUnit) * 4);
        uint8_t* dst = in_buf + st * kStreamInBytes;
        mbar_expect_tx(&full_bar[st], (RMS ? 4u : 3u) * bytes);
        bulk_load(dst, op.theta + e, bytes, &full_bar[st], keep);
        bulk_load(dst + kStreamUnit * 4, op.grad + e, bytes, &full_bar[st], drop);
        bulk_load(dst + 2 * kStreamUnit * 4, xi + e, bytes, &full_bar[st], drop);
        if (RMS) bulk_load(dst + 3 * kStreamUnit * 4, op.v + e, bytes, &full_bar[st], keep);
      }
    }
    return;
  }
  // ===== compute warps =====
  const float pc = op.prior_coef;
  uint32_t it = 0;
  for (int64_t u = first; u < n_units; u += stride, ++it) {
    const int st = it % kStreamStages, ob = it & 1;
    const int64_t wt = u * 8 + warp;                     // this warp's 256-element tile
    const int64_t e_warp = wt * 256;
    const bool active = e_warp < n_params;               // (the last unit may be short)
    const uint32_t tpc = op.tiles_per_chain;
    const int64_t c = active ? wt / tpc : 0;
    const uint32_t t = (uint32_t)(wt - c * tpc);
    const float ns = op.scale_for(c);
    const float s = FMT == 1 ? __ldg(op.scale + c) : 1.0f;
    mbar_wait(&full_bar[st], (it / kStreamStages) & 1);
    const uint8_t* src = in_buf + st * kStreamInBytes;
    const int l0 = warp * 256 + lane * 4, l1 = l0 + 128;            // element inside the unit
    const float4 th0 = *reinterpret_cast<const float4*>(src + l0 * 4);
    const float4 th1 = *reinterpret_cast<const float4*>(src + l1 * 4);
    float4 g0 = *reinterpret_cast<const float4*>(src + kStreamUnit * 4 + l0 * 4);
    float4 g1 = *reinterpret_cast<const float4*>(src + kStreamUnit * 4 + l1 * 4);
    const float4 x0 = *reinterpret_cast<const float4*>(src + 2 * kStreamUnit * 4 + l0 * 4);
    const float4 x1 = *reinterpret_cast<const float4*>(src + 2 * kStreamUnit * 4 + l1 * 4);
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (RMS) {
      v0 = *reinterpret_cast<const float4*>(src + 3 * kStreamUnit * 4 + l0 * 4);
      v1 = *reinterpret_cast<const float4*>(src + 3 * kStreamUnit * 4 + l1 * 4);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[st]);          // the stage can be refilled
    if (pc != 0.f) {
      g0.x = fmaf(th0.x, pc, g0.x); g0.y = fmaf(th0.y, pc, g0.y);
      g0.z = fmaf(th0.z, pc, g0.z); g0.w = fmaf(th0.w, pc, g0.w);
      g1.x = fmaf(th1.x, pc, g1.x); g1.y = fmaf(th1.y, pc, g1.y);
      g1.z = fmaf(th1.z, pc, g1.z); g1.w = fmaf(th1.w, pc, g1.w);
    }
    float4 o0, o1;
    o0.x = op.one(th0.x, g0.x, v0.x, x0.x, ns); o0.y = op.one(th0.y, g0.y, v0.y, x0.y, ns);
    o0.z = op.one(th0.z, g0.z, v0.z, x0.z, ns); o0.w = op.one(th0.w, g0.w, v0.w, x0.w, ns);
    o1.x = op.one(th1.x, g1.x, v1.x, x1.x, ns); o1.y = op.one(th1.y, g1.y, v1.y, x1.y, ns);
    o1.z = op.one(th1.z, g1.z, v1.z, x1.z, ns); o1.w = op.one(th1.w, g1.w, v1.w, x1.w, ns);
    // the output buffer of two units ago has been read by its bulk stores
    if (threadIdx.x == 0) bulk_wait_read_1();
    named_bar_sync(1, 256);
    uint8_t* dst = out_buf + ob * kStreamOutBytes;
    *reinterpret_cast<float4*>(dst + l0 * 4) = o0;
    *reinterpret_cast<float4*>(dst + l1 * 4) = o1;
    if (RMS) {
      *reinterpret_cast<float4*>(dst + kStreamUnit * 4 + l0 * 4) = v0;
      *reinterpret_cast<float4*>(dst + kStreamUnit * 4 + l1 * 4) = v1;
    }
    uint8_t* hi = dst + 2 * kStreamUnit * 4;
    uint8_t* lo = hi + kStreamUnit * 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 o = h ? o1 : o0;
      const int l = h ? l1 : l0;
      const float w0 = o.x * s, w1 = o.y * s, w2 = o.z * s, w3 = o.w * s;
      if (FMT == 1) {
        const __half2 h01 = __floats2half2_rn(w0, w1), h23 = __floats2half2_rn(w2, w3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(w0 - f01.x, w1 - f01.y);
        const __half2 l23 = __floats2half2_rn(w2 - f23.x, w3 - f23.y);
        uint2 ph, pl;
        ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
        pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(hi + l * 2) = ph;
        *reinterpret_cast<uint2*>(lo + l * 2) = pl;
      } else {
        const __nv_bfloat162 b01 = __floats2bfloat162_rn(w0, w1), b23 = __floats2bfloat162_rn(w2, w3);
        uint2 pb;
        pb.x = *reinterpret_cast<const uint32_t*>(&b01); pb.y = *reinterpret_cast<const uint32_t*>(&b23);
        *reinterpret_cast<uint2*>(hi + l * 2) = pb;
      }
    }
    fence_proxy_async();
    named_bar_sync(2, 256);
    if (threadIdx.x == 0) {
      const int64_t e = u * kStreamUnit;
      const int64_t left = n_params - e;
      const uint32_t cnt = (uint32_t)(left < kStreamUnit ? left : kStreamUnit);
      bulk_store(op.theta + e, dst, cnt * 4, keep);
      if (RMS) bulk_store(op.v + e, dst + kStreamUnit * 4, cnt * 4, keep);
      if (FMT == 1) {
        bulk_store(reinterpret_cast<__half*>(op.th_hi) + e, hi, cnt * 2, keep);
        bulk_store(op.th_lo + e, lo, cnt * 2, keep);
      } else {
        bulk_store(reinterpret_cast<__nv_bfloat16*>(op.th_hi) + e, hi, cnt * 2, keep);
      }
      bulk_commit_group();
    }
    if (active) {
      // partial statistics of the warp's 256 elements: the association of sq4 / the
      // one-shot kernel, so the prior value comes out bit-identical
      float amax = fmaxf(fmaxf(fmaxf(fabsf(o0.x), fabsf(o0.y)), fmaxf(fabsf(o0.z), fabsf(o0.w))),
                         fmaxf(fmaxf(fabsf(o1.x), fabsf(o1.y)), fmaxf(fabsf(o1.z), fabsf(o1.w))));
      float sum = 0.f;
      if (op.prior_hi > op.prior_lo)
        sum = ((o0.x * o0.x + o0.y * o0.y) + (o0.z * o0.z + o0.w * o0.w)) +
              ((o1.x * o1.x + o1.y * o1.y) + (o1.z * o1.z + o1.w * o1.w));
#pragma unroll
      for (int k = 16; k > 0; k >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, k);
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, k));
      }
      if (lane == 0) op.reduce2(c, t, sum, amax);
    }
  }
  if (threadIdx.x == 0) bulk_wait_all();                 // the stores have been performed
}

template <bool RMS, bool FAST, int FMT>
static int launch_split(cudaStream_t stream, const LeafTable& tab, const uint32_t* keys_in,
                        uint32_t* keys_out, int64_t C, int key_mode, int layout,
                        const SgldOp<RMS, FAST>& base, const SgldSplitOut& so) {
  SgldSplitOp<RMS, FAST, FMT> op;
  static_cast<SgldOp<RMS, FAST>&>(op) = base;
  op.th_hi = so.th_hi; op.th_lo = reinterpret_cast<__half*>(so.th_lo);
  op.scale = so.scale; op.amax_bits = so.amax_bits; op.sumsq_part = so.sumsq_part;
  op.tiles_per_chain = tab.tiles_per_chain;
  op.prior_lo = (uint32_t)so.prior_lo; op.prior_hi = (uint32_t)so.prior_hi;
  op.prior_coef = so.prior_coef; op.grad_rw = so.grad_rw;
  if (so.xi != nullptr && tab.P % 256 == 0 && so.grad_rw == nullptr &&
      (so.prior_hi == so.prior_lo || (so.prior_lo == 0 && so.prior_hi == (int)tab.P)) &&
      (so.prior_coef != 0.f) == (so.prior_hi > so.prior_lo) && option(SGMC_OPT_STREAM_UPDATE)) {
    auto kfn = k_sgld_apply_stream<RMS, FAST, FMT>;
    static bool attr_set = false;
    if (!attr_set) {
      if (check_cuda(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kStreamSmem), "cudaFuncSetAttribute"))
        return 1;
      attr_set = true;
    }
    const int64_t n_params = C * (int64_t)tab.P;
    const int64_t n_units = (n_params + kStreamUnit - 1) / kStreamUnit;
    const unsigned grid = (unsigned)std::min<int64_t>(n_units, 2 * sm_count());
    launch_pdl(kfn, dim3(grid), dim3(kStreamThreads), (size_t)kStreamSmem, stream, op, so.xi,
               (int64_t)tab.P, n_units, n_params);
    return post_launch("k_sgld_apply_stream");
  }
  if (so.xi != nullptr) {
    const int64_t n_tiles = C * (int64_t)tab.tiles_per_chain;
    launch_pdl(k_sgld_apply_split<RMS, FAST, FMT>, dim3((unsigned)((n_tiles + 7) / 8)), dim3(256),
               0, stream, op, so.xi, (int64_t)tab.P, n_tiles);
    return post_launch("k_sgld_apply_split");
  }
  return launch_noise_pass(stream, tab, keys_in, keys_out, C, key_mode, layout, op,
                           "sgmc_sgld_update<split>");
}

int sgld_split_tiles_per_chain(int64_t P) { return (int)(((P / 2 + 3) / 4 + 31) / 32); }

int sgld_update_split(cudaStream_t stream, float* theta, float* v, const float* grad,
                      const uint32_t* keys_in, uint32_t* keys_out, int64_t n_chains, int64_t P,
                      float step_size, float temperature, const float* temp_per_chain,
                      float alpha, float lmbd, int prng_layout, const SgldSplitOut& so) {
  SGMC_REQUIRE(so.fmt == 1 || so.fmt == 2, "unknown split format");
  SGMC_REQUIRE(P % 8 == 0, "split update needs P %% 8 == 0");
  LeafTable tab;
  const int64_t sizes[1] = {P};
  if (int e = build_leaf_table(&tab, sizes, 1, n_chains,
                               aligned16({theta, v, grad, so.th_hi, so.th_lo}))) return e;
  SGMC_REQUIRE(tab.vec_ok[0] && !tab.shifted[0], "split update needs 16-byte aligned buffers");
  const float eps = step_size;
  const float ns = sqrtf((2.0f * temperature) * eps);
  const bool rms = v != nullptr;
  const int key_mode = so.noise_keys ? kKeyCached : kKeySplit2;
  const uint32_t* kin = so.noise_keys ? so.noise_keys : keys_in;
  uint32_t* kout = so.noise_keys ? nullptr : keys_out;
#define SGMC_SPLIT_CASE(R, F)                                                              \
  {                                                                                        \
    SgldOp<R, F> base{theta, v, grad, temp_per_chain, eps, -eps, ns, alpha, 1.0f - alpha,   \
                      lmbd};                                                               \
    if (so.fmt == 1)                                                                       \
      return launch_split<R, F, 1>(stream, tab, kin, kout, n_chains, key_mode, prng_layout, \
                                   base, so);                                              \
    return launch_split<R, F, 2>(stream, tab, kin, kout, n_chains, key_mode, prng_layout,   \
                                 base, so);                                                \
  }
  if (rms && option(SGMC_OPT_EXACT_UPDATE_MATH)) SGMC_SPLIT_CASE(true, false)
  if (rms) SGMC_SPLIT_CASE(true, true)
  SGMC_SPLIT_CASE(false, false)
#undef SGMC_SPLIT_CASE
}

}  // namespace sgmc

extern "C" {

int sgmc_sgld_update(void* stream, float* theta, const float* grad,
                     const uint32_t* keys_in, uint32_t* keys_out,
                     int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                     float step_size, float temperature,
                     const float* temp_per_chain, int prng_layout) {
  return sgld_common(stream, theta, nullptr, grad, keys_in, keys_out, n_chains,
                     leaf_sizes, n_leaves, step_size, temperature,
                     temp_per_chain, 0.f, 0.f, prng_layout, false);
}

int sgmc_sgld_rms_update(void* stream, float* theta, float* v,
                         const float* grad, const uint32_t* keys_in,
                         uint32_t* keys_out, int64_t n_chains,
                         const int64_t* leaf_sizes, int n_leaves,
                         float step_size, float temperature,
                         const float* temp_per_chain, float alpha, float lmbd,
                         int prng_layout) {
  return sgld_common(stream, theta, v, grad, keys_in, keys_out, n_chains,
                     leaf_sizes, n_leaves, step_size, temperature,
                     temp_per_chain, alpha, lmbd, prng_layout, true);
}

int sgmc_sghmc_begin(void* stream, float* theta, float* momentum,
                     const uint32_t* keys_in, uint32_t* keys_out,
                     int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                     float step_size, const float* mass, int prng_layout) {
  SGMC_REQUIRE(keys_in != keys_out, "keys_out must not alias keys_in");
  LeafTable tab;
  if (int e = build_leaf_table(&tab, leaf_sizes, n_leaves, n_chains,
                               aligned16({theta, momentum}))) return e;
  SghmcBeginOp op{theta, momentum, mass, step_size};
  return launch_noise_pass((cudaStream_t)stream, tab, keys_in, keys_out,
                           n_chains, kKeySplit2, prng_layout, op,
                           "sgmc_sghmc_begin");
}

int sgmc_sghmc_step(void* stream, float* theta, float* momentum,
                    const float* grad, const uint32_t* keys_in,
                    uint32_t* keys_out, int64_t n_chains,
                    const int64_t* leaf_sizes, int n_leaves, float step_size,
                    float friction_scalar, const float* friction,
                    const float* mass, int last, int prng_layout) {
  SGMC_REQUIRE(keys_in != keys_out, "keys_out must not alias keys_in");
  LeafTable tab;
  if (int e = build_leaf_table(&tab, leaf_sizes, n_leaves, n_chains,
                               aligned16({theta, momentum, grad}))) return e;
  const float eps = step_size;
  SghmcStepOp op{theta, momentum, grad, friction, mass, eps, -eps,
                 sqrtf(2.0f * eps), friction_scalar, last};
  op.noise_mul = g_noise_mul;
  return launch_noise_pass((cudaStream_t)stream, tab, keys_in, keys_out,
                           n_chains, kKeySplit2, prng_layout, op,
                           "sgmc_sghmc_step");
}

// The SGHMC step with a Fisher noise model (friction_leapfrog(noise_model=...),
// integrator.py:632-650): the injected noise is sqrt(2 eps) * (cb_diff_sqrt o xi) with
// cb_diff_sqrt f32[n_chains][P] from sgmc_glm_fisher_diag, instead of friction * sqrt(2 eps) xi.
int sgmc_sghmc_step_noise_model(void* stream, float* theta, float* momentum,
                                const float* grad, const uint32_t* keys_in,
                                uint32_t* keys_out, int64_t n_chains,
                                const int64_t* leaf_sizes, int n_leaves, float step_size,
                                float friction_scalar, const float* friction,
                                const float* mass, const float* cb_diff_sqrt, int last,
                                int prng_layout) {
  SGMC_REQUIRE(cb_diff_sqrt != nullptr, "null noise model");
  g_noise_mul = cb_diff_sqrt;
  const int rc = sgmc_sghmc_step(stream, theta, momentum, grad, keys_in, keys_out, n_chains,
                                 leaf_sizes, n_leaves, step_size, friction_scalar, friction,
                                 mass, last, prng_layout);
  g_noise_mul = nullptr;
  return rc;
}

int sgmc_obabo_pass_a(void* stream, float* theta, float* momentum,
                      const float* grad, float* ke_start,
                      const uint32_t* keys_in, uint32_t* keys_out,
                      int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                      float step_size, float temperature, float friction,
                      const float* mass, int prng_layout) {
  SGMC_REQUIRE(keys_in != keys_out, "keys_out must not alias keys_in");
  LeafTable tab;
  if (int e = build_leaf_table(&tab, leaf_sizes, n_leaves, n_chains,
                               aligned16({theta, momentum, grad}))) return e;
  // integrator.py:195-199
  const float a = (float)exp((double)(-friction * step_size));  // f64 libm, rounded once
  const float o_noise = sqrtf((1.0f - a) * temperature);
  float* part = energy_scratch((cudaStream_t)stream, (size_t)n_chains * tab.tiles_per_chain);
  SGMC_REQUIRE(part != nullptr || n_chains == 0, "sgmc_obabo_pass_a: no memory for the energy partials");
  ObaboAOp op{theta, momentum, grad, part, mass, step_size, sqrtf(a),
              o_noise, -1.0f * (0.5f * step_size)};
  op.mass_inv = g_adapted.inv;
  op.mass_sqrt = g_adapted.sqrt;
  return launch_energy_pass((cudaStream_t)stream, tab, keys_in, keys_out, n_chains, kKeySplit3A,
                            prng_layout, op, part, ke_start, 0.5f, "sgmc_obabo_pass_a");
}

int sgmc_obabo_pass_b(void* stream, float* momentum, const float* grad,
                      float* ke_end, const uint32_t* keys_in,
                      int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                      float step_size, float temperature, float friction,
                      const float* mass, int prng_layout) {
  LeafTable tab;
  if (int e = build_leaf_table(&tab, leaf_sizes, n_leaves, n_chains,
                               aligned16({momentum, grad}))) return e;
  const float a = (float)exp((double)(-friction * step_size));  // f64 libm, rounded once
  const float o_noise = sqrtf((1.0f - a) * temperature);
  float* part = energy_scratch((cudaStream_t)stream, (size_t)n_chains * tab.tiles_per_chain);
  SGMC_REQUIRE(part != nullptr || n_chains == 0, "sgmc_obabo_pass_b: no memory for the energy partials");
  ObaboBOp op{momentum, grad, part, mass, sqrtf(a), o_noise,
              -1.0f * (0.5f * step_size)};
  op.mass_inv = g_adapted.inv;
  op.mass_sqrt = g_adapted.sqrt;
  return launch_energy_pass((cudaStream_t)stream, tab, keys_in, nullptr, n_chains, kKeySplit3B,
                            prng_layout, op, part, ke_end, 0.5f, "sgmc_obabo_pass_b");
}

int sgmc_revleapfrog_step(void* stream, float* theta, float* momentum, const float* grad,
                          float* energy, const uint32_t* keys_in, uint32_t* keys_out,
                          int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                          float step_size, float friction, const float* mass, int last,
                          int prng_layout) {
  SGMC_REQUIRE(keys_in != keys_out, "keys_out must not alias keys_in");
  LeafTable tab;
  if (int e = build_leaf_table(&tab, leaf_sizes, n_leaves, n_chains,
                               aligned16({theta, momentum, grad}))) return e;
  // integrator.py:423-446: weak-typed python scalars times the f32 step size
  const float ef = step_size * friction;
  float* part = energy_scratch((cudaStream_t)stream, (size_t)n_chains * tab.tiles_per_chain);
  SGMC_REQUIRE(part != nullptr || n_chains == 0, "sgmc_revleapfrog_step: no memory for the energy partials");
  RevLeapfrogOp op{theta, momentum, grad, part, mass,
                   1.0f - ef, -1.0f * step_size, sqrtf((4.0f * friction) * step_size),
                   1.0f / (1.0f + ef), 0.5f * step_size,
                   last ? 0.5f * step_size : step_size};
  op.mass_inv = g_adapted.inv;
  op.mass_sqrt = g_adapted.sqrt;
  return launch_energy_pass((cudaStream_t)stream, tab, keys_in, keys_out, n_chains, kKeySplit2,
                            prng_layout, op, part, energy, 0.5f * step_size, "sgmc_revleapfrog_step");
}

// The same three passes with an ADAPTED mass matrix (adaption.mass_matrix, diagonal): every
// chain carries its own M^-1 and M^1/2 (f32[C][P] each, as MassMatrix(inv, sqrt) hands them
// to the integrators, integrator.py:177-200, :395-446) instead of one constant mass vector.
int sgmc_obabo_pass_a_adapted(void* stream, float* theta, float* momentum,
                              const float* grad, float* ke_start,
                              const uint32_t* keys_in, uint32_t* keys_out,
                              int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                              float step_size, float temperature, float friction,
                              const float* mass_inv, const float* mass_sqrt, int prng_layout) {
  SGMC_REQUIRE(mass_inv && mass_sqrt, "null mass matrix");
  AdaptedMassScope scope(mass_inv, mass_sqrt);
  return sgmc_obabo_pass_a(stream, theta, momentum, grad, ke_start, keys_in, keys_out, n_chains,
                           leaf_sizes, n_leaves, step_size, temperature, friction, nullptr,
                           prng_layout);
}

int sgmc_obabo_pass_b_adapted(void* stream, float* momentum, const float* grad,
                              float* ke_end, const uint32_t* keys_in,
                              int64_t n_chains, const int64_t* leaf_sizes, int n_leaves,
                              float step_size, float temperature, float friction,
                              const float* mass_inv, const float* mass_sqrt, int prng_layout) {
  SGMC_REQUIRE(mass_inv && mass_sqrt, "null mass matrix");
  AdaptedMassScope scope(mass_inv, mass_sqrt);
  return sgmc_obabo_pass_b(stream, momentum, grad, ke_end, keys_in, n_chains, leaf_sizes,
                           n_leaves, step_size, temperature, friction, nullptr, prng_layout);
}

int sgmc_revleapfrog_step_adapted(void* stream, float* theta, float* momentum,
                                  const float* grad, float* energy, const uint32_t* keys_in,
                                  uint32_t* keys_out, int64_t n_chains,
                                  const int64_t* leaf_sizes, int n_leaves, float step_size,
                                  float friction, const float* mass_inv,
                                  const float* mass_sqrt, int last, int prng_layout) {
  SGMC_REQUIRE(mass_inv && mass_sqrt, "null mass matrix");
  AdaptedMassScope scope(mass_inv, mass_sqrt);
  return sgmc_revleapfrog_step(stream, theta, momentum, grad, energy, keys_in, keys_out,
                               n_chains, leaf_sizes, n_leaves, step_size, friction, nullptr,
                               last, prng_layout);
}

}  // extern "C"
