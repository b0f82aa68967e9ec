// One warp-tile (256 consecutive parameters of a chain) of the carried step's update when
// the step's noise already exists (xi): theta', v', the fp16 hi/lo (or bf16) operand form of
// theta' and the tile's row statistics.  Shared by the stand-alone update kernel
// (k_sgld_apply_split, update_kernels.cu) and the update phase at the tail of the potential
// kernel (k_glm_tc_pair, glm_tc.cu) so both produce the same bits.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "sgld_math.cuh"

namespace sgmc {

__device__ __forceinline__ float4 ld4(const float* p, int64_t i) {
  return *reinterpret_cast<const float4*>(p + i);
}
__device__ __forceinline__ void st4(float* p, int64_t i, float4 v) {
  *reinterpret_cast<float4*>(p + i) = v;
}
// L2 eviction-priority hints (createpolicy + .L2::cache_hint): the per-step working set
// of the carried Langevin step (~120 MB) is about the size of the L2, so the persistent
// state (theta, v, the operand split) asks to stay and the transient streams (gradient,
// noise) give their lines up at their last read.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ld4_hint(const float* p, int64_t i, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i), "l"(pol));
  return v;
}
__device__ __forceinline__ void st4_hint(float* p, int64_t i, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%4], {%0,%1,%2,%3}, %5;" ::"f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "l"(p + i), "l"(pol) : "memory");
}
__device__ __forceinline__ void st2u_hint(void* p, uint2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.u32 [%2], {%0,%1}, %3;" ::"r"(v.x), "r"(v.y), "l"(p),
               "l"(pol) : "memory");
}

// What a tile needs (plain data: also a kernel parameter of k_glm_tc_pair).
struct ApplyTileArgs {
  float* theta; float* v; const float* grad; const float* xi;
  void* th_hi; __half* th_lo;          // operand form of theta' (fp16 hi / lo, or bf16 in th_hi)
  const float* scale;                  // f32[C]: power-of-two scale of this update's split
  uint32_t* amax_bits;                 // u32[C]: max |theta'| per row (atomicMax on the bits)
  float* sumsq_part;                   // f32[C][tiles_per_chain]
  int64_t P;
  uint32_t tiles_per_chain;
  int prior_on;                        // the gaussian prior covers the whole sample
  float prior_coef;                    // theta * coef completes the gradient (0: nothing to add)
  float noise_scale, neg_eps, alpha, one_m_alpha, lmbd;
};

// fp16 hi/lo (FMT 1) or bf16 (FMT 2) of four scaled values -> global memory
template <int FMT>
__device__ __forceinline__ void emit_split4(const ApplyTileArgs& a, int64_t i, const float4& t,
                                            float s, uint64_t pol) {
  const float v0 = t.x * s, v1 = t.y * s, v2 = t.z * s, v3 = t.w * s;
  if (FMT == 1) {
    const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y);
    const __half2 l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
    uint2 ph, pl;
    ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
    pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
    st2u_hint(reinterpret_cast<__half*>(a.th_hi) + i, ph, pol);
    st2u_hint(a.th_lo + i, pl, pol);
  } else {
    const __nv_bfloat162 b01 = __floats2bfloat162_rn(v0, v1), b23 = __floats2bfloat162_rn(v2, v3);
    uint2 pb;
    pb.x = *reinterpret_cast<const uint32_t*>(&b01); pb.y = *reinterpret_cast<const uint32_t*>(&b23);
    st2u_hint(reinterpret_cast<__nv_bfloat16*>(a.th_hi) + i, pb, pol);
  }
}

// The whole tile [256 t, 256 t + 256) of chain c lies inside the sample and the prior range
// (or there is no prior): all eight loads first, no per-element range tests.
template <bool RMS, bool FAST, int FMT>
__device__ __forceinline__ void apply_tile_fast(const ApplyTileArgs& a, int64_t c, uint32_t t,
                                                int lane, uint64_t keep, uint64_t drop) {
  const float ns = a.noise_scale;
  const float s = FMT == 1 ? __ldg(a.scale + c) : 1.0f;
  const int64_t i0 = c * a.P + t * 256u + (uint32_t)lane * 4u, i1 = i0 + 128;
  const float4 th0 = ld4_hint(a.theta, i0, keep), th1 = ld4_hint(a.theta, i1, keep);
  float4 g0 = ld4_hint(a.grad, i0, drop), g1 = ld4_hint(a.grad, i1, drop);
  const float4 x0 = ld4_hint(a.xi, i0, drop), x1 = ld4_hint(a.xi, i1, drop);
  float4 v0 = RMS ? ld4_hint(a.v, i0, keep) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 v1 = RMS ? ld4_hint(a.v, i1, keep) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float pc = a.prior_coef;
  if (pc != 0.f) {
    g0.x = fmaf(th0.x, pc, g0.x); g0.y = fmaf(th0.y, pc, g0.y);
    g0.z = fmaf(th0.z, pc, g0.z); g0.w = fmaf(th0.w, pc, g0.w);
    g1.x = fmaf(th1.x, pc, g1.x); g1.y = fmaf(th1.y, pc, g1.y);
    g1.z = fmaf(th1.z, pc, g1.z); g1.w = fmaf(th1.w, pc, g1.w);
  }
  auto one = [&](float tv, float gv, float& vv, float xv) {
    return sgld_one<RMS, FAST>(tv, gv, vv, xv, ns, a.neg_eps, a.alpha, a.one_m_alpha, a.lmbd);
  };
  float4 o0, o1;
  o0.x = one(th0.x, g0.x, v0.x, x0.x); o0.y = one(th0.y, g0.y, v0.y, x0.y);
  o0.z = one(th0.z, g0.z, v0.z, x0.z); o0.w = one(th0.w, g0.w, v0.w, x0.w);
  o1.x = one(th1.x, g1.x, v1.x, x1.x); o1.y = one(th1.y, g1.y, v1.y, x1.y);
  o1.z = one(th1.z, g1.z, v1.z, x1.z); o1.w = one(th1.w, g1.w, v1.w, x1.w);
  st4_hint(a.theta, i0, o0, keep);
  st4_hint(a.theta, i1, o1, keep);
  if (RMS) {
    st4_hint(a.v, i0, v0, keep);
    st4_hint(a.v, i1, v1, keep);
  }
  emit_split4<FMT>(a, i0, o0, s, keep);
  emit_split4<FMT>(a, i1, o1, s, keep);
  float amax = fmaxf(fmaxf(fmaxf(fabsf(o0.x), fabsf(o0.y)), fmaxf(fabsf(o0.z), fabsf(o0.w))),
                     fmaxf(fmaxf(fabsf(o1.x), fabsf(o1.y)), fmaxf(fabsf(o1.z), fabsf(o1.w))));
  // per-tile sum of squares over the prior range, fixed association (it feeds the prior
  // value of the next potential bit for bit)
  float sum = 0.f;
  if (a.prior_on)
    sum = ((o0.x * o0.x + o0.y * o0.y) + (o0.z * o0.z + o0.w * o0.w)) +
          ((o1.x * o1.x + o1.y * o1.y) + (o1.z * o1.z + o1.w * o1.w));
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, k);
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, k));
  }
  if (lane == 0) {
    a.sumsq_part[c * a.tiles_per_chain + t] = sum;
    atomicMax(a.amax_bits + c, __float_as_uint(amax));   // amax >= 0: uint order == float order
  }
}

}  // namespace sgmc
