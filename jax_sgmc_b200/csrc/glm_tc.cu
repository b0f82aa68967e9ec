// GLM stochastic potential + gradient on the 5th-generation tensor cores
// (tcgen05 / TMEM / TMA), paths 1 and 2 of sgmc_glm_potential_grad.
//
// Replaces potential.minibatch_potential + its reverse-mode gradient
// (jax_sgmc/potential.py:159-214; integrator.py:166,593,792) for the logistic
// family when many chains share one minibatch (the reference default: every
// chain's data key is PRNGKey(0), data/numpy_loader.py:124).  The work is two
// dense contractions chains x features x batch:
//     GEMM1  Z[C,n] = Theta[C,d] . Xb[n,d]^T     epilogue: ell, R = cot*dl/dz
//     GEMM2  G[C,d] = R[C,n]     . Xb[n,d]       epilogue: + prior gradient
// path 1 ("parity"): every f32 operand x is split as x*s = hi + lo with fp16
//   hi, lo (s a power of two: per chain row for Theta, per tensor for Xb and
//   R); three MMAs per k-step (hi*hi + hi*lo + lo*hi) accumulate in fp32 in
//   TMEM -> ~2^-22 relative accuracy (the dropped lo*lo term), i.e. fp32-level
//   parity with the reference at 3 tensor-core passes.
// path 2 ("throughput"): single bf16 pass.
//
// Kernel anatomy (one 128x256 output tile per CTA, K streamed in 64-element
// blocks): warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier
// expect_tx), warp 1 = single-thread tcgen05.mma issuer (accumulator 128 lanes
// x 256 fp32 columns in TMEM), warp 2 = TMEM allocator; afterwards all 16 warps
// run the epilogue: tcgen05.ld 32x32b.x32 -> registers -> link function /
// per-row likelihood statistics -> 32x32 transpose through the (now idle)
// pipeline shared memory -> fully coalesced global stores.
// Roofline: tensor pipe; algorithmic FLOPs 2*C*n*d per GEMM.
#include "glm.cuh"
#include "noise_pass.cuh"
#include "sgld_math.cuh"
#include "sgld_apply_tile.cuh"
#include "sgld_split.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <mutex>
#include <vector>

namespace sgmc {

#ifndef SGMC_TC_BK
#define SGMC_TC_BK 64
#endif
constexpr int BM = 128, BN = 256, BK = SGMC_TC_BK;   // tile (elements); BK*2 B = swizzle span
static_assert(BK == 64 || BK == 32, "BK must be 64 (128B swizzle) or 32 (64B swizzle)");
constexpr int kTcThreads = 512;               // 16 warps
constexpr int kTcWarps = kTcThreads / 32;
// EPI 2 (gradient + fused SGLD update): the 16 warps generate the step's Gaussian
// noise while the tensor pipe works, so the pipeline (TMA + MMA issue) moves to
// one thread of a 17th warp.
constexpr int kTcFusedThreads = kTcThreads + 32;
constexpr uint32_t kTmemCols = 256;

// Phase timers (globaltimer ns) of CTA (0,0), thread 64 -- compiled in only with
// -DSGMC_TC_DEBUG (SGMC_TC_DEBUG=1 python -m jax_sgmc_b200.build), read back with
// sgmc_debug_tc_timers (tools/dbg_tc_timers.py, tools/dbg_fused_timers.py).
__device__ unsigned long long g_tc_dbg[10];
#ifdef SGMC_TC_DEBUG
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TC_DBG(cond, i) do { if (cond) g_tc_dbg[i] = gtime(); } while (0)
#else
#define TC_DBG(cond, i) do { } while (0)
#endif

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return make_smem_desc_k<BK>(smem_addr);
}

// ---------------------------------------------------------------------------
// Epilogues
// ---------------------------------------------------------------------------
// Per-row partial likelihood statistics of one CTA tile (combined by the
// last-arriver block of the link epilogue): {count, mean, M2, masked_sum}.
constexpr int kStatFields = 4;

struct TcLinkEpi {      // GEMM1: z -> ell statistics, R = cot*mask*dl/dz as fp16 hi/lo
  const float* theta; int64_t P; int aux_off;
  const float* y; const int32_t* idx; const float* mask;
  const float* row_scale;   // f32[C]: scale applied to Theta rows (1 for bf16)
  const float* b_scale;     // device scalar: scale applied to Xb
  float cot;
  float r_scale;            // scale applied to R before the fp16 split
  float* ell;               // f32[C][n] or null (only when the caller wants it)
  float* stats;             // f32[C][parts][4]
  int parts;
  // last-arriver finalisation (U, var) per 128-row block
  uint32_t* counters;       // u32[2 ceil(C/128)], zeroed by k_prepare_all
  const float* row_sumsq;   // f32[C]: sum theta^2 over the gaussian-prior range
  float* potential; float* variance;
  float n_obs, inv_temperature, prior_half_inv;   // N, 1/T, 0.5/scale^2 (0: no gaussian prior)
  __half* r_hi; __half* r_lo;          // fp16[C][n]  (path 1)
  __nv_bfloat16* r_bf;                 // bf16[C][n]  (path 2)
  int C, n;
};

struct TcGradEpi {      // GEMM2: G -> grad (adds -grad(prior)/T)
  const float* theta; float* grad; int64_t P; int w_off; int d; int C;
  int prior_lo, prior_hi;   // flat index range of the gaussian prior (empty if none)
  float prior_coef;         // 1 / (scale^2 * T)
  const float* xt_scale;    // device scalar: scale applied to XbT
  float r_scale;            // scale applied to R
  // EPI 2: the SGLD / pSGLD update of integrator.py:860-922 applied in the epilogue
  float* theta_rw;          // f32[C][P], updated in place
  float* v;                 // f32[C][P] RMSprop state or null (plain SGLD)
  const uint32_t* keys_in;  // u32[C][2]
  uint32_t* keys_out;       // u32[C][2]
  float neg_eps, noise_scale, alpha, one_m_alpha, lmbd;
  uint32_t half;            // d / 2: element j shares its threefry block with j + half
};

// Side job of the GEMM kernels for sgmc_glm_sgld_step: while the tensor pipe
// works, the otherwise idle warps 2..15 generate part of the step's Gaussian
// noise (integrator.random_tree, single leaf of d elements per chain) into
// xi[C][d]; k_sgld_apply then consumes it.  Work unit = a warp-tile of
// k_noise_pass: 32 groups x 4 threefry pairs = 256 elements of one chain.
struct TcNoiseJob {
  float* xi;                // f32[C][d] or null (no job)
  const uint32_t* keys_in;  // u32[C][2]
  uint32_t* keys_out;       // u32[C][2]: split(key)[0], written by the tile owner
  int tile0, tile_end;      // this launch's range of warp-tiles
  int tiles_per_cta;
  int d;
};
constexpr int kNoiseJobWarps = kTcWarps - 2;
constexpr int kNoiseJobMaxChains = 128;

// logistic link with SFU-based exp / log / reciprocal (abs. error ~1e-7):
//   ell = y z - softplus(z),  dz = y - sigmoid(z)
__device__ __forceinline__ void logistic_link_fast(float z, float y, float& ell, float& dz) {
  // e = exp(-|z|) in (0, 1], den = 1 + e in (1, 2]: the flush-to-zero SFU forms need
  // no sub-normal fix-up code around them
  float e, lg, rden;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(z) * -1.4426950408889634f));
  const float den = 1.0f + e;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(den));
  // log(1+e) = lg2(den) * ln 2: abs. error <= 1 ulp(1) = 6e-8 (|ell| >= 7 when e < 1e-3)
  ell = y * z - fmaf(lg, 0.6931471805599453f, fmaxf(z, 0.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rden) : "f"(den));
  dz = y - (z >= 0.0f ? rden : 1.0f - rden);
}

// ---------------------------------------------------------------------------
// The GEMM kernel.  TERMS = 3: operands (A_hi, A_lo) x (B_hi, B_lo), MMAs
// lo*hi + hi*lo + hi*hi.  TERMS = 1: single operand pair.
// EPI = 0: link epilogue, EPI = 1: gradient epilogue.
// ---------------------------------------------------------------------------
template <int TERMS>
struct TcSmem {
  static constexpr int kNA = TERMS == 3 ? 2 : 1;
  static constexpr int kStageBytes = kNA * (BM * BK * 2) + kNA * (BN * BK * 2);
  static constexpr int kStages = (TERMS == 3 ? 2 : 4) * (64 / BK);
  static constexpr int kPipeBytes = kStages * kStageBytes;
  static constexpr int kAuxBytes = 256 /*barriers*/ + 3 * BN * 4 /*y, mask, rm*/ +
                                   4 * BM * kStatFields * 4 /*row stats*/ +
                                   kNoiseJobMaxChains * 8 /*noise-job keys*/;
  static constexpr int kBytes = kPipeBytes + kAuxBytes + 1024 /*alignment slack*/;
  static_assert(kPipeBytes >= kTcWarps * 32 * 33 * 4, "staging must fit in the pipeline smem");
};

template <int TERMS, int EPI, int ABFMT>
__global__ void __launch_bounds__(EPI == 2 ? kTcFusedThreads : kTcThreads, 1)
k_glm_tc_gemm(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
              const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
              int num_k_blocks, const TcLinkEpi link, const TcGradEpi gradp,
              const TcNoiseJob job) {
  using S = TcSmem<TERMS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kPipeBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full_bar = empty_bar + S::kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_y = reinterpret_cast<float*>(smem + S::kPipeBytes + 256);
  float* s_mask = s_y + BN;
  float* s_rm = s_mask + BN;                       // cot * mask * r_scale per column
  float* s_stats = s_rm + BN;                      // [4 col groups][BM][4]
  Key* s_nk = reinterpret_cast<Key*>(s_stats + 4 * BM * kStatFields);   // noise-job keys

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // EPI 2 tiles pair feature j with j + d/2 (the two outputs of one threefry
  // block): accumulator columns [0,128) = features n0a.., [128,256) = n0b..
  const int n0a = blockIdx.x * (BN / 2), n0b = (int)gradp.half + n0a;
  Key* s_lk = reinterpret_cast<Key*>(s_y);         // EPI 2: per-row noise keys
  TC_DBG(blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (TERMS == 3) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, kTmemCols);
  // Everything above (descriptor prefetch, barrier init, TMEM allocation) overlaps
  // the previous kernel's tail under programmatic dependent launch.
  pdl_launch_dependents();
  pdl_wait();
  if (EPI == 0 && threadIdx.x < BN) {   // per-column observation data of this tile
    const int col = n0 + threadIdx.x;
    float yv = 0.f, mv = 0.f;
    if (col < link.n) {
      yv = link.y[link.idx ? link.idx[col] : col];
      mv = link.mask ? link.mask[col] : 1.0f;
    }
    s_y[threadIdx.x] = yv;
    s_mask[threadIdx.x] = mv;
    s_rm[threadIdx.x] = link.cot * mv * link.r_scale;
  }
  // noise job: tiles [jt0, jt1) of this CTA, keys of the chains they touch
  int jt0 = 0, jt1 = 0, jtpc = 1, jc_lo = 0;
  if (EPI != 2 && job.xi != nullptr) {
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    jtpc = job.d / 256;
    jt0 = min(job.tile0 + cta * job.tiles_per_cta, job.tile_end);
    jt1 = min(jt0 + job.tiles_per_cta, job.tile_end);
    if (jt0 < jt1) {
      jc_lo = jt0 / jtpc;
      const int n_ch = (jt1 - 1) / jtpc - jc_lo + 1;
      if ((int)threadIdx.x < n_ch) {
        const int64_t c = jc_lo + threadIdx.x;
        const Key k{job.keys_in[2 * c], job.keys_in[2 * c + 1]};
        Key newk, sub;
        split2(k, 0, newk, sub);                       // integrator.py:871
        s_nk[threadIdx.x] = split_key(sub, 0u, 1u, 0); // random_tree: split(sub, 1)[0]
        if (c * jtpc >= jt0) {                         // owner of the chain's first tile
          job.keys_out[2 * c] = newk.k0;
          job.keys_out[2 * c + 1] = newk.k1;
        }
      }
    }
  }
  if (EPI == 2 && threadIdx.x < BM) {
    // key', sub = split(key); noise key of the single leaf = split(sub, 1)[0]
    // (integrator.py:871, :131-133)
    const int64_t c = m0 + threadIdx.x;
    const Key k{gradp.keys_in[2 * c], gradp.keys_in[2 * c + 1]};
    Key newk, sub;
    split2(k, 0, newk, sub);
    s_lk[threadIdx.x] = split_key(sub, 0u, 1u, 0);
    if (blockIdx.x == 0) {
      gradp.keys_out[2 * c] = newk.k0;
      gradp.keys_out[2 * c + 1] = newk.k1;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const bool dbg = blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64;
  TC_DBG(dbg, 1);

  float nzA[32], nzB[32];   // EPI 2: noise of (row q*32+r, feature n0a/n0b + cg*32 + lane)
  if (EPI == 2) {
    if (warp == kTcWarps) {
      if (lane == 0) {
        // ===== TMA producer + MMA issuer in one thread (both only enqueue work) =====
        constexpr uint32_t idesc = make_idesc(ABFMT, BM, BN);
        auto load_block = [&](int kb) {
          const int s = kb % S::kStages;
          uint8_t* st = tiles + s * S::kStageBytes;
          mbar_expect_tx(&full_bar[s], S::kStageBytes);
          tma_load_2d(st, &tmA0, &full_bar[s], kb * BK, m0);
          if (TERMS == 3) tma_load_2d(st + BM * BK * 2, &tmA1, &full_bar[s], kb * BK, m0);
          uint8_t* sb = st + S::kNA * BM * BK * 2;
          tma_load_2d(sb, &tmB0, &full_bar[s], kb * BK, n0a);
          tma_load_2d(sb + (BN / 2) * BK * 2, &tmB0, &full_bar[s], kb * BK, n0b);
          if (TERMS == 3) {
            tma_load_2d(sb + BN * BK * 2, &tmB1, &full_bar[s], kb * BK, n0a);
            tma_load_2d(sb + BN * BK * 2 + (BN / 2) * BK * 2, &tmB1, &full_bar[s], kb * BK, n0b);
          }
        };
        for (int kb = 0; kb < S::kStages && kb < num_k_blocks; ++kb) load_block(kb);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          const int s = kb % S::kStages;
          mbar_wait(&full_bar[s], (kb / S::kStages) & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(tiles + s * S::kStageBytes);
          const uint32_t a1 = a0 + BM * BK * 2;
          const uint32_t b0 = a0 + S::kNA * BM * BK * 2;
          const uint32_t b1 = b0 + BN * BK * 2;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da0 = make_smem_desc(a0 + k * 32), db0 = make_smem_desc(b0 + k * 32);
            if (TERMS == 3) {
              const uint64_t da1 = make_smem_desc(a1 + k * 32), db1 = make_smem_desc(b1 + k * 32);
              umma_f16(tmem_base, da1, db0, idesc, (kb | k) != 0);
              umma_f16(tmem_base, da0, db1, idesc, 1);
              umma_f16(tmem_base, da0, db0, idesc, 1);
            } else {
              umma_f16(tmem_base, da0, db0, idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty_bar[s]);
          // refill the stage of the previous block once its MMAs have retired
          const int kn = kb - 1 + S::kStages;
          if (kb >= 1 && kn < num_k_blocks) {
            mbar_wait(&empty_bar[(kb - 1) % S::kStages], ((kb - 1) / S::kStages) & 1);
            load_block(kn);
          }
        }
        umma_commit(tmem_full_bar);
      }
    } else {
      // ===== the step's Gaussian noise, generated under the mainloop =====
      // lane = feature inside the warp's 32-column slice, r = chain row: the
      // layout the coalesced epilogue below consumes.
      const int q_ = warp & 3, cg_ = warp >> 2;
      const uint32_t jA = (uint32_t)(n0a + cg_ * 32 + lane);
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const Key lk = s_lk[q_ * 32 + r];
        uint32_t wa = jA, wb = jA + gradp.half;
        threefry2x32(lk, wa, wb);
        NormalPartial pa, pb;
        nzA[r] = normal_main(wa, pa);
        nzB[r] = normal_main(wb, pb);
        if (normal_is_tail(pa) | normal_is_tail(pb)) {
          if (normal_is_tail(pa)) nzA[r] = normal_tail(pa);
          if (normal_is_tail(pb)) nzB[r] = normal_tail(pb);
        }
      }
      TC_DBG(dbg, 8);
    }
  } else if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    for (int kb = 0; kb < num_k_blocks; ++kb) {
      const int s = kb % S::kStages;
      const uint32_t ph = (kb / S::kStages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* st = tiles + s * S::kStageBytes;
      mbar_expect_tx(&full_bar[s], S::kStageBytes);
      tma_load_2d(st, &tmA0, &full_bar[s], kb * BK, m0);
      if (TERMS == 3) tma_load_2d(st + BM * BK * 2, &tmA1, &full_bar[s], kb * BK, m0);
      uint8_t* sb = st + S::kNA * BM * BK * 2;
      tma_load_2d(sb, &tmB0, &full_bar[s], kb * BK, n0);
      if (TERMS == 3) tma_load_2d(sb + BN * BK * 2, &tmB1, &full_bar[s], kb * BK, n0);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(ABFMT, BM, BN);
    for (int kb = 0; kb < num_k_blocks; ++kb) {
      const int s = kb % S::kStages;
      const uint32_t ph = (kb / S::kStages) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t a0 = smem_u32(tiles + s * S::kStageBytes);
      const uint32_t a1 = a0 + BM * BK * 2;
      const uint32_t b0 = a0 + S::kNA * BM * BK * 2;
      const uint32_t b1 = b0 + BN * BK * 2;
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) {
        // advancing K by 16 elements (32 B) inside the 128 B swizzle atom
        const uint64_t da0 = make_smem_desc(a0 + k * 32), db0 = make_smem_desc(b0 + k * 32);
        if (TERMS == 3) {
          const uint64_t da1 = make_smem_desc(a1 + k * 32), db1 = make_smem_desc(b1 + k * 32);
          umma_f16(tmem_base, da1, db0, idesc, (kb | k) != 0);   // lo*hi
          umma_f16(tmem_base, da0, db1, idesc, 1);               // hi*lo
          umma_f16(tmem_base, da0, db0, idesc, 1);               // hi*hi
        } else {
          umma_f16(tmem_base, da0, db0, idesc, (kb | k) != 0);
        }
      }
      umma_commit(&empty_bar[s]);          // frees the smem stage when the MMAs retire
    }
    umma_commit(tmem_full_bar);            // accumulator complete
  }

  // ===== noise job of the idle warps, under the mainloop =====
  if (EPI != 2 && warp >= 2 && jt0 < jt1) {
    const uint32_t half = (uint32_t)job.d >> 1;
    for (int t = jt0 + (warp - 2); t < jt1; t += kNoiseJobWarps) {
      const int c = t / jtpc;
      const uint32_t j0 = (uint32_t)((t - c * jtpc) * 32 + lane) * 4u;
      float nA[4], nB[4];
      group_noise<0, true>(s_nk[c - jc_lo], j0, half, (uint32_t)job.d, nA, nB);
      float* row = job.xi + (int64_t)c * job.d;
      *reinterpret_cast<float4*>(row + j0) = make_float4(nA[0], nA[1], nA[2], nA[3]);
      *reinterpret_cast<float4*>(row + half + j0) = make_float4(nB[0], nB[1], nB[2], nB[3]);
    }
  }

  // ===== epilogue: all 16 warps =====
  __syncwarp();
  if (EPI != 2 || warp < kTcWarps) {
  const int q = warp & 3;                  // TMEM lane quarter this warp may access
  const int cg = warp >> 2;                // column group: 64 columns
  // GEMM2: the prior-gradient operand theta[row][col] does not depend on the
  // accumulator -> fetch it (coalesced, 64 loads in flight per thread) while the
  // tensor pipe is still busy, so the epilogue itself only stores.
  float thp[2][32];
  if (EPI == 1) {
#pragma unroll
    for (int c2 = 0; c2 < 2; ++c2) {
      const int col0 = n0 + cg * 64 + c2 * 32, row0 = m0 + q * 32;
      const int p = gradp.w_off + col0 + lane;
      const bool want = (row0 + 32 <= gradp.C) && (col0 + 32 <= gradp.d) &&
                        p >= gradp.prior_lo && p < gradp.prior_hi;
      const float* tp = gradp.theta + (int64_t)row0 * gradp.P + p;
#pragma unroll
      for (int r = 0; r < 32; ++r) thp[c2][r] = want ? __ldg(tp + (int64_t)r * gradp.P) : 0.f;
    }
  }
  mbar_wait(tmem_full_bar, 0);
  tc_fence_after();
  TC_DBG(dbg, 2);
  // All TMA writes have landed and all MMAs have read them: the pipeline smem
  // is free and becomes 16 private 32x33 f32 transpose buffers.
  float* stage = reinterpret_cast<float*>(tiles) + warp * (32 * 33);
  const int row = m0 + q * 32 + lane;      // this thread's accumulator row
  const int rsub = lane >> 4, csub = (lane & 15) * 2;   // packed 2-row store mapping

  if (EPI == 0) {
    const float inv = row < link.C ? 1.0f / (link.row_scale[row] * __ldg(link.b_scale)) : 0.f;
    const float bias = (link.aux_off >= 0 && row < link.C)
                           ? link.theta[(int64_t)row * link.P + link.aux_off] : 0.0f;
    float cnt = 0.f, shift = 0.f, s1 = 0.f, s2 = 0.f, sm = 0.f;
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      const int ct = cg * 64 + c;                       // column inside the tile
      const int col0 = n0 + ct;
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ct, acc);
      TC_DBG(dbg && c == 0, 4);
      if (col0 >= link.n) continue;                      // warp-uniform
      float ellv[32];
      const bool full_cols = col0 + 32 <= link.n;        // warp-uniform
      if (full_cols) {
        // interior chunk: no per-element bounds logic; the running shift of the
        // one-pass variance is the first likelihood this thread sees
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float z = fmaf(__uint_as_float(acc[j]), inv, bias);
          float l, dz;
          logistic_link_fast(z, s_y[ct + j], l, dz);
          ellv[j] = l;
          stage[lane * 33 + j] = dz * s_rm[ct + j];
        }
        if (cnt == 0.f) shift = ellv[0];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float dl = ellv[j] - shift;
          s1 += dl;
          s2 = fmaf(dl, dl, s2);
          sm = fmaf(ellv[j], s_mask[ct + j], sm);
        }
        cnt += 32.f;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float z = fmaf(__uint_as_float(acc[j]), inv, bias);
          float l, dz;
          logistic_link_fast(z, s_y[ct + j], l, dz);
          const float m = s_mask[ct + j];
          ellv[j] = l;
          stage[lane * 33 + j] = dz * s_rm[ct + j];
          if (col0 + j < link.n) {
            if (cnt == 0.f) shift = l;
            const float dl = l - shift;
            cnt += 1.f; s1 += dl; s2 = fmaf(dl, dl, s2); sm = fmaf(l, m, sm);
          }
        }
      }
      __syncwarp();
      TC_DBG(dbg && c == 0, 5);
      // coalesced R stores: two rows per instruction, two columns per lane
      const int64_t ro0 = (int64_t)(m0 + q * 32 + rsub) * link.n + col0 + csub;
      const bool full_tile = (m0 + q * 32 + 32 <= link.C) && (col0 + 32 <= link.n);
#pragma unroll
      for (int r = 0; r < 32; r += 2) {
        const int rr = r + rsub, grow = m0 + q * 32 + rr, gcol = col0 + csub;
        const float v0 = stage[rr * 33 + csub], v1 = stage[rr * 33 + csub + 1];
        if (full_tile || (grow < link.C && gcol < link.n)) {   // n % 8 == 0: pairs never straddle
          const int64_t o = ro0 + (int64_t)r * link.n;
          if (TERMS == 3) {
            const __half2 h = __floats2half2_rn(v0, v1);
            const float2 hf = __half22float2(h);
            *reinterpret_cast<__half2*>(link.r_hi + o) = h;
            *reinterpret_cast<__half2*>(link.r_lo + o) = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
          } else {
            *reinterpret_cast<__nv_bfloat162*>(link.r_bf + o) = __floats2bfloat162_rn(v0, v1);
          }
        }
      }
      TC_DBG(dbg && c == 0, 6);
      if (link.ell) {                                    // optional per-observation output
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = ellv[j];
        __syncwarp();
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const int grow = m0 + q * 32 + r;
          if (grow < link.C && col0 + lane < link.n)
            link.ell[(int64_t)grow * link.n + col0 + lane] = stage[r * 33 + lane];
        }
      }
      __syncwarp();
    }
    TC_DBG(dbg, 7);
    // per-row statistics of this warp's 64 columns -> smem -> combine 4 groups
    {
      float mean = 0.f, m2 = 0.f;
      if (cnt > 0.f) {
        mean = shift + s1 / cnt;
        m2 = fmaxf(s2 - s1 * s1 / cnt, 0.f);
      }
      float* st = s_stats + ((cg * BM) + q * 32 + lane) * kStatFields;
      st[0] = cnt; st[1] = mean; st[2] = m2; st[3] = sm;
    }
    __syncthreads();
    if (threadIdx.x < BM && m0 + threadIdx.x < link.C) {
      float n_t = 0.f, mean_t = 0.f, m2_t = 0.f, sm_t = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {                      // Chan et al. pairwise update
        const float* st = s_stats + ((g * BM) + threadIdx.x) * kStatFields;
        const float nb = st[0];
        if (nb > 0.f) {
          const float nn = n_t + nb, delta = st[1] - mean_t;
          mean_t += delta * (nb / nn);
          m2_t += st[2] + delta * delta * (n_t * nb / nn);
          n_t = nn;
        }
        sm_t += st[3];
      }
      float* o = link.stats + ((int64_t)(m0 + threadIdx.x) * link.parts + blockIdx.x) * kStatFields;
      o[0] = n_t; o[1] = mean_t; o[2] = m2_t; o[3] = sm_t;
    }
    // The CTA that completes a 128-row block last combines its column tiles into
    // U = (L - prior)/T (potential.py:183-185, :210) and var(ell) (integrator.py:880).
    __shared__ int s_is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
      s_is_last = atomicAdd(&link.counters[blockIdx.y], 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_is_last && threadIdx.x < BM && m0 + threadIdx.x < link.C) {
      __threadfence();
      const int c = m0 + threadIdx.x;
      float n_t = 0.f, mean_t = 0.f, m2_t = 0.f, sm_t = 0.f;
      for (int g = 0; g < link.parts; ++g) {
        const float4 st = __ldcg(reinterpret_cast<const float4*>(
            link.stats + ((int64_t)c * link.parts + g) * kStatFields));
        if (st.x > 0.f) {
          const float nn = n_t + st.x, delta = st.y - mean_t;
          mean_t += delta * (st.x / nn);
          m2_t += st.z + delta * delta * (n_t * st.x / nn);
          n_t = nn;
        }
        sm_t += st.w;
      }
      const float L = link.mask ? (-link.n_obs / (float)link.n) * sm_t : -link.n_obs * mean_t;
      const float prior = -link.prior_half_inv * link.row_sumsq[c];
      link.potential[c] = (L - prior) * link.inv_temperature;
      if (link.variance) link.variance[c] = m2_t / (float)link.n;
    }
  } else if (EPI == 2) {
    // Gradient tile -> SGLD / pSGLD update in place.  Per 32x32 block: accumulator
    // -> transpose through smem -> lane = feature, loop over the 32 chain rows with
    // coalesced theta / v accesses; the noise is already in registers.
    const float inv_scale = 1.0f / (gradp.r_scale * __ldg(gradp.xt_scale));
    const bool rms = gradp.v != nullptr;
    const int row0 = m0 + q * 32;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ct = h * (BN / 2) + cg * 32;                 // accumulator column
      const int p = (h ? n0b : n0a) + cg * 32 + lane;        // flat parameter index (w_off = 0)
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ct, acc);
      TC_DBG(dbg && h == 0, 4);
#pragma unroll
      for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = __uint_as_float(acc[j]) * inv_scale;
      __syncwarp();
      TC_DBG(dbg && h == 0, 5);
      const float coef = (p >= gradp.prior_lo && p < gradp.prior_hi) ? gradp.prior_coef : 0.f;
      const int64_t o0 = (int64_t)row0 * gradp.P + p;
      float* tp = gradp.theta_rw + o0;
      float* vp = rms ? gradp.v + o0 : nullptr;
      float* gp = gradp.grad ? gradp.grad + o0 : nullptr;
#pragma unroll
      for (int rb = 0; rb < 32; rb += 8) {
        float t8[8], v8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          t8[k] = tp[(int64_t)(rb + k) * gradp.P];
          v8[k] = rms ? vp[(int64_t)(rb + k) * gradp.P] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = rb + k;
          const float g = fmaf(t8[k], coef, stage[r * 33 + lane]);
          const float xi = h ? nzB[r] : nzA[r];
          float tn;
          if (rms) {
            tn = sgld_one<true, true>(t8[k], g, v8[k], xi, gradp.noise_scale, gradp.neg_eps,
                                      gradp.alpha, gradp.one_m_alpha, gradp.lmbd);
            vp[(int64_t)r * gradp.P] = v8[k];
          } else {
            tn = sgld_one<false, false>(t8[k], g, v8[k], xi, gradp.noise_scale, gradp.neg_eps,
                                        0.f, 0.f, 0.f);
          }
          tp[(int64_t)r * gradp.P] = tn;
          if (gp) gp[(int64_t)r * gradp.P] = g;
        }
      }
      __syncwarp();
      TC_DBG(dbg && h == 0, 6);
    }
  } else {
    const float inv_scale = 1.0f / (gradp.r_scale * __ldg(gradp.xt_scale));
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      const int ct = cg * 64 + c;
      const int col0 = n0 + ct;
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ct, acc);
      TC_DBG(dbg && c == 0, 4);
      if (col0 >= gradp.d) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = __uint_as_float(acc[j]) * inv_scale;
      __syncwarp();
      TC_DBG(dbg && c == 0, 5);
      const int gcol = col0 + lane;
      const int p = gradp.w_off + gcol;
      const bool in_prior = p >= gradp.prior_lo && p < gradp.prior_hi;
      const bool col_ok = gcol < gradp.d;
      const int row0 = m0 + q * 32;
      const int64_t o0 = (int64_t)row0 * gradp.P + p;
      const bool full_tile = (row0 + 32 <= gradp.C) && (col0 + 32 <= gradp.d);   // warp-uniform
      if (full_tile) {
        // interior tile: theta was prefetched before the accumulator wait
        float* gp = gradp.grad + o0;
        if (c == 0) {
#pragma unroll
          for (int r = 0; r < 32; ++r)
            gp[(int64_t)r * gradp.P] = fmaf(thp[0][r], gradp.prior_coef, stage[r * 33 + lane]);
        } else {
#pragma unroll
          for (int r = 0; r < 32; ++r)
            gp[(int64_t)r * gradp.P] = fmaf(thp[1][r], gradp.prior_coef, stage[r * 33 + lane]);
        }
      } else {
        for (int r = 0; r < 32; ++r) {
          const int grow = row0 + r;
          if (grow < gradp.C && col_ok) {
            const int64_t o = (int64_t)grow * gradp.P + p;
            float g = stage[r * 33 + lane];
            if (in_prior) g = fmaf(gradp.theta[o], gradp.prior_coef, g);
            gradp.grad[o] = g;
          }
        }
      }
      __syncwarp();
      TC_DBG(dbg && c == 0, 6);
    }
  }
  TC_DBG(dbg, 7);
  }  // epilogue warps
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
  TC_DBG(dbg, 3);
}

// ---------------------------------------------------------------------------
// The fused potential kernel on CTA pairs (the default tensor-core path).
//
// Why pairs: with fp16 hi/lo operands the mainloop of a 128 x 256 tile per SM
// asks the L2 for 96 KB per 64-wide k-block per SM -- 201 MB per GEMM at C2,
// which is what the LTS fabric delivers in the 17.7 us the round-1 kernel
// measured (~11.4 TB/s): that kernel was L2-bound, not MMA-bound.  Two SMs of a
// TPC running ONE tcgen05.mma.cta_group::2 of M = 256, N = 256 share the B
// operand (each loads half of it), so a pair moves 128 KB per k-block for twice
// the work: 134 MB per GEMM, and the tensor pipe becomes the limiter.
//
// ONE launch evaluates both contractions of the stochastic potential:
//   tiles [0, T1)        GEMM1  Z = Theta . Xb^T  -> link epilogue (ell stats, R)
//   tiles [T1, T1 + T2)  GEMM2  G = R . Xb        -> gradient epilogue
// A pair walks the tile list with stride (number of pairs), row-block-major, so
// a GEMM2 tile only ever waits for GEMM1 tiles EARLIER in the list: a per-row-
// block arrival counter (release / acquire + proxy fences) hands R from the
// epilogue's TMA stores to the TMA loads of the consumer.  Warp roles per CTA:
//   warps 0..15   epilogue: tcgen05.ld a 32x32 sub-block -> link / gradient math
//                 in registers -> swizzled 16-byte st.shared into a private 4 KB
//                 staging buffer -> one TMA store (the TMA engine writes whole
//                 rows; edges are clipped by the tensor map)
//   warp 16       TMA producer (one thread), 4-stage ring of 32-wide k-blocks
//   warp 17       tcgen05.mma issuer (one thread of the LEADER CTA) + TMEM allocator
// The accumulator is double-buffered in TMEM (2 x 256 columns).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Timeline of k_glm_tc_pair (globaltimer ns), one row of 16 stamps per pair; written only
// when PairSched::dbg != 0 (sgmc_set_option(SGMC_OPT_TC_TIMELINE, 1)), read back with
// sgmc_debug_pair_timeline (tools/r2_timeline.py).
constexpr int kDbgPairs = 80, kDbgSlots = 32;
__device__ unsigned long long g_pair_dbg[kDbgPairs * kDbgSlots];
__device__ __forceinline__ void pair_stamp(int enabled, int pair, int slot) {
  if (enabled && pair < kDbgPairs && slot < kDbgSlots) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_pair_dbg[pair * kDbgSlots + slot] = t;
  }
}

constexpr int kPrEpiWarps = 16;
constexpr int kPrEpiThreads = kPrEpiWarps * 32;
constexpr int kPrThreads = kPrEpiThreads + 64;
constexpr int kPrBK = 64;                       // k-block: 64 elements = 128-byte swizzle span
constexpr int kPrSwz = 128;                     // (TMA delivers a fixed number of box ROWS per
                                                // cycle: 128-byte rows halve the row count)
constexpr int kPrStageOut = 4096;               // epilogue staging bytes per warp

struct PairMaps {
  CUtensorMap a[2][2];    // loads, [gemm][hi / lo]: Theta (GEMM1), R (GEMM2); box 128 x 32
  CUtensorMap b[2][2];    // loads, [gemm][hi / lo]: Xb (GEMM1), XbT (GEMM2); box (BN/CG) x 32
  CUtensorMap r[2];       // stores: R hi / lo, box 32 x 32 halves (64-byte swizzle)
  CUtensorMap g;          // store: grad, box 32 x 32 floats (128-byte swizzle)
};

struct PairSched {
  int mt;                 // row blocks of 128 * CG rows
  int nt1, nt2;           // BN-column tiles of GEMM1 (over n) / GEMM2 (over d)
  int kb1, kb2;           // 32-wide k-blocks of GEMM1 (over d) / GEMM2 (over n)
  int tiles1, tiles_total;
  int dbg;                // record the timeline (g_pair_dbg)
};

// BN = accumulator columns per CTA and tile: 256 (one tile per pair and GEMM at C2: the
// least L2 traffic, but the epilogues are exposed) or 128 (two to four tiles per pair: every
// epilogue but the last runs under the next tile's mainloop).
template <int TERMS, int CG, int BN>
struct PrSmem {
  static_assert(BN == 128 || BN == 256, "tile width");
  static constexpr int kNA = TERMS == 3 ? 2 : 1;
  static constexpr int kBRows = BN / CG;
  static constexpr int kABytes = BM * kPrBK * 2;
  static constexpr int kBBytes = kBRows * kPrBK * 2;
  static constexpr int kStageBytes = kNA * (kABytes + kBBytes);
  // 192 KB of operand stages.  The epilogue's 64 KB of staging buffers ALIAS the last
  // kStgStages stages: the producer hands them over per tile (epi_done barrier), which
  // costs nothing when a pair runs one tile per GEMM (the epilogue and the next
  // mainloop cannot overlap there anyway: GEMM2 waits for R).
  static constexpr int kStages = 196608 / kStageBytes > 8 ? 8 : 196608 / kStageBytes;
  static constexpr int kPipeBytes = kStages * kStageBytes;
  static constexpr int kOutBytes = kPrEpiWarps * kPrStageOut;
  static constexpr int kStgStages = (kOutBytes + kStageBytes - 1) / kStageBytes;
  static constexpr int kAuxBytes = 256 /*barriers, tmem ptr, flags*/ + 2 * 3 * BN * 4 +
                                   2 * BM * 4 * 16 /*likelihood partials per (row, column group)*/;
  static constexpr int kBytes = kPipeBytes + kAuxBytes + 1024 /*alignment slack*/;
  static_assert(kStages > kStgStages, "need at least one stage that is never handed over");
  static_assert(2 * kStages + 5 <= 24, "barrier block is 256 bytes");
  static_assert(kBytes <= 232448, "exceeds the 227 KB of shared memory per CTA");
};

// Side job of the epilogue warps while the tensor pipe works (sgmc_glm_sgld_step in
// carried mode): the step's Gaussian noise (integrator.random_tree, one leaf of d
// elements per chain, original threefry layout) is generated into xi[C][d] from the
// per-chain noise keys k_prepare_all derived; k_sgld_apply_split consumes it.  The
// threefry + erf_inv arithmetic is issue-bound (about 120 instructions per normal)
// and would otherwise be the bulk of the update kernel; here it runs in issue slots
// the GEMM mainloops leave idle.  Unit of work: 32 consecutive elements j of a chain's
// first half and their block partners j + d/2 (64 normals, two coalesced 128-byte
// stores per warp).
struct PairNoise {
  float* xi;                 // f32[C][d] or null (no job)
  const uint32_t* noise_keys;
  int d;
  int units_per_chain;       // d / 64
  int units_total;           // C * units_per_chain
};

// Update phase at the tail of the launch (sgmc_glm_sgld_step in carried mode, opt-in
// SGMC_OPT_FUSED_PAIR_UPDATE): once EVERY CTA has stored its gradient tiles and its share of
// the noise (one grid-wide arrival counter; all CTAs are resident), the sixteen epilogue
// warps of every CTA run the pSGLD update of k_sgld_apply_split over warp-tiles dealt round
// robin -- same tile body (sgld_apply_tile.cuh), same bits, no second launch.
struct PairUpdate {
  ApplyTileArgs t;
  uint32_t* grid_counter;     // zeroed by k_prepare_all
  int64_t n_tiles;            // C * tiles_per_chain warp-tiles of 256 parameters
  int enabled, rms;
};

__device__ __forceinline__ void pair_noise_unit(const PairNoise& nz, int u, int lane) {
  const int c = u / nz.units_per_chain;
  const uint32_t half = (uint32_t)nz.d >> 1;
  const uint32_t j = (uint32_t)(u - c * nz.units_per_chain) * 32u + (uint32_t)lane;
  const Key lk{__ldg(nz.noise_keys + 2 * c), __ldg(nz.noise_keys + 2 * c + 1)};
  uint32_t wa = j, wb = j + half;
  threefry2x32(lk, wa, wb);
  NormalPartial pa, pb;
  float na = normal_main(wa, pa), nb = normal_main(wb, pb);
  if (normal_is_tail(pa) | normal_is_tail(pb)) {
    if (normal_is_tail(pa)) na = normal_tail(pa);
    if (normal_is_tail(pb)) nb = normal_tail(pb);
  }
  float* row = nz.xi + (int64_t)c * nz.d;
  row[j] = na;
  row[half + j] = nb;
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

template <int TERMS, int ABFMT, int CG, int BN>
__global__ void __launch_bounds__(kPrThreads, 1)
k_glm_tc_pair(const __grid_constant__ PairMaps maps, const PairSched sch,
              const TcLinkEpi link, const TcGradEpi gradp, const PairNoise nz,
              const PairUpdate upd) {
  using S = PrSmem<TERMS, CG, BN>;
  constexpr int kChunks = BN / 128;                // 32-column chunks per epilogue warp and tile
  // R is handed to GEMM2 per column HALF of a tile (256-wide tiles): the epilogue warps
  // work through columns [0, 128) first and publish them while they compute [128, 256)
  constexpr int kHalves = kChunks;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;
  // epilogue staging (16 x 4 KB) aliases the last stage(s) of the operand ring
  uint8_t* out_stage = smem + S::kPipeBytes - S::kOutBytes;
  uint8_t* aux = smem + S::kPipeBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* acc_full = empty_bar + S::kStages;      // [2]
  uint64_t* acc_empty = acc_full + 2;               // [2]
  uint64_t* epi_done = acc_empty + 2;               // staging buffers free again (per tile)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(aux + 224);
  int* s_last = reinterpret_cast<int*>(tmem_ptr + 1);   // [2]
  int* s_noise_next = s_last + 2;                       // next noise unit of this CTA
  float* s_col = reinterpret_cast<float*>(aux + 256);   // [2][3][BN]
  float4* s_part_all = reinterpret_cast<float4*>(aux + 256 + 2 * 3 * BN * 4);   // [2][BM][4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int pair = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  if (threadIdx.x == 0) pair_stamp(sch.dbg && leader, pair, 7);      // kernel entry
  const int n_pairs = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const bool have2 = sch.tiles_total > sch.tiles1;
  if (warp == kPrEpiWarps && lane == 0) {
    tma_prefetch_desc(&maps.a[0][0]);
    tma_prefetch_desc(&maps.b[0][0]);
    tma_prefetch_desc(&maps.r[0]);
    if (TERMS == 3) {
      tma_prefetch_desc(&maps.a[0][1]);
      tma_prefetch_desc(&maps.b[0][1]);
      tma_prefetch_desc(&maps.r[1]);
    }
    if (have2) {
      tma_prefetch_desc(&maps.a[1][0]);
      tma_prefetch_desc(&maps.b[1][0]);
      tma_prefetch_desc(&maps.g);
      if (TERMS == 3) {
        tma_prefetch_desc(&maps.a[1][1]);
        tma_prefetch_desc(&maps.b[1][1]);
      }
    }
  }
  // this CTA's share of the noise job
  const int nu0 = nz.xi ? (int)((int64_t)nz.units_total * blockIdx.x / gridDim.x) : 0;
  const int nu1 = nz.xi ? (int)((int64_t)nz.units_total * (blockIdx.x + 1) / gridDim.x) : 0;
  if (threadIdx.x == 0) *s_noise_next = nu0;
  if (warp == kPrEpiWarps + 1) {
    if (lane == 0) {
      for (int s = 0; s < S::kStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&acc_full[i], 1);
        mbar_init(&acc_empty[i], CG * kPrEpiWarps);
      }
      mbar_init(epi_done, kPrEpiWarps);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_cg<CG>(tmem_ptr, 2 * BN);
  }
  // Everything above overlaps the previous kernel's tail (programmatic dependent launch).
  pdl_launch_dependents();
  pdl_wait();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int dbg = sch.dbg && leader;
  if (threadIdx.x == 0) pair_stamp(dbg, pair, 0);

  if (warp == kPrEpiWarps) {
    // ===== TMA producer (both CTAs of the pair; bytes are credited to the leader's barrier) =====
    if (lane == 0) {
      uint32_t uses[S::kStages];                     // fills of each stage so far (phase = use & 1)
#pragma unroll
      for (int s = 0; s < S::kStages; ++s) uses[s] = 0;
      uint32_t it = 0;
      for (int t = pair; t < sch.tiles_total; t += n_pairs, ++it) {
        const int g2 = t >= sch.tiles1 ? 1 : 0;
        const int tt = g2 ? t - sch.tiles1 : t;
        const int ntl = g2 ? sch.nt2 : sch.nt1;
        const int rb = tt / ntl, tn = tt - rb * ntl;
        const int m0 = rb * (BM * CG) + (int)rank * BM;
        const int n0 = tn * BN + (int)rank * S::kBRows;
        const int kbs = g2 ? sch.kb2 : sch.kb1;
        // GEMM2 consumes R in the order the GEMM1 epilogues publish it: first the k-blocks
        // of the FIRST column half of every R tile, then those of the second halves
        // (kHalves = 2; with 128-wide tiles a tile is one half and the order is the natural one)
        constexpr int kKbTile = BN / kPrBK;            // k-blocks of GEMM2 per R tile
        constexpr int kKbHalf = kKbTile / kHalves;
        int count0 = kbs;
        if (g2 && kHalves == 2) {
          const int full = kbs / kKbTile, rem = kbs - full * kKbTile;
          count0 = full * kKbHalf + (rem < kKbHalf ? rem : kKbHalf);
        }
        if (g2) pair_stamp(dbg, pair, 2);
#pragma unroll 1
        for (int i = 0; i < kbs; ++i) {
          int kb = i;
          if (g2) {
            if (kHalves == 2) {
              const int h = i >= count0 ? 1 : 0, ii = i - h * count0;
              kb = (ii / kKbHalf) * kKbTile + h * kKbHalf + ii % kKbHalf;
            }
            if (i == 0 || (kHalves == 2 && i == count0)) {
              // every R tile of this row block has stored this column half (GEMM1
              // epilogues of tiles earlier in the list, on this pair or another one)
              const uint32_t* cnt = &link.counters[rb * kHalves + (i == 0 ? 0 : 1)];
              while (ld_acquire_gpu(cnt) < (uint32_t)(CG * sch.nt1)) __nanosleep(40);
              fence_proxy_async_global();
              if (i == 0) pair_stamp(dbg, pair, 3);
            }
          }
          const int s = i % S::kStages;                // the ring restarts with every tile
          if (it > 0 && i == s && s >= S::kStages - S::kStgStages)
            mbar_wait(epi_done, (it - 1) & 1);         // previous tile's epilogue left the staging
          uint32_t use = 0;
#pragma unroll
          for (int q = 0; q < S::kStages; ++q)
            if (q == s) { use = uses[q]; uses[q] = use + 1; }
          mbar_wait(&empty_bar[s], (use & 1) ^ 1);
          uint8_t* st = tiles + s * S::kStageBytes;
          uint8_t* sb = st + S::kNA * S::kABytes;
          if (CG == 2) {
            if (leader) mbar_expect_tx(&full_bar[s], 2 * S::kStageBytes);
            tma_load_2d_pair(st, &maps.a[g2][0], &full_bar[s], kb * kPrBK, m0);
            if (TERMS == 3) tma_load_2d_pair(st + S::kABytes, &maps.a[g2][1], &full_bar[s], kb * kPrBK, m0);
            tma_load_2d_pair(sb, &maps.b[g2][0], &full_bar[s], kb * kPrBK, n0);
            if (TERMS == 3) tma_load_2d_pair(sb + S::kBBytes, &maps.b[g2][1], &full_bar[s], kb * kPrBK, n0);
          } else {
            mbar_expect_tx(&full_bar[s], S::kStageBytes);
            tma_load_2d(st, &maps.a[g2][0], &full_bar[s], kb * kPrBK, m0);
            if (TERMS == 3) tma_load_2d(st + S::kABytes, &maps.a[g2][1], &full_bar[s], kb * kPrBK, m0);
            tma_load_2d(sb, &maps.b[g2][0], &full_bar[s], kb * kPrBK, n0);
            if (TERMS == 3) tma_load_2d(sb + S::kBBytes, &maps.b[g2][1], &full_bar[s], kb * kPrBK, n0);
          }
        }
      }
    }
  } else if (warp == kPrEpiWarps + 1) {
    // ===== MMA issuer: one thread of the leader CTA drives both SMs' tensor cores =====
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(ABFMT, BM * CG, BN);
      uint32_t uses[S::kStages];
#pragma unroll
      for (int s = 0; s < S::kStages; ++s) uses[s] = 0;
      uint32_t it = 0;
      for (int t = pair; t < sch.tiles_total; t += n_pairs, ++it) {
        const int kbs = t >= sch.tiles1 ? sch.kb2 : sch.kb1;
        const uint32_t par = it & 1, aph = (it >> 1) & 1;
        mbar_wait_cluster(&acc_empty[par], aph ^ 1);   // both CTAs' epilogues drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + par * BN;
#pragma unroll 1
        for (int kb = 0; kb < kbs; ++kb) {
          const int s = kb % S::kStages;
          uint32_t use = 0;
#pragma unroll
          for (int q = 0; q < S::kStages; ++q)
            if (q == s) { use = uses[q]; uses[q] = use + 1; }
          mbar_wait_cluster(&full_bar[s], use & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(tiles + s * S::kStageBytes);
          const uint32_t a1 = a0 + S::kABytes;
          const uint32_t b0 = a0 + S::kNA * S::kABytes;
          const uint32_t b1 = b0 + S::kBBytes;
#pragma unroll
          for (int k = 0; k < kPrBK / 16; ++k) {
            // advancing K by 16 elements (32 B) inside the 128 B swizzle atom
            const uint64_t da0 = make_smem_desc_k<kPrBK>(a0 + k * 32);
            const uint64_t db0 = make_smem_desc_k<kPrBK>(b0 + k * 32);
            if (TERMS == 3) {
              const uint64_t da1 = make_smem_desc_k<kPrBK>(a1 + k * 32);
              const uint64_t db1 = make_smem_desc_k<kPrBK>(b1 + k * 32);
              umma_f16_cg<CG>(d_tmem, da1, db0, idesc, (kb | k) != 0);   // lo*hi
              umma_f16_cg<CG>(d_tmem, da0, db1, idesc, 1);               // hi*lo
              umma_f16_cg<CG>(d_tmem, da0, db0, idesc, 1);               // hi*hi
            } else {
              umma_f16_cg<CG>(d_tmem, da0, db0, idesc, (kb | k) != 0);
            }
          }
          umma_commit_cg<CG>(&empty_bar[s]);   // frees the stage in both CTAs when the MMAs retire
        }
        umma_commit_cg<CG>(&acc_full[par]);    // accumulator complete (both CTAs)
        pair_stamp(dbg, pair, 4 + (it > 0 ? 1 : 0));   // last MMA of the tile ISSUED
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;                          // TMEM lane quarter
    const int cg = warp >> 2;                        // column group: BN / 4 columns
    uint8_t* sbuf = out_stage + warp * kPrStageOut;  // private staging buffer
    // wait for an accumulator; meanwhile generate noise units (64 normals each)
    auto wait_acc = [&](uint64_t* bar, uint32_t parity) {
      if (nz.xi != nullptr) {
        for (;;) {
          int done = 0, u = 0;
          if (lane == 0) {
            done = mbar_test(bar, parity) ? 1 : 0;
            if (!done) u = atomicAdd(s_noise_next, 1);
          }
          done = __shfl_sync(0xffffffffu, done, 0);
          if (done) break;
          u = __shfl_sync(0xffffffffu, u, 0);
          if (u >= nu1) break;
          pair_noise_unit(nz, u, lane);
        }
      }
      mbar_wait(bar, parity);
    };
    const int parts = sch.nt1;                       // one likelihood partial per (row, R tile)
    uint32_t it = 0;
    for (int t = pair; t < sch.tiles_total; t += n_pairs, ++it) {
      const int g2 = t >= sch.tiles1 ? 1 : 0;
      const int tt = g2 ? t - sch.tiles1 : t;
      const int ntl = g2 ? sch.nt2 : sch.nt1;
      const int rb = tt / ntl, tn = tt - rb * ntl;
      const int m0 = rb * (BM * CG) + (int)rank * BM;
      const int n0 = tn * BN;                        // accumulator column 0 of this tile
      const uint32_t par = it & 1, aph = (it >> 1) & 1;
      const int row0 = m0 + q * 32;
      const int row = row0 + lane;                   // this thread's accumulator row
      const uint32_t tmem_row = tmem_base + par * BN + ((uint32_t)(q * 32) << 16);

      if (!g2) {
        // ---- link epilogue: z -> ell statistics, R = cot * mask * dl/dz (fp16 hi/lo) ----
        float* cy = s_col + par * 3 * BN;
        float* cm = cy + BN;
        float* crm = cm + BN;
        if ((int)threadIdx.x < BN) {              // per-column observation data of this tile
          const int col = n0 + threadIdx.x;
          float yv = 0.f, mv = 0.f;
          if (col < link.n) {
            yv = link.y[link.idx ? link.idx[col] : col];
            mv = link.mask ? link.mask[col] : 1.0f;
          }
          cy[threadIdx.x] = yv;
          cm[threadIdx.x] = mv;
          crm[threadIdx.x] = link.cot * mv * link.r_scale;
        }
        const bool row_ok = row < link.C;
        const float inv = row_ok ? 1.0f / (link.row_scale[row] * __ldg(link.b_scale)) : 0.f;
        named_bar_sync(1, kPrEpiThreads);
        if (threadIdx.x == 0) pair_stamp(dbg, pair, 8 + 5 * (int)it);
        wait_acc(&acc_full[par], aph);
        tc_fence_after();
        if (threadIdx.x == 0) pair_stamp(dbg, pair, 9 + 5 * (int)it);
        float n_w = 0.f, mean_w = 0.f, m2_w = 0.f, sm_w = 0.f;   // this row over the warp's chunks
#pragma unroll 1
        for (int ch = 0; ch < kChunks; ++ch) {
          const int ct = (ch * 4 + cg) * 32;               // column inside the tile (half ch)
          const int col0 = n0 + ct;
          uint32_t acc[32];
          tmem_ld32(tmem_row + (uint32_t)ct, acc);
          if (ch == kChunks - 1) {                             // accumulator drained: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&acc_empty[par]);
          }
          float cnt = 0.f, shift = 0.f, s1 = 0.f, s2 = 0.f, sm = 0.f;
          if (col0 < link.n) {                       // warp-uniform
            const bool full_cols = col0 + 32 <= link.n;
            float* ep = (link.ell && row_ok) ? link.ell + (int64_t)row * link.n + col0 : nullptr;
            // the statistics run on the fly (shifted by the first likelihood of the
            // chunk) so no per-element state beyond R stays live: more chains of
            // exp / log / rcp in flight per thread
            if (full_cols && link.mask == nullptr && link.ell == nullptr) {
              const float rmc = link.cot * link.r_scale;       // mask == 1 everywhere
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float z = __uint_as_float(acc[j]) * inv;
                float l, dz;
                logistic_link_fast(z, cy[ct + j], l, dz);
                if (j == 0) shift = l;
                const float dl = l - shift;
                s1 += dl;
                s2 = fmaf(dl, dl, s2);
                acc[j] = __float_as_uint(dz * rmc);
              }
              cnt = 32.f;
              sm = fmaf(32.f, shift, s1);                       // sum(ell * 1)
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float z = __uint_as_float(acc[j]) * inv;
                float l, dz;
                logistic_link_fast(z, cy[ct + j], l, dz);
                if (j == 0) shift = l;
                if (col0 + j < link.n) {
                  const float dl = l - shift;
                  cnt += 1.f; s1 += dl; s2 = fmaf(dl, dl, s2); sm = fmaf(l, cm[ct + j], sm);
                  if (ep) ep[j] = l;                           // optional per-observation output
                }
                acc[j] = __float_as_uint(dz * crm[ct + j]);
              }
            }
            // R tile -> swizzled staging (row = lane, 64 bytes per row: hi at +0, lo at
            // +2048) -> TMA store; rows >= C and columns >= n are clipped by the map
            if (lane == 0) bulk_wait_read_all();     // the previous store has read the buffer
            __syncwarp();
            const float* v = reinterpret_cast<const float*>(acc);
#pragma unroll
            for (int c16 = 0; c16 < 4; ++c16) {
              const uint32_t pos = (uint32_t)lane * 64u + (uint32_t)((c16 ^ ((lane >> 1) & 3)) * 16);
              uint4 hi, lo;
              uint32_t* hp = reinterpret_cast<uint32_t*>(&hi);
              uint32_t* lp = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float v0 = v[c16 * 8 + 2 * e], v1 = v[c16 * 8 + 2 * e + 1];
                if (TERMS == 3) {
                  const __half2 h = __floats2half2_rn(v0, v1);
                  const float2 hf = __half22float2(h);
                  hp[e] = *reinterpret_cast<const uint32_t*>(&h);
                  lp[e] = pack_half2(v0 - hf.x, v1 - hf.y);
                } else {
                  const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                  hp[e] = *reinterpret_cast<const uint32_t*>(&h);
                }
              }
              *reinterpret_cast<uint4*>(sbuf + pos) = hi;
              if (TERMS == 3) *reinterpret_cast<uint4*>(sbuf + 2048 + pos) = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&maps.r[0], sbuf, col0, row0);
              if (TERMS == 3) tma_store_2d(&maps.r[1], sbuf + 2048, col0, row0);
              bulk_commit_group();
            }
          }
          if (cnt > 0.f) {
            // statistics of this (row, 32-column chunk) folded into the warp's running
            // ones (Chan et al., chunk order)
            const float mean = shift + s1 / cnt;
            const float m2 = fmaxf(s2 - s1 * s1 / cnt, 0.f);
            const float nn = n_w + cnt, delta = mean - mean_w;
            mean_w += delta * (cnt / nn);
            m2_w += m2 + delta * delta * (n_w * cnt / nn);
            n_w = nn;
          }
          sm_w += sm;
          if (ch < kChunks - 1) {
            // publish this column half of R: the TMA stores of all sixteen warps are
            // performed and visible to the TMA loads of other SMs before the half's
            // counter of the row block moves (GEMM2 of this row block may start on it)
            if (lane == 0) bulk_wait_all();
            __syncwarp();
            __threadfence();
            fence_proxy_async_global();
            if (warp == 0) {
              named_bar_sync(4, kPrEpiThreads);
              if (lane == 0) atom_add_release_gpu(&link.counters[rb * kHalves + ch], 1u);
            } else {
              named_bar_arrive(4, kPrEpiThreads);
            }
          }
        }
        // publish: R (TMA stores complete) + stats of this tile are visible -- also to the
        // TMA loads of other SMs -- before the row-block counter moves
        if (threadIdx.x == 0) pair_stamp(dbg, pair, 10 + 5 * (int)it);
        if (lane == 0) {
          bulk_wait_all();
          mbar_arrive(epi_done);                     // this warp's staging buffer is free again
        }
        // the four column groups of a row meet in shared memory: one partial per (row, tile)
        // goes to global memory, so the CTA that finishes a row block folds nt1 partials per
        // chain, not 8 nt1 (that serial fold used to hold its gradient epilogue back by 20 us)
        float4* s_part = s_part_all + par * (BM * 4);
        s_part[(q * 32 + lane) * 4 + cg] = make_float4(n_w, mean_w, m2_w, sm_w);
        __threadfence();
        fence_proxy_async_global();
        if (threadIdx.x == 0) pair_stamp(dbg, pair, 11 + 5 * (int)it);
        if (warp < 4) {
          named_bar_sync(2, kPrEpiThreads);
          if (row0 - q * 32 + (int)threadIdx.x < link.C) {
            float n_t = 0.f, mean_t = 0.f, m2_t = 0.f, sm_t = 0.f;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 st = s_part[threadIdx.x * 4 + g];
              if (st.x > 0.f) {
                const float nn = n_t + st.x, delta = st.y - mean_t;
                mean_t += delta * (st.x / nn);
                m2_t += st.z + delta * delta * (n_t * st.x / nn);
                n_t = nn;
              }
              sm_t += st.w;
            }
            *reinterpret_cast<float4*>(
                link.stats + ((int64_t)(m0 + (int)threadIdx.x) * parts + tn) * kStatFields) =
                make_float4(n_t, mean_t, m2_t, sm_t);
          }
          __threadfence();
          named_bar_sync(3, 128);
          if (threadIdx.x == 0) {
            const uint32_t prev = atom_add_release_gpu(&link.counters[rb * kHalves + kHalves - 1], 1u);
            s_last[par] = prev == (uint32_t)(CG * sch.nt1) - 1u;
          }
          named_bar_sync(5, 128);
          if (s_last[par]) {
            // last tile of the row block: U = (L - prior)/T (potential.py:183-185, :210)
            // and var(ell) (integrator.py:880) from the partials, Chan et al. in fixed order
            __threadfence();
            for (int hf = 0; hf < CG; ++hf) {
              const int c = rb * (BM * CG) + hf * BM + (int)threadIdx.x;
              if (c >= link.C) continue;
              float n_t = 0.f, mean_t = 0.f, m2_t = 0.f, sm_t = 0.f;
              const float4* sp = reinterpret_cast<const float4*>(
                  link.stats + (int64_t)c * parts * kStatFields);
              for (int g0 = 0; g0 < parts; g0 += 8) {
                float4 stv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                  stv[u] = g0 + u < parts ? __ldcg(sp + g0 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const float4 st = stv[u];
                  if (st.x > 0.f) {
                    const float nn = n_t + st.x, delta = st.y - mean_t;
                    mean_t += delta * (st.x / nn);
                    m2_t += st.z + delta * delta * (n_t * st.x / nn);
                    n_t = nn;
                  }
                  sm_t += st.w;
                }
              }
              const float L = link.mask ? (-link.n_obs / (float)link.n) * sm_t : -link.n_obs * mean_t;
              const float prior = -link.prior_half_inv * link.row_sumsq[c];
              link.potential[c] = (L - prior) * link.inv_temperature;
              if (link.variance) link.variance[c] = m2_t / (float)link.n;
            }
          }
          if (threadIdx.x == 0) pair_stamp(dbg, pair, 12 + 5 * (int)it);
        } else {
          named_bar_arrive(2, kPrEpiThreads);
        }
      } else {
        // ---- gradient epilogue: G * 1/scale (+ theta * prior_coef) -> grad ----
        const float inv_scale = 1.0f / (gradp.r_scale * __ldg(gradp.xt_scale));
        const bool row_ok = row < gradp.C;
        if (threadIdx.x == 0) pair_stamp(dbg, pair, 8 + 5 * (int)it);
        wait_acc(&acc_full[par], aph);
        tc_fence_after();
        if (threadIdx.x == 0) pair_stamp(dbg, pair, 9 + 5 * (int)it);
#pragma unroll 1
        for (int ch = 0; ch < kChunks; ++ch) {
          const int ct = cg * (BN / 4) + ch * 32;
          const int col0 = n0 + ct;
          const int p0 = gradp.w_off + col0;
          uint32_t acc[32];
          tmem_ld32(tmem_row + (uint32_t)ct, acc);
          if (ch == kChunks - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&acc_empty[par]);
          }
          if (col0 >= gradp.d) continue;             // warp-uniform
          float* o = reinterpret_cast<float*>(acc);
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(acc[j]) * inv_scale;
          // prior gradient in the epilogue (stand-alone potential calls; the carried step
          // folds it into the update kernel instead, see SgldSplitOp): theta is read
          // COALESCED, one 128-byte row segment per load (lane = column), issued before
          // the staging buffer is written, and added to the staged tile afterwards
          const bool with_prior = p0 < gradp.prior_hi && p0 + 32 > gradp.prior_lo;
          float tv[32];
          if (with_prior) {
            const int p = p0 + lane;
            const bool col_ok = col0 + lane < gradp.d && p >= gradp.prior_lo && p < gradp.prior_hi;
            const float* tp = gradp.theta + (int64_t)row0 * gradp.P + p;
#pragma unroll
            for (int r = 0; r < 32; ++r)
              tv[r] = (col_ok && row0 + r < gradp.C) ? __ldg(tp + (int64_t)r * gradp.P) : 0.f;
          }
          if (lane == 0) bulk_wait_read_all();
          __syncwarp();
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16) {
            const uint32_t pos = (uint32_t)lane * 128u + (uint32_t)((c16 ^ (lane & 7)) * 16);
            *reinterpret_cast<float4*>(sbuf + pos) =
                make_float4(o[4 * c16], o[4 * c16 + 1], o[4 * c16 + 2], o[4 * c16 + 3]);
          }
          if (with_prior) {
            __syncwarp();
            // element (r, lane) of the tile lives in 16-byte chunk (lane >> 2) ^ (r & 7) of
            // staging row r: 32 distinct words per row, no bank conflicts
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              float* q = reinterpret_cast<float*>(
                  sbuf + (uint32_t)r * 128u + (uint32_t)((((lane >> 2) ^ (r & 7)) * 16) + (lane & 3) * 4));
              *q = fmaf(tv[r], gradp.prior_coef, *q);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&maps.g, sbuf, p0, row0);
            bulk_commit_group();
          }
        }
        if (lane == 0) {
          bulk_wait_read_all();
          mbar_arrive(epi_done);
        }
        if (threadIdx.x == 0) pair_stamp(dbg, pair, 10 + 5 * (int)it);
      }
    }
    if (lane == 0) bulk_wait_all();                  // staging buffers are read, stores performed
    if (threadIdx.x == 0) pair_stamp(dbg, pair, 6);
    if (nz.xi != nullptr) {                          // whatever is left of the noise job
      for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(s_noise_next, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= nu1) break;
        pair_noise_unit(nz, u, lane);
      }
    }
    if (threadIdx.x == 0) pair_stamp(dbg, pair, 1);
    if (upd.enabled) {
      // ---- update phase: G and xi of every CTA are complete and visible first ----
      __threadfence();
      fence_proxy_async_global();
      named_bar_sync(6, kPrEpiThreads);
      if (threadIdx.x == 0) {
        atom_add_release_gpu(upd.grid_counter, 1u);
        while (ld_acquire_gpu(upd.grid_counter) < gridDim.x) __nanosleep(64);
        pair_stamp(dbg, pair, 28);
      }
      named_bar_sync(7, kPrEpiThreads);
      constexpr int kFmt = TERMS == 3 ? 1 : 2;
      const uint64_t keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
      const uint32_t tpc = upd.t.tiles_per_chain;
      for (int64_t tile = (int64_t)blockIdx.x * kPrEpiWarps + warp; tile < upd.n_tiles;
           tile += (int64_t)gridDim.x * kPrEpiWarps) {
        const int64_t c = tile / tpc;
        const uint32_t t = (uint32_t)(tile - c * tpc);
        if (upd.rms) apply_tile_fast<true, true, kFmt>(upd.t, c, t, lane, keep, drop);
        else apply_tile_fast<false, true, kFmt>(upd.t, c, t, lane, keep, drop);
      }
      if (threadIdx.x == 0) pair_stamp(dbg, pair, 29);
    }
  }
  __syncwarp();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == kPrEpiWarps + 1) tmem_dealloc_cg<CG>(tmem_base, 2 * BN);
  if (threadIdx.x == 0) pair_stamp(dbg, pair, 31);                   // kernel exit
}

// ---------------------------------------------------------------------------
// Operand preparation / finalisation kernels
// ---------------------------------------------------------------------------
__device__ __forceinline__ float pow2_scale_for(float absmax) {
  // power of two s with absmax*s in [2^12, 2^13); 1 for absmax == 0 / non-finite
  if (!(absmax > 0.0f) || isinf(absmax)) return 1.0f;
  int e;
  frexpf(absmax, &e);                     // absmax = m * 2^e, m in [0.5, 1)
  return ldexpf(1.0f, 13 - e);
}

// |X[idx]| max over the minibatch: one warp per gathered row.
__global__ void k_absmax_gather(const float* __restrict__ X, const int32_t* __restrict__ idx,
                                int n, int d, uint32_t* __restrict__ out_bits) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* src = X + (int64_t)(idx ? idx[warp] : warp) * d;
  float m = 0.f;
  for (int j = lane; j < d; j += 32) m = fmaxf(m, fabsf(src[j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) atomicMax(out_bits, __float_as_uint(m));  // m >= 0: uint order == float order
}

// Both operand preparations in ONE launch (they are independent): blocks
// [0, theta_blocks) convert Theta rows (one warp per chain row, 128-bit loads,
// 64-bit packed stores), the remaining blocks gather / split / transpose the
// minibatch in 64x64 tiles.  Also clears the per-row-block tile counters used by
// GEMM1's last-arriver finalisation.
struct PrepareArgs {
  const float* theta; int64_t P; int w_off, d, C, prior_lo, prior_hi;
  void* th_hi; __half* th_lo; float* row_scale; float* row_sumsq;
  const float* X; const int32_t* idx; int n;
  const uint32_t* absmax_bits; float static_absmax;
  void* xb_hi; __half* xb_lo; void* xt_hi; __half* xt_lo; float* x_scale;
  int theta_blocks, x_tiles_x;
  uint32_t* tile_counters; int n_counters;
  // carried split (sgmc_glm_sgld_step, SGMC_STEP_CARRY*): theta_mode 0 = convert every
  // row; 2 = the previous update wrote the split with next_scale[row] -- validate it
  // against the row's new |max| and only re-split rows outside the safe window
  int theta_mode;
  float* next_scale; uint32_t* amax_bits; const float* sumsq_part; int tiles_per_chain;
  // noise-key cache of the step's update: chain keys in, key' out, noise key out
  const uint32_t* keys_in; uint32_t* keys_out; uint32_t* noise_keys; int prng_layout;
  int key_blocks;           // leading blocks that derive the noise keys (one thread per chain)
};

constexpr int kPrepTile = 64;   // minibatch tile: 64 observations x 64 features

template <bool SPLIT>
__global__ void __launch_bounds__(256) k_prepare_all(const PrepareArgs a) {
  __shared__ float tile[kPrepTile][kPrepTile + 1];
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  if ((int)blockIdx.x < a.key_blocks) {
    // key', sub = split(key) (integrator.py:871); the single leaf's noise key is
    // split(sub, 1)[0] (integrator.py:131-133): derived here, one THREAD per chain (four
    // dependent threefry evaluations), so the update kernel starts without its serial
    // key prologue
    const int c = (int)blockIdx.x * 256 + (int)threadIdx.x;
    if (c < a.C) {
      const Key k{a.keys_in[2 * c], a.keys_in[2 * c + 1]};
      Key newk, sub;
      split2(k, a.prng_layout, newk, sub);
      const Key nk = split_key(sub, 0u, 1u, a.prng_layout);
      a.noise_keys[2 * c] = nk.k0; a.noise_keys[2 * c + 1] = nk.k1;
      a.keys_out[2 * c] = newk.k0; a.keys_out[2 * c + 1] = newk.k1;
    }
    return;
  }
  const int bid = (int)blockIdx.x - a.key_blocks;
  if (bid < a.theta_blocks) {
    if (bid == 0)
      for (int i = threadIdx.x; i < a.n_counters; i += 256) a.tile_counters[i] = 0u;
    const int row = bid * 8 + (threadIdx.x >> 5);
    if (row >= a.C) return;
    if (a.theta_mode == 2) {
      const float s_used = a.next_scale[row];
      const float amax = __uint_as_float(a.amax_bits[row]);
      const float m = amax * s_used;
      // fp16 split: the row maximum must stay well inside the normal range of hi AND
      // leave lo = x*s - hi normal for every element that matters
      const bool ok = !SPLIT || amax == 0.0f || (m > 64.0f && m < 65000.0f);
      if (ok) {
        if (lane == 0) {
          float sq = 0.f;
          for (int t = 0; t < a.tiles_per_chain; ++t)
            sq += a.sumsq_part[(int64_t)row * a.tiles_per_chain + t];
          a.row_scale[row] = s_used;
          a.row_sumsq[row] = sq;
          a.next_scale[row] = SPLIT ? pow2_scale_for(amax) : 1.0f;
          a.amax_bits[row] = 0u;
        }
        return;
      }
    }
    const float* src = a.theta + (int64_t)row * a.P + a.w_off;
    const bool vec = (a.d & 3) == 0 && (a.P & 3) == 0 && (a.w_off & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(a.theta) & 15u) == 0;
    const int64_t ob = (int64_t)row * a.d;
    auto in_prior4 = [&](int j) { return a.w_off + j >= a.prior_lo && a.w_off + j + 3 < a.prior_hi; };
    auto sumsq4 = [&](const float4& x, int j, float sq) {
      if (in_prior4(j)) return sq + (x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w);
      const int p0 = a.w_off + j;
      if (p0 >= a.prior_lo && p0 < a.prior_hi) sq = fmaf(x.x, x.x, sq);
      if (p0 + 1 >= a.prior_lo && p0 + 1 < a.prior_hi) sq = fmaf(x.y, x.y, sq);
      if (p0 + 2 >= a.prior_lo && p0 + 2 < a.prior_hi) sq = fmaf(x.z, x.z, sq);
      if (p0 + 3 >= a.prior_lo && p0 + 3 < a.prior_hi) sq = fmaf(x.w, x.w, sq);
      return sq;
    };
    auto store4 = [&](const float4& x, int j, float s) {
      const float v0 = x.x * s, v1 = x.y * s, v2 = x.z * s, v3 = x.w * s;
      if (SPLIT) {
        const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y);
        const __half2 l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
        uint2 ph, pl;
        ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
        pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.th_hi) + ob + j) = ph;
        *reinterpret_cast<uint2*>(a.th_lo + ob + j) = pl;
      } else {
        const __nv_bfloat162 b01 = __floats2bfloat162_rn(v0, v1), b23 = __floats2bfloat162_rn(v2, v3);
        uint2 pb;
        pb.x = *reinterpret_cast<const uint32_t*>(&b01); pb.y = *reinterpret_cast<const uint32_t*>(&b23);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.th_hi) + ob + j) = pb;
      }
    };
    float m = 0.f, sq = 0.f;
    if (vec && a.d <= 1024) {
      // the whole row lives in registers: 8 independent 128-bit loads per lane in
      // flight, one pass over memory
      float4 x[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = lane * 4 + k * 128;
        x[k] = j < a.d ? *reinterpret_cast<const float4*>(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = lane * 4 + k * 128;
        if (j < a.d) {
          m = fmaxf(fmaxf(m, fmaxf(fabsf(x[k].x), fabsf(x[k].y))),
                    fmaxf(fabsf(x[k].z), fabsf(x[k].w)));
          sq = sumsq4(x[k], j, sq);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
      }
      const float s = SPLIT ? pow2_scale_for(m) : 1.0f;
      if (lane == 0) {
        a.row_scale[row] = s;
        a.row_sumsq[row] = sq;
        if (a.next_scale) {
          a.next_scale[row] = s;
          a.amax_bits[row] = 0u;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = lane * 4 + k * 128;
        if (j < a.d) store4(x[k], j, s);
      }
      return;
    }
    if (vec) {
      for (int j = lane * 4; j < a.d; j += 128) {
        const float4 x = *reinterpret_cast<const float4*>(src + j);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(x.x), fabsf(x.y))), fmaxf(fabsf(x.z), fabsf(x.w)));
        sq = sumsq4(x, j, sq);
      }
    } else {
      for (int j = lane; j < a.d; j += 32) {
        const float x = src[j];
        m = fmaxf(m, fabsf(x));
        if (a.w_off + j >= a.prior_lo && a.w_off + j < a.prior_hi) sq = fmaf(x, x, sq);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const float s = SPLIT ? pow2_scale_for(m) : 1.0f;
    if (lane == 0) {
      a.row_scale[row] = s;
      a.row_sumsq[row] = sq;
      if (a.next_scale) {
        a.next_scale[row] = s;
        a.amax_bits[row] = 0u;
      }
    }
    if (vec) {
      for (int j = lane * 4; j < a.d; j += 128)
        store4(*reinterpret_cast<const float4*>(src + j), j, s);   // L1 / L2 hit
    } else {
      for (int j = lane; j < a.d; j += 32) {
        const float x = src[j] * s;
        if (SPLIT) {
          const __half h = __float2half_rn(x);
          reinterpret_cast<__half*>(a.th_hi)[ob + j] = h;
          a.th_lo[ob + j] = __float2half_rn(x - __half2float(h));
        } else {
          reinterpret_cast<__nv_bfloat16*>(a.th_hi)[ob + j] = __float2bfloat16_rn(x);
        }
      }
    }
    return;
  }
  // ---- minibatch tile: gather, scale, split; row-major and transposed copies --------
  // (n and d are multiples of 8 on this path: element pairs never straddle an edge)
  const int t = bid - a.theta_blocks;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
  const int n = a.n, d = a.d;
  float s = 1.0f;
  if (SPLIT)
    s = pow2_scale_for(a.static_absmax > 0.f ? a.static_absmax : __uint_as_float(*a.absmax_bits));
  if (t == 0 && threadIdx.x == 0) *a.x_scale = s;
  const int r0 = (t / a.x_tiles_x) * kPrepTile, c0 = (t % a.x_tiles_x) * kPrepTile;
  const bool x8 = (reinterpret_cast<uintptr_t>(a.X) & 7u) == 0;
  auto store2 = [&](void* hi, __half* lo, int64_t o, float v0, float v1) {
    if (SPLIT) {
      const __half2 h = __floats2half2_rn(v0, v1);
      const float2 hf = __half22float2(h);
      *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(hi) + o) = h;
      *reinterpret_cast<__half2*>(lo + o) = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    } else {
      *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(hi) + o) =
          __floats2bfloat162_rn(v0, v1);
    }
  };
  {
    const int c = c0 + 2 * tx;
    float2 val[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {                   // 8 independent 64-bit loads in flight
      const int r = r0 + ty + 8 * k;
      val[k] = make_float2(0.f, 0.f);
      if (r < n && c < d) {
        const int64_t row = a.idx ? a.idx[r] : r;
        const float* px = a.X + row * d + c;
        if (x8) val[k] = *reinterpret_cast<const float2*>(px);
        else val[k] = make_float2(px[0], px[1]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = ty + 8 * k, r = r0 + i;
      const float v0 = val[k].x * s, v1 = val[k].y * s;
      if (r < n && c < d) store2(a.xb_hi, a.xb_lo, (int64_t)r * d + c, v0, v1);
      tile[i][2 * tx] = v0;
      tile[i][2 * tx + 1] = v1;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int cc = ty + 8 * k, c = c0 + cc, r = r0 + 2 * tx;   // transposed: row = feature c
    if (c < d && r < n)
      store2(a.xt_hi, a.xt_lo, (int64_t)c * n + r, tile[2 * tx][cc], tile[2 * tx + 1][cc]);
  }
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// 2-D row-major [rows][cols] matrix of 2-byte elements, box = [box_rows][64].
static int make_map(CUtensorMap* m, const void* ptr, int bf16, int64_t rows, int64_t cols,
                    int box_rows) {
  EncodeTiledFn fn = encode_fn();
  SGMC_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                  2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SGMC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// General 2-D row-major map: dtype 0 fp16, 1 bf16, 2 f32; box = [box_rows][box_cols];
// swizzle in bytes (64 or 128) = box_cols * element size.
static int make_map_ex(CUtensorMap* m, const void* ptr, int dtype, int64_t rows, int64_t cols,
                       int box_rows, int box_cols, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  SGMC_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable");
  const int esize = dtype == 2 ? 4 : 2;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                              : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(m, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SGMC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

struct TcWorkspace {
  void* th_hi; __half* th_lo; float* row_scale; float* row_sumsq;
  void* xb_hi; __half* xb_lo; void* xt_hi; __half* xt_lo;
  void* r_hi; __half* r_lo;
  float* stats;
  uint32_t* absmax_bits; float* x_scale;
  uint32_t* counters;
  float* xi;                // f32[C][d]: noise of the pending update (sgmc_glm_sgld_step)
  // carried split of sgmc_glm_sgld_step
  float* next_scale; uint32_t* amax_row; float* sumsq_part; uint32_t* noise_keys;
  int32_t* idx2;            // second minibatch index buffer of the pipelined scans
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t carve(TcWorkspace* w, uint8_t* base, int64_t C, int64_t n, int64_t d,
                    int slot = 0) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  // per-(row, 32-column chunk) likelihood partials of the fused kernel (tiles of up to
  // 256 columns, so n is padded to the next multiple of 256)
  const int64_t parts = ((n + 255) / 256) * 8;
  void* th_hi = take((size_t)C * d * 2);
  void* th_lo = take((size_t)C * d * 2);
  void* rs = take((size_t)C * 4);
  void* rq = take((size_t)C * 4);
  // the minibatch operands exist twice: sgmc_glm_prepare_minibatch stages the NEXT
  // step's minibatch (slot 1 - s) while the current step's kernels read slot s
  void *xb_hi = nullptr, *xb_lo = nullptr, *xt_hi = nullptr, *xt_lo = nullptr, *am = nullptr;
  for (int sl = 0; sl < 2; ++sl) {
    void* a0 = take((size_t)n * d * 2);
    void* a1 = take((size_t)n * d * 2);
    void* a2 = take((size_t)n * d * 2);
    void* a3 = take((size_t)n * d * 2);
    void* a4 = take(256);
    if (sl == slot) { xb_hi = a0; xb_lo = a1; xt_hi = a2; xt_lo = a3; am = a4; }
  }
  void* r_hi = take((size_t)C * n * 2);
  void* r_lo = take((size_t)C * n * 2);
  void* stats = take((size_t)C * parts * kStatFields * 4);
  // two hand-over counters per row block + the grid-wide arrival counter of the update phase
  void* cnt = take((size_t)(2 * ((C + BM - 1) / BM) + 8) * 4);
  void* xi = take((size_t)C * d * 4);
  void* nsc = take((size_t)C * 4);
  void* amr = take((size_t)C * 4);
  void* ssp = take((size_t)C * sgld_split_tiles_per_chain(d) * 4);
  void* nks = take((size_t)C * 8);
  void* idx2 = take((size_t)n * 4);
  if (w) {
    w->idx2 = (int32_t*)idx2;
    w->next_scale = (float*)nsc; w->amax_row = (uint32_t*)amr;
    w->sumsq_part = (float*)ssp; w->noise_keys = (uint32_t*)nks;
    w->counters = (uint32_t*)cnt;
    w->xi = (float*)xi;
    w->th_hi = th_hi; w->th_lo = (__half*)th_lo;
    w->row_scale = (float*)rs; w->row_sumsq = (float*)rq;
    w->xb_hi = xb_hi; w->xb_lo = (__half*)xb_lo; w->xt_hi = xt_hi; w->xt_lo = (__half*)xt_lo;
    w->r_hi = r_hi; w->r_lo = (__half*)r_lo;
    w->stats = (float*)stats;
    w->absmax_bits = (uint32_t*)am; w->x_scale = (float*)am + 1;
  }
  return off;
}

// The update half of sgmc_glm_sgld_step when the noise was generated by the GEMM
// kernels: one elementwise pass theta, v <- sgld(theta, v, grad, xi)
// (integrator.py:882-914, adaption.py:254-291); 24 B per parameter, HBM/L2-bound.
template <bool RMS>
__global__ void __launch_bounds__(256) k_sgld_apply(float* __restrict__ theta, float* __restrict__ v,
                                                    const float* __restrict__ grad,
                                                    const float* __restrict__ xi, int64_t n4,
                                                    float neg_eps, float ns, float alpha,
                                                    float one_m_alpha, float lmbd) {
  pdl_launch_dependents();
  pdl_wait();
  float4* t4 = reinterpret_cast<float4*>(theta);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float4* g4 = reinterpret_cast<const float4*>(grad);
  const float4* x4 = reinterpret_cast<const float4*>(xi);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 t = t4[i];
    const float4 g = g4[i], x = x4[i];
    float4 vv = RMS ? v4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    t.x = sgld_one<RMS, RMS>(t.x, g.x, vv.x, x.x, ns, neg_eps, alpha, one_m_alpha, lmbd);
    t.y = sgld_one<RMS, RMS>(t.y, g.y, vv.y, x.y, ns, neg_eps, alpha, one_m_alpha, lmbd);
    t.z = sgld_one<RMS, RMS>(t.z, g.z, vv.z, x.z, ns, neg_eps, alpha, one_m_alpha, lmbd);
    t.w = sgld_one<RMS, RMS>(t.w, g.w, vv.w, x.w, ns, neg_eps, alpha, one_m_alpha, lmbd);
    t4[i] = t;
    if (RMS) v4[i] = vv;
  }
}

int glm_pair_timeline_read(unsigned long long* out, int n) {
  const size_t cnt = (size_t)std::min(n, kDbgPairs * kDbgSlots);
  return check_cuda(cudaMemcpyFromSymbol(out, g_pair_dbg, sizeof(unsigned long long) * cnt), "dbg");
}

int glm_tc_debug_read(unsigned long long* out) {
  return check_cuda(cudaMemcpyFromSymbol(out, g_tc_dbg, sizeof(unsigned long long) * 10), "dbg");
}

int32_t* glm_tc_spare_idx(float* tc_ws, int64_t n_chains, int64_t batch_size, int64_t d) {
  TcWorkspace w;
  uint8_t* base = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(tc_ws) + 255) & ~(uintptr_t)255);
  carve(&w, base, n_chains, batch_size, d);
  return w.idx2;
}

size_t glm_tc_workspace_bytes(int64_t n_chains, int64_t batch_size, int64_t d, int) {
  return carve(nullptr, nullptr, n_chains, batch_size, d) + 256;
}

template <int TERMS, int EPI, int ABFMT>
static int launch_gemm(cudaStream_t stream, const CUtensorMap& a0, const CUtensorMap& a1,
                       const CUtensorMap& b0, const CUtensorMap& b1, int M, int N, int K,
                       const TcLinkEpi& link, const TcGradEpi& gradp, const char* name,
                       TcNoiseJob job = TcNoiseJob{}) {
  using S = TcSmem<TERMS>;
  auto kfn = k_glm_tc_gemm<TERMS, EPI, ABFMT>;
  static bool attr_set = false;
  if (!attr_set) {
    if (check_cuda(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        S::kBytes), "cudaFuncSetAttribute"))
      return 1;
    attr_set = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  if (job.xi != nullptr) {
    const int ctas = (int)(grid.x * grid.y);
    job.tiles_per_cta = (job.tile_end - job.tile0 + ctas - 1) / ctas;
  }
  launch_pdl(kfn, grid, dim3(EPI == 2 ? kTcFusedThreads : kTcThreads), S::kBytes, stream, a0, a1, b0, b1,
             (int)((K + BK - 1) / BK), link, gradp, job);
  return post_launch(name);
}

struct MapKey {
  const void* base; const void* grad; int64_t C, n, P; int d, cg, bn, split;
  bool operator==(const MapKey& o) const {
    return base == o.base && grad == o.grad && C == o.C && n == o.n && P == o.P && d == o.d &&
           cg == o.cg && bn == o.bn && split == o.split;
  }
};
struct MapCacheEntry { MapKey key; PairMaps maps; };
static std::mutex g_map_mu;
static std::vector<MapCacheEntry> g_map_cache;
static bool map_cache_get(const MapKey& k, PairMaps* out) {
  std::lock_guard<std::mutex> lock(g_map_mu);
  for (const MapCacheEntry& e : g_map_cache)
    if (e.key == k) { *out = e.maps; return true; }
  return false;
}
static void map_cache_put(const MapKey& k, const PairMaps& m) {
  std::lock_guard<std::mutex> lock(g_map_mu);
  if (g_map_cache.size() >= 32) g_map_cache.erase(g_map_cache.begin());
  g_map_cache.push_back(MapCacheEntry{k, m});
}

template <int TERMS, int ABFMT, int CG, int BN>
static int launch_pair(cudaStream_t stream, const PairMaps& maps, const PairSched& sch,
                       const TcLinkEpi& link, const TcGradEpi& gradp, const PairNoise& nz,
                       const PairUpdate& upd, const char* name) {
  using S = PrSmem<TERMS, CG, BN>;
  auto kfn = k_glm_tc_pair<TERMS, ABFMT, CG, BN>;
  static bool attr_set = false;
  static int max_pairs = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kPrThreads);
  cfg.dynamicSmemBytes = S::kBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = option(SGMC_OPT_SERIAL_LAUNCH) ? 0 : 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = CG;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (!attr_set) {
    if (check_cuda(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        S::kBytes), "cudaFuncSetAttribute"))
      return 1;
    // every pair of the grid must be resident at once (the row-block hand-over
    // between GEMM1 and GEMM2 tiles spins on other pairs' progress)
    max_pairs = sm_count() / CG;
    if (CG == 2) {
      cfg.gridDim = dim3(2 * (sm_count() / 2));
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kfn, &cfg) == cudaSuccess && n > 0)
        max_pairs = std::min(max_pairs, n);
      else
        cudaGetLastError();
    }
    attr_set = true;
  }
  // with a noise job every SM takes part, also pairs without a tile
  const int cap = option(SGMC_OPT_TC_MAX_PAIRS) > 0 ? std::min(max_pairs, option(SGMC_OPT_TC_MAX_PAIRS))
                                                    : max_pairs;
  const int pairs = nz.xi ? cap : std::min(cap, sch.tiles_total);
  cfg.gridDim = dim3(CG * pairs);
  if (check_cuda(cudaLaunchKernelEx(&cfg, kfn, maps, sch, link, gradp, nz, upd), name)) return 1;
  return post_launch(name);
}

int glm_tc(cudaStream_t stream, const GlmArgs& a, int path) {
  const int64_t C = a.C, n = a.n;
  const int d = a.spec.d;
  SGMC_REQUIRE(a.spec.family == kFamilyLogistic,
               "tensor-core path supports the logistic family (use path 0)");
  SGMC_REQUIRE(a.spec.aux_off < 0, "tensor-core path: bias term not supported (use path 0)");
  SGMC_REQUIRE(a.spec.prior != kPriorInvSigma, "tensor-core path: prior not supported");
  SGMC_REQUIRE(d % 8 == 0 && n % 8 == 0, "tensor-core path needs d %% 8 == 0 and n %% 8 == 0");
  const bool split = path == 1;
  TcWorkspace w;
  uint8_t* base = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(a.tc_ws) + 255) & ~(uintptr_t)255);
  carve(&w, base, C, n, d, a.x_slot);
  const bool gauss_prior = a.spec.prior == kPriorGaussian;
  const int prior_lo = gauss_prior ? a.spec.prior_off : 0;
  const int prior_hi = gauss_prior ? a.spec.prior_off + a.spec.prior_size : 0;

  // sgmc_glm_sgld_step: the update's Gaussian noise is generated by the idle warps
  // of the two GEMMs (half each) and applied by k_sgld_apply afterwards.
  const FusedSgld& fu = a.fused;
  const bool aligned = ((reinterpret_cast<uintptr_t>(fu.theta_rw) | reinterpret_cast<uintptr_t>(fu.v) |
                         reinterpret_cast<uintptr_t>(a.grad)) & 15u) == 0;
  const bool fusable = fu.requested && fu.layout == 0 && a.P == d && a.spec.w_off == 0 &&
                       d % 256 == 0 && a.grad != nullptr && !option(SGMC_OPT_EXACT_UPDATE_MATH);
  const bool epilogue_update = fusable && option(SGMC_OPT_FUSED_STEP_EPILOGUE) && C % BM == 0;
  const bool noise_job = fusable && !epilogue_update && aligned &&
                         option(SGMC_OPT_STEP_NOISE_IN_GEMM);
  TcNoiseJob job1{}, job2{};
  if (noise_job) {
    const int tiles = (int)(C * (d / 256));
    job1.xi = job2.xi = w.xi;
    job1.keys_in = job2.keys_in = fu.keys_in;
    job1.keys_out = job2.keys_out = fu.keys_out;
    job1.d = job2.d = d;
    job1.tile0 = 0; job1.tile_end = tiles / 2;
    job2.tile0 = tiles / 2; job2.tile_end = tiles;
  }

  const bool legacy = option(SGMC_OPT_TC_LEGACY) || epilogue_update || noise_job ||
                      ((reinterpret_cast<uintptr_t>(a.theta) | reinterpret_cast<uintptr_t>(a.grad) |
                        reinterpret_cast<uintptr_t>(a.ell_requested ? a.ell : nullptr)) & 15u) != 0;
  // carried split: the previous sgmc_glm_sgld_step's update left Theta's operand form
  // in the workspace (and this step's update will do so for the next one)
  CarryCtx* cc = a.carry;
  const bool carry = cc != nullptr && cc->mode != 0 && !legacy && a.P == d && a.spec.w_off == 0 &&
                     a.grad != nullptr;

  // ---- operand preparation (one launch) -------------------------------------
  const float x_absmax = a.spec.x_absmax > 0.f ? a.spec.x_absmax : 0.f;
  if (split && x_absmax == 0.f && !a.x_prepared) {   // no data-set bound given: reduce over the minibatch
    if (check_cuda(cudaMemsetAsync(w.absmax_bits, 0, 4, stream), "memset")) return 1;
    k_absmax_gather<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(a.X, a.idx, (int)n, d,
                                                                w.absmax_bits);
    if (post_launch("k_absmax_gather")) return 1;
  }
  {
    PrepareArgs pa{};
    pa.theta = a.theta; pa.P = a.P; pa.w_off = a.spec.w_off; pa.d = d; pa.C = (int)C;
    pa.prior_lo = prior_lo; pa.prior_hi = prior_hi;
    pa.th_hi = w.th_hi; pa.th_lo = w.th_lo; pa.row_scale = w.row_scale; pa.row_sumsq = w.row_sumsq;
    pa.X = a.X; pa.idx = a.idx; pa.n = (int)n;
    pa.absmax_bits = w.absmax_bits; pa.static_absmax = x_absmax;
    pa.xb_hi = w.xb_hi; pa.xb_lo = w.xb_lo; pa.xt_hi = w.xt_hi; pa.xt_lo = w.xt_lo;
    pa.x_scale = w.x_scale;
    pa.theta_blocks = a.only_x ? 0 : (int)((C + 7) / 8);
    pa.x_tiles_x = (d + kPrepTile - 1) / kPrepTile;
    pa.tile_counters = w.counters; pa.n_counters = (int)(2 * ((C + BM - 1) / BM) + 8);
    if (carry) {
      pa.theta_mode = cc->mode == 2 ? 2 : 0;
      pa.next_scale = w.next_scale; pa.amax_bits = w.amax_row; pa.sumsq_part = w.sumsq_part;
      pa.tiles_per_chain = sgld_split_tiles_per_chain(d);
      pa.keys_in = cc->keys_in; pa.keys_out = cc->keys_out; pa.noise_keys = w.noise_keys;
      pa.prng_layout = cc->prng_layout;
      cc->active = true;
      cc->out.fmt = split ? 1 : 2;
      cc->out.th_hi = w.th_hi; cc->out.th_lo = w.th_lo;
      cc->out.scale = w.next_scale; cc->out.amax_bits = w.amax_row;
      cc->out.sumsq_part = w.sumsq_part;
      cc->out.prior_lo = prior_lo; cc->out.prior_hi = prior_hi;
      cc->out.noise_keys = w.noise_keys;
    }
    pa.key_blocks = pa.keys_in ? (int)((C + 255) / 256) : 0;
    // the minibatch tiles are skipped when sgmc_glm_prepare_minibatch staged them already
    const int64_t x_tiles = a.x_prepared ? 0 : pa.x_tiles_x * ((n + kPrepTile - 1) / kPrepTile);
    const unsigned grid = (unsigned)(pa.key_blocks + pa.theta_blocks + x_tiles);
    if (grid > 0) {
      if (split) launch_pdl(k_prepare_all<true>, dim3(grid), dim3(256), 0, stream, pa);
      else launch_pdl(k_prepare_all<false>, dim3(grid), dim3(256), 0, stream, pa);
      if (post_launch("k_prepare_all")) return 1;
    }
  }
  if (a.only_x) return 0;
  prof_mark(stream, 1);
  // The Xb scale is computed on the device (k_prepare_all) and read by the
  // epilogues through a device scalar, so the whole op stays sync-free.
  // R scale: |R| <= |cot| for the logistic family (|dl/dz| <= 1, mask in [0,1]).
  float r_scale = 1.0f;
  if (split) {
    int e;
    frexpf(fabsf(a.cot) > 0.f ? fabsf(a.cot) : 1.0f, &e);
    r_scale = ldexpf(1.0f, 12 - e);
  }

  TcLinkEpi link{};
  link.theta = a.theta; link.P = a.P; link.aux_off = a.spec.aux_off;
  link.y = a.y; link.idx = a.idx; link.mask = a.mask;
  link.row_scale = w.row_scale; link.b_scale = w.x_scale;
  link.cot = a.cot; link.r_scale = r_scale;
  link.ell = a.ell_requested ? a.ell : nullptr;
  link.stats = w.stats; link.parts = (int)((n + BN - 1) / BN);
  link.counters = w.counters; link.row_sumsq = w.row_sumsq;
  link.potential = a.potential; link.variance = a.variance;
  link.n_obs = (float)a.N; link.inv_temperature = 1.0f / a.spec.temperature;
  link.prior_half_inv = gauss_prior ? 0.5f / (a.spec.prior_scale * a.spec.prior_scale) : 0.f;
  link.r_hi = (__half*)w.r_hi; link.r_lo = w.r_lo; link.r_bf = (__nv_bfloat16*)w.r_hi;
  link.C = (int)C; link.n = (int)n;
  TcGradEpi gradp{};
  gradp.theta = a.theta; gradp.grad = a.grad; gradp.P = a.P; gradp.w_off = a.spec.w_off;
  gradp.d = d; gradp.C = (int)C;
  gradp.prior_lo = prior_lo; gradp.prior_hi = prior_hi;
  gradp.prior_coef = gauss_prior
      ? (1.0f / (a.spec.prior_scale * a.spec.prior_scale)) / a.spec.temperature : 0.f;
  gradp.xt_scale = w.x_scale;
  gradp.r_scale = r_scale;

  // ---- default: both contractions in one persistent launch -----------------------
  if (!legacy) {
    const int cgn = option(SGMC_OPT_TC_CTA_GROUP) == 1 ? 1 : 2;
    const int bn = option(SGMC_OPT_TC_TILE_N) == 128 ? 128 : 256;
    PairMaps maps;
    PairSched sch;
    sch.mt = (int)((C + BM * cgn - 1) / (BM * cgn));
    sch.nt1 = (int)((n + bn - 1) / bn);
    sch.nt2 = a.grad ? (d + bn - 1) / bn : 0;
    sch.kb1 = (d + kPrBK - 1) / kPrBK;
    sch.kb2 = (int)((n + kPrBK - 1) / kPrBK);
    sch.tiles1 = sch.mt * sch.nt1;
    sch.tiles_total = sch.tiles1 + sch.mt * sch.nt2;
    sch.dbg = option(SGMC_OPT_TC_TIMELINE);
    // The maps only depend on the workspace / gradient addresses and the shapes:
    // encode them once per (workspace, shape) instead of once per step.
    const MapKey mkey{base, a.grad, C, n, a.P, d, cgn, bn + a.x_slot, split ? 1 : 0};
    if (!map_cache_get(mkey, &maps)) {
      const int dt = split ? 0 : 1, brows = bn / cgn;
      if (make_map_ex(&maps.a[0][0], w.th_hi, dt, C, d, BM, kPrBK, kPrSwz)) return 2;
      if (make_map_ex(&maps.b[0][0], w.xb_hi, dt, n, d, brows, kPrBK, kPrSwz)) return 2;
      if (make_map_ex(&maps.a[1][0], w.r_hi, dt, C, n, BM, kPrBK, kPrSwz)) return 2;
      if (make_map_ex(&maps.b[1][0], w.xt_hi, dt, d, n, brows, kPrBK, kPrSwz)) return 2;
      if (make_map_ex(&maps.r[0], w.r_hi, dt, C, n, 32, 32, 64)) return 2;
      if (split) {
        if (make_map_ex(&maps.a[0][1], w.th_lo, 0, C, d, BM, kPrBK, kPrSwz)) return 2;
        if (make_map_ex(&maps.b[0][1], w.xb_lo, 0, n, d, brows, kPrBK, kPrSwz)) return 2;
        if (make_map_ex(&maps.a[1][1], w.r_lo, 0, C, n, BM, kPrBK, kPrSwz)) return 2;
        if (make_map_ex(&maps.b[1][1], w.xt_lo, 0, d, n, brows, kPrBK, kPrSwz)) return 2;
        if (make_map_ex(&maps.r[1], w.r_lo, 0, C, n, 32, 32, 64)) return 2;
      } else {
        maps.a[0][1] = maps.a[0][0]; maps.b[0][1] = maps.b[0][0];
        maps.a[1][1] = maps.a[1][0]; maps.b[1][1] = maps.b[1][0];
        maps.r[1] = maps.r[0];
      }
      if (a.grad) {
        if (make_map_ex(&maps.g, a.grad, 2, C, a.P, 32, 32, 128)) return 2;
      } else {
        maps.g = maps.r[0];
      }
      map_cache_put(mkey, maps);
    }
    PairNoise nz{};
    if (carry) {   // the update kernel adds the prior gradient (it reads theta anyway)
      cc->out.prior_coef = gradp.prior_coef;
      gradp.prior_lo = gradp.prior_hi = 0;
      // the step's noise is generated under the mainloops when the layout allows
      if (cc->prng_layout == 0 && d % 64 == 0 && !option(SGMC_OPT_NO_SHADOW_NOISE)) {
        nz.xi = w.xi; nz.noise_keys = w.noise_keys; nz.d = d;
        nz.units_per_chain = d / 64;
        nz.units_total = (int)(C * (d / 64));
        cc->out.xi = w.xi;
      }
    }
    // the update of sgmc_glm_sgld_step at the tail of the same launch
    PairUpdate upd{};
    if (carry && nz.xi != nullptr && fu.requested && !fu.write_grad && fu.layout == 0 &&
        option(SGMC_OPT_FUSED_PAIR_UPDATE) && !option(SGMC_OPT_EXACT_UPDATE_MATH) && d % 256 == 0 &&
        (prior_hi == prior_lo || (prior_lo == 0 && prior_hi == d)) &&
        (cc->out.prior_coef != 0.f) == (prior_hi > prior_lo) && aligned) {
      upd.enabled = 1;
      upd.rms = fu.v != nullptr;
      upd.grid_counter = w.counters + 2 * ((C + BM - 1) / BM);
      upd.n_tiles = C * (int64_t)sgld_split_tiles_per_chain(d);
      ApplyTileArgs& t = upd.t;
      t.theta = fu.theta_rw; t.v = fu.v; t.grad = a.grad; t.xi = w.xi;
      t.th_hi = w.th_hi; t.th_lo = w.th_lo; t.scale = w.next_scale;
      t.amax_bits = w.amax_row; t.sumsq_part = w.sumsq_part; t.P = a.P;
      t.tiles_per_chain = (uint32_t)sgld_split_tiles_per_chain(d);
      t.prior_on = prior_hi > prior_lo ? 1 : 0;
      t.prior_coef = cc->out.prior_coef;
      t.noise_scale = sqrtf((2.0f * fu.temperature) * fu.step_size);   // integrator.py:882-884
      t.neg_eps = -fu.step_size; t.alpha = fu.alpha; t.one_m_alpha = 1.0f - fu.alpha;
      t.lmbd = fu.lmbd;
      *fu.applied = true;
    }
#define SGMC_PAIR_CASE(T, F, G, B, NAME) \
    if (cgn == G && bn == B) return launch_pair<T, F, G, B>(stream, maps, sch, link, gradp, nz, upd, NAME)
    if (split) {
      SGMC_PAIR_CASE(3, 0, 2, 128, "k_glm_tc_pair<split,2,128>");
      SGMC_PAIR_CASE(3, 0, 2, 256, "k_glm_tc_pair<split,2,256>");
      SGMC_PAIR_CASE(3, 0, 1, 128, "k_glm_tc_pair<split,1,128>");
      SGMC_PAIR_CASE(3, 0, 1, 256, "k_glm_tc_pair<split,1,256>");
    } else {
      SGMC_PAIR_CASE(1, 1, 2, 128, "k_glm_tc_pair<bf16,2,128>");
      SGMC_PAIR_CASE(1, 1, 2, 256, "k_glm_tc_pair<bf16,2,256>");
      SGMC_PAIR_CASE(1, 1, 1, 128, "k_glm_tc_pair<bf16,1,128>");
      SGMC_PAIR_CASE(1, 1, 1, 256, "k_glm_tc_pair<bf16,1,256>");
    }
#undef SGMC_PAIR_CASE
    SGMC_REQUIRE(false, "no kernel for cta_group %d, tile %d", cgn, bn);
  }

  CUtensorMap mA0, mA1, mB0, mB1;
  // ---- legacy two-kernel sequence (SGMC_OPT_TC_LEGACY, the opt-in fused-update modes) ----
  // ---- GEMM1: Z[C,n] = Theta[C,d] . Xb[n,d]^T ------------------------------
  if (make_map(&mA0, w.th_hi, !split, C, d, BM)) return 2;
  if (make_map(&mB0, w.xb_hi, !split, n, d, BN)) return 2;
  if (split) {
    if (make_map(&mA1, w.th_lo, 0, C, d, BM)) return 2;
    if (make_map(&mB1, w.xb_lo, 0, n, d, BN)) return 2;
    if (launch_gemm<3, 0, 0>(stream, mA0, mA1, mB0, mB1, (int)C, (int)n, d, link, gradp,
                             "k_glm_tc_gemm<split,link>", job1)) return 1;
  } else {
    if (launch_gemm<1, 0, 1>(stream, mA0, mA0, mB0, mB0, (int)C, (int)n, d, link, gradp,
                             "k_glm_tc_gemm<bf16,link>", job1)) return 1;
  }
  // U and var(ell) were finalised inside GEMM1 by the last CTA of every row block.
  // ---- GEMM2 with the SGLD / pSGLD update in its epilogue -----------------------------
  // Needs whole tiles, the sample to be exactly the weight vector (one leaf, so
  // feature j shares its threefry block with j + d/2) and the default (fast)
  // preconditioner arithmetic; otherwise the caller runs the stand-alone update.
  if (epilogue_update) {
    gradp.theta_rw = fu.theta_rw; gradp.v = fu.v;
    gradp.keys_in = fu.keys_in; gradp.keys_out = fu.keys_out;
    gradp.neg_eps = -fu.step_size;
    gradp.noise_scale = sqrtf((2.0f * fu.temperature) * fu.step_size);   // integrator.py:882-884
    gradp.alpha = fu.alpha; gradp.one_m_alpha = 1.0f - fu.alpha; gradp.lmbd = fu.lmbd;
    gradp.half = (uint32_t)(d / 2);
    if (!fu.write_grad) gradp.grad = nullptr;
    if (make_map(&mA0, w.r_hi, !split, C, n, BM)) return 2;
    if (make_map(&mB0, w.xt_hi, !split, d, n, BN / 2)) return 2;
    if (split) {
      if (make_map(&mA1, w.r_lo, 0, C, n, BM)) return 2;
      if (make_map(&mB1, w.xt_lo, 0, d, n, BN / 2)) return 2;
      if (launch_gemm<3, 2, 0>(stream, mA0, mA1, mB0, mB1, (int)C, d, (int)n, link, gradp,
                               "k_glm_tc_gemm<split,grad+sgld>")) return 1;
    } else {
      if (launch_gemm<1, 2, 1>(stream, mA0, mA0, mB0, mB0, (int)C, d, (int)n, link, gradp,
                               "k_glm_tc_gemm<bf16,grad+sgld>")) return 1;
    }
    *fu.applied = true;
    return 0;
  }
  if (!a.grad) return 0;
  // ---- GEMM2: G[C,d] = R[C,n] . XbT[d,n]^T ---------------------------------------
  if (make_map(&mA0, w.r_hi, !split, C, n, BM)) return 2;
  if (make_map(&mB0, w.xt_hi, !split, d, n, BN)) return 2;
  if (split) {
    if (make_map(&mA1, w.r_lo, 0, C, n, BM)) return 2;
    if (make_map(&mB1, w.xt_lo, 0, d, n, BN)) return 2;
    if (launch_gemm<3, 1, 0>(stream, mA0, mA1, mB0, mB1, (int)C, d, (int)n, link, gradp,
                             "k_glm_tc_gemm<split,grad>", job2)) return 1;
  } else {
    if (launch_gemm<1, 1, 1>(stream, mA0, mA0, mB0, mB0, (int)C, d, (int)n, link, gradp,
                             "k_glm_tc_gemm<bf16,grad>", job2)) return 1;
  }
  if (noise_job) {
    const int64_t n4 = C * (int64_t)d / 4;
    const unsigned grid = (unsigned)std::min<int64_t>((n4 + 255) / 256, (int64_t)sm_count() * 8);
    const float ns = sqrtf((2.0f * fu.temperature) * fu.step_size);   // integrator.py:882-884
    if (fu.v)
      launch_pdl(k_sgld_apply<true>, dim3(grid), dim3(256), 0, stream, fu.theta_rw, fu.v,
                 (const float*)a.grad, (const float*)w.xi, n4, -fu.step_size, ns, fu.alpha,
                 1.0f - fu.alpha, fu.lmbd);
    else
      launch_pdl(k_sgld_apply<false>, dim3(grid), dim3(256), 0, stream, fu.theta_rw,
                 (float*)nullptr, (const float*)a.grad, (const float*)w.xi, n4, -fu.step_size,
                 ns, 0.f, 0.f, 0.f);
    if (post_launch("k_sgld_apply")) return 1;
    *fu.applied = true;
  }
  return 0;
}

}  // namespace sgmc
