// Shared pieces of the fixed-base MSM kernels (msm.cu: XYZZ accumulation, msm_ba.cuh: batched-affine accumulation).
#pragma once
#include "g1.cuh"
#include "fpinv.cuh"
#include "kernels.h"
#include "recode.cuh"

namespace lw {

#ifndef LWKZG_MSM_THREADS
#define LWKZG_MSM_THREADS 128
#endif
#ifndef LWKZG_MSM_MIN_BLOCKS
#define LWKZG_MSM_MIN_BLOCKS 3
#endif
constexpr int MSM_THREADS = LWKZG_MSM_THREADS;


__device__ __forceinline__ G1Affine load_entry(const uint4* __restrict__ table, size_t idx) {
  const uint4* p = table + idx * 6;
  uint4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3), v4 = __ldg(p + 4), v5 = __ldg(p + 5);
  G1Affine e;
  e.x.l[0] = v0.x; e.x.l[1] = v0.y; e.x.l[2] = v0.z; e.x.l[3] = v0.w;
  e.x.l[4] = v1.x; e.x.l[5] = v1.y; e.x.l[6] = v1.z; e.x.l[7] = v1.w;
  e.x.l[8] = v2.x; e.x.l[9] = v2.y; e.x.l[10] = v2.z; e.x.l[11] = v2.w;
  e.y.l[0] = v3.x; e.y.l[1] = v3.y; e.y.l[2] = v3.z; e.y.l[3] = v3.w;
  e.y.l[4] = v4.x; e.y.l[5] = v4.y; e.y.l[6] = v4.z; e.y.l[7] = v4.w;
  e.y.l[8] = v5.x; e.y.l[9] = v5.y; e.y.l[10] = v5.z; e.y.l[11] = v5.w;
  return e;
}

// XYZZ <-> shared memory in limb-major (SoA) layout: word w of thread t lives at
// smem[w * stride + t]  -> conflict-free for a warp.
__device__ __forceinline__ void xyzz_to_smem(uint32_t* smem, int stride, int t, const G1Xyzz& p) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 48; i++) smem[i * stride + t] = w[i];
}
__device__ __forceinline__ G1Xyzz xyzz_from_smem(const uint32_t* smem, int stride, int t) {
  G1Xyzz p;
  uint32_t* w = reinterpret_cast<uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 48; i++) w[i] = smem[i * stride + t];
  return p;
}

// Block-wide sum of the per-thread XYZZ accumulators; result valid in thread 0.  THREADS is a power of two.  The
// upper half of the block holds sums of q-halves: psi (X -> beta X, Y -> -Y; ZZ, ZZZ unchanged) is applied to them
// at the first tree level, where they meet the m-half sums of the lower threads.
LW_COLD void xyzz_psi_ni(G1Xyzz& p) {
  Fp beta;
  for (int i = 0; i < 12; i++) beta.l[i] = k::FP_BETA[i];
  p.x = fp_mul(p.x, beta);
  p.y = fp_neg(p.y);
}
template <int THREADS>
__device__ __forceinline__ void block_reduce_xyzz_glv(G1Xyzz& acc, uint32_t* red /* 48 * THREADS / 2 words */) {
  static_assert((THREADS & (THREADS - 1)) == 0 && THREADS >= 2, "power-of-two block");
  constexpr int H = THREADS / 2;
  const int tid = threadIdx.x;
  for (int s = H; s > 0; s >>= 1) {
    if (tid >= s && tid < 2 * s) xyzz_to_smem(red, H, tid - s, acc);
    __syncthreads();
    if (tid < s) {
      G1Xyzz o = xyzz_from_smem(red, H, tid);
      if (s == H) xyzz_psi_ni(o);
      xyzz_add_ni(acc, o);
    }
    __syncthreads();
  }
}

// One scalar -> the 128-bit GLV half this thread works on (4 little-endian limbs).  BE: raw big-endian blob word,
// reduced mod r (App. A.1); else canonical little-endian limbs (reduced again so that a caller's garbage can never
// produce a digit outside the table).
template <bool BE>
__device__ __forceinline__ void load_scalar_half(uint32_t* h4, const uint8_t* __restrict__ sc, int pi, int half) {
  const uint4* sp = reinterpret_cast<const uint4*>(sc + (size_t)pi * 32);
  const uint4 a = __ldg(sp), b = __ldg(sp + 1);
  Fr k;
  if (BE) {
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    k = fr_canon_from_be_words(w);   // bswap + reduce mod r
  } else {
    k.l[0] = a.x; k.l[1] = a.y; k.l[2] = a.z; k.l[3] = a.w;
    k.l[4] = b.x; k.l[5] = b.y; k.l[6] = b.z; k.l[7] = b.w;
    mod_reduce_small<FrCfg, 2>(k.l);
  }
  uint32_t q[4], m[4];
  glv_split_barrett(q, m, k.l);
#pragma unroll
  for (int i = 0; i < 4; i++) h4[i] = half ? q[i] : m[i];
}

// table entry of digit magnitude `mag` >= 1 of window j, point pi
__device__ __forceinline__ uint32_t entry_index(int c, int nwin, uint32_t cnt_top, int j, int pi, int mag) {
  const uint32_t cnt = (j == nwin - 1) ? cnt_top : (1u << (c - 1));
  return (((uint32_t)j * N_POINTS) << (c - 1)) + (uint32_t)pi * cnt + (uint32_t)(mag - 1);
}

}  // namespace lw
