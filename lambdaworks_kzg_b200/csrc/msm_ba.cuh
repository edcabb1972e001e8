// Batched-affine accumulation kernel of the fixed-base MSM (large batches).  One translation unit per variant
// (msm_ba_v*.cu) so that the variants compile in parallel.
#pragma once
#include "msm_common.cuh"

namespace lw {

// ---------------------------------------------------------------------------
// Batched-affine variant of the gather kernel (large batches).
//
// The XYZZ mixed addition above costs 8 M + 2 S on the integer-multiply pipe, and
// that pipe is the bound (profiles/r01_ncu_msm_summary.md, r01_multiplier_experiments.md).
// An AFFINE addition costs 2 M + 1 S plus one inversion of (x2 - x1); with
// Montgomery's trick K independent additions share one inversion for 3 M each:
// 5 M + 1 S per accumulated table entry.  Each thread therefore keeps K affine
// accumulators (in an L2-resident scratch area, interleaved so that a warp's
// 128-bit accesses are contiguous) and consumes its table entries K at a time:
//   pass 1: d_k = T_k.x - A_k.x, exclusive prefix products (stored)
//   one inversion of the total product (fpinv.cuh: binary GCD, mostly ALU work)
//   pass 2 (k descending): 1/d_k from the running inverse and the stored prefix,
//           A_k <- A_k + T_k
// Rare cases (accumulator at infinity, equal x) are flagged per slot and never
// enter the product.  At the end the K accumulators are folded into one XYZZ
// sum and the block reduces as before, so partials / finalize are unchanged.
LW_COLD Fp fp_inv_gcd_ni(Fp y) { return fp_inv_gcd(y); }
LW_COLD G1Affine g1a_dbl_ni(G1Affine p) { return xyzz_to_affine(xyzz_dbl_affine(p)); }

constexpr uint32_t BA_NONE = 0xffffffffu;

template <int TH>
__device__ __forceinline__ Fp load_fp_scratch(const uint4* p /* 3 words, stride TH */) {
  uint4 v0 = __ldcg(p), v1 = __ldcg(p + TH), v2 = __ldcg(p + 2 * TH);
  Fp e;
  e.l[0] = v0.x; e.l[1] = v0.y; e.l[2] = v0.z; e.l[3] = v0.w;
  e.l[4] = v1.x; e.l[5] = v1.y; e.l[6] = v1.z; e.l[7] = v1.w;
  e.l[8] = v2.x; e.l[9] = v2.y; e.l[10] = v2.z; e.l[11] = v2.w;
  return e;
}
template <int TH>
__device__ __forceinline__ void store_fp_scratch(uint4* p, const Fp& e) {
  __stcg(p, make_uint4(e.l[0], e.l[1], e.l[2], e.l[3]));
  __stcg(p + TH, make_uint4(e.l[4], e.l[5], e.l[6], e.l[7]));
  __stcg(p + 2 * TH, make_uint4(e.l[8], e.l[9], e.l[10], e.l[11]));
}
__device__ __forceinline__ Fp load_entry_x(const uint4* __restrict__ table, size_t idx) {
  const uint4* p = table + idx * 6;
  uint4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
  Fp e;
  e.l[0] = v0.x; e.l[1] = v0.y; e.l[2] = v0.z; e.l[3] = v0.w;
  e.l[4] = v1.x; e.l[5] = v1.y; e.l[6] = v1.z; e.l[7] = v1.w;
  e.l[8] = v2.x; e.l[9] = v2.y; e.l[10] = v2.z; e.l[11] = v2.w;
  return e;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_entry_l2(const uint4* table, size_t idx) {
  const uint4* p = table + idx * 6;
  prefetch_l2(p);
  prefetch_l2(p + 4);
}
template <int TH>
__device__ __forceinline__ void prefetch_fp_scratch(const uint4* p) {
  prefetch_l2(p); prefetch_l2(p + TH); prefetch_l2(p + 2 * TH);
}
template <int TH>
__device__ __forceinline__ void prefetch_slot_l2(const uint4* p) {
#pragma unroll
  for (int w = 0; w < 9; w++) prefetch_l2(p + w * TH);
}

template <bool BE, int K, int MINB, int TH>
__global__ void __launch_bounds__(TH, MINB)
msm_gather_ba_kernel(G1Xyzz* __restrict__ partials, const uint4* __restrict__ table, const uint8_t* __restrict__ scalars,
                     uint4* __restrict__ scratch, int c, int nwin, uint32_t cnt_top) {
  static_assert(K <= 64, "slot masks are 64 bits wide");
  constexpr int HT = TH / 2;              // threads per GLV half
  constexpr int NPT = N_POINTS / HT;      // points per thread
  __shared__ uint32_t sk[4][TH];
  __shared__ uint32_t sidx[K][TH];   // entry index | sign << 31, or BA_NONE
  __shared__ uint32_t red[48 * HT];

  const int tid = threadIdx.x;
  const int half = tid / HT, hl = tid % HT;
  const int blob = blockIdx.x;
  const uint8_t* sc = scalars + (size_t)blob * BLOB_BYTES;
  auto limb = [&](int w) { return sk[w][tid]; };
  // slot k: words 0-2 = A.x, 3-5 = A.y, 6-8 = exclusive prefix product
  uint4* my = scratch + (size_t)blob * (K * 9 * TH) + tid;

  // A round consumes the 2^c-ary digits of PPR points: slot p * nwin + j <-> (point p of the round, window j).
  // Zero digits (probability 2^-c for uniform scalars) leave their slot idle for the round.
  const int PPR = K / nwin;
  const int rounds = (NPT + PPR - 1) / PPR;
  uint64_t infmask = K >= 64 ? ~0ull : ((1ull << K) - 1ull);   // accumulators at infinity
  for (int k = PPR * nwin; k < K; k++) sidx[k][tid] = BA_NONE;

#pragma unroll 1
  for (int rnd = 0; rnd < rounds; rnd++) {
    // ------------------------------------------------------------ pass 1a: digits -> entry indices; their table
    // lines and the accumulator rows are requested into L2 so that pass 1b finds them there
#pragma unroll 1
    for (int p = 0; p < PPR; p++) {
      const int t_idx = rnd * PPR + p;
      const bool have = t_idx < NPT;
      const int pi = hl + HT * t_idx;
      if (have) {
        uint32_t h4[4];
        load_scalar_half<BE>(h4, sc, pi, half);
#pragma unroll
        for (int i = 0; i < 4; i++) sk[i][tid] = h4[i];
      }
      int carry = 0;
#pragma unroll 1
      for (int j = 0; j < nwin; j++) {
        const int k = p * nwin + j;
        uint32_t e = BA_NONE;
        if (have) {
          const int d = glv_digit(limb, c, nwin, j, carry);
          if (d != 0) e = entry_index(c, nwin, cnt_top, j, pi, d < 0 ? -d : d) | (d < 0 ? 0x80000000u : 0u);
        }
        sidx[k][tid] = e;
        if (e != BA_NONE) {
          prefetch_l2(table + (size_t)(e & 0x7fffffffu) * 6);
          if (!((infmask >> k) & 1ull)) prefetch_fp_scratch<TH>(my + (k * 9) * TH);
        }
      }
    }
    // ------------------------------------------------------------ pass 1b: differences and prefix products
    Fp prod = fp_one();
    uint64_t specmask = 0;   // slots with T.x == A.x
    uint32_t e_next = sidx[0][tid];
    Fp tx_next = fp_zero(), ax_next = fp_zero();
    if (e_next != BA_NONE) {
      tx_next = load_entry_x(table, e_next & 0x7fffffffu);
      if (!(infmask & 1ull)) ax_next = load_fp_scratch<TH>(my);
    }
#pragma unroll 1
    for (int k = 0; k < K; k++) {
      uint32_t e = e_next;
      const Fp tx = tx_next, ax = ax_next;
      if (k + 1 < K) {   // operands of the next slot are in flight while this one multiplies
        e_next = sidx[k + 1][tid];
        if (e_next != BA_NONE) {
          tx_next = load_entry_x(table, e_next & 0x7fffffffu);
          if (!((infmask >> (k + 1)) & 1ull)) ax_next = load_fp_scratch<TH>(my + ((k + 1) * 9) * TH);
        }
      }
      if (e == BA_NONE) continue;
      if (fp_is_zero(tx)) {
        // (0, 0) encodes infinity in the table (hand-built setups); x == 0 with y != 0 is a curve point
        const G1Affine t = load_entry(table, e & 0x7fffffffu);
        if (fp_is_zero(t.y)) { sidx[k][tid] = BA_NONE; continue; }
      }
      if ((infmask >> k) & 1ull) continue;
      const Fp d = fp_sub(tx, ax);
      if (fp_is_zero(d)) {
        specmask |= 1ull << k;
      } else {
        store_fp_scratch<TH>(my + (k * 9 + 6) * TH, prod);
        prod = fp_mul_nv(prod, d);
      }
    }
    // ------------------------------------------------------------ shared inversion
    Fp inv = fp_inv_gcd_ni(prod);
    // ------------------------------------------------------------ pass 2
    for (int k = K - 1; k >= K - 2 && k >= 0; k--) {
      const uint32_t e = sidx[k][tid];
      if (e != BA_NONE) { prefetch_entry_l2(table, e & 0x7fffffffu); prefetch_slot_l2<TH>(my + (k * 9) * TH); }
    }
#pragma unroll 1
    for (int k = K - 1; k >= 0; k--) {
      if (k >= 2) {
        const uint32_t e2 = sidx[k - 2][tid];
        if (e2 != BA_NONE) { prefetch_entry_l2(table, e2 & 0x7fffffffu); prefetch_slot_l2<TH>(my + ((k - 2) * 9) * TH); }
      }
      const uint32_t e = sidx[k][tid];
      if (e == BA_NONE) continue;
      G1Affine t = load_entry(table, e & 0x7fffffffu);
      t.y = fp_cneg(t.y, (e >> 31) != 0);
      uint4* slot = my + (k * 9) * TH;
      if ((infmask >> k) & 1ull) {
        store_fp_scratch<TH>(slot, t.x);
        store_fp_scratch<TH>(slot + 3 * TH, t.y);
        infmask &= ~(1ull << k);
        continue;
      }
      const Fp ax = load_fp_scratch<TH>(slot), ay = load_fp_scratch<TH>(slot + 3 * TH);
      if ((specmask >> k) & 1ull) {
        if (fp_eq(t.y, ay)) {
          const G1Affine dd = g1a_dbl_ni(t);
          store_fp_scratch<TH>(slot, dd.x);
          store_fp_scratch<TH>(slot + 3 * TH, dd.y);
        } else {
          infmask |= 1ull << k;   // T == -A
        }
        continue;
      }
      const Fp ex = load_fp_scratch<TH>(slot + 6 * TH);
      const Fp d = fp_sub(t.x, ax);
      const Fp dinv = fp_mul_nv(inv, ex);
      inv = fp_mul_nv(inv, d);
      const Fp lam = fp_mul_nv(fp_sub(t.y, ay), dinv);
      const Fp x3 = fp_sub(fp_sub(fp_sqr_nv(lam), ax), t.x);
      const Fp y3 = fp_sub(fp_mul_nv(lam, fp_sub(ax, x3)), ay);
      store_fp_scratch<TH>(slot, x3);
      store_fp_scratch<TH>(slot + 3 * TH, y3);
    }
  }

  // fold the K accumulators, then the block
  G1Xyzz acc = xyzz_inf();
#pragma unroll 1
  for (int k = 0; k < K; k++) {
    if ((infmask >> k) & 1ull) continue;
    G1Affine a;
    a.x = load_fp_scratch<TH>(my + (k * 9) * TH);
    a.y = load_fp_scratch<TH>(my + (k * 9 + 3) * TH);
    xyzz_madd_hot(acc, a);
  }
  block_reduce_xyzz_glv<TH>(acc, red);
  if (tid == 0) partials[blob] = acc;
}

template <int K, int MINB, int TH>
static void launch_ba(void* d_partials, const void* d_table, int c, const void* d_scalars, bool be_input, int n_blobs,
                      void* d_scratch, cudaStream_t st) {
  const int nwin = glv_num_windows(c);
  const uint32_t cnt_top = glv_top_max(c) + 1u;
  if (be_input)
    msm_gather_ba_kernel<true, K, MINB, TH><<<n_blobs, TH, 0, st>>>((G1Xyzz*)d_partials, (const uint4*)d_table, (const uint8_t*)d_scalars, (uint4*)d_scratch, c, nwin, cnt_top);
  else
    msm_gather_ba_kernel<false, K, MINB, TH><<<n_blobs, TH, 0, st>>>((G1Xyzz*)d_partials, (const uint4*)d_table, (const uint8_t*)d_scalars, (uint4*)d_scratch, c, nwin, cnt_top);
}

}  // namespace lw
