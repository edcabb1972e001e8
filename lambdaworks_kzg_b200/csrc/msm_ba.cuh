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

#ifndef LWKZG_BA_PREFETCH
#define LWKZG_BA_PREFETCH false
#endif
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

// ---- operand staging through shared memory (cp.async)
// Every table entry, accumulator and prefix product a slot needs is requested one slot AHEAD with cp.async into a
// per-thread column of shared memory.  Two reasons it has to be cp.async and not loads into registers: (1) the
// multiplier is an out-of-line function and a CALL waits for every outstanding register load, so register
// prefetching stalls at the first product of each slot (ncu: `CALL fp_mul_nv` and the first use of the loaded
// words were the top stall sites, 13 % of all warp samples); (2) the 60 registers a slot's operands occupy are
// not available under the 168-register budget of three blocks per SM.
__device__ __forceinline__ void cp_async16(uint4* smem_dst, const uint4* gsrc) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int TH>
__device__ __forceinline__ Fp fp_from_stage(const uint4* p /* 3 words, stride TH */) {
  const uint4 v0 = p[0], v1 = p[TH], v2 = p[2 * TH];
  Fp e;
  e.l[0] = v0.x; e.l[1] = v0.y; e.l[2] = v0.z; e.l[3] = v0.w;
  e.l[4] = v1.x; e.l[5] = v1.y; e.l[6] = v1.z; e.l[7] = v1.w;
  e.l[8] = v2.x; e.l[9] = v2.y; e.l[10] = v2.z; e.l[11] = v2.w;
  return e;
}

template <int TH>
__device__ __forceinline__ void fp_to_stage(uint4* p /* 3 words, stride TH */, const Fp& e) {
  p[0] = make_uint4(e.l[0], e.l[1], e.l[2], e.l[3]);
  p[TH] = make_uint4(e.l[4], e.l[5], e.l[6], e.l[7]);
  p[2 * TH] = make_uint4(e.l[8], e.l[9], e.l[10], e.l[11]);
}

constexpr int BA_STAGE_WORDS = 15;   // T.x T.y (6) + A.x A.y (6) + prefix product (3) 128-bit words per thread
template <int K, int TH>
constexpr size_t ba_smem_bytes() { return (size_t)BA_STAGE_WORDS * TH * 16 + (size_t)K * TH * 2 + 4 * TH * 4; }

// A slot's digit is kept as 16 bits (two's complement for the signed windows -- 0x8000 can only be +2^15, the
// most negative digit is -2^15 + 1 -- and plain unsigned for the top window); 0 = nothing to add this round.
__device__ __forceinline__ uint32_t ba_entry_of(uint32_t code, int c, int nwin, uint32_t cnt_top, int j, int pi) {
  if (code == 0) return BA_NONE;
  int d = (int)code;
  if (j < nwin - 1 && code > 0x8000u) d = (int)code - 65536;
  return entry_index(c, nwin, cnt_top, j, pi, d < 0 ? -d : d) | (d < 0 ? 0x80000000u : 0u);
}

template <bool BE, int K, int TH, bool PF>
__device__ __forceinline__ void msm_gather_ba_body(G1Xyzz* __restrict__ partials, const uint4* __restrict__ table, const uint8_t* __restrict__ scalars,
                                                   uint4* __restrict__ scratch, int c, int nwin, uint32_t cnt_top, int seg_stride, int split) {
  static_assert(K <= 64, "slot masks are 64 bits wide");
  static_assert(48 * (TH / 2) * 4 <= BA_STAGE_WORDS * TH * 16, "the block-reduction scratch aliases the staging area");
  constexpr int HT = TH / 2;              // threads per GLV half
  constexpr int NPT = N_POINTS / HT;      // points per thread
  extern __shared__ uint4 dyn_smem[];
  uint4* stage = dyn_smem + threadIdx.x;                                                  // word w at stage[w * TH]
  uint16_t (*sdig)[TH] = reinterpret_cast<uint16_t (*)[TH]>(dyn_smem + BA_STAGE_WORDS * TH);   // digit codes of the round
  uint32_t (*sk)[TH] = reinterpret_cast<uint32_t (*)[TH]>(reinterpret_cast<uint8_t*>(dyn_smem + BA_STAGE_WORDS * TH) + (size_t)K * TH * 2);
  uint32_t* red = reinterpret_cast<uint32_t*>(dyn_smem);

  const int tid = threadIdx.x;
  const int half = tid / HT, hl = tid % HT;
  // split > 1 (small batches): `split` blocks share a blob, block `part` takes the points p in [part, part + 1) * NPT / split
  // of every thread's residue class and writes its own partial sum; the finalize kernel adds them
  const int blob = blockIdx.x / split, part = blockIdx.x % split;
  const int npt = NPT / split, p_begin = part * npt;
  const uint8_t* sc = scalars + (size_t)blob * BLOB_BYTES;
  auto limb = [&](int w) { return sk[w][tid]; };
  // slot k: words 0-2 = A.x, 3-5 = A.y, 6-8 = exclusive prefix product
  uint4* my = scratch + (size_t)blockIdx.x * (K * 9 * TH) + tid;

  // A round consumes the 2^c-ary digits of PPR points: slot p * nwin + j <-> (point p of the round, window j).
  // Zero digits (probability 2^-c for uniform scalars) leave their slot idle for the round.
  const int PPR = K / nwin;
  const int rounds = (npt + PPR - 1) / PPR;
  uint64_t infmask = K >= 64 ? ~0ull : ((1ull << K) - 1ull);   // accumulators at infinity
  for (int k = PPR * nwin; k < K; k++) sdig[k][tid] = 0;

  // requests of pass 1b (T.x, A.x) and pass 2 (T, A, prefix product) for slot k with entry e
  auto request1 = [&](int k, uint32_t e) {
    if (e != BA_NONE) {
      const uint4* tp = table + (size_t)(e & 0x7fffffffu) * 6;
      cp_async16(stage, tp); cp_async16(stage + TH, tp + 1); cp_async16(stage + 2 * TH, tp + 2);
      if (!((infmask >> k) & 1ull)) {
        const uint4* ap = my + (k * 9) * TH;
        cp_async16(stage + 6 * TH, ap); cp_async16(stage + 7 * TH, ap + TH); cp_async16(stage + 8 * TH, ap + 2 * TH);
      }
    }
    cp_async_commit();
  };
  auto request2 = [&](int k, uint32_t e) {
    if (e != BA_NONE) {
      const uint4* tp = table + (size_t)(e & 0x7fffffffu) * 6;
#pragma unroll
      for (int w = 0; w < 6; w++) cp_async16(stage + w * TH, tp + w);
      if (!((infmask >> k) & 1ull)) {
        const uint4* ap = my + (k * 9) * TH;
#pragma unroll
        for (int w = 0; w < 9; w++) cp_async16(stage + (6 + w) * TH, ap + w * TH);
      }
    }
    cp_async_commit();
  };

#pragma unroll 1
  for (int rnd = 0; rnd < rounds; rnd++) {
    const int pi0 = hl + HT * (p_begin + rnd * PPR);   // point of slot group p: pi0 + HT * p
    // ------------------------------------------------------------ pass 1a: digits of the round's points
#pragma unroll 1
    for (int p = 0; p < PPR; p++) {
      const bool have = rnd * PPR + p < npt;
      const int pi = pi0 + HT * p;
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
        int d = 0;
        if (have) d = glv_digit(limb, c, nwin, j, carry);
        sdig[k][tid] = (uint16_t)d;
        if (PF && d != 0) {   // table line and accumulator row on their way into L2
          prefetch_l2(table + (size_t)entry_index(c, nwin, cnt_top, j, pi, d < 0 ? -d : d) * 6);
          if (!((infmask >> k) & 1ull)) prefetch_fp_scratch<TH>(my + (k * 9) * TH);
        }
      }
    }
    // ------------------------------------------------------------ pass 1b: differences and prefix products
    Fp prod = fp_one();
    uint64_t specmask = 0;   // slots with T.x == A.x
    uint32_t e_next = ba_entry_of(sdig[0][tid], c, nwin, cnt_top, 0, pi0);
    request1(0, e_next);
    int jn = 0, pin = pi0;   // (window, point) of slot k + 1
#pragma unroll 1
    for (int k = 0; k < K; k++) {
      const uint32_t e = e_next;
      cp_async_wait_all();
      const bool inf = (infmask >> k) & 1ull;
      Fp d = fp_zero();
      bool tx_zero = false;
      if (e != BA_NONE) {
        const Fp tx = fp_from_stage<TH>(stage);
        tx_zero = fp_is_zero(tx);
        if (!inf) d = fp_sub(tx, fp_from_stage<TH>(stage + 6 * TH));
      }
      // the staged words are consumed (d, tx_zero depend on all of them): the column may be overwritten
      if (k + 1 < K) {
        if (++jn == nwin) { jn = 0; pin += HT; }
        e_next = ba_entry_of(sdig[k + 1][tid], c, nwin, cnt_top, jn, pin);
        request1(k + 1, e_next);
      }
      if (e == BA_NONE) continue;
      if (tx_zero) {
        // (0, 0) encodes infinity in the table (hand-built setups); x == 0 with y != 0 is a curve point
        const G1Affine t = load_entry(table, e & 0x7fffffffu);
        if (fp_is_zero(t.y)) { sdig[k][tid] = 0; continue; }
      }
      if (inf) continue;
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
    // (window, point) of slot K - 1: slots >= PPR * nwin are empty whatever their coordinates
    jn = (K - 1) % nwin;
    pin = pi0 + HT * ((K - 1) / nwin);
    e_next = ba_entry_of(sdig[K - 1][tid], c, nwin, cnt_top, jn, pin);
    request2(K - 1, e_next);
#pragma unroll 1
    for (int k = K - 1; k >= 0; k--) {
      const uint32_t e = e_next;
      if (k > 0) {
        if (--jn < 0) { jn = nwin - 1; pin -= HT; }
        e_next = ba_entry_of(sdig[k - 1][tid], c, nwin, cnt_top, jn, pin);
      }
      cp_async_wait_all();
      if (e == BA_NONE) {
        if (k > 0) request2(k - 1, e_next);
        continue;
      }
      G1Affine t;
      t.x = fp_from_stage<TH>(stage);
      t.y = fp_cneg(fp_from_stage<TH>(stage + 3 * TH), (e >> 31) != 0);
      uint4* slot = my + (k * 9) * TH;
      if ((infmask >> k) & 1ull) {
        store_fp_scratch<TH>(slot, t.x);
        store_fp_scratch<TH>(slot + 3 * TH, t.y);
        infmask &= ~(1ull << k);
        if (k > 0) request2(k - 1, e_next);
        continue;
      }
      const Fp ax = fp_from_stage<TH>(stage + 6 * TH), ay = fp_from_stage<TH>(stage + 9 * TH);
      if ((specmask >> k) & 1ull) {
        if (fp_eq(t.y, ay)) {
          const G1Affine dd = g1a_dbl_ni(t);
          store_fp_scratch<TH>(slot, dd.x);
          store_fp_scratch<TH>(slot + 3 * TH, dd.y);
        } else {
          infmask |= 1ull << k;   // T == -A
        }
        if (k > 0) request2(k - 1, e_next);
        continue;
      }
      const Fp d = fp_sub(t.x, ax);
      const Fp dy = fp_sub(t.y, ay);
      const Fp dinv = fp_mul_nv(inv, fp_from_stage<TH>(stage + 12 * TH));
      // every staged word of this slot now sits in a register: request the next slot while the remaining
      // four products run
      if (k > 0) request2(k - 1, e_next);
      const Fp lam = fp_mul_nv(dy, dinv);   // (before inv * d: dy and dinv die here, five field elements live across either call)
      inv = fp_mul_nv(inv, d);
      const Fp x3 = fp_sub(fp_sub(fp_sqr_nv(lam), ax), t.x);
      const Fp y3 = fp_sub(fp_mul_nv(lam, fp_sub(ax, x3)), ay);
      store_fp_scratch<TH>(slot, x3);
      store_fp_scratch<TH>(slot + 3 * TH, y3);
    }
  }

  // fold the K accumulators, then the block.  The running XYZZ sum lives in the (now idle) staging column of the
  // thread, not in registers: with it there the fold needs five live field elements instead of nine (1024-blob launch
  // 19.05 -> 18.9 ms).  Measured with it: the kernel at 128 registers / four blocks per SM, 180 bytes of spills left --
  // 19.1 ms, no gain from the fourth block (profiles/r02_ba_128reg_variant.log).
  // words 0-2 x, 3-5 y, 6-8 zz, 9-11 zzz
  cp_async_wait_all();
  bool acc_inf = true;
#pragma unroll 1
  for (int k = 0; k < K; k++) {
    if ((infmask >> k) & 1ull) continue;
    const Fp px = load_fp_scratch<TH>(my + (k * 9) * TH), py = load_fp_scratch<TH>(my + (k * 9 + 3) * TH);
    if (acc_inf) {
      const Fp one = fp_one();
      fp_to_stage<TH>(stage, px); fp_to_stage<TH>(stage + 3 * TH, py); fp_to_stage<TH>(stage + 6 * TH, one); fp_to_stage<TH>(stage + 9 * TH, one);
      acc_inf = false;
      continue;
    }
    const Fp Pd = fp_sub(fp_mul_nv(px, fp_from_stage<TH>(stage + 6 * TH)), fp_from_stage<TH>(stage));
    const Fp Rd = fp_sub(fp_mul_nv(py, fp_from_stage<TH>(stage + 9 * TH)), fp_from_stage<TH>(stage + 3 * TH));
    if (fp_is_zero(Pd)) {   // same x: doubling or cancellation (cold)
      G1Xyzz a4;
      a4.x = fp_from_stage<TH>(stage); a4.y = fp_from_stage<TH>(stage + 3 * TH); a4.zz = fp_from_stage<TH>(stage + 6 * TH); a4.zzz = fp_from_stage<TH>(stage + 9 * TH);
      G1Affine p;
      p.x = px; p.y = py;
      xyzz_madd_rare(a4, p);
      if (xyzz_is_inf(a4)) {
        acc_inf = true;
      } else {
        fp_to_stage<TH>(stage, a4.x); fp_to_stage<TH>(stage + 3 * TH, a4.y); fp_to_stage<TH>(stage + 6 * TH, a4.zz); fp_to_stage<TH>(stage + 9 * TH, a4.zzz);
      }
      continue;
    }
    const Fp PP = fp_sqr_nv(Pd), PPP = fp_mul_nv(Pd, PP);
    const Fp Q = fp_mul_nv(fp_from_stage<TH>(stage), PP);
    const Fp X3 = fp_sub(fp_sub(fp_sqr_nv(Rd), PPP), fp_dbl(Q));
    fp_to_stage<TH>(stage, X3);
    const Fp Y3 = fp_sub(fp_mul_nv(Rd, fp_sub(Q, X3)), fp_mul_nv(fp_from_stage<TH>(stage + 3 * TH), PPP));
    fp_to_stage<TH>(stage + 3 * TH, Y3);
    fp_to_stage<TH>(stage + 6 * TH, fp_mul_nv(fp_from_stage<TH>(stage + 6 * TH), PP));
    fp_to_stage<TH>(stage + 9 * TH, fp_mul_nv(fp_from_stage<TH>(stage + 9 * TH), PPP));
  }
  G1Xyzz acc = xyzz_inf();
  if (!acc_inf) {
    acc.x = fp_from_stage<TH>(stage); acc.y = fp_from_stage<TH>(stage + 3 * TH); acc.zz = fp_from_stage<TH>(stage + 6 * TH); acc.zzz = fp_from_stage<TH>(stage + 9 * TH);
  }
  __syncthreads();   // the reduction scratch aliases the staging columns
  if (seg_stride > 0) {
    // segmented form (FK20 cell proofs, cells.cu): thread hl of either half holds the sum over the points
    // pi = hl (mod HT), which the caller arranged to be ONE small MSM; the two GLV halves of segment hl meet
    // (psi on the q-half) and the block writes HT results instead of one: partials[hl * seg_stride + blob]
    if (half == 1) xyzz_to_smem(red, HT, hl, acc);
    __syncthreads();
    if (half == 0) {
      G1Xyzz o = xyzz_from_smem(red, HT, hl);
      xyzz_psi_ni(o);
      xyzz_add_ni(acc, o);
      partials[(size_t)hl * seg_stride + blob] = acc;
    }
    return;
  }
  block_reduce_xyzz_glv<TH>(acc, red);
  if (tid == 0) partials[blockIdx.x] = acc;
}

// Two ways to fix the register budget: MINB > 0 -> __launch_bounds__(TH, MINB) (ptxas picks the count),
// MINB < 0 -> __maxnreg__(-MINB) (the two qualifiers cannot be combined).
template <bool BE, int K, int MINB, int TH, bool PF>
__global__ void __launch_bounds__(TH, MINB)
msm_gather_ba_kernel(G1Xyzz* __restrict__ partials, const uint4* __restrict__ table, const uint8_t* __restrict__ scalars,
                     uint4* __restrict__ scratch, int c, int nwin, uint32_t cnt_top, int seg_stride, int split) {
  msm_gather_ba_body<BE, K, TH, PF>(partials, table, scalars, scratch, c, nwin, cnt_top, seg_stride, split);
}
template <bool BE, int K, int REGS, int TH, bool PF>
__global__ void __maxnreg__(REGS)
msm_gather_ba_kernel_r(G1Xyzz* __restrict__ partials, const uint4* __restrict__ table, const uint8_t* __restrict__ scalars,
                       uint4* __restrict__ scratch, int c, int nwin, uint32_t cnt_top, int seg_stride, int split) {
  msm_gather_ba_body<BE, K, TH, PF>(partials, table, scalars, scratch, c, nwin, cnt_top, seg_stride, split);
}

template <int K, int MINB, int TH>
static void launch_ba(void* d_partials, const void* d_table, int c, const void* d_scalars, bool be_input, int n_blobs,
                      void* d_scratch, cudaStream_t st, int seg_stride = 0, int split = 1) {
  const int nwin = glv_num_windows(c);
  const uint32_t cnt_top = glv_top_max(c) + 1u;
  constexpr size_t smem = ba_smem_bytes<K, TH>();
  static bool attr_set = false;
  void (*kbe)(G1Xyzz*, const uint4*, const uint8_t*, uint4*, int, int, uint32_t, int, int);
  void (*kle)(G1Xyzz*, const uint4*, const uint8_t*, uint4*, int, int, uint32_t, int, int);
  if constexpr (MINB > 0) {
    kbe = msm_gather_ba_kernel<true, K, MINB, TH, LWKZG_BA_PREFETCH>;
    kle = msm_gather_ba_kernel<false, K, MINB, TH, LWKZG_BA_PREFETCH>;
  } else {
    kbe = msm_gather_ba_kernel_r<true, K, -MINB, TH, LWKZG_BA_PREFETCH>;
    kle = msm_gather_ba_kernel_r<false, K, -MINB, TH, LWKZG_BA_PREFETCH>;
  }
  if (!attr_set) {
    cudaFuncSetAttribute(kbe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kle, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kbe, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(kle, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_set = true;
  }
  if (be_input)
    kbe<<<n_blobs * split, TH, smem, st>>>((G1Xyzz*)d_partials, (const uint4*)d_table, (const uint8_t*)d_scalars, (uint4*)d_scratch, c, nwin, cnt_top, seg_stride, split);
  else
    kle<<<n_blobs * split, TH, smem, st>>>((G1Xyzz*)d_partials, (const uint4*)d_table, (const uint8_t*)d_scalars, (uint4*)d_scratch, c, nwin, cnt_top, seg_stride, split);
}

}  // namespace lw
