// Batched fixed-base G1 MSM over the 4096-point SRS: the commitment / proof
// hot loop (replaces lambdaworks-math `msm::pippenger::msm` as called by
// KZG::commit / KZG::open -- /root/reference/src/lib.rs:241-243, 269-270, 329,
// 394).
//
// B200-first design: the SRS never changes, and the GPU has 180 GB of HBM, so
// instead of per-call bucket accumulation + bucket reduction we precompute the
// FULL signed-digit table
//        T[j][i][d-1] = d * 2^(c j) * P_i ,  d = 1 .. 2^(c-1)       (affine, 96 B)
// once at setup (c = 13: 20 x 4096 x 4096 entries = 32 GiB) and an MSM becomes
//        C = sum_{i,j} sign(d_ij) * T[j][i][|d_ij|-1]
// i.e. 4096 * ceil(256/c) mixed additions with NO buckets, NO sorting, NO
// reduction tree over buckets and perfectly uniform work per thread.  Each
// thread owns a slice of points, streams their scalars with 128-bit loads,
// recodes them to signed digits in registers, gathers the table entries
// (3 x 32 B sectors each, random in HBM/L2) and accumulates in XYZZ
// coordinates; a block-level tree in shared memory and a tiny finalize kernel
// combine the per-thread sums.  The kernel is bound by the integer (IMAD)
// pipe: ~10 Fp multiplications per gathered entry against 96 B of traffic.
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
#ifndef LWKZG_MSM_BA_K
#define LWKZG_MSM_BA_K 16
#endif
constexpr int MSM_THREADS = LWKZG_MSM_THREADS;

int msm_threads_per_block() { return MSM_THREADS; }

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

// Block-wide sum of per-thread XYZZ accumulators; result valid in thread 0.  THREADS need not be a
// power of two: the tree starts at the next power of two and skips partners beyond the block.
__host__ __device__ constexpr int reduce_half(int threads) {
  int s = 1;
  while (2 * s < threads) s *= 2;
  return s;
}
template <int THREADS>
__device__ __forceinline__ void block_reduce_xyzz(G1Xyzz& acc, uint32_t* red /* 48 * reduce_half(THREADS) words */) {
  constexpr int H = reduce_half(THREADS);
  const int tid = threadIdx.x;
  for (int s = H; s > 0; s >>= 1) {
    if (tid >= s && tid < 2 * s && tid < THREADS) xyzz_to_smem(red, H, tid - s, acc);
    __syncthreads();
    if (tid < s && tid + s < THREADS) {
      G1Xyzz o = xyzz_from_smem(red, H, tid);
      xyzz_add_ni(acc, o);
    }
    __syncthreads();
  }
}

// digit j of the scalar held in shared memory (word w of thread t at sk[w][t]);
// c is a run-time value here, so limb indices are dynamic -> shared memory.
__device__ __forceinline__ int smem_digit(const uint32_t (*sk)[MSM_THREADS], int tid, int c, int j, int& carry) {
  const int bit = j * c;
  const int w = bit >> 5, s = bit & 31;
  const uint32_t lo = sk[w][tid];
  const uint32_t hi = (w + 1 < 8) ? sk[w + 1][tid] : 0u;
  uint32_t raw = __funnelshift_r(lo, hi, s) & ((1u << c) - 1u);
  int d = (int)raw + carry;
  if (d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else { carry = 0; }
  return d;
}

template <int TH>
__device__ __forceinline__ int smem_digit_t(const uint32_t (*sk)[TH], int tid, int c, int j, int& carry) {
  const int bit = j * c;
  const int w = bit >> 5, s = bit & 31;
  const uint32_t lo = sk[w][tid];
  const uint32_t hi = (w + 1 < 8) ? sk[w + 1][tid] : 0u;
  uint32_t raw = __funnelshift_r(lo, hi, s) & ((1u << c) - 1u);
  int d = (int)raw + carry;
  if (d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else { carry = 0; }
  return d;
}

template <bool BE>
__global__ void __launch_bounds__(MSM_THREADS, LWKZG_MSM_MIN_BLOCKS)
msm_gather_kernel(G1Xyzz* __restrict__ partials, const uint4* __restrict__ table, const uint8_t* __restrict__ scalars,
                  int c, int pt_threads, int wsplit) {
  __shared__ uint32_t sk[8][MSM_THREADS];
  __shared__ uint32_t red[48 * reduce_half(MSM_THREADS)];

  const int W = 255 / c + 1;
  const int tid = threadIdx.x;
  const int blob = blockIdx.y;
  const int q = blockIdx.x * MSM_THREADS + tid;
  const int pl = q % pt_threads;   // point lane
  const int wg = q / pt_threads;   // window group (only > 0 for tiny batches)
  const uint8_t* sc = scalars + (size_t)blob * BLOB_BYTES;

  G1Xyzz acc = xyzz_inf();

  for (int pi = pl; pi < N_POINTS; pi += pt_threads) {
    // ---- scalar: two coalesced 128-bit loads (a warp reads 1 KiB contiguous)
    const uint4* sp = reinterpret_cast<const uint4*>(sc + (size_t)pi * 32);
    uint4 a = __ldg(sp), b = __ldg(sp + 1);
    Fr k;
    if (BE) {
      uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      k = fr_canon_from_be_words(w);   // bswap + reduce mod r (App. A.1)
    } else {
      k.l[0] = a.x; k.l[1] = a.y; k.l[2] = a.z; k.l[3] = a.w;
      k.l[4] = b.x; k.l[5] = b.y; k.l[6] = b.z; k.l[7] = b.w;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) sk[i][tid] = k.l[i];   // private column: no sync needed
    // ---- recode on the fly, gather, accumulate; the entry for window j+1 is
    // requested before the mixed addition for window j is issued.
    const size_t pbase = (size_t)pi << (c - 1);
    int carry = 0;
    int d = 0;
    G1Affine cur = g1a_inf();
    for (int j = 0; j <= W; j++) {
      int dn = 0;
      G1Affine nxt = g1a_inf();
      if (j < W) {
        dn = smem_digit(sk, tid, c, j, carry);
        if ((j % wsplit) != wg) dn = 0;
        if (dn != 0) nxt = load_entry(table, (((size_t)j * N_POINTS) << (c - 1)) + pbase + (size_t)((dn < 0 ? -dn : dn) - 1));
      }
      if (d != 0) {
        cur.y = fp_cneg(cur.y, d < 0);
        xyzz_madd_hot(acc, cur);
      }
      cur = nxt; d = dn;
    }
  }

  block_reduce_xyzz<MSM_THREADS>(acc, red);
  if (tid == 0) partials[(size_t)blob * gridDim.x + blockIdx.x] = acc;
}

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
                     uint4* __restrict__ scratch, int c, int pt_threads) {
  __shared__ uint32_t sk[8][TH];
  __shared__ uint32_t sidx[K][TH];   // entry index | sign << 31, or BA_NONE
  __shared__ uint32_t red[48 * reduce_half(TH)];

  const int W = 255 / c + 1;
  const int tid = threadIdx.x;
  const int blob = blockIdx.y;
  const int pl = blockIdx.x * TH + tid;   // point lane, < pt_threads <= N_POINTS
  const uint8_t* sc = scalars + (size_t)blob * BLOB_BYTES;
  // slot k: words 0-2 = A.x, 3-5 = A.y, 6-8 = exclusive prefix product
  uint4* my = scratch + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * (K * 9 * TH) + tid;

  int pi = pl, j = W, carry = 0;        // j == W: the next scalar has to be fetched
  size_t pbase = 0;
  bool ended = pi >= N_POINTS;
  uint64_t infmask = K >= 64 ? ~0ull : ((1ull << K) - 1ull);   // accumulators at infinity

  while (!ended) {
    // ------------------------------------------------------------ pass 1a: the next K non-zero digits
    // of this thread's (point, window) stream -> entry indices; their table lines and the accumulator
    // rows are requested into L2 so that pass 1b finds them there
#pragma unroll 1
    for (int k = 0; k < K; k++) {
      uint32_t e = BA_NONE;
      while (!ended) {
        if (j == W) {
          const uint4* sp = reinterpret_cast<const uint4*>(sc + (size_t)pi * 32);
          uint4 a = __ldg(sp), b = __ldg(sp + 1);
          Fr kk;
          if (BE) {
            uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            kk = fr_canon_from_be_words(w);
          } else {
            kk.l[0] = a.x; kk.l[1] = a.y; kk.l[2] = a.z; kk.l[3] = a.w;
            kk.l[4] = b.x; kk.l[5] = b.y; kk.l[6] = b.z; kk.l[7] = b.w;
          }
#pragma unroll
          for (int i = 0; i < 8; i++) sk[i][tid] = kk.l[i];
          j = 0; carry = 0;
          pbase = (size_t)pi << (c - 1);
        }
        const int d = smem_digit_t<TH>(sk, tid, c, j, carry);
        const size_t wbase = ((size_t)j * N_POINTS) << (c - 1);
        j++;
        if (j == W) { pi += pt_threads; if (pi >= N_POINTS) ended = true; }
        if (d != 0) {
          e = (uint32_t)(wbase + pbase + (size_t)((d < 0 ? -d : d) - 1)) | (d < 0 ? 0x80000000u : 0u);
          break;
        }
      }
      sidx[k][tid] = e;
      if (e != BA_NONE) {
        prefetch_l2(table + (size_t)(e & 0x7fffffffu) * 6);
        if (!((infmask >> k) & 1ull)) prefetch_fp_scratch<TH>(my + (k * 9) * TH);
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
  block_reduce_xyzz<TH>(acc, red);
  if (tid == 0) partials[(size_t)blob * gridDim.x + blockIdx.x] = acc;
}

// variant = accumulators per thread (K), threads per blob (= block size) and register budget
struct BaVariant { int k, threads; };
static const BaVariant BA_VARIANTS[] = {{64, 128}, {32, 128}, {64, 64}, {32, 64}, {64, 32}, {16, 128}, {64, 96}, {48, 96}};
static int g_ba_variant = 0;
void msm_ba_set_variant(int v) { if (v >= 0 && v < (int)(sizeof(BA_VARIANTS) / sizeof(BA_VARIANTS[0]))) g_ba_variant = v; }
int msm_ba_threads() { return BA_VARIANTS[g_ba_variant].threads; }
int msm_ba_slots() { return BA_VARIANTS[g_ba_variant].k; }
size_t msm_ba_scratch_bytes(int n_blobs) {   // sized for the largest variant: 64 slots x 9 words x 128 threads
  return (size_t)n_blobs * 64 * 9 * 128 * sizeof(uint4);
}

template <int K, int MINB, int TH>
static void launch_ba(void* d_partials, const void* d_table, int c, const void* d_scalars, bool be_input, int n_blobs,
                      void* d_scratch, cudaStream_t st) {
  dim3 grid(1, n_blobs);
  if (be_input)
    msm_gather_ba_kernel<true, K, MINB, TH><<<grid, TH, 0, st>>>((G1Xyzz*)d_partials, (const uint4*)d_table, (const uint8_t*)d_scalars, (uint4*)d_scratch, c, TH);
  else
    msm_gather_ba_kernel<false, K, MINB, TH><<<grid, TH, 0, st>>>((G1Xyzz*)d_partials, (const uint4*)d_table, (const uint8_t*)d_scalars, (uint4*)d_scratch, c, TH);
}

// one block per blob; partials: one XYZZ per blob
void launch_msm_gather_ba(void* d_partials, const void* d_table, int c, const void* d_scalars, bool be_input, int n_blobs,
                          void* d_scratch, cudaStream_t st) {
  if (n_blobs <= 0) return;
  switch (g_ba_variant) {
    case 1: launch_ba<32, 3, 128>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st); break;
    case 2: launch_ba<64, 6, 64>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st); break;
    case 3: launch_ba<32, 6, 64>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st); break;
    case 4: launch_ba<64, 12, 32>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st); break;
    case 5: launch_ba<16, 3, 128>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st); break;
    case 6: launch_ba<64, 4, 96>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st); break;
    case 7: launch_ba<48, 4, 96>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st); break;
    default: launch_ba<64, 3, 128>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st); break;
  }
  count_launch();
}

// Sum the per-block partials of one blob, normalise (one inversion), compress.
__global__ void __launch_bounds__(32)
msm_finalize_kernel(uint8_t* __restrict__ out48, G1Affine* __restrict__ aff_out, const G1Xyzz* __restrict__ partials,
                    int parts, int n) {
  const int blob = blockIdx.x * blockDim.x + threadIdx.x;
  if (blob >= n) return;
  const G1Xyzz* p = partials + (size_t)blob * parts;
  G1Xyzz acc = p[0];
  for (int i = 1; i < parts; i++) {
    G1Xyzz o = p[i];
    xyzz_add_ni(acc, o);
  }
  G1Affine a = xyzz_to_affine(acc);
  if (aff_out) aff_out[blob] = a;
  if (out48) g1_compress(out48 + (size_t)blob * 48, a);
}

void launch_msm_gather(void* d_partials, const void* d_table, int c, const void* d_scalars, bool be_input, int n_blobs,
                       int blocks_per_blob, cudaStream_t st) {
  if (n_blobs <= 0) return;
  int total = blocks_per_blob * MSM_THREADS;
  int pt = total < N_POINTS ? total : N_POINTS;
  int ws = total / pt;
  dim3 grid(blocks_per_blob, n_blobs);
  if (be_input)
    msm_gather_kernel<true><<<grid, MSM_THREADS, 0, st>>>((G1Xyzz*)d_partials, (const uint4*)d_table, (const uint8_t*)d_scalars, c, pt, ws);
  else
    msm_gather_kernel<false><<<grid, MSM_THREADS, 0, st>>>((G1Xyzz*)d_partials, (const uint4*)d_table, (const uint8_t*)d_scalars, c, pt, ws);
  count_launch();
}

// Warp-per-blob variant for many partials (single-blob / tiny-batch calls use up to
// 128 blocks per blob): lanes sum strided partials, then a 5-level tree in shared memory.
__global__ void __launch_bounds__(32) msm_finalize_warp_kernel(uint8_t* __restrict__ out48, G1Affine* __restrict__ aff_out,
                                                                const G1Xyzz* __restrict__ partials, int parts, int n) {
  __shared__ uint32_t red[48 * 16];
  const int blob = blockIdx.x, lane = threadIdx.x;
  const G1Xyzz* p = partials + (size_t)blob * parts;
  G1Xyzz acc = xyzz_inf();
  for (int i = lane; i < parts; i += 32) {
    G1Xyzz o = p[i];
    xyzz_add_ni(acc, o);
  }
  for (int s = 16; s > 0; s >>= 1) {
    if (lane >= s && lane < 2 * s) xyzz_to_smem(red, 16, lane - s, acc);
    __syncwarp();
    if (lane < s) {
      G1Xyzz o = xyzz_from_smem(red, 16, lane);
      xyzz_add_ni(acc, o);
    }
    __syncwarp();
  }
  if (lane == 0) {
    G1Affine a = xyzz_to_affine(acc);
    if (aff_out) aff_out[blob] = a;
    if (out48) g1_compress(out48 + (size_t)blob * 48, a);
  }
}

void launch_msm_finalize(void* d_out48, void* d_aff_out, const void* d_partials, int parts_per_blob, int n_blobs, cudaStream_t st) {
  if (n_blobs <= 0) return;
  if (parts_per_blob >= 8)
    msm_finalize_warp_kernel<<<n_blobs, 32, 0, st>>>((uint8_t*)d_out48, (G1Affine*)d_aff_out, (const G1Xyzz*)d_partials, parts_per_blob, n_blobs);
  else
    msm_finalize_kernel<<<(n_blobs + 31) / 32, 32, 0, st>>>((uint8_t*)d_out48, (G1Affine*)d_aff_out, (const G1Xyzz*)d_partials, parts_per_blob, n_blobs);
  count_launch();
}

}  // namespace lw
