// Variable-base G1 multi-scalar multiplication (the reference's g1_lincomb =
// lambdaworks-math msm::pippenger::msm, /root/reference/src/lib.rs:241-243, and the
// three `msm` calls of verify_kzg_proof_batch, lib.rs:679-685; BASELINE config 5:
// N = 2^12 .. 2^22).
//
// Bucket method laid out for the GPU, no comparison sort, one pipeline for every size:
//   1. prep: points -> Montgomery affine (on-curve check when they come from the host);
//      scalars reduced mod r.  When the caller vouches that the points are in G1 every
//      scalar is GLV-split into two 128-bit halves, k = m + q x^2 with psi(P) = (beta x,
//      -y) = [x^2]P, so the job becomes 2n "items" with 128-bit scalars: half the windows
//      to reduce and half the doubling chain at the end -- what bounds the latency of a
//      small MSM -- for the same number of bucket additions.  Arbitrary curve points
//      (lwkzg_g1_lincomb's default) keep whole 255-bit scalars.  Signed c-bit digits,
//      unsigned top window; per (window, bucket) histogram with atomics.
//   2. exclusive scan per window, scatter of item indices into bucket order (the order
//      inside a bucket is irrelevant: group addition is exact)
//   3. accumulate: S lanes per (window, bucket), S a power of two chosen so that the
//      launch is about one warp per SM sub-partition -- a small MSM has few buckets and
//      a thread per bucket would leave a long serial chain on an empty GPU; the S partial
//      sums are folded with warp shuffles.  Buckets with very long runs (skewed digits)
//      go to a block-per-bucket kernel.
//   4. per window sum_b b * B_b: threads own short runs of buckets (running sums + one
//      small double-and-add), block tree, several blocks per window when there are many
//      buckets
//   5. chain: one warp per window multiplies its window sum by 2^(c j) -- a Jacobian
//      doubling whose independent field products are spread over three lanes (three
//      products deep instead of seven) -- and the last warp to finish adds the W
//      results, normalises and writes the compressed point (or affine coordinates for
//      the pairing).
// The signed digit set, the split and the window size are internal choices; the result is
// the same group element as the reference's unsigned-window Pippenger.
#include "g1.cuh"
#include "kernels.h"
#include "recode.cuh"

#include <algorithm>

namespace lw {

constexpr int VM_RED_THREADS = 128;
constexpr int VM_RED_PER = 8;           // buckets per thread of the window reduction (at most)

// measured on B200 (tools/vm_tune.py, profiles/r02_vm_tune.log): with the split, windows that tile the 128 bits
// exactly (16 x 8, 8 x 16) win at both ends; in between 12 and 15 bits
__host__ __device__ inline int vm_window_bits_for(size_t n, bool glv) {
  int lg = 0;
  while ((size_t(1) << (lg + 1)) <= n) lg++;
  int c;
  if (glv) {
    if (lg <= 8) c = 6;
    else if (lg <= 15) c = 8;
    else if (lg <= 16) c = 12;
    else if (lg <= 18) c = 15;
    else c = 16;
  } else {
    c = lg - 3;
    if (c > 16) c = 16;
  }
  if (c < 4) c = 4;
  return c;
}
// tuning hooks, read on every call (tools/vm_tune.py flips them between launches): LWKZG_VM_C = window bits,
// LWKZG_VM_GLV = 0 / 1, LWKZG_VM_S = lanes per bucket
static int g_vm_force_c = 0, g_vm_force_glv = -1, g_vm_force_s = 0;
static void vm_env() {
  const char* e = getenv("LWKZG_VM_C");
  g_vm_force_c = e ? atoi(e) : 0;
  e = getenv("LWKZG_VM_GLV");
  g_vm_force_glv = e ? atoi(e) : -1;
  e = getenv("LWKZG_VM_S");
  g_vm_force_s = e ? atoi(e) : 0;
}

struct VmLayout {
  size_t pts, halves, counts, offsets, cursors, heavy, done, bad, idx, buckets, partials, chain, total;
  size_t zero_begin, zero_end;   // one memset: counts, cursors, heavy count, done ticket, bad flag
  size_t items;
  bool glv;
  int c, W, B, NL;  // B = 2^(c-1) buckets per signed window (bucket b holds digit magnitude b+1), 2B for the unsigned
                    // top window (bucket slots are j * B + b, so the top window simply extends past W * B);
                    // NL = limbs per scalar
  int S, rblocks, per;
};
// in_g1: the caller vouches that every point lies in the r-torsion (decoded with a subgroup check, or built from
// SRS points) -- psi(P) = [x^2]P only holds there, so arbitrary curve points never take the split
static VmLayout vm_layout(size_t n, bool in_g1) {
  vm_env();
  VmLayout L;
  const size_t nn = n ? n : 1;
  L.glv = in_g1 && (g_vm_force_glv >= 0 ? g_vm_force_glv != 0 : true);
  L.c = vm_window_bits_for(nn, L.glv);
  if (g_vm_force_c >= 4 && g_vm_force_c <= 16) L.c = g_vm_force_c;
  L.items = L.glv ? 2 * nn : nn;
  L.NL = L.glv ? 4 : 8;
  // windows 0 .. W-2 carry signed digits in [-2^(c-1)+1, 2^(c-1)]; the top window is unsigned (raw bits + the last
  // carry <= 2^c), so no carry leaves it and no window is spent on a carry alone
  const int bits = L.glv ? 128 : 255;
  L.W = (bits + L.c - 1) / L.c;
  L.B = 1 << (L.c - 1);
  // lanes per bucket: at most ~6 warps per SM sub-partition (148 x 4), at least ~4 entries per lane
  const size_t nb = (size_t)(L.W + 1) * L.B;
  const size_t avg = L.items / (size_t)L.B + 1;
  int S = 1;
  while (S < 32 && nb * (size_t)(2 * S) <= (size_t)592 * 32 * 6 && (size_t)(2 * S) * 4 <= avg) S <<= 1;
  if (g_vm_force_s >= 1 && g_vm_force_s <= 32 && !(g_vm_force_s & (g_vm_force_s - 1))) S = g_vm_force_s;
  L.S = S;
  L.rblocks = (2 * L.B + VM_RED_THREADS * VM_RED_PER - 1) / (VM_RED_THREADS * VM_RED_PER);   // sized for the top window
  L.per = (2 * L.B + VM_RED_THREADS * L.rblocks - 1) / (VM_RED_THREADS * L.rblocks);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
  L.pts = take(L.items * 96);
  L.halves = take(L.items * (size_t)L.NL * 4);
  L.zero_begin = off;
  L.counts = take(nb * 4);
  L.cursors = take(nb * 4);
  L.heavy = take((nb + 1) * 4);  // [0] = count, then bucket ids whose run is too long for one lane group
  L.done = take(4);
  L.bad = take(4);
  L.zero_end = off;
  L.offsets = take(nb * 4);
  L.idx = take(L.items * (size_t)L.W * 4);
  L.buckets = take(nb * sizeof(G1Xyzz));
  L.partials = take((size_t)L.W * L.rblocks * sizeof(G1Xyzz));
  L.chain = take((size_t)L.W * sizeof(G1Xyzz));
  L.total = off;
  return L;
}
size_t var_msm_scratch_bytes(size_t n) {   // sized for either plan
  const size_t a = vm_layout(n, false).total, b = vm_layout(n, true).total;
  return a > b ? a : b;
}

// signed digit j of an NL-limb little-endian scalar; carry threaded from digit j - 1
__device__ __forceinline__ int vm_digit(const uint32_t* k, int nl, int c, int W, int j, int& carry) {
  const int bit = j * c;
  const int w = bit >> 5, s = bit & 31;
  const uint32_t lo = w < nl ? k[w] : 0u;
  const uint32_t hi = (w + 1 < nl) ? k[w + 1] : 0u;
  const uint32_t raw = __funnelshift_r(lo, hi, s) & ((1u << c) - 1u);
  int d = (int)raw + carry;
  if (j < W - 1 && d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else { carry = 0; }
  return d;
}

// ---- 1. prep.  FROM_BE: points canonical big-endian affine x || y (all-zero = infinity), scalars 32 big-endian
// bytes (reduced mod r like the reference's from_bytes_be); else points Montgomery affine, scalars canonical limbs.
template <bool FROM_BE>
__global__ void __launch_bounds__(128) vm_prep_kernel(G1Affine* __restrict__ pts, uint32_t* __restrict__ halves, uint32_t* __restrict__ counts,
                                                      int* __restrict__ bad, const uint8_t* __restrict__ pts_in, const uint8_t* __restrict__ sc_in,
                                                      size_t n, int glv, int c, int W, int B) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = g1a_inf();
  Fr k;
  if (FROM_BE) {
    const uint8_t* b = pts_in + i * 96;
    bool z = true;
    for (int t = 0; t < 96; t++) z = z && (b[t] == 0);
    if (!z) {
      p.x = fp_from_be48(b);
      p.y = fp_from_be48(b + 48);
      if (!g1a_on_curve(p)) atomicExch(bad, 1);
    }
    k = fr_canon_from_be32(sc_in + i * 32);
  } else {
    p = reinterpret_cast<const G1Affine*>(pts_in)[i];
    const uint32_t* s = reinterpret_cast<const uint32_t*>(sc_in) + i * 8;
    for (int t = 0; t < 8; t++) k.l[t] = s[t];
  }
  const bool inf = g1a_is_inf(p);
  if (inf) for (int t = 0; t < 8; t++) k.l[t] = 0;   // contributes nothing: no digits, no bucket entries
  pts[i] = p;
  const int NL = glv ? 4 : 8;
  uint32_t h[2][8];
  if (glv) {
    glv_split_barrett(h[1], h[0], k.l);   // k = h0 + h1 x^2
    G1Affine p2 = p;
    if (!inf) {
      Fp beta;
      for (int t = 0; t < 12; t++) beta.l[t] = k::FP_BETA[t];
      p2.x = fp_mul(p.x, beta);
      p2.y = fp_neg(p.y);
    }
    pts[n + i] = p2;
  } else {
    for (int t = 0; t < 8; t++) h[0][t] = k.l[t];
  }
  for (int half = 0; half < (glv ? 2 : 1); half++) {
    uint32_t* dst = halves + ((size_t)half * n + i) * NL;
    for (int t = 0; t < NL; t++) dst[t] = h[half][t];
    int carry = 0;
    for (int j = 0; j < W; j++) {
      const int d = vm_digit(h[half], NL, c, W, j, carry);
      if (d != 0) atomicAdd(&counts[(size_t)j * B + (uint32_t)((d < 0 ? -d : d) - 1)], 1u);
    }
  }
}

// ---- 2. exclusive scan of one window's B counts (one block per window), then the scatter
__global__ void __launch_bounds__(1024) vm_scan_kernel(uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts, int B0) {
  __shared__ uint32_t part[1024];
  const int j = blockIdx.x, t = threadIdx.x;
  const int B = j == gridDim.x - 1 ? 2 * B0 : B0;   // the unsigned top window
  const int per = (B + 1023) / 1024;
  const uint32_t* c = counts + (size_t)j * B0;
  uint32_t* o = offsets + (size_t)j * B0;
  uint32_t s = 0;
  for (int k = 0; k < per; k++) {
    int b = t * per + k;
    if (b < B) s += c[b];
  }
  part[t] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    uint32_t v = (t >= d) ? part[t - d] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = part[t] - s;
  for (int k = 0; k < per; k++) {
    int b = t * per + k;
    if (b < B) { o[b] = run; run += c[b]; }
  }
}

__global__ void __launch_bounds__(128) vm_scatter_kernel(uint32_t* __restrict__ cursors, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ idx,
                                                         const uint32_t* __restrict__ halves, size_t items, int NL, int c, int W, int B) {
  const size_t h = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= items) return;
  uint32_t k[8];
  for (int t = 0; t < NL; t++) k[t] = halves[h * NL + t];
  int carry = 0;
  for (int j = 0; j < W; j++) {
    const int d = vm_digit(k, NL, c, W, j, carry);
    if (d == 0) continue;
    const size_t slot = (size_t)j * B + (uint32_t)((d < 0 ? -d : d) - 1);
    const uint32_t pos = atomicAdd(&cursors[slot], 1u);
    idx[(size_t)j * items + offsets[slot] + pos] = (uint32_t)h | (d < 0 ? 0x80000000u : 0u);
  }
}

// Compact group law for the latency-bound stages (a warp or two per SM): every field product is a CALL to the
// one out-of-line multiplier, so an addition is ~1 KB of code instead of ~70 KB with the products inlined -- a
// lone warp walking through straight-line code of that size runs at instruction-fetch speed.
__device__ __noinline__ void vm_add_c(G1Xyzz& a, const G1Xyzz& b) {
  if (!xyzz_is_inf(b)) {
    if (xyzz_is_inf(a)) {
      a = b;
    } else {
      const Fp U1 = fp_mul_nv(a.x, b.zz), U2 = fp_mul_nv(b.x, a.zz);
      const Fp S1 = fp_mul_nv(a.y, b.zzz), S2 = fp_mul_nv(b.y, a.zzz);
      const Fp Pd = fp_sub(U2, U1), Rd = fp_sub(S2, S1);
      if (fp_is_zero(Pd)) {
        xyzz_add_ni(a, b);   // doubling / cancellation: generic formulas
      } else {
        const Fp PP = fp_sqr_nv(Pd), PPP = fp_mul_nv(Pd, PP), Q = fp_mul_nv(U1, PP);
        const Fp X3 = fp_sub(fp_sub(fp_sqr_nv(Rd), PPP), fp_dbl(Q));
        const Fp Y3 = fp_sub(fp_mul_nv(Rd, fp_sub(Q, X3)), fp_mul_nv(S1, PPP));
        a.zz = fp_mul_nv(fp_mul_nv(a.zz, b.zz), PP);
        a.zzz = fp_mul_nv(fp_mul_nv(a.zzz, b.zzz), PPP);
        a.x = X3;
        a.y = Y3;
      }
    }
  }
}
__device__ __noinline__ void vm_dbl_c(G1Xyzz& p) {
  if (!xyzz_is_inf(p)) {
    const Fp U = fp_dbl(p.y), V = fp_sqr_nv(U), W = fp_mul_nv(U, V), S = fp_mul_nv(p.x, V), X2 = fp_sqr_nv(p.x);
    const Fp M = fp_add(fp_dbl(X2), X2);
    const Fp X3 = fp_sub(fp_sqr_nv(M), fp_dbl(S));
    p.y = fp_sub(fp_mul_nv(M, fp_sub(S, X3)), fp_mul_nv(W, p.y));
    p.x = X3;
    p.zz = fp_mul_nv(V, p.zz);
    p.zzz = fp_mul_nv(W, p.zzz);
  }
}

// ---- 3. accumulate
__device__ __forceinline__ G1Xyzz vm_shfl_down(const G1Xyzz& p, int d) {
  G1Xyzz r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&p);
  uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 48; i++) o[i] = __shfl_down_sync(0xffffffffu, s[i], d);
  return r;
}

// S lanes per (window, bucket).  Runs much longer than the average (skewed digit distributions: the top window
// only holds the last carry, equal scalars, ...) are deferred to vm_heavy_kernel, which puts a whole block on each
// of them: one such bucket on one lane group would be the critical path of a small MSM.
constexpr int VM_HEAVY_THREADS = 128;
__global__ void __launch_bounds__(128, 3) vm_accumulate_kernel(G1Xyzz* __restrict__ buckets, uint32_t* __restrict__ heavy,
                                                               const uint32_t* __restrict__ idx, const uint32_t* __restrict__ offsets,
                                                               const uint32_t* __restrict__ counts, const G1Affine* __restrict__ pts, size_t items,
                                                               size_t nbuckets, int B, int W, int S, uint32_t heavy_cap) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t g = t / S;
  const int s = (int)(t % S);
  const bool live = g < nbuckets;   // every lane stays for the shuffles
  G1Xyzz acc = xyzz_inf();
  bool is_heavy = false;
  if (live) {
    const int j = min((int)(g / B), W - 1);
    const uint32_t* run = idx + (size_t)j * items + offsets[g];
    const uint32_t cnt = counts[g];
    if (cnt > heavy_cap) {
      is_heavy = true;
      if (s == 0) {
        uint32_t slot = atomicAdd(&heavy[0], 1u);
        heavy[1 + slot] = (uint32_t)g;
      }
    } else {
      for (uint32_t k = s; k < cnt; k += S) {
        uint32_t e = run[k];
        G1Affine p = pts[e & 0x7fffffffu];
        p.y = fp_cneg(p.y, (e >> 31) != 0);
        xyzz_madd_hot(acc, p);
      }
    }
  }
  for (int d = S >> 1; d > 0; d >>= 1) {
    G1Xyzz o = vm_shfl_down(acc, d);
    if (s < d) vm_add_c(acc, o);
  }
  if (live && s == 0 && !is_heavy) buckets[g] = acc;
}

__device__ __forceinline__ void vm_to_smem(uint32_t* smem, int stride, int t, const G1Xyzz& p) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 48; i++) smem[i * stride + t] = w[i];
}
__device__ __forceinline__ G1Xyzz vm_from_smem(const uint32_t* smem, int stride, int t) {
  G1Xyzz p;
  uint32_t* w = reinterpret_cast<uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 48; i++) w[i] = smem[i * stride + t];
  return p;
}

// one block per heavy bucket (grid-stride over the heavy list)
__global__ void __launch_bounds__(VM_HEAVY_THREADS, 3) vm_heavy_kernel(G1Xyzz* __restrict__ buckets, const uint32_t* __restrict__ heavy,
                                                                       const uint32_t* __restrict__ idx, const uint32_t* __restrict__ offsets,
                                                                       const uint32_t* __restrict__ counts, const G1Affine* __restrict__ pts,
                                                                       size_t items, int B, int W) {
  __shared__ uint32_t red[48 * (VM_HEAVY_THREADS / 2)];
  const uint32_t nheavy = heavy[0];
  const int tid = threadIdx.x;
  for (uint32_t h = blockIdx.x; h < nheavy; h += gridDim.x) {
    const uint32_t t = heavy[1 + h];
    const int j = min((int)(t / (uint32_t)B), W - 1);
    const uint32_t* run = idx + (size_t)j * items + offsets[t];
    const uint32_t cnt = counts[t];
    G1Xyzz acc = xyzz_inf();
    for (uint32_t k = tid; k < cnt; k += VM_HEAVY_THREADS) {
      uint32_t e = run[k];
      G1Affine p = pts[e & 0x7fffffffu];
      p.y = fp_cneg(p.y, (e >> 31) != 0);
      xyzz_madd_hot(acc, p);
    }
    for (int s = VM_HEAVY_THREADS / 2; s > 0; s >>= 1) {
      if (tid >= s && tid < 2 * s) vm_to_smem(red, VM_HEAVY_THREADS / 2, tid - s, acc);
      __syncthreads();
      if (tid < s) {
        G1Xyzz o = vm_from_smem(red, VM_HEAVY_THREADS / 2, tid);
        vm_add_c(acc, o);
      }
      __syncthreads();
    }
    if (tid == 0) buckets[t] = acc;
  }
}

// ---- 4. window sums: partial[j][blk] = sum over the block's buckets of (b + 1) * bucket[j][b]
__global__ void __launch_bounds__(VM_RED_THREADS) vm_reduce_kernel(G1Xyzz* __restrict__ partials, const G1Xyzz* __restrict__ buckets, int B0, int per) {
  __shared__ uint32_t red[48 * (VM_RED_THREADS / 2)];
  const int j = blockIdx.y, t = threadIdx.x;
  const int B = j == gridDim.y - 1 ? 2 * B0 : B0;   // the unsigned top window
  const G1Xyzz* bk = buckets + (size_t)j * B0;
  per = (B + VM_RED_THREADS * (int)gridDim.x - 1) / (VM_RED_THREADS * (int)gridDim.x);   // buckets per thread in THIS window
  const int lo = min(B, (blockIdx.x * VM_RED_THREADS + t) * per), hi = min(B, lo + per);  // buckets [lo, hi): weights lo+1 .. hi
  G1Xyzz run = xyzz_inf(), acc = xyzz_inf();
  for (int b = hi - 1; b >= lo; b--) {
    G1Xyzz v = bk[b];
    vm_add_c(run, v);
    vm_add_c(acc, run);  // acc = sum (b - lo + 1) * B_b
  }
  // + lo * run   (double-and-add from the top set bit of lo)
  if (lo > 0 && lo < hi && !xyzz_is_inf(run)) {
    G1Xyzz m = run;
    for (int bit = 30 - __clz(lo); bit >= 0; bit--) {
      vm_dbl_c(m);
      if ((lo >> bit) & 1) vm_add_c(m, run);
    }
    vm_add_c(acc, m);
  }
  for (int s = VM_RED_THREADS / 2; s > 0; s >>= 1) {
    if (t >= s && t < 2 * s) vm_to_smem(red, VM_RED_THREADS / 2, t - s, acc);
    __syncthreads();
    if (t < s) {
      G1Xyzz o = vm_from_smem(red, VM_RED_THREADS / 2, t);
      vm_add_c(acc, o);
    }
    __syncthreads();
  }
  if (t == 0) partials[(size_t)j * gridDim.x + blockIdx.x] = acc;
}

// ---- 5. the doubling chains and the final sum
__device__ __forceinline__ Fp vm_bcast(const Fp& v, int src) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
  return r;
}
// One Jacobian doubling (a = 0; A = X^2, B = Y^2, C = B^2, D = 2((X+B)^2 - A - C), E = 3A, F = E^2, X3 = F - 2D,
// Y3 = E(D - X3) - 8C, Z3 = 2YZ) executed by a whole warp that holds the same point in every lane: the three
// independent products of the first two levels go to the lanes with role 0, 1, 2 and are broadcast, the last
// product is computed by everybody -- three multiplications deep instead of seven.  Z = 0 (infinity) stays 0.
__device__ __forceinline__ void vm_jac_dbl_coop(Fp& X, Fp& Y, Fp& Z, int role) {
  const Fp a1 = role == 0 ? X : Y;
  const Fp b1 = role == 0 ? X : (role == 1 ? Y : Z);
  const Fp p1 = fp_mul_nv(a1, b1);
  const Fp A = vm_bcast(p1, 0), Bq = vm_bcast(p1, 1), YZ = vm_bcast(p1, 2);
  const Fp E = fp_add(fp_dbl(A), A);
  const Fp a2 = role == 0 ? Bq : (role == 1 ? fp_add(X, Bq) : E);
  const Fp p2 = fp_sqr_nv(a2);
  const Fp C = vm_bcast(p2, 0), T = vm_bcast(p2, 1), F = vm_bcast(p2, 2);
  const Fp D = fp_dbl(fp_sub(fp_sub(T, A), C));
  const Fp X3 = fp_sub(F, fp_dbl(D));
  const Fp C8 = fp_dbl(fp_dbl(fp_dbl(C)));
  Y = fp_sub(fp_mul_nv(E, fp_sub(D, X3)), C8);
  X = X3;
  Z = fp_dbl(YZ);
}

__device__ __forceinline__ G1Xyzz vm_shfl_xyzz(const G1Xyzz& p, int src) {
  G1Xyzz r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&p);
  uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 48; i++) o[i] = __shfl_sync(0xffffffffu, s[i], src);
  return r;
}

// out_kind 0: 48-byte compressed point; 1: canonical big-endian affine x || y (96 bytes, all-zero = infinity)
__global__ void __launch_bounds__(32) vm_chain_kernel(uint8_t* __restrict__ out, int out_kind, G1Xyzz* __restrict__ chain, uint32_t* __restrict__ done,
                                                      const G1Xyzz* __restrict__ partials, int rblocks, int c, int W) {
  const int j = blockIdx.x, lane = threadIdx.x;
  // window sum: lanes add the block partials, shuffle tree
  G1Xyzz acc = xyzz_inf();
  for (int b = lane; b < rblocks; b += 32) {
    G1Xyzz v = partials[(size_t)j * rblocks + b];
    vm_add_c(acc, v);
  }
  if (rblocks > 1) {
    for (int d = 16; d > 0; d >>= 1) {
      G1Xyzz o = vm_shfl_down(acc, d);
      if (lane < d) vm_add_c(acc, o);
    }
  }
  acc = vm_shfl_xyzz(acc, 0);
  // 2^(c j) * window sum.  XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2) -> Jacobian with Z = ZZ: (X ZZ, Y ZZZ, ZZ)
  if (j > 0) {
    Fp X = fp_mul_nv(acc.x, acc.zz), Y = fp_mul_nv(acc.y, acc.zzz), Z = acc.zz;
    const int role = lane % 3;   // lanes 0, 1, 2 are the broadcast sources; the others compute copies
    for (int k = 0; k < c * j; k++) vm_jac_dbl_coop(X, Y, Z, role);
    acc.x = X; acc.y = Y;
    acc.zz = fp_sqr_nv(Z);
    acc.zzz = fp_mul_nv(acc.zz, Z);
  }
  if (lane == 0) chain[j] = acc;
  __threadfence();
  uint32_t ticket = 0;
  if (lane == 0) ticket = atomicAdd(done, 1u);
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket != (uint32_t)(W - 1)) return;
  // last warp: sum of the W chain results
  __threadfence();
  G1Xyzz tot = xyzz_inf();
  for (int b = lane; b < W; b += 32) {
    G1Xyzz v;   // written by other blocks: read past L1
    const uint4* src = reinterpret_cast<const uint4*>(chain + b);
    uint4* dst = reinterpret_cast<uint4*>(&v);
#pragma unroll
    for (int q = 0; q < 12; q++) dst[q] = __ldcg(src + q);
    vm_add_c(tot, v);
  }
  for (int d = 16; d > 0; d >>= 1) {
    G1Xyzz o = vm_shfl_down(tot, d);
    if (lane < d) vm_add_c(tot, o);
  }
  if (lane == 0) {
    const G1Affine a = xyzz_to_affine(tot);
    if (out_kind == 0) {
      g1_compress(out, a);
    } else if (g1a_is_inf(a)) {
      for (int k = 0; k < 96; k++) out[k] = 0;
    } else {
      fp_canon_to_be48(out, fp_from_mont(a.x));
      fp_canon_to_be48(out + 48, fp_from_mont(a.y));
    }
    *done = 0;
  }
}

static void vm_run(void* d_out, int out_kind, const void* d_points, const void* d_scalars, bool from_be, bool in_g1, size_t n, void* d_scratch,
                   cudaStream_t st) {
  const VmLayout L = vm_layout(n, in_g1);
  uint8_t* base = (uint8_t*)d_scratch;
  G1Affine* pts = (G1Affine*)(base + L.pts);
  uint32_t* halves = (uint32_t*)(base + L.halves);
  uint32_t* counts = (uint32_t*)(base + L.counts);
  uint32_t* offsets = (uint32_t*)(base + L.offsets);
  uint32_t* cursors = (uint32_t*)(base + L.cursors);
  uint32_t* idx = (uint32_t*)(base + L.idx);
  G1Xyzz* buckets = (G1Xyzz*)(base + L.buckets);
  G1Xyzz* partials = (G1Xyzz*)(base + L.partials);
  G1Xyzz* chain = (G1Xyzz*)(base + L.chain);
  uint32_t* heavy = (uint32_t*)(base + L.heavy);
  uint32_t* done = (uint32_t*)(base + L.done);
  int* bad = (int*)(base + L.bad);
  cudaMemsetAsync(base + L.zero_begin, 0, L.zero_end - L.zero_begin, st);
  const size_t nb = (size_t)(L.W + 1) * L.B;
  if (n) {
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (from_be)
      vm_prep_kernel<true><<<blocks, 128, 0, st>>>(pts, halves, counts, bad, (const uint8_t*)d_points, (const uint8_t*)d_scalars, n, L.glv ? 1 : 0, L.c, L.W, L.B);
    else
      vm_prep_kernel<false><<<blocks, 128, 0, st>>>(pts, halves, counts, bad, (const uint8_t*)d_points, (const uint8_t*)d_scalars, n, L.glv ? 1 : 0, L.c, L.W, L.B);
    vm_scan_kernel<<<L.W, 1024, 0, st>>>(offsets, counts, L.B);
    vm_scatter_kernel<<<(unsigned)((L.items + 127) / 128), 128, 0, st>>>(cursors, offsets, idx, halves, L.items, L.NL, L.c, L.W, L.B);
    count_launch(3);
  }
  // a lane group takes a bucket of up to ~3x the average run (at least 16 entries per lane); longer runs get a block
  const size_t avg = L.items / (size_t)L.B + 1;
  const uint32_t heavy_cap = (uint32_t)std::min<size_t>(std::max<size_t>(4 * avg, (size_t)16 * L.S), (size_t)384 * L.S);
  vm_accumulate_kernel<<<(unsigned)((nb * L.S + 127) / 128), 128, 0, st>>>(buckets, heavy, idx, offsets, counts, pts, L.items, nb, L.B, L.W, L.S, heavy_cap);
  vm_heavy_kernel<<<1184, VM_HEAVY_THREADS, 0, st>>>(buckets, heavy, idx, offsets, counts, pts, L.items, L.B, L.W);
  vm_reduce_kernel<<<dim3(L.rblocks, L.W), VM_RED_THREADS, 0, st>>>(partials, buckets, L.B, L.per);
  vm_chain_kernel<<<L.W, 32, 0, st>>>((uint8_t*)d_out, out_kind, chain, done, partials, L.rblocks, L.c, L.W);
  count_launch(4);
}

void launch_var_msm(void* d_out48, const void* d_points_xy_be, const void* d_scalars_be, size_t n, void* d_scratch, cudaStream_t st, bool points_in_g1) {
  vm_run(d_out48, 0, d_points_xy_be, d_scalars_be, true, points_in_g1, n, d_scratch, st);
}
void launch_var_msm_mont(void* d_out_affine_be96, const void* d_points_mont, const void* d_scalars_canon8, size_t n, void* d_scratch, cudaStream_t st) {
  vm_run(d_out_affine_be96, 1, d_points_mont, d_scalars_canon8, false, true, n, d_scratch, st);
}

// ---- synthetic inputs for the size sweep (BASELINE config 5): point t is the
// fixed-base table entry number (t * 2654435761) mod n_entries, i.e. a valid
// curve point with a KNOWN discrete log d * 2^(c j) * tau^i when the setup's tau
// is known (tests), exported in the public input format (canonical big-endian
// affine); scalar t = the synthetic blob word generator with blob id `seed`.
__host__ __device__ inline uint64_t vm_splitmix(uint64_t& state) {
  state += 0x9E3779B97F4A7C15ull;
  uint64_t z = state;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void vm_synth_kernel(uint8_t* __restrict__ pts_be, uint8_t* __restrict__ sc_be, const uint4* __restrict__ table,
                                unsigned long long n_entries, unsigned long long seed, size_t n) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  unsigned long long e = ((unsigned long long)t * 2654435761ull) % n_entries;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(table + e * 6);
  G1Affine p;
  for (int k = 0; k < 12; k++) { p.x.l[k] = w[k]; p.y.l[k] = w[12 + k]; }
  uint8_t* o = pts_be + t * 96;
  if (g1a_is_inf(p)) {
    for (int k = 0; k < 96; k++) o[k] = 0;
  } else {
    fp_canon_to_be48(o, fp_from_mont(p.x));
    fp_canon_to_be48(o + 48, fp_from_mont(p.y));
  }
  uint64_t st = 0xB2004844ull ^ (seed * 4096ull + (uint64_t)t);
  uint8_t* sc = sc_be + t * 32;
  for (int u = 0; u < 4; u++) {
    uint64_t v = vm_splitmix(st);
    for (int b = 0; b < 8; b++) sc[8 * u + b] = (uint8_t)(v >> (56 - 8 * b));
  }
  sc[0] &= 0x3f;
}
void launch_var_msm_synth(void* d_pts_be, void* d_sc_be, const void* d_table, unsigned long long n_entries, unsigned long long seed, size_t n,
                          cudaStream_t st) {
  if (!n) return;
  vm_synth_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((uint8_t*)d_pts_be, (uint8_t*)d_sc_be, (const uint4*)d_table, n_entries, seed, n);
  count_launch();
}
int var_msm_window_bits(size_t n) { return vm_layout(n ? n : 1, false).c; }

// offset of the "a point was not on the curve" flag inside the scratch buffer
size_t var_msm_bad_flag_offset(size_t n, bool points_in_g1) { return vm_layout(n, points_in_g1).bad; }

}  // namespace lw
