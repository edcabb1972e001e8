// PeerDAS / EIP-7594 cells and FK20 cell proofs.
//
// The reference stops before this: it loads all 65 G2 points of the trusted setup but only ever reads two
// (/root/reference/src/srs.rs:274; constants at src/lib.rs:60-92) and has no G1 FFT.  What is computed here follows
// consensus-specs `specs/fulu/polynomial-commitments-sampling.md`, restated in oracle/py/cells.py:
//
//   cells   : p in coefficient form (inverse FFT of the blob), evaluated on the 8192-point domain; cell i = the 64
//             evaluations on the coset h_i <w64>, h_i = w8192^brp7(i).  The even half of the domain IS the blob, so
//             only the odd half -- one 4096-point FFT of f_n w8192^n -- is computed.
//   proofs  : pi_i = [ (p - I_i)(tau) / (tau^64 - h_i^64) ] G for all 128 cosets at once (Feist-Khovratovich):
//             with l = 64, H_t = sum_k f_(k + 64 (t + 1)) [tau^k]G (t < 63) and  pi = DFT_128(H, 0 ...)  in
//             natural order of the coset shifts w8192^m.  The H_t are 64 upper-triangular Toeplitz products (one per
//             offset b = k mod 64), each embedded in a circulant of size 128:
//                 Hhat_j = sum_b DFT_128(c^b)_j * X^b_j ,   X^b = DFT_128(s_(64 (62 - v) + b))_v   (fixed, set-up time)
//             i.e. per blob 64 scalar FFTs of size 128, ONE fixed-base MSM per frequency j over 64 points (128 MSMs,
//             8192 points in all, served by a GLV digit table exactly like the commitment MSM's), one inverse and one
//             forward G1 FFT of size 128.
//
// B200 mapping: the scalar FFTs run out of shared memory (a blob's 4096 coefficients = 128 KB, limb-major so that a
// warp's accesses are conflict-free); the MSM is one warp per (blob, frequency) gathering table entries; the G1 FFTs
// are batched ACROSS blobs -- a warp holds the same butterfly of 32 different blobs, so the twiddle (a fixed 128th
// root of unity, recoded once into signed digits of its two GLV halves) drives a warp-uniform double-and-add ladder.
#include "kernels_cells.h"
#include "msm_common.cuh"

namespace lw {

namespace {

__device__ __forceinline__ Fr fr_const(const uint32_t* c) { Fr r; for (int i = 0; i < 8; i++) r.l[i] = c[i]; return r; }
__device__ __forceinline__ Fp fp_beta() { Fp b; for (int i = 0; i < 12; i++) b.l[i] = k::FP_BETA[i]; return b; }
__device__ __forceinline__ uint32_t brp7(uint32_t i) { return __brev(i) >> 25; }
__device__ __forceinline__ Fr fr_pow_small(Fr base, uint32_t e) {
  Fr acc = fr_one();
  for (int bit = 31; bit >= 0; bit--) {
    acc = fr_sqr(acc);
    if ((e >> bit) & 1u) acc = fr_mul(acc, base);
  }
  return acc;
}

// ------------------------------------------------------------------ setup
__global__ void cell_twiddles_kernel(Fr* __restrict__ tw) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < EXT_POINTS) tw[i] = fr_pow_small(fr_const(k::FR_ROOT_8192), i);
}

// Width-4 non-adjacent-form recoding of the GLV halves of the 128th roots of unity nu^e, nu = w8192^64:
// [nu^e]P = [m]P + [q](beta x, -y), digits in {0, +-1, +-3, +-5, +-7}, at most one non-zero digit in any four
// consecutive positions.  Layout per e: 160 int8 digits of m, then 160 of q, least significant first.
__global__ void cell_twiddle_naf_kernel(int8_t* __restrict__ naf) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 128) return;
  Fr w = fr_from_mont(fr_pow_small(fr_const(k::FR_ROOT_8192), 64u * e));
  uint32_t q[4], m[4];
  glv_split(q, m, w.l);
  for (int half = 0; half < 2; half++) {
    uint32_t kk[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) kk[i] = half ? q[i] : m[i];
    int8_t* out = naf + (size_t)e * CELL_NAF_BYTES + half * 160;
    for (int bit = 0; bit < 160; bit++) {
      int d = 0;
      if (kk[0] & 1u) {
        d = (int)(kk[0] & 15u);
        if (d >= 8) d -= 16;
        // k -= d
        if (d > 0) {
          kk[0] -= (uint32_t)d;          // the low four bits are exactly d: no borrow
        } else {
          uint32_t carry = (uint32_t)(-d);
          for (int i = 0; i < 5 && carry; i++) {
            const uint32_t t = kk[i] + carry;
            carry = t < kk[i] ? 1u : 0u;
            kk[i] = t;
          }
        }
      }
      out[bit] = (int8_t)d;
      for (int i = 0; i < 4; i++) kk[i] = (kk[i] >> 1) | (kk[i + 1] << 31);
      kk[4] >>= 1;
    }
  }
}

__global__ void cell_srs_columns_kernel(G1Xyzz* __restrict__ pts, const G1Affine* __restrict__ srs) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= EXT_POINTS) return;
  const int v = t / 64, b = t % 64;
  pts[t] = v <= 62 ? xyzz_from_affine(srs[64 * (62 - v) + b]) : xyzz_inf();
}

__global__ void cell_fk20_points_kernel(G1Affine* __restrict__ out, const G1Xyzz* __restrict__ pts) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= EXT_POINTS) return;
  const int j = t / 64, b = t % 64;
  // MSM order: half v = j / 64 has its own 4096-point table, point b * 64 + (j % 64) -- so the points of ONE frequency
  // are those congruent to j mod 64, which is what a thread of the batched-affine kernel sums (msm_ba.cuh)
  out[(j / 64) * N_POINTS + b * 64 + (j % 64)] = xyzz_to_affine(pts[brp7(j) * 64 + b]);
}

// ------------------------------------------------------------------ G1 FFT stage
// Compact group law for the twiddle ladders: every field multiplication is a CALL to the one out-of-line multiplier
// (fp_mul_nv / fp_sqr_nv), so a doubling or an addition is a few hundred bytes of SASS instead of ~40 KB -- a stage runs
// only a handful of warps per SM, each at its own place in the ladder, and the force-inlined formulas thrashed the
// instruction cache.
__device__ __noinline__ void cell_dbl_c(G1Xyzz& p) {
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
// acc += (x, y), an affine point that is not infinity
__device__ __noinline__ void cell_madd_c(G1Xyzz& acc, const Fp& x, const Fp& y) {
  if (xyzz_is_inf(acc)) {
    acc.x = x; acc.y = y; acc.zz = fp_one(); acc.zzz = fp_one();
  } else {
    const Fp Pd = fp_sub(fp_mul_nv(x, acc.zz), acc.x);
    const Fp Rd = fp_sub(fp_mul_nv(y, acc.zzz), acc.y);
    if (fp_is_zero(Pd)) {
      G1Affine a;
      a.x = x; a.y = y;
      xyzz_madd_rare(acc, a);
    } else {
      const Fp PP = fp_sqr_nv(Pd), PPP = fp_mul_nv(Pd, PP), Q = fp_mul_nv(acc.x, PP);
      const Fp X3 = fp_sub(fp_sub(fp_sqr_nv(Rd), PPP), fp_dbl(Q));
      acc.y = fp_sub(fp_mul_nv(Rd, fp_sub(Q, X3)), fp_mul_nv(acc.y, PPP));
      acc.x = X3;
      acc.zz = fp_mul_nv(acc.zz, PP);
      acc.zzz = fp_mul_nv(acc.zzz, PPP);
    }
  }
}
// a += b (both XYZZ), b negated first if neg_b
__device__ __noinline__ void cell_add_c(G1Xyzz& a, const G1Xyzz& b, bool neg_b) {
  if (!xyzz_is_inf(b)) {
    const Fp by = fp_cneg(b.y, neg_b);
    if (xyzz_is_inf(a)) {
      // unrolled limb copies, not a struct assignment: see the ptxas note in DESIGN 3.5
#pragma unroll
      for (int i = 0; i < 12; i++) { a.x.l[i] = b.x.l[i]; a.y.l[i] = by.l[i]; a.zz.l[i] = b.zz.l[i]; a.zzz.l[i] = b.zzz.l[i]; }
    } else {
      const Fp U1 = fp_mul_nv(a.x, b.zz), S1 = fp_mul_nv(a.y, b.zzz);
      const Fp Pd = fp_sub(fp_mul_nv(b.x, a.zz), U1), Rd = fp_sub(fp_mul_nv(by, a.zzz), S1);
      if (fp_is_zero(Pd)) {
        if (fp_is_zero(Rd)) cell_dbl_c(a);
        else a = xyzz_inf();
      } else {
        const Fp PP = fp_sqr_nv(Pd), PPP = fp_mul_nv(Pd, PP), Q = fp_mul_nv(U1, PP);
        const Fp X3 = fp_sub(fp_sub(fp_sqr_nv(Rd), PPP), fp_dbl(Q));
        a.y = fp_sub(fp_mul_nv(Rd, fp_sub(Q, X3)), fp_mul_nv(S1, PPP));
        a.x = X3;
        a.zz = fp_mul_nv(fp_mul_nv(a.zz, b.zz), PP);
        a.zzz = fp_mul_nv(fp_mul_nv(a.zzz, b.zzz), PPP);
      }
    }
  }
}

// Jacobian coordinates for the twiddle ladders: a doubling is 2 M + 5 S (1 770 MAC32) against 6 M + 3 S (2 502) in XYZZ,
// and a ladder is mostly doublings.  Z = 0 is infinity (and stays so under the doubling formulas).
struct G1Jac {
  Fp x, y, z;
};
__device__ __noinline__ void jac_dbl_c(G1Jac& p) {
  const Fp A = fp_sqr_nv(p.x), B = fp_sqr_nv(p.y), C = fp_sqr_nv(B);
  const Fp D = fp_dbl(fp_sub(fp_sub(fp_sqr_nv(fp_add(p.x, B)), A), C));
  const Fp E = fp_add(fp_dbl(A), A);
  const Fp X3 = fp_sub(fp_sqr_nv(E), fp_dbl(D));
  const Fp Z3 = fp_dbl(fp_mul_nv(p.y, p.z));
  p.y = fp_sub(fp_mul_nv(E, fp_sub(D, X3)), fp_dbl(fp_dbl(fp_dbl(C))));
  p.x = X3;
  p.z = Z3;
}
// p += (x2, +-y2), an affine point that is not infinity (7 M + 4 S)
__device__ __noinline__ void jac_madd_c(G1Jac& p, const Fp& x2, const Fp& y2in, bool neg) {
  const Fp y2 = fp_cneg(y2in, neg);
  if (fp_is_zero(p.z)) {
    p.x = x2; p.y = y2; p.z = fp_one();
  } else {
    const Fp Z1Z1 = fp_sqr_nv(p.z), U2 = fp_mul_nv(x2, Z1Z1), S2 = fp_mul_nv(fp_mul_nv(y2, p.z), Z1Z1);
    const Fp H = fp_sub(U2, p.x), R0 = fp_sub(S2, p.y);
    if (fp_is_zero(H)) {
      if (fp_is_zero(R0)) {   // the same point: double it
        p.x = x2; p.y = y2; p.z = fp_one();
        jac_dbl_c(p);
      } else {
        p.z = fp_zero();
      }
    } else {
      const Fp HH = fp_sqr_nv(H), I = fp_dbl(fp_dbl(HH)), J = fp_mul_nv(H, I), R2 = fp_dbl(R0), V = fp_mul_nv(p.x, I);
      const Fp X3 = fp_sub(fp_sub(fp_sqr_nv(R2), J), fp_dbl(V));
      const Fp Z3 = fp_sub(fp_sub(fp_sqr_nv(fp_add(p.z, H)), Z1Z1), HH);
      p.y = fp_sub(fp_mul_nv(R2, fp_sub(V, X3)), fp_dbl(fp_mul_nv(p.y, J)));
      p.x = X3;
      p.z = Z3;
    }
  }
}

// t <- [nu^e] t: width-4 NAF ladder over the two GLV halves (warp-uniform: every lane of a warp has the same e).
// 129 doublings + ~52 mixed additions from a table of T, 3T, 5T, 7T (and psi of them: beta x, -y).
__device__ __noinline__ void g1_mul_root(G1Xyzz& t, const int8_t* __restrict__ naf_e) {
  if (!xyzz_is_inf(t)) {
    Fp tx[4], ty[4], bx[4];
    {
      const G1Affine a = xyzz_to_affine(t);
      tx[0] = a.x; ty[0] = a.y;
      // 2T, 3T = 2T + T, 4T, 5T = 4T + T, 8T, 7T = 8T - T: doublings and mixed additions only
      G1Jac d;
      d.x = a.x; d.y = a.y; d.z = fp_one();
      jac_dbl_c(d);
      G1Jac p3 = d;
      jac_madd_c(p3, a.x, a.y, false);
      jac_dbl_c(d);
      G1Jac p5 = d;
      jac_madd_c(p5, a.x, a.y, false);
      jac_dbl_c(d);
      jac_madd_c(d, a.x, a.y, true);   // 7T
      // one inversion for the three Z's (none is zero: T has prime order r > 8)
      const Fp z35 = fp_mul_nv(p3.z, p5.z);
      Fp inv = fp_inv(fp_mul_nv(z35, d.z));
      const Fp iz7 = fp_mul_nv(inv, z35);
      inv = fp_mul_nv(inv, d.z);                 // 1 / (z3 z5)
      const Fp iz3 = fp_mul_nv(inv, p5.z), iz5 = fp_mul_nv(inv, p3.z);
      Fp i2 = fp_sqr_nv(iz3);
      tx[1] = fp_mul_nv(p3.x, i2); ty[1] = fp_mul_nv(p3.y, fp_mul_nv(i2, iz3));
      i2 = fp_sqr_nv(iz5);
      tx[2] = fp_mul_nv(p5.x, i2); ty[2] = fp_mul_nv(p5.y, fp_mul_nv(i2, iz5));
      i2 = fp_sqr_nv(iz7);
      tx[3] = fp_mul_nv(d.x, i2); ty[3] = fp_mul_nv(d.y, fp_mul_nv(i2, iz7));
      const Fp beta = fp_beta();
      for (int i = 0; i < 4; i++) bx[i] = fp_mul_nv(tx[i], beta);
    }
    const int8_t* dm = naf_e;
    const int8_t* dq = naf_e + 160;
    int top = 159;
    while (top > 0 && dm[top] == 0 && dq[top] == 0) top--;
    G1Jac acc;
    acc.x = fp_zero(); acc.y = fp_zero(); acc.z = fp_zero();
    for (int bit = top; bit >= 0; bit--) {
      jac_dbl_c(acc);
      const int a = dm[bit], b = dq[bit];
      if (a) { const int i = ((a < 0 ? -a : a) - 1) >> 1; jac_madd_c(acc, tx[i], ty[i], a < 0); }
      if (b) { const int i = ((b < 0 ? -b : b) - 1) >> 1; jac_madd_c(acc, bx[i], ty[i], b > 0); }   // psi negates y
    }
    // Jacobian -> XYZZ: zz = Z^2, zzz = Z^3
    t.x = acc.x;
    t.y = acc.y;
    t.zz = fp_sqr_nv(acc.z);
    t.zzz = fp_mul_nv(t.zz, acc.z);
    if (fp_is_zero(acc.z)) t = xyzz_inf();
  }
}

#ifndef LWKZG_CELL_FFT_MIN_BLOCKS
#define LWKZG_CELL_FFT_MIN_BLOCKS 8   // 128 registers, 16 warps per SM: +5 % over 6 blocks / 168 registers (B200, profiles/r02_cells_chunk_sweep.log)
#endif
__global__ void __launch_bounds__(64, LWKZG_CELL_FFT_MIN_BLOCKS)
cell_g1_fft_stage_kernel(G1Xyzz* __restrict__ pts, int batch, int batch_pad, int half, int dif, int inverse, int upper_half_zero,
                         const int8_t* __restrict__ naf) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int item = (int)(gt % batch_pad), bf = (int)(gt / batch_pad);
  if (bf >= 64 || item >= batch) return;
  const int j = bf & (half - 1);
  const int i0 = ((bf - j) << 1) + j, i1 = i0 + half;
  int e = j * (64 / half);
  if (inverse) e = (128 - e) & 127;
  G1Xyzz* p0 = pts + (size_t)i0 * batch + item;
  G1Xyzz* p1 = pts + (size_t)i1 * batch + item;
  if (dif) {
    // (P, Q) -> (P + Q, [w](P - Q)).  upper_half_zero on the last stage: the odd outputs are H_t, t >= 64, which the
    // caller discards -- they are written as infinity instead of being computed
    const G1Xyzz Q = *p1;
    G1Xyzz s = *p0;
    if (upper_half_zero && half == 1) {
      cell_add_c(s, Q, false);
      *p0 = s;
      *p1 = xyzz_inf();
    } else {
      G1Xyzz d = s;
      cell_add_c(s, Q, false);
      *p0 = s;
      cell_add_c(d, Q, true);
      if (e) g1_mul_root(d, naf + (size_t)e * CELL_NAF_BYTES);
      *p1 = d;
    }
  } else {
    // (P, Q) -> (P + [w]Q, P - [w]Q)
    G1Xyzz t = *p1;
    if (e) g1_mul_root(t, naf + (size_t)e * CELL_NAF_BYTES);
    G1Xyzz s = *p0;
    G1Xyzz d = s;
    cell_add_c(s, t, false);
    *p0 = s;
    cell_add_c(d, t, true);
    *p1 = d;
  }
}

// ------------------------------------------------------------------ scalar FFTs in shared memory
// limb-major: limb l of element i at sm[l * N + i]
template <int N>
__device__ __forceinline__ Fr sm_load(const uint32_t* sm, int i) {
  Fr r;
#pragma unroll
  for (int l = 0; l < 8; l++) r.l[l] = sm[l * N + i];
  return r;
}
template <int N>
__device__ __forceinline__ void sm_store(uint32_t* sm, int i, const Fr& v) {
#pragma unroll
  for (int l = 0; l < 8; l++) sm[l * N + i] = v.l[l];
}

// bit-reversed in -> natural out, inverse twiddles, unscaled.  tw = powers of w8192; the N-th root is w8192^(8192/N)
template <int N>
__device__ void sm_dit_inverse(uint32_t* sm, const Fr* __restrict__ tw) {
  for (int half = 1; half < N; half <<= 1) {
    const int step = (EXT_POINTS / 2) / half;   // exponent of w8192 per unit j: 8192 / (2 half)
    for (int t = threadIdx.x; t < N / 2; t += blockDim.x) {
      const int j = t & (half - 1);
      const int i0 = ((t - j) << 1) + j, i1 = i0 + half;
      Fr u = sm_load<N>(sm, i0), v = sm_load<N>(sm, i1);
      if (j) v = fr_mul(v, tw[(EXT_POINTS - step * j) & (EXT_POINTS - 1)]);
      sm_store<N>(sm, i0, fr_add(u, v));
      sm_store<N>(sm, i1, fr_sub(u, v));
    }
    __syncthreads();
  }
}
// natural in -> bit-reversed out, forward twiddles
template <int N>
__device__ void sm_dif_forward(uint32_t* sm, const Fr* __restrict__ tw) {
  for (int half = N / 2; half >= 1; half >>= 1) {
    const int step = (EXT_POINTS / 2) / half;
    for (int t = threadIdx.x; t < N / 2; t += blockDim.x) {
      const int j = t & (half - 1);
      const int i0 = ((t - j) << 1) + j, i1 = i0 + half;
      Fr u = sm_load<N>(sm, i0), v = sm_load<N>(sm, i1);
      sm_store<N>(sm, i0, fr_add(u, v));
      Fr d = fr_sub(u, v);
      if (j) d = fr_mul(d, tw[step * j]);
      sm_store<N>(sm, i1, d);
    }
    __syncthreads();
  }
}

// 32 bytes of a blob / cell <-> canonical field element in the mode's byte order (mode 1: little-endian)
__device__ __forceinline__ Fr fr_from_wire(const uint8_t* p, int mode) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  Fr r;
  if (mode == 1) {
    for (int i = 0; i < 8; i++) r.l[i] = w[i];
    mod_reduce_small<FrCfg, 2>(r.l);
  } else {
    r = fr_canon_from_be_words(w);
  }
  return r;
}
__device__ __forceinline__ void fr_to_wire(uint8_t* p, const Fr& canon, int mode) {
  uint4 a, b;
  if (mode == 1) {
    a = make_uint4(canon.l[0], canon.l[1], canon.l[2], canon.l[3]);
    b = make_uint4(canon.l[4], canon.l[5], canon.l[6], canon.l[7]);
  } else {
    a = make_uint4(bswap32(canon.l[7]), bswap32(canon.l[6]), bswap32(canon.l[5]), bswap32(canon.l[4]));
    b = make_uint4(bswap32(canon.l[3]), bswap32(canon.l[2]), bswap32(canon.l[1]), bswap32(canon.l[0]));
  }
  reinterpret_cast<uint4*>(p)[0] = a;
  reinterpret_cast<uint4*>(p)[1] = b;
}

constexpr int POLY_THREADS = 512;
constexpr int POLY_SMEM = N_POINTS * 32;

// One block per blob.  from_blob: parse the blob (and turn evaluations into coefficients in the Lagrange modes),
// write the coefficient form; else read it from d_coef.  Then, if cells are wanted, the extended evaluations.
__global__ void __launch_bounds__(POLY_THREADS) cell_poly_kernel(Fr* __restrict__ coef, uint8_t* __restrict__ cells, const uint8_t* __restrict__ blobs,
                                                               int mode, int from_blob, const Fr* __restrict__ tw) {
  extern __shared__ uint32_t sm[];
  const int blob = blockIdx.x;
  Fr* cf = coef + (size_t)blob * N_POINTS;
  uint8_t* out = cells ? cells + (size_t)blob * N_CELLS * CELL_BYTES : nullptr;
  const bool evals_in = from_blob && mode != 0;
  if (from_blob) {
    const uint8_t* src = blobs + (size_t)blob * BLOB_BYTES;
    for (int i = threadIdx.x; i < N_POINTS; i += blockDim.x) sm_store<N_POINTS>(sm, i, fr_to_mont(fr_from_wire(src + (size_t)i * 32, mode)));
    __syncthreads();
    if (evals_in) {
      sm_dit_inverse<N_POINTS>(sm, tw);
      const Fr ninv = fr_const(k::FR_N_INV);
      for (int i = threadIdx.x; i < N_POINTS; i += blockDim.x) cf[i] = fr_mul(sm_load<N_POINTS>(sm, i), ninv);
    } else {
      for (int i = threadIdx.x; i < N_POINTS; i += blockDim.x) cf[i] = sm_load<N_POINTS>(sm, i);
    }
    __syncthreads();   // cf is re-read below by other threads of this block
  }
  if (!out) return;
  // first half of the extended blob: p on the 4096th roots of unity in bit-reversed order
  if (evals_in) {
    const uint4* s4 = reinterpret_cast<const uint4*>(blobs + (size_t)blob * BLOB_BYTES);
    uint4* d4 = reinterpret_cast<uint4*>(out);
    for (int i = threadIdx.x; i < BLOB_BYTES / 16; i += blockDim.x) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = threadIdx.x; i < N_POINTS; i += blockDim.x) sm_store<N_POINTS>(sm, i, cf[i]);
    __syncthreads();
    sm_dif_forward<N_POINTS>(sm, tw);
    for (int i = threadIdx.x; i < N_POINTS; i += blockDim.x) fr_to_wire(out + (size_t)i * 32, fr_from_mont(sm_load<N_POINTS>(sm, i)), mode);
    __syncthreads();
  }
  // second half: p on w8192 * (the same roots)
  for (int i = threadIdx.x; i < N_POINTS; i += blockDim.x) sm_store<N_POINTS>(sm, i, fr_mul(cf[i], tw[i]));
  __syncthreads();
  sm_dif_forward<N_POINTS>(sm, tw);
  for (int i = threadIdx.x; i < N_POINTS; i += blockDim.x)
    fr_to_wire(out + (size_t)(N_POINTS + i) * 32, fr_from_mont(sm_load<N_POINTS>(sm, i)), mode);
}

// FK20 scalars.  Block = 4 offsets b x 64 threads; grid (16, n_blobs).
constexpr int TOEP_COLS = 4;
__global__ void __launch_bounds__(64 * TOEP_COLS) cell_toeplitz_kernel(uint32_t* __restrict__ scalars, const Fr* __restrict__ coef, const Fr* __restrict__ tw) {
  __shared__ uint32_t smem[TOEP_COLS][8 * 128];
  const int col = threadIdx.x / 64, t = threadIdx.x % 64;
  const int b = blockIdx.x * TOEP_COLS + col, blob = blockIdx.y;
  const Fr* cf = coef + (size_t)blob * N_POINTS;
  uint32_t* sm = smem[col];
  // c_0 = f_(64 * 63 + b); c_u = 0 for 1 <= u <= 65; c_u = f_(64 (u - 65) + b) for 66 <= u <= 127
  for (int u = t; u < 128; u += 64) {
    Fr v = fr_zero();
    if (u == 0) v = cf[64 * 63 + b];
    else if (u >= 66) v = cf[64 * (u - 65) + b];
    sm_store<128>(sm, u, v);
  }
  __syncthreads();
  for (int half = 64; half >= 1; half >>= 1) {
    const int step = (EXT_POINTS / 2) / half;
    const int j = t & (half - 1);
    const int i0 = ((t - j) << 1) + j, i1 = i0 + half;
    Fr u = sm_load<128>(sm, i0), v = sm_load<128>(sm, i1);
    sm_store<128>(sm, i0, fr_add(u, v));
    Fr d = fr_sub(u, v);
    if (j) d = fr_mul(d, tw[step * j]);
    sm_store<128>(sm, i1, d);
    __syncthreads();
  }
  const Fr inv128 = fr_const(k::FR_INV_128);
  for (int pos = t; pos < 128; pos += 64) {
    const Fr v = fr_from_mont(fr_mul(sm_load<128>(sm, pos), inv128));
    const int j = brp7(pos);
    uint4* dst = reinterpret_cast<uint4*>(scalars + ((((size_t)(j / 64) * gridDim.y + blob) * N_POINTS) + b * 64 + (j % 64)) * 8);
    dst[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    dst[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
  }
}

// One warp per (blob, frequency): 64 points, lanes 0..15 sum the m-halves, lanes 16..31 the q-halves.  The path for
// small batches; from CELL_BA_MIN_BLOBS blobs up the batched-affine kernel of the commitment MSM runs in its
// segmented form (msm_ba.cuh: 5 M + 1 S per table entry instead of 8 M + 2 S).
__global__ void __launch_bounds__(32) cell_msm_kernel(G1Xyzz* __restrict__ out, const uint4* __restrict__ table, size_t half_words, const uint8_t* __restrict__ scalars,
                                                     int c, int nwin, uint32_t cnt_top, int n_blobs) {
  __shared__ uint32_t sk[4][32];
  __shared__ uint32_t red[48 * 16];
  const int tid = threadIdx.x;
  const int j = blockIdx.x % N_CELLS, blob = blockIdx.x / N_CELLS;
  const int v = j / 64, jl = j % 64;
  const uint4* tab = table + (size_t)v * half_words;
  const uint8_t* sc = scalars + ((size_t)v * n_blobs + blob) * BLOB_BYTES;
  const int half = tid / 16, pl = tid % 16;
  auto limb = [&](int w) { return sk[w][tid]; };
  G1Xyzz acc = xyzz_inf();
  for (int b = pl; b < 64; b += 16) {
    const int pi = b * 64 + jl;
    uint32_t h4[4];
    load_scalar_half<false>(h4, sc, pi, half);
#pragma unroll
    for (int i = 0; i < 4; i++) sk[i][tid] = h4[i];
    int carry = 0, d = 0;
    G1Affine cur = g1a_inf();
    for (int w = 0; w <= nwin; w++) {
      int dn = 0;
      G1Affine nxt = g1a_inf();
      if (w < nwin) {
        dn = glv_digit(limb, c, nwin, w, carry);
        if (dn != 0) nxt = load_entry(tab, entry_index(c, nwin, cnt_top, w, pi, dn < 0 ? -dn : dn));
      }
      if (d != 0) {
        cur.y = fp_cneg(cur.y, d < 0);
        xyzz_madd_hot(acc, cur);
      }
      cur = nxt; d = dn;
    }
  }
  block_reduce_xyzz_glv<32>(acc, red);
  if (tid == 0) out[(size_t)j * n_blobs + blob] = acc;
}

__global__ void cell_proofs_finalize_kernel(uint8_t* __restrict__ proofs, const G1Xyzz* __restrict__ pts, int n_blobs) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)n_blobs * N_CELLS) return;
  const int blob = (int)(t % n_blobs), i = (int)(t / n_blobs);
  const G1Affine a = xyzz_to_affine(pts[(size_t)brp7(i) * n_blobs + blob]);
  g1_compress(proofs + ((size_t)blob * N_CELLS + i) * 48, a);
}

}  // namespace

void launch_cell_twiddles(void* d_tw, cudaStream_t st) {
  cell_twiddles_kernel<<<EXT_POINTS / 128, 128, 0, st>>>((Fr*)d_tw);
  count_launch();
}
void launch_cell_twiddle_naf(void* d_naf, cudaStream_t st) {
  cell_twiddle_naf_kernel<<<1, 128, 0, st>>>((int8_t*)d_naf);
  count_launch();
}
void launch_cell_srs_columns(void* d_pts, const void* d_srs, cudaStream_t st) {
  cell_srs_columns_kernel<<<EXT_POINTS / 128, 128, 0, st>>>((G1Xyzz*)d_pts, (const G1Affine*)d_srs);
  count_launch();
}
void launch_cell_fk20_points(void* d_aff, const void* d_pts, cudaStream_t st) {
  cell_fk20_points_kernel<<<EXT_POINTS / 64, 64, 0, st>>>((G1Affine*)d_aff, (const G1Xyzz*)d_pts);
  count_launch();
}
void launch_cell_g1_fft_stage(void* d_pts, int batch, int half, bool dif, bool inverse, bool upper_half_zero, const void* d_naf, cudaStream_t st) {
  if (batch <= 0) return;
  const int pad = (batch + 31) / 32 * 32;
  const long threads = (long)pad * 64;
  cell_g1_fft_stage_kernel<<<(unsigned)((threads + 63) / 64), 64, 0, st>>>((G1Xyzz*)d_pts, batch, pad, half, dif ? 1 : 0, inverse ? 1 : 0,
                                                                          upper_half_zero ? 1 : 0, (const int8_t*)d_naf);
  count_launch();
}
static void poly_launch(void* d_coef, void* d_cells, const void* d_blobs, int n, int mode, int from_blob, const void* d_tw, cudaStream_t st) {
  if (n <= 0) return;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(cell_poly_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POLY_SMEM);
    attr_set[dev] = true;
  }
  cell_poly_kernel<<<n, POLY_THREADS, POLY_SMEM, st>>>((Fr*)d_coef, (uint8_t*)d_cells, (const uint8_t*)d_blobs, mode, from_blob, (const Fr*)d_tw);
  count_launch();
}
void launch_cell_poly(void* d_coef, void* d_cells, const void* d_blobs, int n, int mode, const void* d_tw, cudaStream_t st) {
  poly_launch(d_coef, d_cells, d_blobs, n, mode, 1, d_tw, st);
}
void launch_cell_coef_to_cells(void* d_cells, const void* d_coef, int n, int mode, const void* d_tw, cudaStream_t st) {
  poly_launch(const_cast<void*>(d_coef), d_cells, nullptr, n, mode, 0, d_tw, st);
}
void launch_cell_toeplitz(void* d_scalars, const void* d_coef, int n, const void* d_tw, cudaStream_t st) {
  if (n <= 0) return;
  cell_toeplitz_kernel<<<dim3(64 / TOEP_COLS, n), 64 * TOEP_COLS, 0, st>>>((uint32_t*)d_scalars, (const Fr*)d_coef, (const Fr*)d_tw);
  count_launch();
}
void launch_ba_segmented(void* d_out, const void* d_table, int c, const void* d_scalars, int n_blobs, void* d_scratch, int seg_stride, cudaStream_t st);   // msm_ba_v0.cu
size_t cell_table_half_entries(int c) { return (size_t)glv_table_entries(c, N_POINTS); }
size_t cell_msm_scratch_bytes(int n) { return n >= CELL_BA_MIN_BLOBS ? msm_ba_scratch_bytes(n) : 16; }
void launch_cell_msm(void* d_pts, const void* d_table, int c, const void* d_scalars, int n, void* d_scratch, cudaStream_t st) {
  if (n <= 0) return;
  const size_t half_entries = cell_table_half_entries(c);
  if (n >= CELL_BA_MIN_BLOBS && glv_num_windows(c) <= 64) {
    for (int v = 0; v < 2; v++)
      launch_ba_segmented((G1Xyzz*)d_pts + (size_t)v * 64 * n, (const uint8_t*)d_table + (size_t)v * half_entries * AFFINE_BYTES, c,
                          (const uint8_t*)d_scalars + (size_t)v * n * BLOB_BYTES, n, d_scratch, n, st);
    count_launch(2);
    return;
  }
  cell_msm_kernel<<<(unsigned)n * N_CELLS, 32, 0, st>>>((G1Xyzz*)d_pts, (const uint4*)d_table, half_entries * 6, (const uint8_t*)d_scalars, c, glv_num_windows(c),
                                                     glv_top_max(c) + 1u, n);
  count_launch();
}
void launch_cell_proofs_finalize(void* d_proofs48, const void* d_pts, int n, cudaStream_t st) {
  if (n <= 0) return;
  const long total = (long)n * N_CELLS;
  cell_proofs_finalize_kernel<<<(unsigned)((total + 63) / 64), 64, 0, st>>>((uint8_t*)d_proofs48, (const G1Xyzz*)d_pts, n);
  count_launch();
}

}  // namespace lw
