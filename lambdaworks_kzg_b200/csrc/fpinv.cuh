// Fp inversion by an approximate binary extended GCD (Pornin, "Optimized Binary
// GCD for Modular Inversion", ePrint 2020/972, algorithm 2, k = 31), branch-free
// inside a round so that the 32 lanes of a warp stay converged.
//
// Why it exists: Fermat inversion (a^(p-2), ~460 Montgomery products on the
// integer-multiply pipe, which is the pipe the MSM is bound by) is too expensive
// to share between a handful of affine additions.  This routine needs ~120 wide
// MACs per round and ~20 rounds (~8 product-equivalents); the rest of its work
// is ALU instructions, a pipe the MSM kernel leaves ~80 % idle.  It is what makes
// batched-affine accumulation (g1_batch.cuh) profitable.
//
// Each round: take 62-bit approximations (top 32 bits + low 30 bits) of a and b,
// run 30 divsteps on them while recording the 2 x 2 transition matrix
// (f0 g0; f1 g1), then apply the matrix to the full-size (a, b) -- an exact
// division by 2^30 -- and to the cofactors (u, v) modulo p with one Montgomery
// step (division by 2^32).  Invariant: a = 4^t u y, b = 4^t v y (mod p) after t
// rounds, so when a reaches 0 (b = gcd = 1): y^-1 = 4^t v.
//
// Replaces FieldElement::inv of the un-vendored lambdaworks-math dependency
// (used by to_affine in /root/reference/src/compression.rs:33-60 and inside
// the dependency's MSM); same value as fp_inv (field.cuh), 0 -> 0.
#pragma once
#include "field.cuh"

namespace lw {
namespace gcdinv {

constexpr int STEPS = 30;        // divsteps per round (k - 1)
constexpr int MAX_ROUNDS = 30;   // 26 rounds cover the 2 * 381 - 1 divstep bound; observed maximum 26, typical 18-19

#if defined(LWKZG_HOST_EMUL)
inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
inline uint32_t shl_pair_hi(uint32_t lo, uint32_t hi, int s) { return s ? ((hi << s) | (lo >> (32 - s))) : hi; }
inline bool warp_all(bool p) { return p; }
#else
LW_INL int clz32(uint32_t x) { return __clz((int)x); }
LW_INL uint32_t shl_pair_hi(uint32_t lo, uint32_t hi, int s) { return __funnelshift_l(lo, hi, s); }
LW_INL bool warp_all(bool p) { return __all_sync(__activemask(), p); }
#endif

// X (13 limbs) = a (12 limbs) * w
LW_INL void mul_word(uint32_t* X, const uint32_t* a, uint32_t w) {
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    acc = (uint64_t)a[i] * w + (acc >> 32);
    X[i] = (uint32_t)acc;
  }
  X[12] = (uint32_t)(acc >> 32);
}

// r = |a f + b g| / 2^30 (the division is exact), returns all-ones when a f + b g < 0.
// 0 <= a, b < 2^382; |f|, |g| <= 2^30.  Out of line (called twice per round): the inversion runs
// concurrently with other warps' multiplication code and must not evict it from the instruction cache.
LW_INL uint32_t lincomb_shr_inl(uint32_t* r, const uint32_t* a, uint32_t f, const uint32_t* b, uint32_t g) {
  const uint32_t mf = (uint32_t)((int32_t)f >> 31), mg = (uint32_t)((int32_t)g >> 31);
  const uint32_t af = (f ^ mf) - mf, ag = (g ^ mg) - mg;
  uint32_t X[13], Y[13];
  mul_word(X, a, af);
  mul_word(Y, b, ag);
  // D = X + Y when f and g have the same sign, X - Y otherwise (two's complement, 14 limbs);
  // a f + b g = D for f >= 0 and -D for f < 0
  const uint32_t m = mf ^ mg;
  uint32_t D[14];
  ptx::add_cc(m, m);  // carry-in = m & 1
  D[0] = ptx::addc_cc(X[0], Y[0] ^ m);
#pragma unroll
  for (int i = 1; i < 13; i++) D[i] = ptx::addc_cc(X[i], Y[i] ^ m);
  D[13] = ptx::addc(0, m);
  const uint32_t sd = (uint32_t)((int32_t)D[13] >> 31);
  uint32_t s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = (D[i] >> 30) | (D[i + 1] << 2);
  ptx::add_cc(sd, sd);
  r[0] = ptx::addc_cc(s[0] ^ sd, 0);
#pragma unroll
  for (int i = 1; i < 12; i++) r[i] = ptx::addc_cc(s[i] ^ sd, 0);
  return sd ^ mf;
}

// r = (u f + v g) / 2^32 mod p, in [0, p).  0 <= u, v <= p; |f|, |g| <= 2^30.
LW_INL void cofactor_update_inl(uint32_t* r, const uint32_t* u, uint32_t f, const uint32_t* v, uint32_t g) {
  const uint32_t mf = (uint32_t)((int32_t)f >> 31), mg = (uint32_t)((int32_t)g >> 31);
  const uint32_t af = (f ^ mf) - mf, ag = (g ^ mg) - mg;
  const uint32_t* mod = k::FP_MOD;
  uint32_t uu[12], vv[12], t[12];
  limbs_sub<12>(t, mod, u);  // -u mod p (p itself when u == 0: still == 0 mod p and within the bounds)
#pragma unroll
  for (int i = 0; i < 12; i++) uu[i] = mf ? t[i] : u[i];
  limbs_sub<12>(t, mod, v);
#pragma unroll
  for (int i = 0; i < 12; i++) vv[i] = mg ? t[i] : v[i];
  uint32_t X[13], Y[13], Z[13];
  mul_word(X, uu, af);
  mul_word(Y, vv, ag);
  X[0] = ptx::add_cc(X[0], Y[0]);
#pragma unroll
  for (int i = 1; i < 12; i++) X[i] = ptx::addc_cc(X[i], Y[i]);
  X[12] = ptx::addc(X[12], Y[12]);  // < p 2^31
  const uint32_t q = X[0] * k::FP_INV;
  mul_word(Z, mod, q);
  X[0] = ptx::add_cc(X[0], Z[0]);  // == 0
#pragma unroll
  for (int i = 1; i < 12; i++) X[i] = ptx::addc_cc(X[i], Z[i]);
  X[12] = ptx::addc(X[12], Z[12]);  // (p 2^31 + p 2^32) / 2^32 < 1.5 p after dropping limb 0
  const uint32_t borrow = limbs_sub<12>(t, X + 1, mod);
#pragma unroll
  for (int i = 0; i < 12; i++) r[i] = borrow ? X[i + 1] : t[i];
}

// by-value wrappers: operands and results travel in registers (pointer arguments would force the
// caller's arrays into local memory)
struct FpSigned { Fp v; uint32_t neg; };
LW_COLD FpSigned lincomb_shr(Fp a, uint32_t f, Fp b, uint32_t g) {
  FpSigned r;
  r.neg = lincomb_shr_inl(r.v.l, a.l, f, b.l, g);
  return r;
}
LW_COLD Fp cofactor_update(Fp u, uint32_t f, Fp v, uint32_t g) {
  Fp r;
  cofactor_update_inl(r.l, u.l, f, v.l, g);
  return r;
}

}  // namespace gcdinv

// Montgomery form in, Montgomery form out; 0 -> 0.
LW_INL Fp fp_inv_gcd(const Fp& y) {
  using namespace gcdinv;
  Fp A = y, B, U = fp_zero(), V = fp_zero();
#pragma unroll
  for (int i = 0; i < 12; i++) B.l[i] = k::FP_MOD[i];
  U.l[0] = 1;
  uint32_t *a = A.l, *b = B.l;
  int rounds = 0;
#pragma unroll 1
  for (; rounds < MAX_ROUNDS; rounds++) {
    if (warp_all(limbs_is_zero<12>(a))) break;
    // ---- 62-bit approximations: the two limbs below the common leading limb of max(a, b)
    uint32_t ah = 0, al = 0, bh = 0, bl = 0, found = 0, at1 = 0;
#pragma unroll
    for (int i = 11; i >= 1; i--) {
      const uint32_t nz = (a[i] | b[i]) != 0 ? ~0u : 0u;
      const uint32_t take = nz & ~found;
      ah = (a[i] & take) | (ah & ~take);
      al = (a[i - 1] & take) | (al & ~take);
      bh = (b[i] & take) | (bh & ~take);
      bl = (b[i - 1] & take) | (bl & ~take);
      if (i == 1) at1 = take;
      found |= nz;
    }
    const int c = clz32(ah | bh);
    // bit length n = 32 h + 32 - c; below 62 bits the values are used exactly
    const bool exact = !found || (at1 && c >= 3);
    const uint32_t ta = shl_pair_hi(al, ah, c & 31), tb = shl_pair_hi(bl, bh, c & 31);
    uint64_t abar = exact ? (((uint64_t)a[1] << 32) | a[0]) : (((uint64_t)ta << 30) | (a[0] & 0x3fffffffu));
    uint64_t bbar = exact ? (((uint64_t)b[1] << 32) | b[0]) : (((uint64_t)tb << 30) | (b[0] & 0x3fffffffu));
    // ---- 30 divsteps on the approximations
    uint32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 2
    for (int j = 0; j < STEPS; j++) {
      const uint64_t odd = 0ull - (abar & 1ull);
      const uint64_t sw = odd & (abar < bbar ? ~0ull : 0ull);
      const uint64_t tx = (abar ^ bbar) & sw;
      abar ^= tx; bbar ^= tx;
      const uint32_t sw32 = (uint32_t)sw, odd32 = (uint32_t)odd;
      const uint32_t tf = (f0 ^ f1) & sw32, tg = (g0 ^ g1) & sw32;
      f0 ^= tf; f1 ^= tf; g0 ^= tg; g1 ^= tg;
      abar -= bbar & odd;
      f0 -= f1 & odd32;
      g0 -= g1 & odd32;
      abar >>= 1;
      f1 <<= 1;
      g1 <<= 1;
    }
    // ---- apply the transition matrix
    const FpSigned na = lincomb_shr(A, f0, B, g0);
    const FpSigned nb = lincomb_shr(A, f1, B, g1);
    const uint32_t sa = na.neg, sb = nb.neg;
    f0 = (f0 ^ sa) - sa; g0 = (g0 ^ sa) - sa;  // a < 0: (a, f0, g0) <- (-a, -f0, -g0)
    f1 = (f1 ^ sb) - sb; g1 = (g1 ^ sb) - sb;
    const Fp nu = cofactor_update(U, f0, V, g0);
    const Fp nv = cofactor_update(U, f1, V, g1);
    A = na.v; B = nb.v; U = nu; V = nv;
  }
  // y^-1 = 4^rounds v as integers; Montgomery form of the inverse = y^-1 R^2 = mont_mul(v, 4^rounds R^3)
  Fp kk;
#pragma unroll
  for (int i = 0; i < 12; i++) kk.l[i] = k::FP_GCDINV_SCALE[rounds][i];
  return fp_mul_nv(V, kk);
}

// The inversion every cold caller uses (to_affine, Fp2 inverse, table build): same value as fp_inv_fermat.
LW_COLD Fp fp_inv(const Fp& a) { return fp_inv_gcd(a); }

}  // namespace lw
