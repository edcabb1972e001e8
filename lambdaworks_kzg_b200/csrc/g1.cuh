// BLS12-381 G1 (y^2 = x^3 + 4) point arithmetic.
//
// Replaces lambdaworks-math's ShortWeierstrassProjectivePoint<BLS12381Curve>
// (operate_with / operate_with_self / neg / to_affine; call sites
// /root/reference/src/lib.rs:36-37, 241-243, 664-688, src/compression.rs:22-27).
// The group element computed is the same whatever the coordinate system, so we
// use the cheapest ones for a GPU:
//   * affine (x, y) for stored bases / table entries (96 B); infinity = (0, 0),
//     which is not on the curve and therefore unambiguous;
//   * XYZZ (X, Y, ZZ, ZZZ), x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, for accumulators
//     (mixed add 8M+2S, add 12M+2S, double 6M+3S; EFD "xyzz", a = 0).
//     Infinity <=> ZZ == 0.
// Every formula handles the exceptional cases (P = Q, P = -Q, infinity) so the
// result is exact for arbitrary inputs -- needed for bit-parity on edge blobs.
#pragma once
#include "field.cuh"

namespace lw {

struct G1Affine {
  Fp x, y;
};
struct G1Xyzz {
  Fp x, y, zz, zzz;
};

LW_INL bool g1a_is_inf(const G1Affine& p) { return fp_is_zero(p.x) && fp_is_zero(p.y); }
LW_INL G1Affine g1a_inf() { G1Affine r; r.x = fp_zero(); r.y = fp_zero(); return r; }
LW_INL G1Affine g1a_neg(const G1Affine& p) { G1Affine r; r.x = p.x; r.y = fp_neg(p.y); return r; }
LW_INL G1Affine g1a_cneg(const G1Affine& p, bool neg) { G1Affine r; r.x = p.x; r.y = fp_cneg(p.y, neg); return r; }
LW_INL G1Affine g1a_generator() {
  G1Affine g;
  for (int i = 0; i < 12; i++) { g.x.l[i] = k::G1_GEN_X[i]; g.y.l[i] = k::G1_GEN_Y[i]; }
  return g;
}
// y^2 == x^3 + 4 (infinity (0,0) is NOT on the curve: mirrors from_affine in
// /root/reference/src/srs.rs:155-172)
LW_INL bool g1a_on_curve(const G1Affine& p) {
  Fp b; for (int i = 0; i < 12; i++) b.l[i] = k::FP_B[i];
  Fp lhs = fp_sqr(p.y);
  Fp rhs = fp_add(fp_mul(fp_sqr(p.x), p.x), b);
  return fp_eq(lhs, rhs);
}

LW_INL bool xyzz_is_inf(const G1Xyzz& p) { return fp_is_zero(p.zz); }
LW_INL G1Xyzz xyzz_inf() { G1Xyzz r; r.x = fp_zero(); r.y = fp_zero(); r.zz = fp_zero(); r.zzz = fp_zero(); return r; }
LW_INL G1Xyzz xyzz_from_affine(const G1Affine& p) {
  G1Xyzz r;
  if (g1a_is_inf(p)) return xyzz_inf();
  r.x = p.x; r.y = p.y; r.zz = fp_one(); r.zzz = fp_one();
  return r;
}
LW_INL G1Xyzz xyzz_neg(const G1Xyzz& p) { G1Xyzz r = p; r.y = fp_neg(p.y); return r; }

// 2 * (affine p), p != infinity
LW_INL G1Xyzz xyzz_dbl_affine(const G1Affine& p) {
  G1Xyzz r;
  Fp U = fp_dbl(p.y);
  Fp V = fp_sqr(U);
  Fp W = fp_mul(U, V);
  Fp S = fp_mul(p.x, V);
  Fp X2 = fp_sqr(p.x);
  Fp M = fp_add(fp_dbl(X2), X2);
  r.x = fp_sub(fp_sqr(M), fp_dbl(S));
  r.y = fp_sub(fp_mul(M, fp_sub(S, r.x)), fp_mul(W, p.y));
  r.zz = V;
  r.zzz = W;
  return r;
}

LW_INL G1Xyzz xyzz_dbl(const G1Xyzz& p) {
  if (xyzz_is_inf(p)) return p;
  G1Xyzz r;
  Fp U = fp_dbl(p.y);
  Fp V = fp_sqr(U);
  Fp W = fp_mul(U, V);
  Fp S = fp_mul(p.x, V);
  Fp X2 = fp_sqr(p.x);
  Fp M = fp_add(fp_dbl(X2), X2);
  r.x = fp_sub(fp_sqr(M), fp_dbl(S));
  r.y = fp_sub(fp_mul(M, fp_sub(S, r.x)), fp_mul(W, p.y));
  r.zz = fp_mul(V, p.zz);
  r.zzz = fp_mul(W, p.zzz);
  return r;  // y == 0 (2-torsion) gives zz = 0 = infinity, as it should
}

// acc += p   (p affine)
// Single exit on purpose (also in xyzz_add below): these bodies end up inside out-of-line
// functions, and ptxas 12.9 mis-allocates uniform registers around a divergent early
// `return` (the struct-copy loop of one path reused the uniform register that held an
// operand address of the other path -> illegal local reads; found with compute-sanitizer,
// profiles/r01_sanitizer.md).  Structured if/else keeps the paths inside one
// convergence region.
LW_INL void xyzz_madd(G1Xyzz& acc, const G1Affine& p) {
  if (!g1a_is_inf(p)) {
    if (xyzz_is_inf(acc)) {
      acc.x = p.x; acc.y = p.y; acc.zz = fp_one(); acc.zzz = fp_one();
    } else {
      Fp U2 = fp_mul(p.x, acc.zz);
      Fp S2 = fp_mul(p.y, acc.zzz);
      Fp Pd = fp_sub(U2, acc.x);
      Fp Rd = fp_sub(S2, acc.y);
      if (fp_is_zero(Pd)) {
        if (fp_is_zero(Rd)) acc = xyzz_dbl_affine(p);
        else acc = xyzz_inf();
      } else {
        Fp PP = fp_sqr(Pd);
        Fp PPP = fp_mul(Pd, PP);
        Fp Q = fp_mul(acc.x, PP);
        Fp X3 = fp_sub(fp_sub(fp_sqr(Rd), PPP), fp_dbl(Q));
        Fp Y3 = fp_sub(fp_mul(Rd, fp_sub(Q, X3)), fp_mul(acc.y, PPP));
        acc.zz = fp_mul(acc.zz, PP);
        acc.zzz = fp_mul(acc.zzz, PPP);
        acc.x = X3;
        acc.y = Y3;
      }
    }
  }
}

// a += b   (both XYZZ)
LW_INL void xyzz_add(G1Xyzz& a, const G1Xyzz& b) {
  if (!xyzz_is_inf(b)) {
    if (xyzz_is_inf(a)) {
#pragma unroll
      for (int i = 0; i < 12; i++) { a.x.l[i] = b.x.l[i]; a.y.l[i] = b.y.l[i]; a.zz.l[i] = b.zz.l[i]; a.zzz.l[i] = b.zzz.l[i]; }
    } else {
      Fp U1 = fp_mul(a.x, b.zz);
      Fp U2 = fp_mul(b.x, a.zz);
      Fp S1 = fp_mul(a.y, b.zzz);
      Fp S2 = fp_mul(b.y, a.zzz);
      Fp Pd = fp_sub(U2, U1);
      Fp Rd = fp_sub(S2, S1);
      if (fp_is_zero(Pd)) {
        if (fp_is_zero(Rd)) a = xyzz_dbl(a);
        else a = xyzz_inf();
      } else {
        Fp PP = fp_sqr(Pd);
        Fp PPP = fp_mul(Pd, PP);
        Fp Q = fp_mul(U1, PP);
        Fp X3 = fp_sub(fp_sub(fp_sqr(Rd), PPP), fp_dbl(Q));
        Fp Y3 = fp_sub(fp_mul(Rd, fp_sub(Q, X3)), fp_mul(S1, PPP));
        a.zz = fp_mul(fp_mul(a.zz, b.zz), PP);
        a.zzz = fp_mul(fp_mul(a.zzz, b.zzz), PPP);
        a.x = X3;
        a.y = Y3;
      }
    }
  }
}

// Hot-loop variant of the mixed addition.  Identical arithmetic, but every field
// multiplication is a CALL to one out-of-line multiplier whose operands and
// result travel in registers (by-value structs): the loop body shrinks from
// ~69 KB to ~12 KB of SASS, which removes the instruction-cache misses that ncu
// showed as 19 % stall_no_inst (profiles/r01_ncu_msm_summary.md), and the rare
// equal-x cases go through the out-of-line generic formulas.
LW_COLD void xyzz_madd_rare(G1Xyzz& acc, const G1Affine& p) { xyzz_madd(acc, p); }
LW_INL void xyzz_madd_hot(G1Xyzz& acc, const G1Affine& p) {
  if (!g1a_is_inf(p)) {
    if (xyzz_is_inf(acc)) {
      acc.x = p.x; acc.y = p.y; acc.zz = fp_one(); acc.zzz = fp_one();
    } else {
      Fp U2 = fp_mul_nv(p.x, acc.zz);
      Fp S2 = fp_mul_nv(p.y, acc.zzz);
      Fp Pd = fp_sub(U2, acc.x);
      Fp Rd = fp_sub(S2, acc.y);
      if (fp_is_zero(Pd)) {  // same x: doubling or cancellation
        xyzz_madd_rare(acc, p);
      } else {
        Fp PP = fp_sqr_nv(Pd);
        Fp PPP = fp_mul_nv(Pd, PP);
        Fp Q = fp_mul_nv(acc.x, PP);
        Fp X3 = fp_sub(fp_sub(fp_sqr_nv(Rd), PPP), fp_dbl(Q));
        Fp Y3 = fp_sub(fp_mul_nv(Rd, fp_sub(Q, X3)), fp_mul_nv(acc.y, PPP));
        acc.zz = fp_mul_nv(acc.zz, PP);
        acc.zzz = fp_mul_nv(acc.zzz, PPP);
        acc.x = X3;
        acc.y = Y3;
      }
    }
  }
}

// Out-of-line copies of the group law for cold callers (scalar-mul ladders,
// verification, setup): the hot MSM loop keeps the force-inlined versions.
// They are COMPACT as well: every field multiplication is a call to the one out-of-line multiplier, so each of
// these is a few hundred bytes of SASS instead of 40-60 KB of inlined Montgomery products (the same formulas, the
// same results).  The cold kernels run one warp per block all over the GPU, next to the hash kernels of the
// pipelines; their instruction footprint is what they compete with.
LW_COLD void xyzz_dbl_ni(G1Xyzz& p) {
  if (!xyzz_is_inf(p)) {
    const Fp U = fp_dbl(p.y), V = fp_sqr_nv(U), W = fp_mul_nv(U, V), S = fp_mul_nv(p.x, V), X2 = fp_sqr_nv(p.x);
    const Fp M = fp_add(fp_dbl(X2), X2);
    const Fp X3 = fp_sub(fp_sqr_nv(M), fp_dbl(S));
    p.y = fp_sub(fp_mul_nv(M, fp_sub(S, X3)), fp_mul_nv(W, p.y));
    p.x = X3;
    p.zz = fp_mul_nv(V, p.zz);      // y == 0 (2-torsion) gives zz = 0 = infinity, as it should
    p.zzz = fp_mul_nv(W, p.zzz);
  }
}
LW_COLD void xyzz_madd_ni(G1Xyzz& acc, const G1Affine& p) {
  if (!g1a_is_inf(p)) {
    if (xyzz_is_inf(acc)) {
      acc.x = p.x; acc.y = p.y; acc.zz = fp_one(); acc.zzz = fp_one();
    } else {
      const Fp Pd = fp_sub(fp_mul_nv(p.x, acc.zz), acc.x);
      const Fp Rd = fp_sub(fp_mul_nv(p.y, acc.zzz), acc.y);
      if (fp_is_zero(Pd)) {
        if (fp_is_zero(Rd)) {
          acc.x = p.x; acc.y = p.y; acc.zz = fp_one(); acc.zzz = fp_one();
          xyzz_dbl_ni(acc);
        } else {
          acc = xyzz_inf();
        }
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
}
LW_COLD void xyzz_add_ni(G1Xyzz& a, const G1Xyzz& b) {
  if (!xyzz_is_inf(b)) {
    if (xyzz_is_inf(a)) {
      // limb by limb, unrolled: ptxas 12.9 rolls a struct copy into a loop whose counter lives in a uniform register
      // that the other side of this divergent branch uses for an operand address
#pragma unroll
      for (int i = 0; i < 12; i++) { a.x.l[i] = b.x.l[i]; a.y.l[i] = b.y.l[i]; a.zz.l[i] = b.zz.l[i]; a.zzz.l[i] = b.zzz.l[i]; }
    } else {
      const Fp U1 = fp_mul_nv(a.x, b.zz), S1 = fp_mul_nv(a.y, b.zzz);
      const Fp Pd = fp_sub(fp_mul_nv(b.x, a.zz), U1), Rd = fp_sub(fp_mul_nv(b.y, a.zzz), S1);
      if (fp_is_zero(Pd)) {
        if (fp_is_zero(Rd)) xyzz_dbl_ni(a);
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

// XYZZ -> affine (one inversion); infinity -> (0,0)
LW_COLD G1Affine xyzz_to_affine(const G1Xyzz& p) {
  if (xyzz_is_inf(p)) return g1a_inf();
  Fp zzz_inv = fp_inv(p.zzz);
  Fp t = fp_mul(p.zz, zzz_inv);  // ZZ/ZZZ = 1/sqrt(ZZ)
  Fp zz_inv = fp_sqr(t);         // ZZ^2/ZZZ^2 = 1/ZZ
  G1Affine r;
  r.x = fp_mul(p.x, zz_inv);
  r.y = fp_mul(p.y, zzz_inv);
  return r;
}

// a == b as group elements
LW_INL bool xyzz_eq(const G1Xyzz& a, const G1Xyzz& b) {
  bool ia = xyzz_is_inf(a), ib = xyzz_is_inf(b);
  if (ia || ib) return ia && ib;
  return fp_eq(fp_mul(a.x, b.zz), fp_mul(b.x, a.zz)) && fp_eq(fp_mul(a.y, b.zzz), fp_mul(b.y, a.zzz));
}

// [k]p for a canonical 256-bit little-endian scalar (8 x u32), MSB-first
// double-and-add.  Cold path (verification, table seeding).
LW_COLD G1Xyzz g1_mul_scalar(const G1Affine& p, const uint32_t* kk, int nlimbs) {
  G1Xyzz acc = xyzz_inf();
  bool started = false;
  for (int w = nlimbs - 1; w >= 0; w--) {
    uint32_t word = kk[w];
    for (int bit = 31; bit >= 0; bit--) {
      if (started) xyzz_dbl_ni(acc);
      if ((word >> bit) & 1u) { xyzz_madd_ni(acc, p); started = true; }
    }
  }
  return acc;
}

// [k]p with the GLV split of BLS12-381: r = x^4 - x^2 + 1 and phi(P) = (beta x, y) = [-x^2]P, so with
// k = q x^2 + m (m < x^2 < 2^128, q < 2^128 for every k < r) one gets [k]P = [m]P + [q](beta x, -y): 128
// doublings and two conditional mixed additions per bit instead of 255 doublings (batched verification's
// r^i-multiples, /root/reference/src/lib.rs:651-685).  k: canonical 256-bit little-endian, < r.
LW_COLD void glv_split(uint32_t* q4, uint32_t* m4, const uint32_t* k8) {
  uint32_t rem[5] = {0, 0, 0, 0, 0};
  uint32_t q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int bit = 255; bit >= 0; bit--) {
    // rem = 2 rem + bit
    for (int i = 4; i > 0; i--) rem[i] = (rem[i] << 1) | (rem[i - 1] >> 31);
    rem[0] = (rem[0] << 1) | ((k8[bit >> 5] >> (bit & 31)) & 1u);
    // rem >= x^2 ?
    uint32_t t[5];
    uint32_t borrow = 0;
    for (int i = 0; i < 5; i++) {
      const uint32_t s = i < 4 ? k::BLS_X2[i] : 0u;
      const uint64_t d = (uint64_t)rem[i] - s - borrow;
      t[i] = (uint32_t)d;
      borrow = (uint32_t)(d >> 63);
    }
    if (!borrow) {
      for (int i = 0; i < 5; i++) rem[i] = t[i];
      q[bit >> 5] |= 1u << (bit & 31);
    }
  }
  for (int i = 0; i < 4; i++) { q4[i] = q[i]; m4[i] = rem[i]; }
}
LW_COLD G1Xyzz g1_mul_scalar_glv(const G1Affine& p, const uint32_t* k8) {
  uint32_t q[4], m[4];
  glv_split(q, m, k8);
  G1Affine p2;
  Fp beta;
  for (int i = 0; i < 12; i++) beta.l[i] = k::FP_BETA[i];
  p2.x = fp_mul(p.x, beta);
  p2.y = fp_neg(p.y);
  G1Xyzz acc = xyzz_inf();
  if (g1a_is_inf(p)) return acc;
  bool started = false;
  for (int w = 3; w >= 0; w--) {
    const uint32_t wm = m[w], wq = q[w];
    for (int bit = 31; bit >= 0; bit--) {
      if (started) xyzz_dbl_ni(acc);
      if ((wm >> bit) & 1u) { xyzz_madd_ni(acc, p); started = true; }
      if ((wq >> bit) & 1u) { xyzz_madd_ni(acc, p2); started = true; }
    }
  }
  return acc;
}

// ---------------------------------------------------------------- codecs
// /root/reference/src/compression.rs:33-60
LW_COLD void g1_compress(uint8_t* out48, const G1Affine& p) {
  if (g1a_is_inf(p)) {
    out48[0] = 0xC0;
    for (int i = 1; i < 48; i++) out48[i] = 0;
    return;
  }
  Fp xc = fp_from_mont(p.x);
  Fp yc = fp_from_mont(p.y);
  fp_canon_to_be48(out48, xc);
  out48[0] |= 0x80;
  if (fp_canon_is_lex_large(yc)) out48[0] |= 0x20;
}

// Subgroup test with the same accept set as the reference's [r]P == O
// (src/compression.rs:22-27): for BLS12-381, P in G1 <=> phi(P) == [-x^2]P
// where phi(x,y) = (beta x, y)  (Scott, "A note on group membership tests for
// G1, G2 and GT on BLS pairing-friendly curves", 2021).  [x^2]P costs two
// 64-bit double-and-add ladders instead of a 255-bit one.
LW_COLD G1Xyzz g1_mul_u64(const G1Xyzz& p, unsigned long long e) {
  G1Xyzz acc = xyzz_inf();
  bool started = false;
  for (int bit = 63; bit >= 0; bit--) {
    if (started) xyzz_dbl_ni(acc);
    if ((e >> bit) & 1ull) { xyzz_add_ni(acc, p); started = true; }
  }
  return acc;
}
LW_COLD bool g1_in_subgroup(const G1Affine& p) {
  if (g1a_is_inf(p)) return true;
  G1Xyzz P = xyzz_from_affine(p);
  G1Xyzz t = g1_mul_u64(g1_mul_u64(P, k::BLS_X_ABS), k::BLS_X_ABS);  // [x^2]P
  Fp beta; for (int i = 0; i < 12; i++) beta.l[i] = k::FP_BETA[i];
  G1Xyzz phi = P;
  phi.x = fp_mul(P.x, beta);
  return xyzz_eq(phi, xyzz_neg(t));
}

// /root/reference/src/compression.rs:62-103.  Returns false
// on any rejection.  *out is affine; infinity -> (0,0).
// check_subgroup = false leaves the r-torsion test to the caller (g1_in_subgroup): batched verification runs it
// beside the blob hashes instead of in front of them
LW_COLD bool g1_decompress(G1Affine& out, const uint8_t* in48, bool check_subgroup = true) {
  uint8_t b0 = in48[0];
  if (!(b0 & 0x80)) return false;
  if (b0 & 0x40) { out = g1a_inf(); return true; }  // remaining bits unchecked, like the reference
  uint8_t tmp[48];
  for (int i = 0; i < 48; i++) tmp[i] = in48[i];
  tmp[0] = b0 & 0x1F;
  Fp x = fp_from_be48(tmp);
  Fp b; for (int i = 0; i < 12; i++) b.l[i] = k::FP_B[i];
  Fp y2 = fp_add(fp_mul(fp_sqr(x), x), b);
  Fp y = fp_sqrt_candidate(y2);
  if (!fp_eq(fp_sqr(y), y2)) return false;
  bool large = fp_canon_is_lex_large(fp_from_mont(y));
  bool want_large = (b0 & 0x20) != 0;
  // y == 0 cannot happen (no 2-torsion); if it did both roots coincide.
  if (large != want_large) y = fp_neg(y);
  out.x = x; out.y = y;
  return check_subgroup ? g1_in_subgroup(out) : true;
}

// Strict (ZCash / blst, as c-kzg-4844 requires) variant for MODE_CKZG_LE:
// infinity must be encoded exactly as c0 00..00 and x must be canonical (< p).
LW_COLD bool g1_decompress_strict(G1Affine& out, const uint8_t* in48, bool check_subgroup = true) {
  uint8_t b0 = in48[0];
  if (!(b0 & 0x80)) return false;
  if (b0 & 0x40) {
    if (b0 != 0xC0) return false;
    for (int i = 1; i < 48; i++) if (in48[i]) return false;
    out = g1a_inf();
    return true;
  }
  // canonical x: big-endian compare with p
  Fp xc;
  for (int i = 0; i < 12; i++) {
    const uint8_t* q = in48 + 44 - 4 * i;
    xc.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
  }
  xc.l[11] &= 0x1FFFFFFFu;
  if (!limbs_lt<12>(xc.l, k::FP_MOD)) return false;
  return g1_decompress(out, in48, check_subgroup);
}

}  // namespace lw
