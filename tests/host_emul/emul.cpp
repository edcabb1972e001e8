// Host emulation build of the DEVICE math headers (test-only).
//
// Compiled with g++ -DLWKZG_HOST_EMUL: the PTX carry-chain primitives of
// csrc/ptx.cuh are replaced by a bit-exact software model, everything above
// them (Montgomery arithmetic, point formulas, codecs, SHA-256, pairing) is the
// very same source the CUDA kernels compile.  This lets the CPU-only test tier
// check the device algorithms against the Python oracle without a GPU.  It is
// NOT a fallback: the shipped library has no host compute path.
#include <cstring>
#include "../../lambdaworks_kzg_b200/csrc/g1.cuh"
#include "../../lambdaworks_kzg_b200/csrc/fpinv.cuh"
#include "../../tools/experiments/fpdp.cuh"
#include "../../tools/experiments/karatsuba.cuh"
#include "../../lambdaworks_kzg_b200/csrc/sha256.cuh"
#include "../../lambdaworks_kzg_b200/csrc/frpoly.cuh"
#include "../../lambdaworks_kzg_b200/csrc/recode.cuh"
#ifdef LWKZG_EMUL_PAIRING
#include "../../lambdaworks_kzg_b200/csrc/pairing.cuh"
#include "../../lambdaworks_kzg_b200/csrc/pairing_warp.cuh"
#endif

using namespace lw;

extern "C" {

void emul_fp_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) { mont_mul<FpCfg>(r, a, b); }
void emul_fp_mul_k(uint32_t* r, const uint32_t* a, const uint32_t* b) { mont_mul_karatsuba<FpCfg>(r, a, b); }
void emul_fr_mul_k(uint32_t* r, const uint32_t* a, const uint32_t* b) { mont_mul_karatsuba<FrCfg>(r, a, b); }
void emul_fp_sqr(uint32_t* r, const uint32_t* a) { mont_sqr<FpCfg>(r, a); }
void emul_fp_mul_dp(uint32_t* r, const uint32_t* a, const uint32_t* b) { dp::mont_mul(r, a, b); }
void emul_fp_sqr_dp(uint32_t* r, const uint32_t* a) { dp::mont_sqr(r, a); }
void emul_fr_sqr(uint32_t* r, const uint32_t* a) { mont_sqr<FrCfg>(r, a); }
void emul_fp_add(uint32_t* r, const uint32_t* a, const uint32_t* b) { mod_add<FpCfg>(r, a, b); }
void emul_fp_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) { mod_sub<FpCfg>(r, a, b); }
void emul_fp_neg(uint32_t* r, const uint32_t* a) { mod_neg<FpCfg>(r, a); }
void emul_fr_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) { mont_mul<FrCfg>(r, a, b); }
void emul_fr_add(uint32_t* r, const uint32_t* a, const uint32_t* b) { mod_add<FrCfg>(r, a, b); }
void emul_fr_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) { mod_sub<FrCfg>(r, a, b); }
void emul_fp_to_mont(uint32_t* r, const uint32_t* a) { Fp x; memcpy(x.l, a, 48); x = fp_to_mont(x); memcpy(r, x.l, 48); }
void emul_fp_from_mont(uint32_t* r, const uint32_t* a) { Fp x; memcpy(x.l, a, 48); x = fp_from_mont(x); memcpy(r, x.l, 48); }
void emul_fp_inv(uint32_t* r, const uint32_t* a) { Fp x; memcpy(x.l, a, 48); x = fp_inv(x); memcpy(r, x.l, 48); }
void emul_fp_inv_fermat(uint32_t* r, const uint32_t* a) { Fp x; memcpy(x.l, a, 48); x = fp_inv_fermat(x); memcpy(r, x.l, 48); }
void emul_fp_inv_gcd(uint32_t* r, const uint32_t* a) { Fp x; memcpy(x.l, a, 48); x = fp_inv_gcd(x); memcpy(r, x.l, 48); }
void emul_fr_inv(uint32_t* r, const uint32_t* a) { Fr x; memcpy(x.l, a, 32); x = fr_inv(x); memcpy(r, x.l, 32); }
void emul_fr_from_be32(uint32_t* r, const uint8_t* b) { Fr x = fr_canon_from_be32(b); memcpy(r, x.l, 32); }
void emul_fr_from_be_words(uint32_t* r, const uint8_t* b) { uint32_t w[8]; memcpy(w, b, 32); Fr x = fr_canon_from_be_words(w); memcpy(r, x.l, 32); }

// points are passed as canonical big-endian affine (x||y, 96 bytes; all-zero = infinity)
static G1Affine load_aff(const uint8_t* b) {
  G1Affine p;
  bool z = true;
  for (int i = 0; i < 96; i++) z = z && b[i] == 0;
  if (z) return g1a_inf();
  p.x = fp_from_be48(b);
  p.y = fp_from_be48(b + 48);
  return p;
}
static void store_aff(uint8_t* b, const G1Affine& p) {
  if (g1a_is_inf(p)) { memset(b, 0, 96); return; }
  fp_canon_to_be48(b, fp_from_mont(p.x));
  fp_canon_to_be48(b + 48, fp_from_mont(p.y));
}

// op: 0 = madd (a XYZZ-from-affine += b), 1 = add (xyzz+xyzz), 2 = dbl(a), 3 = a + b via scaled XYZZ
void emul_g1_op(uint8_t* out, const uint8_t* a, const uint8_t* b, int op) {
  G1Affine A = load_aff(a), B = load_aff(b);
  G1Xyzz acc = xyzz_from_affine(A);
  if (op == 0) {
    xyzz_madd(acc, B);
  } else if (op == 1) {
    G1Xyzz bb = xyzz_from_affine(B);
    xyzz_add(acc, bb);
  } else if (op == 2) {
    acc = xyzz_dbl(acc);
  } else {
    // exercise non-trivial ZZ/ZZZ: (A + A) - A + B etc.
    G1Xyzz a2 = xyzz_dbl(acc);      // 2A, zz != 1
    G1Xyzz t = a2;
    xyzz_madd(t, g1a_neg(A));       // A with non-trivial zz
    G1Xyzz b2 = xyzz_dbl(xyzz_from_affine(B));
    xyzz_madd(b2, g1a_neg(B));      // B with non-trivial zz
    xyzz_add(t, b2);
    acc = t;
  }
  store_aff(out, xyzz_to_affine(acc));
}

void emul_g1_mul(uint8_t* out, const uint8_t* a, const uint32_t* k8) {
  G1Affine A = load_aff(a);
  store_aff(out, xyzz_to_affine(g1_mul_scalar(A, k8, 8)));
}
void emul_g1_mul_glv(uint8_t* out, const uint8_t* a, const uint32_t* k8) {
  G1Affine A = load_aff(a);
  store_aff(out, xyzz_to_affine(g1_mul_scalar_glv(A, k8)));
}
int emul_g1_on_curve(const uint8_t* a) { return g1a_on_curve(load_aff(a)) ? 1 : 0; }
int emul_g1_in_subgroup(const uint8_t* a) { return g1_in_subgroup(load_aff(a)) ? 1 : 0; }
void emul_g1_compress(uint8_t* out48, const uint8_t* a) { g1_compress(out48, load_aff(a)); }
int emul_g1_decompress(uint8_t* out96, const uint8_t* in48) {
  G1Affine p;
  if (!g1_decompress(p, in48)) return 0;
  store_aff(out96, p);
  return 1;
}

void emul_sha256(uint8_t* out32, const uint8_t* msg, size_t len) { sha256_oneshot(out32, msg, len); }

// GLV split + signed-digit recoding of both halves: q4, m4 and digits[2][W] as int32 (half 0 = m, half 1 = q)
int emul_glv_windows(int c) { return glv_num_windows(c); }
uint32_t emul_glv_window_count(int c, int j) { return glv_window_count(c, j); }
void emul_glv_recode(uint32_t* q4, uint32_t* m4, int32_t* digits, const uint32_t* k8, int c) {
  glv_split_barrett(q4, m4, k8);
  const int W = glv_num_windows(c);
  for (int h = 0; h < 2; h++) {
    const uint32_t* v = h ? q4 : m4;
    for (int j = 0, carry = 0; j < W; j++) digits[h * W + j] = glv_digit([&](int w) { return v[w]; }, c, W, j, carry);
  }
}

// Horner + quotient over n canonical coefficients (n*8 u32) at z (canonical)
void emul_poly_eval_quot(uint32_t* y8, uint32_t* q, const uint32_t* coeffs, int n, const uint32_t* z8, int chunks) {
  poly_eval_quot_reference_order(y8, q, coeffs, n, z8, chunks);
}

#ifdef LWKZG_EMUL_PAIRING
// pairing check e(a1, q1) * e(a2, q2) == 1 ; G2 points canonical BE: x.c0,x.c1,y.c0,y.c1 (4*48)
int emul_pairing_check(const uint8_t* a1, const uint8_t* q1, const uint8_t* a2, const uint8_t* q2) {
  G1Affine P[2] = {load_aff(a1), load_aff(a2)};
  G2Affine Q[2];
  const uint8_t* qs[2] = {q1, q2};
  for (int i = 0; i < 2; i++) {
    Q[i].x.c0 = fp_from_be48(qs[i]);
    Q[i].x.c1 = fp_from_be48(qs[i] + 48);
    Q[i].y.c0 = fp_from_be48(qs[i] + 96);
    Q[i].y.c1 = fp_from_be48(qs[i] + 144);
  }
  G2Prepared prep[2];
  g2_prepare(prep[0], Q[0]);
  g2_prepare(prep[1], Q[1]);
  return pairing_product_is_one(P, prep, 2) ? 1 : 0;
}
int emul_g2_decompress(uint8_t* out192, const uint8_t* in96) {
  G2Affine q;
  bool inf;
  if (!g2_decompress(q, inf, in96)) return 0;
  if (inf) { memset(out192, 0, 192); return 2; }
  fp_canon_to_be48(out192, fp_from_mont(q.x.c0));
  fp_canon_to_be48(out192 + 48, fp_from_mont(q.x.c1));
  fp_canon_to_be48(out192 + 96, fp_from_mont(q.y.c0));
  fp_canon_to_be48(out192 + 144, fp_from_mont(q.y.c1));
  return 1;
}
// cyclotomic squaring == generic squaring on the cyclotomic subgroup (after the easy part)
int emul_cyclotomic_sqr_check(const uint8_t* a1, const uint8_t* q1) {
  G1Affine P = load_aff(a1);
  G2Affine Q;
  Q.x.c0 = fp_from_be48(q1); Q.x.c1 = fp_from_be48(q1 + 48);
  Q.y.c0 = fp_from_be48(q1 + 96); Q.y.c1 = fp_from_be48(q1 + 144);
  G2Prepared prep;
  g2_prepare(prep, Q);
  Fp12 f = miller_loop(&P, &prep, 1);
  Fp12 g = fp12_mul(fp12_conj(f), fp12_inv(f));
  g = fp12_mul(fp12_frobenius(fp12_frobenius(g)), g);
  return fp12_eq(fp12_cyclotomic_sqr(g), fp12_sqr(g)) && !fp12_eq(fp12_cyclotomic_sqr(f), fp12_sqr(f)) ? 1 : 0;
}
// warp-cooperative pairing (pairing_warp.cuh), lanes emulated sequentially:
// returns 1 iff (a) a random-ish Fp12 product matches the single-thread product and
// (b) the two-pair warp pairing result is bit-identical to the single-thread one
int emul_warp_pairing_check(const uint8_t* a1, const uint8_t* q1, const uint8_t* a2, const uint8_t* q2, int* is_one) {
  G1Affine P[2] = {load_aff(a1), load_aff(a2)};
  G2Affine Q[2];
  const uint8_t* qs[2] = {q1, q2};
  for (int i = 0; i < 2; i++) {
    Q[i].x.c0 = fp_from_be48(qs[i]); Q[i].x.c1 = fp_from_be48(qs[i] + 48);
    Q[i].y.c0 = fp_from_be48(qs[i] + 96); Q[i].y.c1 = fp_from_be48(qs[i] + 144);
  }
  static G2Prepared prep[2];
  g2_prepare(prep[0], Q[0]);
  g2_prepare(prep[1], Q[1]);
  Fp12 f1 = miller_loop(&P[0], &prep[0], 1), f2 = miller_loop(&P[1], &prep[1], 1);
  // (a) multiplication
  static WarpPairingMem m;
  m.a = *reinterpret_cast<WarpFp12*>(&f1);
  m.b = *reinterpret_cast<WarpFp12*>(&f2);
  wfp12_mul(m.c, m.a, m.b, m.sc);
  Fp12 want = fp12_mul(f1, f2);
  if (!fp12_eq(*reinterpret_cast<Fp12*>(&m.c), want)) return 0;
  wfp12_mul(m.a, m.a, m.a, m.sc);  // aliasing
  Fp12 sq = fp12_sqr(f1);
  if (!fp12_eq(*reinterpret_cast<Fp12*>(&m.a), sq)) return 0;
  wfp12_frobenius(m.t, m.c);
  Fp12 fr = fp12_frobenius(want);
  if (!fp12_eq(*reinterpret_cast<Fp12*>(&m.t), fr)) return 0;
  // (b) full pairing product
  const G2Prepared* pq[2] = {&prep[0], &prep[1]};
  warp_pairing_product(m, P, pq, 2);
  Fp12 ref = final_exponentiation(miller_loop(P, prep, 2));
  if (!fp12_eq(*reinterpret_cast<Fp12*>(&m.f), ref)) return 0;
  *is_one = warp_fp12_is_one(m.f) ? 1 : 0;
  return 1;
}
// raw Fp12 pairing output for debugging / cross-checking with oracle/py/pairing.py
void emul_pairing_gt(uint8_t* out576, const uint8_t* a1, const uint8_t* q1) {
  G1Affine P = load_aff(a1);
  G2Affine Q;
  Q.x.c0 = fp_from_be48(q1); Q.x.c1 = fp_from_be48(q1 + 48);
  Q.y.c0 = fp_from_be48(q1 + 96); Q.y.c1 = fp_from_be48(q1 + 144);
  G2Prepared prep;
  g2_prepare(prep, Q);
  Fp12 f = miller_loop(&P, &prep, 1);
  f = final_exponentiation(f);
  const Fp* c = reinterpret_cast<const Fp*>(&f);
  for (int i = 0; i < 12; i++) fp_canon_to_be48(out576 + 48 * i, fp_from_mont(c[i]));
}

#endif

}  // extern "C"
