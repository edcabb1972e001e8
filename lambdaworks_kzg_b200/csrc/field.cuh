// BLS12-381 base field Fp (12 x u32) and scalar field Fr (8 x u32), Montgomery
// form, plus the byte codecs the reference's ABI uses (big-endian, with silent
// reduction: SURVEY.md App. A.1 / D.1; /root/reference/src/utils.rs:35-37).
#pragma once
#include "constants.cuh"
#include "mont.cuh"

namespace lw {

struct FpCfg {
  static constexpr int N = 12;
  static constexpr uint32_t INV = k::FP_INV;
  LW_INL static const uint32_t* mod() { return k::FP_MOD; }
};
struct FrCfg {
  static constexpr int N = 8;
  static constexpr uint32_t INV = k::FR_INV;
  LW_INL static const uint32_t* mod() { return k::FR_MOD; }
};

struct Fp {
  uint32_t l[12];
};
struct Fr {
  uint32_t l[8];
};

// ------------------------------------------------------------------ Fp
LW_INL Fp fp_zero() { Fp r; for (int i = 0; i < 12; i++) r.l[i] = 0; return r; }
LW_INL Fp fp_one() { Fp r; for (int i = 0; i < 12; i++) r.l[i] = k::FP_ONE[i]; return r; }
LW_INL bool fp_is_zero(const Fp& a) { return limbs_is_zero<12>(a.l); }
LW_INL bool fp_eq(const Fp& a, const Fp& b) { return limbs_eq<12>(a.l, b.l); }
LW_INL Fp fp_add(const Fp& a, const Fp& b) { Fp r; mod_add<FpCfg>(r.l, a.l, b.l); return r; }
LW_INL Fp fp_sub(const Fp& a, const Fp& b) { Fp r; mod_sub<FpCfg>(r.l, a.l, b.l); return r; }
LW_INL Fp fp_neg(const Fp& a) { Fp r; mod_neg<FpCfg>(r.l, a.l); return r; }
LW_INL Fp fp_dbl(const Fp& a) { Fp r; mod_add<FpCfg>(r.l, a.l, a.l); return r; }
LW_INL Fp fp_mul(const Fp& a, const Fp& b) { Fp r; mont_mul<FpCfg>(r.l, a.l, b.l); return r; }
LW_INL Fp fp_sqr(const Fp& a) { Fp r; mont_sqr<FpCfg>(r.l, a.l); return r; }
LW_INL Fp fp_cneg(const Fp& a, bool neg) {
  Fp n = fp_neg(a), r;
  for (int i = 0; i < 12; i++) r.l[i] = neg ? n.l[i] : a.l[i];
  return r;
}

// Out-of-line variants for cold, code-size-heavy callers (tower fields, pow).
// ONE out-of-line multiplier whose operands and result travel in registers
// (by-value structs: ptxas keeps them out of local memory).
#if defined(LWKZG_HOST_EMUL)
inline Fp fp_mul_nv(Fp a, Fp b) { return fp_mul(a, b); }
#else
static __device__ __noinline__ Fp fp_mul_nv(Fp a, Fp b) { Fp r; mont_mul<FpCfg>(r.l, a.l, b.l); return r; }
#endif
#if defined(LWKZG_HOST_EMUL)
inline Fp fp_sqr_nv(Fp a) { return fp_sqr(a); }
#else
static __device__ __noinline__ Fp fp_sqr_nv(Fp a) { Fp r; mont_sqr<FpCfg>(r.l, a.l); return r; }
#endif
LW_INL void fp_mul_ni(Fp& r, const Fp& a, const Fp& b) { r = fp_mul_nv(a, b); }
LW_INL void fp_sqr_ni(Fp& r, const Fp& a) { r = fp_sqr_nv(a); }

LW_COLD Fp fp_pow_const(const Fp& a, const uint32_t* e, int ne) {
  Fp acc = fp_one();
  bool started = false;
  for (int w = ne - 1; w >= 0; w--) {
    uint32_t word = e[w];
    for (int bit = 31; bit >= 0; bit--) {
      if (started) fp_sqr_ni(acc, acc);
      if ((word >> bit) & 1u) {
        fp_mul_ni(acc, acc, a);
        started = true;
      }
    }
  }
  return acc;
}
// Fermat inversion a^(p-2), 0 -> 0.  Kept as the independent cross-check of fp_inv (fpinv.cuh: binary GCD, ~10x
// less latency), which is what every caller uses.
LW_COLD Fp fp_inv_fermat(const Fp& a) { return fp_pow_const(a, k::FP_P_MINUS_2, 12); }
LW_COLD Fp fp_inv(const Fp& a);  // defined in fpinv.cuh (included at the end of this header)
// sqrt candidate a^((p+1)/4); caller must check candidate^2 == a
LW_COLD Fp fp_sqrt_candidate(const Fp& a) { return fp_pow_const(a, k::FP_SQRT_EXP, 12); }

LW_INL Fp fp_to_mont(const Fp& a) { Fp r2, r; for (int i = 0; i < 12; i++) r2.l[i] = k::FP_R2[i]; mont_mul<FpCfg>(r.l, a.l, r2.l); return r; }
LW_INL Fp fp_from_mont(const Fp& a) { Fp one = fp_zero(), r; one.l[0] = 1; mont_mul<FpCfg>(r.l, a.l, one.l); return r; }

// canonical (non-Montgomery) value > (p-1)/2 ?
LW_INL bool fp_canon_is_lex_large(const Fp& canon) {
  uint32_t t[12];
  return limbs_sub<12>(t, k::FP_HALF_P, canon.l) != 0;  // half < canon
}

// 48 big-endian bytes (top three bits already cleared by the caller, so the
// value is < 2^381 < 2p) -> Montgomery Fp, reduced mod p like the reference's
// from_bytes_be.
LW_INL Fp fp_from_be48(const uint8_t* b) {
  Fp a;
  for (int i = 0; i < 12; i++) {
    const uint8_t* q = b + 44 - 4 * i;
    a.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
  }
  mod_reduce_small<FpCfg, 1>(a.l);
  return fp_to_mont(a);
}
// canonical limbs -> 48 big-endian bytes
LW_INL void fp_canon_to_be48(uint8_t* b, const Fp& canon) {
  for (int i = 0; i < 12; i++) {
    uint32_t w = canon.l[i];
    uint8_t* q = b + 44 - 4 * i;
    q[0] = (uint8_t)(w >> 24); q[1] = (uint8_t)(w >> 16); q[2] = (uint8_t)(w >> 8); q[3] = (uint8_t)w;
  }
}

// ------------------------------------------------------------------ Fr
LW_INL Fr fr_zero() { Fr r; for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
LW_INL Fr fr_one() { Fr r; for (int i = 0; i < 8; i++) r.l[i] = k::FR_ONE[i]; return r; }
LW_INL bool fr_is_zero(const Fr& a) { return limbs_is_zero<8>(a.l); }
LW_INL bool fr_eq(const Fr& a, const Fr& b) { return limbs_eq<8>(a.l, b.l); }
LW_INL Fr fr_add(const Fr& a, const Fr& b) { Fr r; mod_add<FrCfg>(r.l, a.l, b.l); return r; }
LW_INL Fr fr_sub(const Fr& a, const Fr& b) { Fr r; mod_sub<FrCfg>(r.l, a.l, b.l); return r; }
LW_INL Fr fr_neg(const Fr& a) { Fr r; mod_neg<FrCfg>(r.l, a.l); return r; }
LW_INL Fr fr_mul(const Fr& a, const Fr& b) { Fr r; mont_mul<FrCfg>(r.l, a.l, b.l); return r; }
LW_INL Fr fr_sqr(const Fr& a) { Fr r; mont_sqr<FrCfg>(r.l, a.l); return r; }
LW_INL Fr fr_to_mont(const Fr& a) { Fr r2, r; for (int i = 0; i < 8; i++) r2.l[i] = k::FR_R2[i]; mont_mul<FrCfg>(r.l, a.l, r2.l); return r; }
LW_INL Fr fr_from_mont(const Fr& a) { Fr one = fr_zero(), r; one.l[0] = 1; mont_mul<FrCfg>(r.l, a.l, one.l); return r; }

LW_COLD Fr fr_pow_const(const Fr& a, const uint32_t* e, int ne) {
  Fr acc = fr_one();
  bool started = false;
  for (int w = ne - 1; w >= 0; w--) {
    uint32_t word = e[w];
    for (int bit = 31; bit >= 0; bit--) {
      if (started) acc = fr_sqr(acc);
      if ((word >> bit) & 1u) { acc = fr_mul(acc, a); started = true; }
    }
  }
  return acc;
}
LW_COLD Fr fr_inv(const Fr& a) { return fr_pow_const(a, k::FR_R_MINUS_2, 8); }

// 8 big-endian u32 words as they sit in memory (w[0] = most significant 4
// bytes, still in memory byte order) -> canonical Fr integer, reduced mod r
// (2^256 / r < 3: two conditional subtractions).
LW_INL uint32_t bswap32(uint32_t x) {
#if defined(LWKZG_HOST_EMUL)
  return __builtin_bswap32(x);
#else
  return __byte_perm(x, 0, 0x0123);
#endif
}
LW_INL Fr fr_canon_from_be_words(const uint32_t* w) {
  Fr a;
  for (int i = 0; i < 8; i++) a.l[i] = bswap32(w[7 - i]);
  mod_reduce_small<FrCfg, 2>(a.l);
  return a;
}
LW_INL Fr fr_canon_from_be32(const uint8_t* b) {
  Fr a;
  for (int i = 0; i < 8; i++) {
    const uint8_t* q = b + 28 - 4 * i;
    a.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
  }
  mod_reduce_small<FrCfg, 2>(a.l);
  return a;
}
LW_INL void fr_canon_to_be32(uint8_t* b, const Fr& canon) {
  for (int i = 0; i < 8; i++) {
    uint32_t w = canon.l[i];
    uint8_t* q = b + 28 - 4 * i;
    q[0] = (uint8_t)(w >> 24); q[1] = (uint8_t)(w >> 16); q[2] = (uint8_t)(w >> 8); q[3] = (uint8_t)w;
  }
}

}  // namespace lw

#include "fpinv.cuh"
