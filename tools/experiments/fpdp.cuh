// EXPERIMENT (not on the product path; see profiles/r01_multiplier_experiments.md for the B200
// measurements that rejected it): Fp Montgomery multiplication on the FP64 pipe (DFMA), radix 2^48.
//
// Why: the MSM hot loop is bound by the integer-multiply pipe (fmaheavy, one
// IMAD.WIDE per clock per SM -- profiles/r01_pipe_probe.md) while the FP64
// pipe, which co-issues with it at full rate (1.84 DFMA / clk / SM), is idle.
// This multiplier produces the SAME bits as mont_mul<FpCfg> (same Montgomery
// radix R = 2^384 = (2^48)^8, fully reduced output) using DFMAs for every limb
// product, so a kernel may send any subset of its multiplications here.
//
// Technique (exact double-precision products; Emmart, Zheng, Weems, "Faster
// modular exponentiation using double precision floating point arithmetic on
// the GPU", ARITH 2018): limbs are 48-bit integers held exactly in doubles.
// For a, b < 2^49
//     hi = fma_rz(a, b, 2^100)              = 2^100 + floor(ab / 2^48) 2^48
//     lo = fma_rz(a, b, (2^100 + 2^52) - hi) = 2^52  + (ab mod 2^48)
// are both exact, and the IEEE bit patterns are  (0x463 << 52) + floor(ab/2^48)
// and (0x433 << 52) + (ab mod 2^48): the mantissa fields ARE the two product
// halves.  They are summed column-wise as raw 64-bit integers (one IADD3 pair
// adds two terms), and the exponent-field biases, whose count per column is
// known at compile time, are cancelled by the initial accumulator values.
//
// Replaces (with mont.cuh) the Fp arithmetic of the un-vendored lambdaworks-math
// dependency (/root/reference/src/lib.rs:12-40 call sites).
#pragma once
#include "../../lambdaworks_kzg_b200/csrc/field.cuh"

#if defined(LWKZG_HOST_EMUL)
#include <cfenv>
#include <cmath>
#include <cstring>
#endif

namespace lw {
namespace dp {

#if defined(LWKZG_HOST_EMUL)
inline double fma_rz(double a, double b, double c) {
  volatile double va = a, vb = b, vc = c;
  const int old = fegetround();
  fesetround(FE_TOWARDZERO);
  volatile double r = std::fma(va, vb, vc);
  fesetround(old);
  return r;
}
inline uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
inline double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
inline uint32_t shr_pair(uint32_t lo, uint32_t hi, int s) { return (lo >> s) | (hi << (32 - s)); }
#else
LW_INL double fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
LW_INL uint64_t bits(double d) { return (uint64_t)__double_as_longlong(d); }
LW_INL double from_bits(uint64_t u) { return __longlong_as_double((long long)u); }
LW_INL uint32_t shr_pair(uint32_t lo, uint32_t hi, int s) { return __funnelshift_r(lo, hi, s); }
#endif

constexpr double C_HI = 0x1p100;             // puts floor(ab / 2^48) into the mantissa field
constexpr double C_LO = 0x1p100 + 0x1p52;    // (C_LO - hi) + ab = 2^52 + (ab mod 2^48)
constexpr uint64_t BIAS_HI = 0x463ull << 52;  // exponent field of 2^100
constexpr uint64_t BIAS_LO = 0x433ull << 52;  // exponent field of 2^52
constexpr uint64_t MASK48 = (1ull << 48) - 1;


// exact double of the 48-bit integer (hi16 << 32) | lo32
LW_INL double u48_to_double(uint32_t lo32, uint32_t hi16) {
  return from_bits(((uint64_t)(hi16 | 0x43300000u) << 32) | lo32) - 0x1p52;
}

// 12 x u32 limbs -> 8 x 48-bit limbs as doubles (3 words = 2 limbs)
LW_INL void to_d48(double* d, const uint32_t* l) {
#pragma unroll
  for (int m = 0; m < 4; m++) {
    d[2 * m] = u48_to_double(l[3 * m], l[3 * m + 1] & 0xffffu);
    d[2 * m + 1] = u48_to_double(shr_pair(l[3 * m + 1], l[3 * m + 2], 16), l[3 * m + 2] >> 16);
  }
}

// lo_acc += (ab mod 2^48) + BIAS_LO ; hi_acc += floor(ab / 2^48) + BIAS_HI
LW_INL void mac(uint64_t& lo_acc, uint64_t& hi_acc, double a, double b) {
  const double hi = fma_rz(a, b, C_HI);
  const double lo = fma_rz(a, b, C_LO - hi);
  hi_acc += bits(hi);
  lo_acc += bits(lo);
}

// number of limb products a_i b_j with i + j == k, 0 <= i, j < 8
LW_HD constexpr int pairs_in_column(int k) { return k < 0 || k > 14 ? 0 : (k < 8 ? k + 1 : 15 - k); }
// exponent-field bias that `rows` full 8 x 8 product grids deposit in column k
LW_HD constexpr uint64_t column_bias(int k, int grids) {
  return (uint64_t)grids * ((uint64_t)pairs_in_column(k) * BIAS_LO + (uint64_t)pairs_in_column(k - 1) * BIAS_HI);
}

// Montgomery reduction of the 16 redundant columns in acc (already carrying
// -column_bias(k, G + 1) where G product grids were accumulated), then
// normalisation, repacking into 12 x u32 and the final conditional subtraction.
LW_INL void redc_columns(uint32_t* r, uint64_t* acc) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint64_t q = (acc[i] * k::FP_INV48) & MASK48;  // only the low 48 bits of the column matter
    const double qd = u48_to_double((uint32_t)q, (uint32_t)(q >> 32));
#pragma unroll
    for (int j = 0; j < 8; j++) mac(acc[i + j], acc[i + j + 1], qd, k::FP_MOD48[j]);
    acc[i + 1] += acc[i] >> 48;  // column i is complete, bias-free and == 0 mod 2^48
  }
#pragma unroll
  for (int kk = 8; kk < 15; kk++) {
    acc[kk + 1] += acc[kk] >> 48;
    acc[kk] &= MASK48;
  }
  uint32_t T[12];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const uint64_t L0 = acc[8 + 2 * m], L1 = acc[9 + 2 * m];
    T[3 * m] = (uint32_t)L0;
    T[3 * m + 1] = (uint32_t)(L0 >> 32) | ((uint32_t)L1 << 16);
    T[3 * m + 2] = (uint32_t)(L1 >> 16);
  }
  uint32_t t[12];
  const uint32_t borrow = limbs_sub<12>(t, T, k::FP_MOD);
#pragma unroll
  for (int i = 0; i < 12; i++) r[i] = borrow ? T[i] : t[i];
}

// r = a b / 2^384 mod p; inputs < p, output < p (bit-identical to mont_mul<FpCfg>)
LW_INL void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  double ad[8], bd[8];
  to_d48(ad, a);
  to_d48(bd, b);
  uint64_t acc[16];
#pragma unroll
  for (int kk = 0; kk < 16; kk++) acc[kk] = 0ull - column_bias(kk, 2);
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) mac(acc[i + j], acc[i + j + 1], ad[i], bd[j]);
  redc_columns(r, acc);
}

// r = a^2 / 2^384 mod p: 36 limb products instead of 64 (the doubled operand
// 2 a_i < 2^49 is still exact, and 2 a_i a_j < 2^97 still splits exactly).
LW_INL void mont_sqr(uint32_t* r, const uint32_t* a) {
  double ad[8], a2[8];
  to_d48(ad, a);
#pragma unroll
  for (int i = 0; i < 8; i++) a2[i] = ad[i] + ad[i];
  uint64_t acc[16];
#pragma unroll
  for (int kk = 0; kk < 16; kk++) {
    // product terms landing in column kk: pairs i <= j with i + j == kk (lo) / kk - 1 (hi)
    const int nlo = (pairs_in_column(kk) + 1) / 2, nhi = (pairs_in_column(kk - 1) + 1) / 2;
    acc[kk] = 0ull - (column_bias(kk, 1) + (uint64_t)nlo * BIAS_LO + (uint64_t)nhi * BIAS_HI);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    mac(acc[2 * i], acc[2 * i + 1], ad[i], ad[i]);
#pragma unroll
    for (int j = i + 1; j < 8; j++) mac(acc[i + j], acc[i + j + 1], a2[i], ad[j]);
  }
  redc_columns(r, acc);
}

}  // namespace dp

LW_INL Fp fp_mul_dp(const Fp& a, const Fp& b) { Fp r; dp::mont_mul(r.l, a.l, b.l); return r; }
LW_INL Fp fp_sqr_dp(const Fp& a) { Fp r; dp::mont_sqr(r.l, a.l); return r; }

}  // namespace lw
