// Scalar recoding for the fixed-base MSM: GLV split + signed base-2^c digits.
//
// BLS12-381 has r = x^4 - x^2 + 1 and the endomorphism psi(X, Y) = (beta X, -Y) = [x^2](X, Y) on G1, so every
// k < r splits as k = m + q x^2 with m = k mod x^2 and q = k div x^2, both < x^2 < 2^128, and
//        [k]P = [m]P + psi([q]P).
// The MSM kernels sum the m-halves and the q-halves of all scalars separately (psi is a homomorphism: it is
// applied ONCE to the finished q-sum), so one table of the multiples d 2^(c j) P_i for 128-bit scalars serves both
// halves: W = ceil(128 / c) windows instead of ceil(255 / c) for the same memory -- which is what lets a 16-bit
// window (8 windows, 100 GiB) replace the 15-bit one (18 windows, 108 GiB): 16 table entries per point instead
// of 17.45.  The reference's Pippenger (lambdaworks-math msm::pippenger, called from /root/reference/src/lib.rs:
// 241-243, 269-270, 329, 394) uses unsigned windows over the whole scalar; the digit set and the split are
// internal choices that cannot change the group element.
//
// Digits: windows 0 .. W-2 are signed, d in [-2^(c-1)+1, 2^(c-1)], with a carry into the next window; the top
// window W-1 is unsigned (digit = raw bits + carry <= top_max + 1) so that no carry ever leaves it.  The table
// therefore holds 2^(c-1) multiples per (window, point) for the lower windows and top_max + 1 for the top one.
#pragma once
#include "constants.cuh"
#include "ptx.cuh"

namespace lw {

LW_HD inline int glv_num_windows(int c) { return (128 + c - 1) / c; }
// largest value of (x^2 - 1) >> (c (W - 1)): the top window's raw digit never exceeds it
LW_HD inline uint32_t glv_top_max(int c) {
  // x^2 - 1 = 0xac45a401 0001a402 00000000 ffffffff
  const uint32_t v[4] = {0xffffffffu, 0x00000000u, 0x0001a402u, 0xac45a401u};
  const int sh = c * (glv_num_windows(c) - 1);
  const int w = sh >> 5, s = sh & 31;
  uint32_t lo = v[w], hi = (w + 1 < 4) ? v[w + 1] : 0u;
  return s == 0 ? lo : ((lo >> s) | (hi << (32 - s)));   // c <= 16 and sh >= 112 - ... : fits 32 bits for every c >= 4
}
// multiples stored per point for window j
LW_HD inline uint32_t glv_window_count(int c, int j) { return j == glv_num_windows(c) - 1 ? glv_top_max(c) + 1u : (1u << (c - 1)); }
// entries of the whole table for n points
LW_HD inline unsigned long long glv_table_entries(int c, int npoints) {
  const int W = glv_num_windows(c);
  return (unsigned long long)npoints * ((unsigned long long)(W - 1) * (1ull << (c - 1)) + glv_top_max(c) + 1ull);
}

// k (8 little-endian limbs, < r) -> m = k mod x^2, q = k div x^2.  Barrett with mu = floor(2^256 / x^2):
// q' = floor(k mu / 2^256) is q or q - 1, fixed up by at most one subtraction (two are coded).
LW_INL void glv_split_barrett(uint32_t* q4, uint32_t* m4, const uint32_t* k8) {
  uint32_t prod[13];
#pragma unroll
  for (int i = 0; i < 13; i++) prod[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t carry = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
      const unsigned long long t = (unsigned long long)k8[i] * k::BLS_X2_MU[j] + prod[i + j] + carry;
      prod[i + j] = (uint32_t)t;
      carry = (uint32_t)(t >> 32);
    }
    prod[i + 5] = carry;
  }
  uint32_t q[4] = {prod[8], prod[9], prod[10], prod[11]};
  // m = k - q x^2 (mod 2^160); the true value is < 2 x^2 < 2^129
  uint32_t qx[5] = {0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t carry = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (i + j < 5) {
        const unsigned long long t = (unsigned long long)q[i] * k::BLS_X2[j] + qx[i + j] + carry;
        qx[i + j] = (uint32_t)t;
        carry = (uint32_t)(t >> 32);
      }
    }
    if (i + 4 < 5) qx[i + 4] = carry;
  }
  uint32_t m[5];
  {
    unsigned long long borrow = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const unsigned long long d = (unsigned long long)k8[i] - qx[i] - borrow;
      m[i] = (uint32_t)d;
      borrow = (d >> 63) & 1ull;
    }
  }
#pragma unroll
  for (int rep = 0; rep < 2; rep++) {
    uint32_t t[5];
    unsigned long long borrow = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const unsigned long long d = (unsigned long long)m[i] - (i < 4 ? k::BLS_X2[i] : 0u) - borrow;
      t[i] = (uint32_t)d;
      borrow = (d >> 63) & 1ull;
    }
    if (!borrow) {   // m >= x^2
#pragma unroll
      for (int i = 0; i < 5; i++) m[i] = t[i];
      unsigned long long c1 = 1;
#pragma unroll
      for (int i = 0; i < 4; i++) { c1 += q[i]; q[i] = (uint32_t)c1; c1 >>= 32; }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) { q4[i] = q[i]; m4[i] = m[i]; }
}

// digit j (0 <= j < W) of a 128-bit half given as 4 little-endian limbs read through `limb(w)`; `carry` is
// threaded from digit j - 1 (start at 0).  Returns the signed digit (top window: unsigned).
template <class LimbFn>
LW_INL int glv_digit(LimbFn limb, int c, int W, int j, int& carry) {
  const int bit = j * c;
  const int w = bit >> 5, s = bit & 31;
  const uint32_t lo = limb(w);
  const uint32_t hi = (w + 1 < 4) ? limb(w + 1) : 0u;
  uint32_t raw = (s == 0) ? lo : ((lo >> s) | (hi << (32 - s)));
  raw &= (1u << c) - 1u;
  int d = (int)raw + carry;
  if (j < W - 1 && d > (1 << (c - 1))) {
    d -= (1 << c);
    carry = 1;
  } else {
    carry = 0;
  }
  return d;
}

}  // namespace lw
