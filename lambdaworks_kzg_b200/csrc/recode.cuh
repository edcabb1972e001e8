// Signed-digit (base 2^c) recoding of a canonical 256-bit scalar.
//
// k = sum_j d_j 2^(c j),  d_j in [-2^(c-1)+1, 2^(c-1)].  Because k < r < 2^255
// W = floor(255/c) + 1 windows always suffice: the top window then holds at
// most c-1 scalar bits, so digit + carry <= 2^(c-1) and no carry leaves it.  The reference's Pippenger uses unsigned
// windows; the digit set is an internal choice that cannot
// change the group element.
#pragma once
#include "ptx.cuh"

namespace lw {

// digit j of k (8 little-endian u32 limbs); `carry` is threaded from digit j-1
// (start at 0).  With c and j compile-time constants after unrolling, all limb
// indices are static.
LW_INL int recode_next_digit(const uint32_t* k8, int c, int j, int& carry) {
  const int bit = j * c;
  const int w = bit >> 5, s = bit & 31;
  uint32_t lo = k8[w];
  uint32_t hi = (w + 1 < 8) ? k8[w + 1] : 0u;
  uint32_t raw = (s == 0) ? lo : ((lo >> s) | (hi << (32 - s)));
  raw &= (1u << c) - 1u;
  int d = (int)raw + carry;
  if (d > (1 << (c - 1))) {
    d -= (1 << c);
    carry = 1;
  } else {
    carry = 0;
  }
  return d;
}

LW_HD inline int recode_num_windows(int c) { return 255 / c + 1; }

}  // namespace lw
