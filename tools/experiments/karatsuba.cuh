// EXPERIMENT (not on the product path): one level of Karatsuba in the product phase of the
// 12 x 32-bit Montgomery multiplication.  Bit-exact (tests/test_host_emul.py), but measured on B200
// at 300 vs 303 clk per warp-multiplication per SM (profiles/r01_multiplier_experiments.md): the 36
// wide MACs it saves are paid back in carry handling, so csrc/mont.cuh keeps the schoolbook product.
#pragma once
#include "../../lambdaworks_kzg_b200/csrc/field.cuh"

namespace lw {

// 2H-limb product of two H-limb numbers, schoolbook, as aligned wide-MAC chains.
template <int H>
LW_INL void limbs_mul_half(uint32_t* T, const uint32_t* a, const uint32_t* b) {
  uint32_t E[2 * H], O[2 * H];  // E[k] = limb k, O[k] = limb k + 1
#pragma unroll
  for (int k = 0; k < 2 * H; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
  for (int i = 0; i < H; i++) {
    {  // j of the same parity as i: even positions
      const int j0 = i & 1;
      E[i + j0] = ptx::mad_lo_cc(a[j0], b[i], E[i + j0]);
      E[i + j0 + 1] = ptx::madc_hi_cc(a[j0], b[i], E[i + j0 + 1]);
      int last = j0;
#pragma unroll
      for (int j = j0 + 2; j < H; j += 2) {
        E[i + j] = ptx::madc_lo_cc(a[j], b[i], E[i + j]);
        E[i + j + 1] = ptx::madc_hi_cc(a[j], b[i], E[i + j + 1]);
        last = j;
      }
      if (i + last + 2 < 2 * H) E[i + last + 2] = ptx::addc(E[i + last + 2], 0);
    }
    {  // opposite parity: odd positions
      const int j0 = 1 - (i & 1);
      O[i + j0 - 1] = ptx::mad_lo_cc(a[j0], b[i], O[i + j0 - 1]);
      O[i + j0] = ptx::madc_hi_cc(a[j0], b[i], O[i + j0]);
      int last = j0;
#pragma unroll
      for (int j = j0 + 2; j < H; j += 2) {
        O[i + j - 1] = ptx::madc_lo_cc(a[j], b[i], O[i + j - 1]);
        O[i + j] = ptx::madc_hi_cc(a[j], b[i], O[i + j]);
        last = j;
      }
      if (i + last + 1 < 2 * H) O[i + last + 1] = ptx::addc(O[i + last + 1], 0);
    }
  }
  T[0] = E[0];
  T[1] = ptx::add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 2 * H; k++) T[k] = ptx::addc_cc(E[k], O[k - 1]);
}

// Montgomery product with one level of Karatsuba in the product phase:
// 3 (N/2)^2 = 108 wide MACs instead of 144 for N = 12, plus ~120 ALU operations
// (the ALU pipe is 80 % idle in the MSM kernel); reduction as in mont_redc_2n.
template <class C>
LW_INL void mont_mul_karatsuba(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = C::N, H = N / 2;
  uint32_t P0[2 * H], P2[2 * H], P1[2 * H];
  limbs_mul_half<H>(P0, a, b);
  limbs_mul_half<H>(P2, a + H, b + H);
  uint32_t sa[H], sb[H];
  sa[0] = ptx::add_cc(a[0], a[H]);
#pragma unroll
  for (int k = 1; k < H; k++) sa[k] = ptx::addc_cc(a[k], a[H + k]);
  const uint32_t ca = ptx::addc(0, 0);
  sb[0] = ptx::add_cc(b[0], b[H]);
#pragma unroll
  for (int k = 1; k < H; k++) sb[k] = ptx::addc_cc(b[k], b[H + k]);
  const uint32_t cb = ptx::addc(0, 0);
  limbs_mul_half<H>(P1, sa, sb);
  // M = (sa + ca X)(sb + cb X) - P0 - P2,  X = 2^(32 H);  2H + 1 limbs (M = a0 b1 + a1 b0 >= 0)
  uint32_t M[2 * H + 1];
#pragma unroll
  for (int k = 0; k < 2 * H; k++) M[k] = P1[k];
  M[2 * H] = ca & cb;
  const uint32_t ma = 0u - ca, mb = 0u - cb;
  M[H] = ptx::add_cc(M[H], sb[0] & ma);
#pragma unroll
  for (int k = 1; k < H; k++) M[H + k] = ptx::addc_cc(M[H + k], sb[k] & ma);
  M[2 * H] = ptx::addc(M[2 * H], 0);
  M[H] = ptx::add_cc(M[H], sa[0] & mb);
#pragma unroll
  for (int k = 1; k < H; k++) M[H + k] = ptx::addc_cc(M[H + k], sa[k] & mb);
  M[2 * H] = ptx::addc(M[2 * H], 0);
  M[0] = ptx::sub_cc(M[0], P0[0]);
#pragma unroll
  for (int k = 1; k < 2 * H; k++) M[k] = ptx::subc_cc(M[k], P0[k]);
  M[2 * H] = ptx::subc(M[2 * H], 0);
  M[0] = ptx::sub_cc(M[0], P2[0]);
#pragma unroll
  for (int k = 1; k < 2 * H; k++) M[k] = ptx::subc_cc(M[k], P2[k]);
  M[2 * H] = ptx::subc(M[2 * H], 0);
  // T = P0 + M X + P2 X^2
  uint32_t T[2 * N];
#pragma unroll
  for (int k = 0; k < 2 * H; k++) { T[k] = P0[k]; T[2 * H + k] = P2[k]; }
  T[H] = ptx::add_cc(T[H], M[0]);
#pragma unroll
  for (int k = 1; k <= 2 * H; k++) T[H + k] = ptx::addc_cc(T[H + k], M[k]);
#pragma unroll
  for (int k = 3 * H + 1; k < 2 * N; k++) T[k] = ptx::addc_cc(T[k], 0);
  mont_redc_2n<C>(r, T);
}

}  // namespace lw
