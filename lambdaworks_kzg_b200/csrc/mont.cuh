// N x 32-bit-limb Montgomery field arithmetic (little-endian limbs), used for
// BLS12-381 Fp (N = 12) and Fr (N = 8).
//
// Replaces the arithmetic the reference takes from its un-vendored dependency
// lambdaworks-math (FieldElement<MontgomeryBackendPrimeField<..>>; call sites
// /root/reference/src/lib.rs:12-40, src/utils.rs:35-37) -- SURVEY.md §2.1.
//
// All values handed between functions are fully reduced (0 <= v < modulus), so
// limb-wise equality is field equality.
#pragma once
#include "ptx.cuh"

namespace lw {

template <int N>
struct Limbs {
  uint32_t l[N];
};

// ---------------------------------------------------------------- raw helpers

template <int N>
LW_INL bool limbs_is_zero(const uint32_t* a) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < N; i++) acc |= a[i];
  return acc == 0;
}

template <int N>
LW_INL bool limbs_eq(const uint32_t* a, const uint32_t* b) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < N; i++) acc |= a[i] ^ b[i];
  return acc == 0;
}

// r = a + b, returns carry-out
template <int N>
LW_INL uint32_t limbs_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = ptx::add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r[i] = ptx::addc_cc(a[i], b[i]);
  return ptx::addc(0, 0);
}

// r = a - b, returns borrow-out (1 = a < b)
template <int N>
LW_INL uint32_t limbs_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = ptx::sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r[i] = ptx::subc_cc(a[i], b[i]);
  return ptx::borrow_flag();
}

// a < b ?
template <int N>
LW_INL bool limbs_lt(const uint32_t* a, const uint32_t* b) {
  uint32_t t[N];
  return limbs_sub<N>(t, a, b) != 0;
}

// ---------------------------------------------------------------- field ops
// C supplies: N, mod() -> const uint32_t*, INV (= -mod^-1 mod 2^32)

template <class C>
LW_INL void mod_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = C::N;
  uint32_t s[N], t[N];
  uint32_t carry = limbs_add<N>(s, a, b);  // modulus < 2^(32N-1): carry is always 0
  (void)carry;
  uint32_t borrow = limbs_sub<N>(t, s, C::mod());
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = borrow ? s[i] : t[i];
}

template <class C>
LW_INL void mod_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = C::N;
  uint32_t s[N], t[N];
  uint32_t borrow = limbs_sub<N>(s, a, b);
  limbs_add<N>(t, s, C::mod());
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = borrow ? t[i] : s[i];
}

template <class C>
LW_INL void mod_neg(uint32_t* r, const uint32_t* a) {
  constexpr int N = C::N;
  uint32_t t[N];
  bool z = limbs_is_zero<N>(a);
  limbs_sub<N>(t, C::mod(), a);
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = z ? 0u : t[i];
}

// r = a * 2 mod m
template <class C>
LW_INL void mod_dbl(uint32_t* r, const uint32_t* a) {
  mod_add<C>(r, a, a);
}

// Montgomery product r = a * b / 2^(32N) mod m.  Inputs < m, output < m.
//
// Operand-scanning CIOS organised for the SASS the hardware wants: ptxas fuses
// a (mad.lo.cc, madc.hi.cc) pair on the same operands into ONE IMAD.WIDE.U32.X
// whose 64-bit accumulator must sit in an aligned register pair.  The running
// total T (N+1 limbs, already divided by 2^(32 i)) is therefore kept as two
// separately aligned accumulators
//        T = E + O * 2^32 ,   E pairs cover limbs (0,1)(2,3)..., O pairs (1,2)(3,4)...
// a_j*b_i with j even accumulates into E, j odd into O, each as ONE carry chain
// of N/2 wide MACs; same for m_i * mod.  After the reduction step limb 0 is
// zero and T >>= 32 is free: O becomes the new E, and E (shifted down one pair
// by writing dest != addend) becomes the new O; only the orphaned high word
// E[1] has to be added into the new limb 0.  T + a*b_i + m_i*mod < 2^(32(N+1))
// for both BLS12-381 moduli, so no carry ever leaves limb N.
template <class C>
LW_INL void mont_mad_redc(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi, bool first) {
  constexpr int N = C::N;
  const uint32_t* m = C::mod();
  if (first) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      ev[j] = ptx::mul_lo(a[j], bi);
      ev[j + 1] = ptx::mul_hi(a[j], bi);
      od[j] = ptx::mul_lo(a[j + 1], bi);
      od[j + 1] = ptx::mul_hi(a[j + 1], bi);
    }
  } else {
    // here `ev` is last step's O (new limbs 0..N-1) and `od` is last step's E,
    // whose pair k+1 becomes new O pair k.
    ev[0] = ptx::add_cc(ev[0], od[1]);
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
      od[j] = ptx::madc_lo_cc(a[j + 1], bi, od[j + 2]);
      od[j + 1] = ptx::madc_hi_cc(a[j + 1], bi, od[j + 3]);
    }
    od[N - 2] = ptx::madc_lo_cc(a[N - 1], bi, 0);
    od[N - 1] = ptx::madc_hi(a[N - 1], bi, 0);
    ev[0] = ptx::mad_lo_cc(a[0], bi, ev[0]);
    ev[1] = ptx::madc_hi_cc(a[0], bi, ev[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      ev[j] = ptx::madc_lo_cc(a[j], bi, ev[j]);
      ev[j + 1] = ptx::madc_hi_cc(a[j], bi, ev[j + 1]);
    }
    od[N - 1] = ptx::addc(od[N - 1], 0);
  }
  const uint32_t mi = ptx::mul_lo(ev[0], C::INV);
  od[0] = ptx::mad_lo_cc(m[1], mi, od[0]);
  od[1] = ptx::madc_hi_cc(m[1], mi, od[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    od[j] = ptx::madc_lo_cc(m[j + 1], mi, od[j]);
    od[j + 1] = ptx::madc_hi_cc(m[j + 1], mi, od[j + 1]);
  }
  ev[0] = ptx::mad_lo_cc(m[0], mi, ev[0]);  // == 0
  ev[1] = ptx::madc_hi_cc(m[0], mi, ev[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    ev[j] = ptx::madc_lo_cc(m[j], mi, ev[j]);
    ev[j + 1] = ptx::madc_hi_cc(m[j], mi, ev[j + 1]);
  }
  od[N - 1] = ptx::addc(od[N - 1], 0);
}

template <class C>
LW_INL void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = C::N;
  static_assert(N % 2 == 0, "even limb count");
  uint32_t X[N], Y[N];
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    mont_mad_redc<C>(X, Y, a, b[i], i == 0);
    mont_mad_redc<C>(Y, X, a, b[i + 1], false);
  }
  // after the last step Y plays E (limb 0 == 0, dropped) and X plays O:
  // result limb k = O[k] + E[k+1]
  uint32_t T[N];
  T[0] = ptx::add_cc(X[0], Y[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) T[k] = ptx::addc_cc(X[k], Y[k + 1]);
  T[N - 1] = ptx::addc(X[N - 1], 0);
  // T < 2m: one conditional subtraction
  uint32_t t[N];
  uint32_t borrow = limbs_sub<N>(t, T, C::mod());
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = borrow ? T[i] : t[i];
}

template <class C>
LW_INL void mont_sqr(uint32_t* r, const uint32_t* a) {
  mont_mul<C>(r, a, a);
}

// Reduce an arbitrary N-limb integer (< 2^(32N)) into [0, m): at most
// floor(2^(32N)/m) conditional subtractions (2 for Fr, 9 for Fp -- callers in
// Fp only ever pass values < 2^381 < 2p... see fp_from_be48).
template <class C, int MAXSUB>
LW_INL void mod_reduce_small(uint32_t* a) {
  constexpr int N = C::N;
#pragma unroll
  for (int k = 0; k < MAXSUB; k++) {
    uint32_t t[N];
    uint32_t borrow = limbs_sub<N>(t, a, C::mod());
#pragma unroll
    for (int i = 0; i < N; i++) a[i] = borrow ? a[i] : t[i];
  }
}

// r = a^e, e given as NE little-endian 32-bit limbs (read from constant memory
// at run time; plain left-to-right square-and-multiply).
template <class C>
LW_DEV inline void mont_pow(uint32_t* r, const uint32_t* a, const uint32_t* e, int ne, const uint32_t* one) {
  constexpr int N = C::N;
  uint32_t acc[N], base[N];
#pragma unroll
  for (int i = 0; i < N; i++) { acc[i] = one[i]; base[i] = a[i]; }
  bool started = false;
  for (int w = ne - 1; w >= 0; w--) {
    uint32_t word = e[w];
    for (int bit = 31; bit >= 0; bit--) {
      if (started) mont_sqr<C>(acc, acc);
      if ((word >> bit) & 1u) {
        mont_mul<C>(acc, acc, base);
        started = true;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = acc[i];
}

}  // namespace lw
