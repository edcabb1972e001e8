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

// Montgomery reduction of a 2N-limb product T (< m^2 ... < m 2^(32N)): r = T / 2^(32N) mod m.
// Same even/odd aligned accumulators as mont_mul; one high limb of T is injected per round.
template <class C>
LW_INL void mont_redc_2n(uint32_t* r, const uint32_t* T) {
  constexpr int N = C::N;
  const uint32_t* m = C::mod();
  uint32_t E[N], O[N];
#pragma unroll
  for (int k = 0; k < N; k++) { E[k] = T[k]; O[k] = 0; }
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (i > 0) {
      // R >>= 32: O becomes the new E; E shifted down one pair becomes the new O,
      // with product limb N+i-1 entering at window position N-1
      const uint32_t orphan = E[1];
      uint32_t nO[N];
#pragma unroll
      for (int j = 0; j < N - 2; j++) nO[j] = E[j + 2];
      nO[N - 2] = T[N + i - 1];
      nO[N - 1] = 0;
#pragma unroll
      for (int j = 0; j < N; j++) E[j] = O[j];
#pragma unroll
      for (int j = 0; j < N; j++) O[j] = nO[j];
      E[0] = ptx::add_cc(E[0], orphan);  // carry flows into the O chain below (mul.lo leaves CC alone)
    }
    const uint32_t mi = ptx::mul_lo(E[0], C::INV);
    if (i > 0) {
      O[0] = ptx::madc_lo_cc(m[1], mi, O[0]);
    } else {
      O[0] = ptx::mad_lo_cc(m[1], mi, O[0]);
    }
    O[1] = ptx::madc_hi_cc(m[1], mi, O[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      O[j] = ptx::madc_lo_cc(m[j + 1], mi, O[j]);
      O[j + 1] = ptx::madc_hi_cc(m[j + 1], mi, O[j + 1]);
    }
    E[0] = ptx::mad_lo_cc(m[0], mi, E[0]);  // == 0
    E[1] = ptx::madc_hi_cc(m[0], mi, E[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      E[j] = ptx::madc_lo_cc(m[j], mi, E[j]);
      E[j + 1] = ptx::madc_hi_cc(m[j], mi, E[j + 1]);
    }
    O[N - 1] = ptx::addc(O[N - 1], 0);
  }
  // result limb k = O[k] + E[k+1], plus the last product limb at position N-1
  uint32_t Rr[N];
  Rr[0] = ptx::add_cc(O[0], E[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) Rr[k] = ptx::addc_cc(O[k], E[k + 1]);
  Rr[N - 1] = ptx::addc(O[N - 1], T[2 * N - 1]);
  uint32_t t[N];
  uint32_t borrow = limbs_sub<N>(t, Rr, m);
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = borrow ? Rr[i] : t[i];
}

// Montgomery square: r = a^2 / 2^(32N) mod m.
//
// Product phase with the symmetric half only: the off-diagonal products a_i a_j
// (i < j) are accumulated once -- j - i odd lands on odd limb positions and goes
// to the accumulator PO (pairs (1,2)(3,4)...), j - i even to PE (pairs (0,1)(2,3)
// ...), so every row is again two aligned carry chains of wide MACs -- then
// doubled with funnel shifts and completed by the N diagonal squares (one more
// aligned chain): N(N+1)/2 wide MACs instead of N^2.  Rows are processed in
// ascending i, so the limb that receives a chain's carry-out holds at most
// earlier carries.  The 2N-limb product is then reduced with the same even/odd
// reduction rounds as mont_mul, injecting one high product limb per round.
template <class C>
LW_INL void mont_sqr(uint32_t* r, const uint32_t* a) {
  constexpr int N = C::N;
  const uint32_t* m = C::mod();
  uint32_t PE[2 * N], PO[2 * N];  // PE[k] = limb k, PO[k] = limb k + 1
#pragma unroll
  for (int k = 0; k < 2 * N; k++) { PE[k] = 0; PO[k] = 0; }
#pragma unroll
  for (int i = 0; i < N - 1; i++) {
    // odd offsets j = i+1, i+3, ... -> PO indices (i+j-1, i+j)
    {
      const int j0 = i + 1;
      PO[i + j0 - 1] = ptx::mad_lo_cc(a[i], a[j0], PO[i + j0 - 1]);
      PO[i + j0] = ptx::madc_hi_cc(a[i], a[j0], PO[i + j0]);
      int last = j0;
#pragma unroll
      for (int j = i + 3; j < N; j += 2) {
        PO[i + j - 1] = ptx::madc_lo_cc(a[i], a[j], PO[i + j - 1]);
        PO[i + j] = ptx::madc_hi_cc(a[i], a[j], PO[i + j]);
        last = j;
      }
      PO[i + last + 1] = ptx::addc(PO[i + last + 1], 0);
    }
    // even offsets j = i+2, i+4, ... -> PE indices (i+j, i+j+1)
    if (i + 2 < N) {
      const int j0 = i + 2;
      PE[i + j0] = ptx::mad_lo_cc(a[i], a[j0], PE[i + j0]);
      PE[i + j0 + 1] = ptx::madc_hi_cc(a[i], a[j0], PE[i + j0 + 1]);
      int last = j0;
#pragma unroll
      for (int j = i + 4; j < N; j += 2) {
        PE[i + j] = ptx::madc_lo_cc(a[i], a[j], PE[i + j]);
        PE[i + j + 1] = ptx::madc_hi_cc(a[i], a[j], PE[i + j + 1]);
        last = j;
      }
      PE[i + last + 2] = ptx::addc(PE[i + last + 2], 0);
    }
  }
  // S = PE + (PO << 32), T = 2 S + sum a_i^2 2^(64 i)
  uint32_t T[2 * N];
  T[0] = PE[0];
  T[1] = ptx::add_cc(PE[1], PO[0]);
#pragma unroll
  for (int k = 2; k < 2 * N; k++) T[k] = ptx::addc_cc(PE[k], PO[k - 1]);
#pragma unroll
  for (int k = 2 * N - 1; k > 0; k--) T[k] = (T[k] << 1) | (T[k - 1] >> 31);
  T[0] <<= 1;
  T[0] = ptx::mad_lo_cc(a[0], a[0], T[0]);
  T[1] = ptx::madc_hi_cc(a[0], a[0], T[1]);
#pragma unroll
  for (int i = 1; i < N; i++) {
    T[2 * i] = ptx::madc_lo_cc(a[i], a[i], T[2 * i]);
    T[2 * i + 1] = ptx::madc_hi_cc(a[i], a[i], T[2 * i + 1]);
  }
  mont_redc_2n<C>(r, T);
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
