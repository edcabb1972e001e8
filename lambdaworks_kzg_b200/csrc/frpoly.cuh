// Polynomial evaluation and quotient over Fr, chunk-parallel.
//
// Replaces lambdaworks-math Polynomial::evaluate (Horner) and the Ruffini
// division inside KZG::open (call sites /root/reference/src/lib.rs:320,329,
// 389,394):
//     y = p(z),   q = (p - y)/(X - z),   q_{k-1} = sum_{j>=k} c_j z^(j-k).
// The quotient coefficients are exactly the Horner partial sums, so one
// downward sweep yields both.  To parallelise, the 4096 coefficients are split
// into T chunks of m; chunk t first computes its local Horner value L_t, a
// suffix scan gives the carry-in I_t = sum_{j >= end_t} c_j z^(j-end_t), and a
// second sweep seeded with I_t emits the exact quotient coefficients.
//
// Arithmetic trick: coefficients stay CANONICAL while z (and its powers) are in
// Montgomery form -- mont_mul(zR, v) = z*v -- so no conversion is ever needed.
#pragma once
#include "field.cuh"

namespace lw {

// Local Horner value of m canonical coefficients c[0..m) (ascending degree):
// sum_k c[k] z^k.  `load(k)` returns coefficient k of the chunk.
template <class Load>
LW_INL Fr chunk_horner(Load load, int m, const Fr& z_mont) {
  Fr v = fr_zero();
  for (int k = m - 1; k >= 0; k--) v = fr_add(load(k), fr_mul(z_mont, v));
  return v;
}

// Second sweep: v = carry_in; for k = m-1..0: v = c[k] + z v; store(k, v).
// store(k, v) receives S_{start+k} = sum_{j >= start+k} c_j z^(j-start-k); the
// caller writes it to q[start+k-1] (and S_0 = y).
template <class Load, class Store>
LW_INL Fr chunk_sweep(Load load, Store store, int m, const Fr& z_mont, const Fr& carry_in) {
  Fr v = carry_in;
  for (int k = m - 1; k >= 0; k--) {
    v = fr_add(load(k), fr_mul(z_mont, v));
    store(k, v);
  }
  return v;
}

// z^(2^s) helper
LW_INL Fr fr_pow2k(Fr z_mont, int s) {
  for (int i = 0; i < s; i++) z_mont = fr_sqr(z_mont);
  return z_mont;
}

#if defined(LWKZG_HOST_EMUL)
// Sequential model of the kernel's chunk decomposition (tests only): the same
// per-chunk functions, the scan done as a loop in the order the warp scan
// composes it.  q gets n canonical coefficients (q[n-1] = 0).
inline void poly_eval_quot_reference_order(uint32_t* y8, uint32_t* q, const uint32_t* coeffs, int n, const uint32_t* z8, int chunks) {
  int m = n / chunks;
  Fr zc; for (int i = 0; i < 8; i++) zc.l[i] = z8[i];
  Fr z = fr_to_mont(zc);
  int s = 0; while ((1 << s) < m) s++;
  Fr zm = fr_pow2k(z, s);  // z^m (Montgomery)
  Fr* L = new Fr[chunks];
  for (int t = 0; t < chunks; t++) {
    const uint32_t* base = coeffs + (size_t)t * m * 8;
    L[t] = chunk_horner([&](int k) { Fr c; for (int i = 0; i < 8; i++) c.l[i] = base[k * 8 + i]; return c; }, m, z);
  }
  // inclusive suffix scan H_t = L_t + z^m H_{t+1} by Hillis-Steele doubling:
  // after step d, H_t covers chunks [t, t+2^(d+1))
  Fr* H = new Fr[chunks];
  Fr* Hn = new Fr[chunks];
  for (int t = 0; t < chunks; t++) H[t] = L[t];
  Fr pw = zm;
  for (int d = 1; d < chunks; d <<= 1) {
    for (int t = 0; t < chunks; t++) Hn[t] = (t + d < chunks) ? fr_add(H[t], fr_mul(pw, H[t + d])) : H[t];
    for (int t = 0; t < chunks; t++) H[t] = Hn[t];
    pw = fr_sqr(pw);
  }
  for (int t = 0; t < chunks; t++) {
    const uint32_t* base = coeffs + (size_t)t * m * 8;
    Fr carry = (t + 1 < chunks) ? H[t + 1] : fr_zero();
    int start = t * m;
    Fr v = chunk_sweep([&](int k) { Fr c; for (int i = 0; i < 8; i++) c.l[i] = base[k * 8 + i]; return c; },
                       [&](int k, const Fr& val) { int g = start + k; if (g >= 1) for (int i = 0; i < 8; i++) q[(size_t)(g - 1) * 8 + i] = val.l[i]; },
                       m, z, carry);
    if (t == 0) for (int i = 0; i < 8; i++) y8[i] = v.l[i];
  }
  for (int i = 0; i < 8; i++) q[(size_t)(n - 1) * 8 + i] = 0;
  delete[] L; delete[] H; delete[] Hn;
}
#endif

}  // namespace lw
