// MODE_CKZG_LE polynomial kernels' per-lane pieces: evaluation-form blobs over the
// bit-reversed 4096th roots of unity (c-kzg-4844's
// evaluate_polynomial_in_evaluation_form / compute_kzg_proof_impl).  The
// reference itself never implements this (its loaders leave the Lagrange
// conversion as a TODO, /root/reference/src/lib.rs:760-770, src/srs.rs:117-124);
// the semantics are pinned by the 208 YAML vectors under tests/*/small.
//
//   y = p(z) = (z^n - 1)/n * sum_i b_i w_i / (z - w_i)        (z not in the domain)
//   y = b_k                                                    (z == w_k)
//   q_i = (b_i - y)/(w_i - z)                                  (i != k)
//   q_k = (1/z) * sum_{i != k} (b_i - y) w_i / (z - w_i)       (z == w_k)
//
// A lane owns a strided subset of the 4096 indices and inverts its (z - w_i) in
// groups of LE_INV_GROUP with Montgomery's simultaneous-inversion trick.
// Blob values stay canonical, roots / z / inverses are in Montgomery form, so
// fr_mul(mont, canon) is already canonical.
#pragma once
#include "field.cuh"

namespace lw {

constexpr int LE_INV_GROUP = 32;

// inv[k] = 1 / d[k] (Montgomery) for k < cnt; entries with d[k] == 0 give inv[k] = 0
LW_INL void fr_batch_inv_group(Fr* inv, const Fr* d, int cnt) {
  Fr pref[LE_INV_GROUP];
  Fr run = fr_one();
  for (int k = 0; k < cnt; k++) {
    pref[k] = run;
    if (!fr_is_zero(d[k])) run = fr_mul(run, d[k]);
  }
  Fr r = fr_inv(run);
  for (int k = cnt - 1; k >= 0; k--) {
    if (fr_is_zero(d[k])) { inv[k] = fr_zero(); continue; }
    inv[k] = fr_mul(r, pref[k]);
    r = fr_mul(r, d[k]);
  }
}

// z^4096 - 1 (Montgomery in, Montgomery out)
LW_INL Fr fr_zn_minus_one(Fr z_mont) {
  for (int i = 0; i < 12; i++) z_mont = fr_sqr(z_mont);
  return fr_sub(z_mont, fr_one());
}

}  // namespace lw
