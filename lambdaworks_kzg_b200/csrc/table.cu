// Setup-time kernels: SRS import (the reference re-hydrates the SRS on EVERY
// call -- kzgsettings_to_structured_reference_string, /root/reference/src/
// srs.rs:258-280; here it happens once per KZGSettings and stays in HBM) and
// the fixed-base digit table used by msm.cu.
#include "g1.cuh"
#include "kernels.h"

namespace lw {

// canonical little-endian limbs (x[12] || y[12]) -> Montgomery affine + checks
// (blst_p1_to_g1_point, src/srs.rs:155-172: from_affine == curve equation, so
// the reference's infinity encoding x = y = 0 is rejected there too).
__global__ void srs_import_kernel(G1Affine* __restrict__ out, const uint32_t* __restrict__ in, int* __restrict__ not_on_curve,
                                  int* __restrict__ not_in_subgroup, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp x, y;
  for (int k = 0; k < 12; k++) { x.l[k] = in[i * 24 + k]; y.l[k] = in[i * 24 + 12 + k]; }
  // from_bytes_be reduces silently (App. D.1): values < 2^384 need up to 9
  // subtractions of p; the loaders only ever store canonical values.
  mod_reduce_small<FpCfg, 9>(x.l);
  mod_reduce_small<FpCfg, 9>(y.l);
  G1Affine p;
  p.x = fp_to_mont(x);
  p.y = fp_to_mont(y);
  bool oc = g1a_on_curve(p);
  out[i] = p;
  not_on_curve[i] = oc ? 0 : 1;
  not_in_subgroup[i] = (oc && g1_in_subgroup(p)) ? 0 : 1;
}

// bases[j * npoints + i] = 2^(c j) * P_i
__global__ void table_bases_kernel(G1Affine* __restrict__ bases, const G1Affine* __restrict__ pts, int c, int nwin, int npoints) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npoints) return;
  G1Affine p = pts[i];
  G1Xyzz acc = xyzz_from_affine(p);
  for (int j = 0; j < nwin; j++) {
    G1Affine a = (j == 0) ? p : xyzz_to_affine(acc);
    bases[(size_t)j * npoints + i] = a;
    if (j + 1 < nwin)
      for (int k = 0; k < c; k++) xyzz_dbl_ni(acc);
  }
}

// One thread per (window, point): running sum d*B in XYZZ, normalised to
// affine in batches of TB with Montgomery's simultaneous-inversion trick.
constexpr int TB = 32;

__device__ __forceinline__ void store_entry(uint4* __restrict__ table, size_t idx, const G1Affine& e) {
  uint4* p = table + idx * 6;
  p[0] = make_uint4(e.x.l[0], e.x.l[1], e.x.l[2], e.x.l[3]);
  p[1] = make_uint4(e.x.l[4], e.x.l[5], e.x.l[6], e.x.l[7]);
  p[2] = make_uint4(e.x.l[8], e.x.l[9], e.x.l[10], e.x.l[11]);
  p[3] = make_uint4(e.y.l[0], e.y.l[1], e.y.l[2], e.y.l[3]);
  p[4] = make_uint4(e.y.l[4], e.y.l[5], e.y.l[6], e.y.l[7]);
  p[5] = make_uint4(e.y.l[8], e.y.l[9], e.y.l[10], e.y.l[11]);
}

__global__ void __launch_bounds__(64) table_fill_kernel(uint4* __restrict__ table, const G1Affine* __restrict__ bases, int c, int n_pairs) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  const G1Affine B = bases[t];
  const size_t out_base = (size_t)t << (c - 1);
  const int half = 1 << (c - 1);
  G1Xyzz acc = xyzz_inf();
  G1Xyzz pts[TB];
  Fp pref[TB];
  for (int d0 = 0; d0 < half; d0 += TB) {
    const int cnt = (half - d0 < TB) ? (half - d0) : TB;
    Fp run = fp_one();
    for (int k = 0; k < cnt; k++) {
      xyzz_madd(acc, B);
      pts[k] = acc;
      pref[k] = run;
      if (!xyzz_is_inf(acc)) run = fp_mul(run, acc.zzz);
    }
    Fp inv = fp_inv(run);
    for (int k = cnt - 1; k >= 0; k--) {
      G1Affine e;
      if (xyzz_is_inf(pts[k])) {
        e = g1a_inf();
      } else {
        Fp zzz_inv = fp_mul(inv, pref[k]);
        inv = fp_mul(inv, pts[k].zzz);
        Fp tt = fp_mul(pts[k].zz, zzz_inv);
        Fp zz_inv = fp_sqr(tt);
        e.x = fp_mul(pts[k].x, zz_inv);
        e.y = fp_mul(pts[k].y, zzz_inv);
      }
      store_entry(table, out_base + (size_t)(d0 + k), e);
    }
  }
}

void launch_srs_import(void* d_aff_out, const void* d_canon_in, int* d_not_on_curve, int* d_not_in_subgroup, int n, cudaStream_t st) {
  srs_import_kernel<<<(n + 63) / 64, 64, 0, st>>>((G1Affine*)d_aff_out, (const uint32_t*)d_canon_in, d_not_on_curve, d_not_in_subgroup, n);
  count_launch();
}
void launch_table_bases(void* d_bases, const void* d_aff, int c, int nwin, int npoints, cudaStream_t st) {
  table_bases_kernel<<<(npoints + 31) / 32, 32, 0, st>>>((G1Affine*)d_bases, (const G1Affine*)d_aff, c, nwin, npoints);
  count_launch();
}
void launch_table_fill(void* d_table, const void* d_bases, int c, int nwin, int npoints, cudaStream_t st) {
  int n_pairs = nwin * npoints;
  table_fill_kernel<<<(n_pairs + 63) / 64, 64, 0, st>>>((uint4*)d_table, (const G1Affine*)d_bases, c, n_pairs);
  count_launch();
}

}  // namespace lw
