// Setup-time kernels: SRS import (the reference re-hydrates the SRS on EVERY
// call -- kzgsettings_to_structured_reference_string, /root/reference/src/
// srs.rs:258-280; here it happens once per KZGSettings and stays in HBM) and
// the fixed-base digit table used by msm.cu.
#include "g1.cuh"
#include "kernels.h"
#include "recode.cuh"

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

// Table layout (csrc/recode.cuh): window j < W - 1 holds 2^(c-1) multiples per point, the top window cnt_top:
//   entry(j, i, d) = ((j * npoints) << (c-1)) + i * cnt_j + (d - 1),  d = 1 .. cnt_j,  value d * 2^(c j) * P_i.
// One thread per (window, point, segment): a segment is a run of consecutive multiples; it starts from
// (d0 - 1) * B (one short double-and-add) and continues with the running sum in XYZZ, normalised to affine in
// batches of TB with Montgomery's simultaneous-inversion trick.
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

__global__ void __launch_bounds__(64) table_fill_kernel(uint4* __restrict__ table, const G1Affine* __restrict__ bases, int c, int nwin,
                                                        int npoints, uint32_t cnt_top, int segs) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_pairs = nwin * npoints;
  if (t >= n_pairs * segs) return;
  const int pair = t % n_pairs, seg = t / n_pairs;   // consecutive threads: consecutive points of one window
  const int j = pair / npoints, i = pair % npoints;
  const uint32_t half = 1u << (c - 1);
  const uint32_t cnt = (j == nwin - 1) ? cnt_top : half;
  const uint32_t per = (cnt + segs - 1) / segs;
  const uint32_t d_lo = seg * per, d_hi = min(cnt, d_lo + per);   // multiples d_lo + 1 .. d_hi
  if (d_lo >= d_hi) return;
  const G1Affine B = bases[pair];
  const size_t out_base = (((size_t)j * npoints) << (c - 1)) + (size_t)i * cnt;
  G1Xyzz acc = xyzz_inf();
  if (d_lo) {
    uint32_t kk[1] = {d_lo};
    acc = g1_mul_scalar(B, kk, 1);
  }
  G1Xyzz pts[TB];
  Fp pref[TB];
  for (uint32_t d0 = d_lo; d0 < d_hi; d0 += TB) {
    const int n = (d_hi - d0 < (uint32_t)TB) ? (int)(d_hi - d0) : TB;
    Fp run = fp_one();
    for (int k = 0; k < n; k++) {
      xyzz_madd(acc, B);
      pts[k] = acc;
      pref[k] = run;
      if (!xyzz_is_inf(acc)) run = fp_mul(run, acc.zzz);
    }
    Fp inv = fp_inv(run);
    for (int k = n - 1; k >= 0; k--) {
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

int table_num_windows(int c) { return glv_num_windows(c); }
uint32_t table_top_count(int c) { return glv_top_max(c) + 1u; }
unsigned long long table_entries(int c, int npoints) { return glv_table_entries(c, npoints); }

void launch_srs_import(void* d_aff_out, const void* d_canon_in, int* d_not_on_curve, int* d_not_in_subgroup, int n, cudaStream_t st) {
  srs_import_kernel<<<(n + 63) / 64, 64, 0, st>>>((G1Affine*)d_aff_out, (const uint32_t*)d_canon_in, d_not_on_curve, d_not_in_subgroup, n);
  count_launch();
}
void launch_table_bases(void* d_bases, const void* d_aff, int c, int nwin, int npoints, cudaStream_t st) {
  table_bases_kernel<<<(npoints + 31) / 32, 32, 0, st>>>((G1Affine*)d_bases, (const G1Affine*)d_aff, c, nwin, npoints);
  count_launch();
}
void launch_table_fill(void* d_table, const void* d_bases, int c, int nwin, int npoints, uint32_t cnt_top, cudaStream_t st) {
  const int n_pairs = nwin * npoints;
  // enough threads to fill the GPU whatever the window (16-bit windows: 8 x 4096 pairs of 32768+ multiples each)
  int segs = 1;
  while (n_pairs * segs < 200000 && (1 << (c - 1)) / (segs * 2) >= 4 * TB) segs *= 2;
  const long total = (long)n_pairs * segs;
  table_fill_kernel<<<(unsigned)((total + 63) / 64), 64, 0, st>>>((uint4*)d_table, (const G1Affine*)d_bases, c, nwin, npoints, cnt_top, segs);
  count_launch();
}

}  // namespace lw
