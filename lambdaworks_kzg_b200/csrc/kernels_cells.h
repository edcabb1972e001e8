// Host-side launch interface of the PeerDAS / EIP-7594 cell kernels (cells.cu, cells_verify.cu).  Internal; the
// public boundary is include/lwkzg.h.  SURVEY §8 f4: the reference carries the 65 G2 points this path needs but
// implements none of it (/root/reference/src/srs.rs:274, src/lib.rs:60-92).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace lw {

constexpr int N_CELLS = 128;
constexpr int CELL_ELEMS = 64;
constexpr int CELL_BYTES = 2048;
constexpr int EXT_POINTS = 8192;          // FIELD_ELEMENTS_PER_EXT_BLOB = FK20 points (64 offsets x 128 frequencies)
constexpr int CELL_NAF_BYTES = 320;       // per 128th root of unity: 160 width-4 NAF digits (int8) of each GLV half

// ---- setup
void launch_cell_twiddles(void* d_tw8192, cudaStream_t st);                       // w^k, k < 8192 (Montgomery), w = FR_ROOT_8192
void launch_cell_twiddle_naf(void* d_naf, cudaStream_t st);                       // 128 x CELL_NAF_BYTES
// pts[v * 64 + b] = s_(64 (62 - v) + b) for v <= 62, infinity above: the 64 reversed, strided, zero-padded SRS columns
void launch_cell_srs_columns(void* d_pts_xyzz, const void* d_srs_monomial_aff, cudaStream_t st);
// out[(j / 64) * 4096 + b * 64 + j % 64] = affine(pts[brp7(j) * 64 + b]): the FK20 points in MSM order
void launch_cell_fk20_points(void* d_aff_out, const void* d_pts_xyzz, cudaStream_t st);

// ---- G1 FFT of size 128 over `batch` independent vectors, one stage per launch; pts[idx * batch + item] (XYZZ).
// A warp works on ONE butterfly index for 32 items, so the (fixed) twiddle's signed-digit ladder is warp-uniform.
// dif: natural in -> bit-reversed out, butterflies (P + Q, [w](P - Q)), half = 64 .. 1
// dit: bit-reversed in -> natural out, butterflies (P + [w]Q, P - [w]Q), half = 1 .. 64
void launch_cell_g1_fft_stage(void* d_pts, int batch, int half, bool dif, bool inverse, bool upper_half_zero, const void* d_naf, cudaStream_t st);

// ---- per blob
// blob bytes -> coefficient form (Montgomery, d_coef: n x 4096 x 32 B) and, if d_cells != NULL, the 128 cells
// (n x 128 x 2048 B, field elements in the mode's byte order).  mode: 0 = coefficients big-endian reduced (reference),
// 1 = evaluations little-endian, 2 = evaluations big-endian (Deneb).  Non-canonical words are the caller's to flag
// (launch_le_blob_check); here they are reduced.
void launch_cell_poly(void* d_coef, void* d_cells, const void* d_blobs, int n, int mode, const void* d_tw8192, cudaStream_t st);
// FK20 scalars: for every offset b the circulant vector of coefficient column b, DFT_128, scaled by 1/128, canonical.
// Frequency j = 64 v + j', offset b  ->  d_scalars[((v * n + blob) * 4096 + b * 64 + j') * 8 .. +8]: each half v is a
// "blob" of 4096 scalars for the fixed-base MSM kernels, over the half's own table of the points X[v][b * 64 + j']
void launch_cell_toeplitz(void* d_scalars, const void* d_coef, int n, const void* d_tw8192, cudaStream_t st);
// Hhat[j * n + blob] = sum_b scalar(blob, j, b) * X(j, b) over the two GLV digit tables (one per half of the frequencies,
// cell_table_half_entries(c) entries each, back to back).  n >= CELL_BA_MIN_BLOBS: the batched-affine kernel in its
// segmented form, two launches; below: one warp per (blob, frequency).  d_scratch: cell_msm_scratch_bytes(n)
constexpr int CELL_BA_MIN_BLOBS = 224;
size_t cell_table_half_entries(int c);
size_t cell_msm_scratch_bytes(int n);
void launch_cell_msm(void* d_pts_xyzz, const void* d_table, int c, const void* d_scalars, int n, void* d_scratch, cudaStream_t st);
// proofs[(blob * 128 + i) * 48] = compress(pts[brp7(i) * n + blob])
void launch_cell_proofs_finalize(void* d_proofs48, const void* d_pts_xyzz, int n, cudaStream_t st);

// ---- verification / recovery (cells_verify.cu)
// cell bytes -> 64 canonical field elements each (8 u32), status 1 if a word >= r
void launch_cell_parse(void* d_evals, int* d_status, const void* d_cells, int n_cells, int mode, cudaStream_t st);
// Scalars of the two MSMs of the verification equation.  Per cell k (one warp each): the interpolation polynomial of
// its 64 evaluations over the coset h_k <w64>, weighted by r^k -> d_wcoef[k][64] (Montgomery scratch); d_rpow[k] = r^k
// (canonical).  Then d_scalars_b (canonical, 8 u32 each):
//   [0, n) r^k h_k^64 | [n, n + nc) w_i = sum of r^k over the cells of commitment i | [n + nc, n + nc + 64) minus the
//   summed interpolation coefficients
// matching the point vector proofs || commitments || g1_monomial[0..64).
void launch_cell_verify_scalars(void* d_wcoef, void* d_rpow, void* d_scalars_b, const void* d_evals, const uint64_t* d_cell_indices,
                                const uint32_t* d_commitment_indices, int n_cells, int n_commitments, const void* d_r, const void* d_tw8192, cudaStream_t st);
// r = H(domain || 4096 || 64 || n_commitments || n_cells || commitments || (commitment_index || cell_index || cell || proof)...)
void launch_cell_batch_challenge(void* d_r, const void* d_commitments48, int n_commitments, const uint32_t* d_commitment_indices, const uint64_t* d_cell_indices,
                                 const void* d_cells, const void* d_proofs48, int n_cells, int mode, cudaStream_t st);
// recovery (one blob): evaluations of >= 64 cells -> coefficient form (Montgomery).  d_coef must hold 4096 + 3 * 8192 field
// elements (the tail is workspace); *d_status |= 1 if the cells are inconsistent with a polynomial of degree < 4096
void launch_cell_recover(void* d_coef, int* d_status, const void* d_evals, const uint64_t* d_cell_indices, int n_cells, const void* d_tw8192, cudaStream_t st);
// coefficient form (Montgomery) -> cells (both halves computed by FFT) -- the tail of launch_cell_poly for callers that already hold coefficients
void launch_cell_coef_to_cells(void* d_cells, const void* d_coef, int n, int mode, const void* d_tw8192, cudaStream_t st);

}  // namespace lw
