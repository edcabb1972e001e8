// MODE_CKZG_LE kernels: the little-endian-era
// c-kzg-4844 semantics that the YAML vectors under /root/reference/tests/*/small
// encode -- canonical little-endian field elements, blob = evaluations over the
// bit-reversed 4096th roots of unity, Lagrange-form SRS (the conversion the
// reference left as a TODO at src/lib.rs:760-770 / src/srs.rs:117-124),
// barycentric evaluation and evaluation-form quotient.  The MSM, SHA-256, codec
// and pairing kernels are shared with the reference mode.
#include "g1.cuh"
#include "kernels.h"
#include "lepoly.cuh"

namespace lw {

__device__ __forceinline__ uint32_t brp12(uint32_t i) { return __brev(i) >> 20; }

__device__ __forceinline__ Fr fr_pow_u32(Fr base_mont, uint32_t e) {
  Fr acc = fr_one();
  for (int bit = 31; bit >= 0; bit--) {
    acc = fr_sqr(acc);
    if ((e >> bit) & 1u) acc = fr_mul(acc, base_mont);
  }
  return acc;
}
__device__ __forceinline__ Fr fr_root() { Fr w; for (int i = 0; i < 8; i++) w.l[i] = k::FR_ROOT_4096[i]; return w; }

// roots[i] = w^brp(i), Montgomery
__global__ void le_roots_kernel(Fr* __restrict__ roots) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N_POINTS) return;
  roots[i] = fr_pow_u32(fr_root(), brp12(i));
}

// Row i of the inverse DFT that maps the monomial SRS to the Lagrange basis in
// bit-reversed order:  L_i(tau) G = sum_j (1/n) w_i^(-j) [tau^j]G.  Rows are written
// as canonical little-endian scalars, i.e. as 4096 "blobs" for the fixed-base MSM.
__global__ void le_idft_rows_kernel(uint32_t* __restrict__ rows) {
  uint32_t i = blockIdx.x;         // row
  uint32_t t = threadIdx.x;        // 256 threads, 16 columns each
  Fr winv = fr_pow_u32(fr_root(), (N_POINTS - brp12(i)) & (N_POINTS - 1));  // w_i^-1
  Fr ninv; for (int k = 0; k < 8; k++) ninv.l[k] = k::FR_N_INV[k];
  Fr cur = fr_mul(fr_pow_u32(winv, t * 16), ninv);  // (1/n) w_i^(-16 t), Montgomery
  for (int u = 0; u < 16; u++) {
    Fr c = fr_from_mont(cur);
    uint32_t* dst = rows + ((size_t)i * N_POINTS + t * 16 + u) * 8;
    for (int k = 0; k < 8; k++) dst[k] = c.l[k];
    cur = fr_mul(cur, winv);
  }
}

// Montgomery affine -> canonical little-endian limbs x[12] || y[12] (all-zero = infinity)
__global__ void affine_to_canon_kernel(uint32_t* __restrict__ out24, const G1Affine* __restrict__ pts, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = pts[i];
  Fp x = fp_zero(), y = fp_zero();
  if (!g1a_is_inf(p)) { x = fp_from_mont(p.x); y = fp_from_mont(p.y); }
  for (int k = 0; k < 12; k++) { out24[i * 24 + k] = x.l[k]; out24[i * 24 + 12 + k] = y.l[k]; }
}

// status[b] = 1 (C_KZG_BADARGS) if any of the 4096 words is >= r.  be = 0: little-endian words (MODE_CKZG_LE),
// be = 1: big-endian words (MODE_DENEB, bytes_to_bls_field of the final spec)
__global__ void __launch_bounds__(128) le_blob_check_kernel(int* __restrict__ status, const uint8_t* __restrict__ blobs, int n, int be) {
  const int b = blockIdx.x;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(blobs + (size_t)b * BLOB_BYTES);
  bool bad = false;
  for (int i = threadIdx.x; i < N_POINTS; i += 128) {
    uint32_t v[8];
    for (int k = 0; k < 8; k++) v[k] = be ? bswap32(w[i * 8 + 7 - k]) : w[i * 8 + k];
    if (!limbs_lt<8>(v, k::FR_MOD)) bad = true;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) status[b] = 1;
}

// 32 little-endian (be = 0) or big-endian (be = 1) bytes -> canonical limbs; status 1 (BADARGS) if >= r
__global__ void le_fr_parse_kernel(uint32_t* __restrict__ out, int* __restrict__ status, const uint8_t* __restrict__ in, int n, int be) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t v[8];
  for (int k = 0; k < 8; k++) {
    const uint8_t* q = in + (size_t)i * 32 + 4 * (be ? 7 - k : k);
    v[k] = be ? ((uint32_t)q[3] | ((uint32_t)q[2] << 8) | ((uint32_t)q[1] << 16) | ((uint32_t)q[0] << 24))
              : ((uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24));
  }
  if (!limbs_lt<8>(v, k::FR_MOD)) { if (status) status[i] = 1; for (int k = 0; k < 8; k++) v[k] = 0; }
  for (int k = 0; k < 8; k++) out[i * 8 + k] = v[k];
}

__device__ __forceinline__ Fr shfl_xor_fr(const Fr& v, int m) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], m);
  return r;
}
__device__ __forceinline__ Fr warp_sum_fr(Fr v) {
  for (int m = 16; m > 0; m >>= 1) v = fr_add(v, shfl_xor_fr(v, m));
  return v;
}

// One warp per blob; lane owns indices i = lane + 32 m.  q_out doubles as the
// scratch that carries 1/(z - w_i) from the first sweep to the second.
constexpr int LE_WARPS = 4;
__global__ void __launch_bounds__(LE_WARPS * 32) le_eval_quot_kernel(uint32_t* __restrict__ q_out, uint32_t* __restrict__ y_out,
                                                                      uint8_t* __restrict__ y_le_out, const uint8_t* __restrict__ blobs,
                                                                      const uint32_t* __restrict__ zs, const Fr* __restrict__ roots, int n, int be) {
  const int lane = threadIdx.x & 31;
  const int blob = blockIdx.x * LE_WARPS + (threadIdx.x >> 5);
  if (blob >= n) return;
  const uint32_t* bw = reinterpret_cast<const uint32_t*>(blobs + (size_t)blob * BLOB_BYTES);
  uint32_t* q = q_out ? q_out + (size_t)blob * N_POINTS * 8 : nullptr;
  Fr zc;
  for (int i = 0; i < 8; i++) zc.l[i] = zs[blob * 8 + i];
  const Fr z = fr_to_mont(zc);

  // blob words are canonical by contract (le_blob_check flags the blobs where they are not; their outputs are discarded)
  auto load_b = [&](int i) { Fr b; const uint4* p = reinterpret_cast<const uint4*>(bw + i * 8); uint4 a = __ldg(p), c = __ldg(p + 1);
                             if (be) { b.l[7] = bswap32(a.x); b.l[6] = bswap32(a.y); b.l[5] = bswap32(a.z); b.l[4] = bswap32(a.w);
                                       b.l[3] = bswap32(c.x); b.l[2] = bswap32(c.y); b.l[1] = bswap32(c.z); b.l[0] = bswap32(c.w); }
                             else { b.l[0] = a.x; b.l[1] = a.y; b.l[2] = a.z; b.l[3] = a.w; b.l[4] = c.x; b.l[5] = c.y; b.l[6] = c.z; b.l[7] = c.w; }
                             return b; };

  // ---- sweep 1: inverses of (z - w_i), barycentric sum, detect z in the domain
  Fr acc = fr_zero();
  int found = -1;
  for (int g = 0; g < N_POINTS / 32 / LE_INV_GROUP; g++) {
    Fr d[LE_INV_GROUP], inv[LE_INV_GROUP];
    for (int k = 0; k < LE_INV_GROUP; k++) {
      int i = lane + 32 * (g * LE_INV_GROUP + k);
      d[k] = fr_sub(z, roots[i]);
      if (fr_is_zero(d[k])) found = i;
    }
    fr_batch_inv_group(inv, d, LE_INV_GROUP);
    for (int k = 0; k < LE_INV_GROUP; k++) {
      int i = lane + 32 * (g * LE_INV_GROUP + k);
      if (q) for (int t = 0; t < 8; t++) q[i * 8 + t] = inv[k].l[t];
      acc = fr_add(acc, fr_mul(fr_mul(roots[i], inv[k]), load_b(i)));  // canonical
    }
  }
  // any lane found z == w_k ?
  int kfound = found;
  for (int m = 16; m > 0; m >>= 1) kfound = max(kfound, __shfl_xor_sync(0xffffffffu, kfound, m));
  Fr y;
  if (kfound >= 0) {
    y = load_b(kfound);
  } else {
    Fr ninv; for (int k = 0; k < 8; k++) ninv.l[k] = k::FR_N_INV[k];
    Fr scale = fr_mul(fr_zn_minus_one(z), ninv);  // Montgomery
    y = fr_mul(scale, warp_sum_fr(acc));          // canonical
  }
  if (lane == 0) {
    if (y_out) for (int t = 0; t < 8; t++) y_out[blob * 8 + t] = y.l[t];
    if (y_le_out) for (int t = 0; t < 8; t++) for (int bb = 0; bb < 4; bb++)
      y_le_out[(size_t)blob * 32 + (be ? 31 - (4 * t + bb) : 4 * t + bb)] = (uint8_t)(y.l[t] >> (8 * bb));
  }
  if (!q_out) return;
  __syncwarp();
  // ---- sweep 2: q_i = (y - b_i) / (z - w_i); collect sum q_i w_i for the z == w_k case
  Fr t_acc = fr_zero();
  for (int m = 0; m < N_POINTS / 32; m++) {
    int i = lane + 32 * m;
    Fr inv;
    for (int t = 0; t < 8; t++) inv.l[t] = q[i * 8 + t];
    Fr qi = fr_mul(inv, fr_sub(y, load_b(i)));  // canonical; 0 when i == kfound (inv == 0)
    for (int t = 0; t < 8; t++) q[i * 8 + t] = qi.l[t];
    if (kfound >= 0) t_acc = fr_add(t_acc, fr_mul(roots[i], qi));
  }
  if (kfound >= 0) {
    // q_k = (1/z) sum_{i != k} (b_i - y) w_i / (z - w_i) = -(1/z) sum_{i != k} q_i w_i
    Fr tot = warp_sum_fr(t_acc);
    if (lane == 0) {
      Fr qk = fr_mul(fr_neg(fr_inv(z)), tot);
      for (int t = 0; t < 8; t++) q[kfound * 8 + t] = qk.l[t];
    }
  }
}

__global__ void write_generator_kernel(G1Affine* out) { if (threadIdx.x == 0 && blockIdx.x == 0) *out = g1a_generator(); }
void launch_write_generator(void* d_out, cudaStream_t st) { write_generator_kernel<<<1, 32, 0, st>>>((G1Affine*)d_out); count_launch(); }
void launch_le_roots(void* d_roots, cudaStream_t st) { le_roots_kernel<<<N_POINTS / 128, 128, 0, st>>>((Fr*)d_roots); count_launch(); }
void launch_le_idft_rows(void* d_rows, cudaStream_t st) { le_idft_rows_kernel<<<N_POINTS, 256, 0, st>>>((uint32_t*)d_rows); count_launch(); }
void launch_affine_to_canon(void* d_out24, const void* d_aff, int n, cudaStream_t st) {
  affine_to_canon_kernel<<<(n + 63) / 64, 64, 0, st>>>((uint32_t*)d_out24, (const G1Affine*)d_aff, n);
  count_launch();
}
void launch_le_blob_check(int* d_status, const void* d_blobs, int n, cudaStream_t st, bool be) {
  if (n <= 0) return;
  LW_SAME_CARVEOUT(le_blob_check_kernel);
  le_blob_check_kernel<<<n, 128, 0, st>>>(d_status, (const uint8_t*)d_blobs, n, be ? 1 : 0);
  count_launch();
}
void launch_le_fr_parse(void* d_out, int* d_status, const void* d_in32, int n, cudaStream_t st, bool be) {
  if (n <= 0) return;
  le_fr_parse_kernel<<<(n + 63) / 64, 64, 0, st>>>((uint32_t*)d_out, d_status, (const uint8_t*)d_in32, n, be ? 1 : 0);
  count_launch();
}
void launch_le_eval_quot(void* d_q, void* d_y, void* d_y_le32, const void* d_blobs, const void* d_z, const void* d_roots, int n, cudaStream_t st, bool be) {
  if (n <= 0) return;
  LW_SAME_CARVEOUT(le_eval_quot_kernel);
  le_eval_quot_kernel<<<(n + LE_WARPS - 1) / LE_WARPS, LE_WARPS * 32, 0, st>>>((uint32_t*)d_q, (uint32_t*)d_y, (uint8_t*)d_y_le32, (const uint8_t*)d_blobs,
                                                                              (const uint32_t*)d_z, (const Fr*)d_roots, n, be ? 1 : 0);
  count_launch();
}

}  // namespace lw
