// Variable-base G1 multi-scalar multiplication (the reference's g1_lincomb =
// lambdaworks-math msm::pippenger::msm, /root/reference/src/lib.rs:241-243,
// 679-685; BASELINE config 5: N = 2^12 .. 2^22).
//
// Bucket method laid out for the GPU, no comparison sort:
//   1. points: canonical big-endian affine -> Montgomery (on-curve check)
//   2. scalars -> signed c-bit digits; per (window, bucket) histogram with
//      atomics, exclusive scan per window, scatter of point indices into
//      bucket order (order inside a bucket is irrelevant: group addition is
//      exact, so the sum is the same element whatever the order)
//   3. one thread per (window, bucket): XYZZ mixed additions over its run
//   4. per window: sum_b b * B_b by chunked running sums (each thread owns a
//      chunk of buckets, adds (lo-1) * chunk_sum with a short double-and-add)
//      and a block tree
//   5. Horner over the windows, normalise, compress
// The signed digit set and window size are internal choices; the result is the
// same group element as the reference's unsigned-window Pippenger.
#include "g1.cuh"
#include "kernels.h"

namespace lw {

constexpr int VM_RED_THREADS = 128;

__host__ __device__ inline int vm_window_bits(size_t n) {
  int lg = 0;
  while ((size_t(1) << (lg + 1)) <= n) lg++;
  int c = lg - 3;
  if (c < 4) c = 4;
  if (c > 16) c = 16;
  return c;
}
__host__ __device__ inline int vm_num_windows(int c) { return 255 / c + 1; }

struct VmLayout {
  size_t pts, counts, offsets, cursors, idx, buckets, winsums, heavy, bad, total;
  int c, W, B;  // B = 2^(c-1) buckets per window (bucket b holds digit magnitude b+1)
};
static VmLayout vm_layout(size_t n) {
  VmLayout L;
  L.c = vm_window_bits(n ? n : 1);
  L.W = vm_num_windows(L.c);
  L.B = 1 << (L.c - 1);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
  L.pts = take((n ? n : 1) * 96);
  L.counts = take((size_t)L.W * L.B * 4);
  L.offsets = take((size_t)L.W * L.B * 4);
  L.cursors = take((size_t)L.W * L.B * 4);
  L.idx = take((n ? n : 1) * (size_t)L.W * 4);
  L.buckets = take((size_t)L.W * L.B * sizeof(G1Xyzz));
  L.winsums = take((size_t)L.W * sizeof(G1Xyzz));
  L.heavy = take(((size_t)L.W * L.B + 1) * 4);  // [0] = count, then bucket ids whose run is too long for one thread
  L.bad = take(256);
  L.total = off;
  return L;
}
size_t var_msm_scratch_bytes(size_t n) { return vm_layout(n).total; }

__device__ __forceinline__ int vm_digit(const uint32_t* k8, int c, int j, int& carry) {
  const int bit = j * c;
  const int w = bit >> 5, s = bit & 31;
  uint32_t lo = k8[w];
  uint32_t hi = (w + 1 < 8) ? k8[w + 1] : 0u;
  uint32_t raw = __funnelshift_r(lo, hi, s) & ((1u << c) - 1u);
  int d = (int)raw + carry;
  if (d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else { carry = 0; }
  return d;
}

__global__ void vm_convert_kernel(G1Affine* __restrict__ pts, int* __restrict__ bad, const uint8_t* __restrict__ pts_be, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* b = pts_be + i * 96;
  bool z = true;
  for (int k = 0; k < 96; k++) z = z && (b[k] == 0);
  G1Affine p = g1a_inf();
  if (!z) {
    p.x = fp_from_be48(b);
    p.y = fp_from_be48(b + 48);
    if (!g1a_on_curve(p)) atomicExch(bad, 1);
  }
  pts[i] = p;
}

// PHASE 0: histogram, PHASE 1: scatter
template <int PHASE>
__global__ void vm_digits_kernel(uint32_t* __restrict__ counts_or_cursors, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ idx,
                                 const uint8_t* __restrict__ sc_be, size_t n, int c, int W, int B) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = fr_canon_from_be32(sc_be + i * 32);  // reduced mod r like the reference's from_bytes_be
  uint32_t kk[8];
  for (int t = 0; t < 8; t++) kk[t] = k.l[t];
  int carry = 0;
  for (int j = 0; j < W; j++) {
    int d = vm_digit(kk, c, j, carry);
    if (d == 0) continue;
    uint32_t b = (uint32_t)((d < 0 ? -d : d) - 1);
    size_t slot = (size_t)j * B + b;
    if (PHASE == 0) {
      atomicAdd(&counts_or_cursors[slot], 1u);
    } else {
      uint32_t pos = atomicAdd(&counts_or_cursors[slot], 1u);
      idx[(size_t)j * n + offsets[slot] + pos] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);
    }
  }
}

// exclusive scan of one window's B counts (one block per window)
__global__ void __launch_bounds__(1024) vm_scan_kernel(uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts, int B) {
  __shared__ uint32_t part[1024];
  const int j = blockIdx.x, t = threadIdx.x;
  const int per = (B + 1023) / 1024;
  const uint32_t* c = counts + (size_t)j * B;
  uint32_t* o = offsets + (size_t)j * B;
  uint32_t s = 0;
  for (int k = 0; k < per; k++) {
    int b = t * per + k;
    if (b < B) s += c[b];
  }
  part[t] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    uint32_t v = (t >= d) ? part[t - d] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = part[t] - s;
  for (int k = 0; k < per; k++) {
    int b = t * per + k;
    if (b < B) { o[b] = run; run += c[b]; }
  }
}

// one thread per (window, bucket).  Runs longer than VM_HEAVY entries (skewed
// digit distributions: the top window only holds 255 mod c scalar bits, equal
// scalars, ...) are deferred to vm_heavy_kernel, which puts a whole block on
// each of them.
constexpr uint32_t VM_HEAVY = 384;
constexpr int VM_HEAVY_THREADS = 128;
__global__ void __launch_bounds__(128, 3) vm_accumulate_kernel(G1Xyzz* __restrict__ buckets, uint32_t* __restrict__ heavy,
                                                               const uint32_t* __restrict__ idx, const uint32_t* __restrict__ offsets,
                                                               const uint32_t* __restrict__ counts, const G1Affine* __restrict__ pts, size_t n,
                                                               int W, int B) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)W * B) return;
  const int j = (int)(t / B);
  const uint32_t* run = idx + (size_t)j * n + offsets[t];
  const uint32_t cnt = counts[t];
  G1Xyzz acc = xyzz_inf();
  if (cnt > VM_HEAVY) {
    uint32_t slot = atomicAdd(&heavy[0], 1u);
    heavy[1 + slot] = (uint32_t)t;
  } else {
    for (uint32_t k = 0; k < cnt; k++) {
      uint32_t e = run[k];
      G1Affine p = pts[e & 0x7fffffffu];
      p.y = fp_cneg(p.y, (e >> 31) != 0);
      xyzz_madd_hot(acc, p);
    }
  }
  buckets[t] = acc;
}

__device__ __forceinline__ void vm_to_smem(uint32_t* smem, int stride, int t, const G1Xyzz& p) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 48; i++) smem[i * stride + t] = w[i];
}
__device__ __forceinline__ G1Xyzz vm_from_smem(const uint32_t* smem, int stride, int t) {
  G1Xyzz p;
  uint32_t* w = reinterpret_cast<uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 48; i++) w[i] = smem[i * stride + t];
  return p;
}

// one block per heavy bucket (grid-stride over the heavy list)
__global__ void __launch_bounds__(VM_HEAVY_THREADS, 3) vm_heavy_kernel(G1Xyzz* __restrict__ buckets, const uint32_t* __restrict__ heavy,
                                                                       const uint32_t* __restrict__ idx, const uint32_t* __restrict__ offsets,
                                                                       const uint32_t* __restrict__ counts, const G1Affine* __restrict__ pts,
                                                                       size_t n, int B) {
  __shared__ uint32_t red[48 * (VM_HEAVY_THREADS / 2)];
  const uint32_t nheavy = heavy[0];
  const int tid = threadIdx.x;
  for (uint32_t h = blockIdx.x; h < nheavy; h += gridDim.x) {
    const uint32_t t = heavy[1 + h];
    const int j = (int)(t / (uint32_t)B);
    const uint32_t* run = idx + (size_t)j * n + offsets[t];
    const uint32_t cnt = counts[t];
    G1Xyzz acc = xyzz_inf();
    for (uint32_t k = tid; k < cnt; k += VM_HEAVY_THREADS) {
      uint32_t e = run[k];
      G1Affine p = pts[e & 0x7fffffffu];
      p.y = fp_cneg(p.y, (e >> 31) != 0);
      xyzz_madd_hot(acc, p);
    }
    for (int s = VM_HEAVY_THREADS / 2; s > 0; s >>= 1) {
      if (tid >= s && tid < 2 * s) vm_to_smem(red, VM_HEAVY_THREADS / 2, tid - s, acc);
      __syncthreads();
      if (tid < s) {
        G1Xyzz o = vm_from_smem(red, VM_HEAVY_THREADS / 2, tid);
        xyzz_add_ni(acc, o);
      }
      __syncthreads();
    }
    if (tid == 0) buckets[t] = acc;
  }
}

// window sum  S_j = sum_b (b+1) * bucket[j][b]; one block per window
__global__ void __launch_bounds__(VM_RED_THREADS) vm_reduce_kernel(G1Xyzz* __restrict__ winsums, const G1Xyzz* __restrict__ buckets, int B) {
  __shared__ uint32_t red[48 * (VM_RED_THREADS / 2)];
  const int j = blockIdx.x, t = threadIdx.x;
  const G1Xyzz* bk = buckets + (size_t)j * B;
  const int per = (B + VM_RED_THREADS - 1) / VM_RED_THREADS;
  const int lo = t * per, hi = min(B, lo + per);  // buckets [lo, hi): weights lo+1 .. hi
  G1Xyzz run = xyzz_inf(), acc = xyzz_inf();
  for (int b = hi - 1; b >= lo; b--) {
    G1Xyzz v = bk[b];
    xyzz_add_ni(run, v);
    xyzz_add_ni(acc, run);  // acc = sum (b - lo + 1) * B_b
  }
  // + lo * run   (double-and-add over the bits of lo)
  if (lo > 0 && lo < hi) {
    G1Xyzz m = xyzz_inf();
    for (int bit = 30; bit >= 0; bit--) {
      xyzz_dbl_ni(m);
      if ((lo >> bit) & 1) xyzz_add_ni(m, run);
    }
    xyzz_add_ni(acc, m);
  }
  for (int s = VM_RED_THREADS / 2; s > 0; s >>= 1) {
    if (t >= s && t < 2 * s) vm_to_smem(red, VM_RED_THREADS / 2, t - s, acc);
    __syncthreads();
    if (t < s) {
      G1Xyzz o = vm_from_smem(red, VM_RED_THREADS / 2, t);
      xyzz_add_ni(acc, o);
    }
    __syncthreads();
  }
  if (t == 0) winsums[j] = acc;
}

// result = sum_j 2^(c j) S_j  (Horner from the top window), normalise, compress
__global__ void vm_final_kernel(uint8_t* __restrict__ out48, const G1Xyzz* __restrict__ winsums, int c, int W) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  G1Xyzz acc = xyzz_inf();
  for (int j = W - 1; j >= 0; j--) {
    for (int k = 0; k < c; k++) xyzz_dbl_ni(acc);
    G1Xyzz s = winsums[j];
    xyzz_add_ni(acc, s);
  }
  g1_compress(out48, xyzz_to_affine(acc));
}

void launch_var_msm(void* d_out48, const void* d_points_xy_be, const void* d_scalars_be, size_t n, void* d_scratch, cudaStream_t st) {
  VmLayout L = vm_layout(n);
  uint8_t* base = (uint8_t*)d_scratch;
  G1Affine* pts = (G1Affine*)(base + L.pts);
  uint32_t* counts = (uint32_t*)(base + L.counts);
  uint32_t* offsets = (uint32_t*)(base + L.offsets);
  uint32_t* cursors = (uint32_t*)(base + L.cursors);
  uint32_t* idx = (uint32_t*)(base + L.idx);
  G1Xyzz* buckets = (G1Xyzz*)(base + L.buckets);
  G1Xyzz* winsums = (G1Xyzz*)(base + L.winsums);
  uint32_t* heavy = (uint32_t*)(base + L.heavy);
  int* bad = (int*)(base + L.bad);
  cudaMemsetAsync(base + L.counts, 0, L.idx - L.counts, st);  // counts, offsets, cursors
  cudaMemsetAsync(bad, 0, sizeof(int), st);
  cudaMemsetAsync(heavy, 0, sizeof(uint32_t), st);
  if (n) {
    unsigned blocks = (unsigned)((n + 127) / 128);
    vm_convert_kernel<<<blocks, 128, 0, st>>>(pts, bad, (const uint8_t*)d_points_xy_be, n);
    vm_digits_kernel<0><<<blocks, 128, 0, st>>>(counts, nullptr, nullptr, (const uint8_t*)d_scalars_be, n, L.c, L.W, L.B);
    vm_scan_kernel<<<L.W, 1024, 0, st>>>(offsets, counts, L.B);
    vm_digits_kernel<1><<<blocks, 128, 0, st>>>(cursors, offsets, idx, (const uint8_t*)d_scalars_be, n, L.c, L.W, L.B);
    count_launch(4);
  }
  size_t nb = (size_t)L.W * L.B;
  vm_accumulate_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(buckets, heavy, idx, offsets, counts, pts, n ? n : 1, L.W, L.B);
  vm_heavy_kernel<<<1184, VM_HEAVY_THREADS, 0, st>>>(buckets, heavy, idx, offsets, counts, pts, n ? n : 1, L.B);
  vm_reduce_kernel<<<L.W, VM_RED_THREADS, 0, st>>>(winsums, buckets, L.B);
  vm_final_kernel<<<1, 32, 0, st>>>((uint8_t*)d_out48, winsums, L.c, L.W);
  count_launch(4);
}
// ---- synthetic inputs for the size sweep (BASELINE config 5): point t is the
// fixed-base table entry number (t * 2654435761) mod n_entries, i.e. a valid
// curve point with a KNOWN discrete log d * 2^(c j) * tau^i when the setup's tau
// is known (tests), exported in the public input format (canonical big-endian
// affine); scalar t = the synthetic blob word generator with blob id `seed`.
__host__ __device__ inline uint64_t vm_splitmix(uint64_t& state) {
  state += 0x9E3779B97F4A7C15ull;
  uint64_t z = state;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void vm_synth_kernel(uint8_t* __restrict__ pts_be, uint8_t* __restrict__ sc_be, const uint4* __restrict__ table,
                                unsigned long long n_entries, unsigned long long seed, size_t n) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  unsigned long long e = ((unsigned long long)t * 2654435761ull) % n_entries;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(table + e * 6);
  G1Affine p;
  for (int k = 0; k < 12; k++) { p.x.l[k] = w[k]; p.y.l[k] = w[12 + k]; }
  uint8_t* o = pts_be + t * 96;
  if (g1a_is_inf(p)) {
    for (int k = 0; k < 96; k++) o[k] = 0;
  } else {
    fp_canon_to_be48(o, fp_from_mont(p.x));
    fp_canon_to_be48(o + 48, fp_from_mont(p.y));
  }
  uint64_t st = 0xB2004844ull ^ (seed * 4096ull + (uint64_t)t);
  uint8_t* sc = sc_be + t * 32;
  for (int u = 0; u < 4; u++) {
    uint64_t v = vm_splitmix(st);
    for (int b = 0; b < 8; b++) sc[8 * u + b] = (uint8_t)(v >> (56 - 8 * b));
  }
  sc[0] &= 0x3f;
}
void launch_var_msm_synth(void* d_pts_be, void* d_sc_be, const void* d_table, unsigned long long n_entries, unsigned long long seed, size_t n,
                          cudaStream_t st) {
  if (!n) return;
  vm_synth_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((uint8_t*)d_pts_be, (uint8_t*)d_sc_be, (const uint4*)d_table, n_entries, seed, n);
  count_launch();
}
int var_msm_window_bits(size_t n) { return vm_window_bits(n ? n : 1); }

// offset of the "a point was not on the curve" flag inside the scratch buffer
size_t var_msm_bad_flag_offset(size_t n) { return vm_layout(n).bad; }

}  // namespace lw
