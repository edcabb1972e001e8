// Synthetic blob generator and the integer-pipe peak probe used
// as the roofline denominator R_int.
#include <atomic>
#include "kernels.h"
#include "ptx.cuh"

namespace lw {

static std::atomic<uint64_t> g_launches{0};
uint64_t launches() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n); }

__host__ __device__ inline uint64_t splitmix64_next(uint64_t& state) {
  state += 0x9E3779B97F4A7C15ull;
  uint64_t z = state;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// word i of blob k: four big-endian u64 from SplitMix64(seed = 0xB2004844 ^ (k*4096+i)),
// then byte[0] &= 0x3f  (value < 2^254 < r: canonical big-endian scalar)
__host__ __device__ inline void synth_word(uint8_t* out32, uint64_t k, uint32_t i) {
  uint64_t st = 0xB2004844ull ^ (k * 4096ull + (uint64_t)i);
  for (int u = 0; u < 4; u++) {
    uint64_t v = splitmix64_next(st);
    for (int b = 0; b < 8; b++) out32[8 * u + b] = (uint8_t)(v >> (56 - 8 * b));
  }
  out32[0] &= 0x3f;
}

__global__ void synth_blobs_kernel(uint8_t* __restrict__ blobs, uint64_t first_blob, size_t n_words) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_words) return;
  uint64_t k = first_blob + g / 4096;
  uint32_t i = (uint32_t)(g % 4096);
  uint8_t w[32];
  synth_word(w, k, i);
  uint4* dst = reinterpret_cast<uint4*>(blobs + g * 32);
  uint32_t* ww = reinterpret_cast<uint32_t*>(w);
  dst[0] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
  dst[1] = make_uint4(ww[4], ww[5], ww[6], ww[7]);
}

void launch_synth_blobs(void* d_blobs, uint64_t first_blob, size_t n, cudaStream_t st) {
  size_t n_words = n * 4096;
  if (!n_words) return;
  synth_blobs_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>((uint8_t*)d_blobs, first_blob, n_words);
  count_launch();
}

extern "C" void lwkzg_synth_blob_host(uint8_t* blob, uint64_t k) {
  for (uint32_t i = 0; i < 4096; i++) synth_word(blob + 32 * i, k, i);
}

// ---- IMAD peak probe (roofline denominator R_int).  Operands are DISTINCT
// registers per MAC, as in a real multi-limb product: with a single shared
// multiplicand pair the operand-reuse cache makes IMAD.WIDE look twice as fast
// as it is in practice (profiles/r01_pipe_probe.md).
// variant 0: a carry chain of fused mad.lo.cc/madc.hi.cc pairs (IMAD.WIDE.U32.X,
//            exactly one row of the Montgomery product in mont.cuh)
// variant 1: carry-less 64-bit column accumulation (IMAD.WIDE.U32)
constexpr int PROBE_LIMBS = 12;
constexpr int PROBE_ITERS = 2048;

__global__ void __launch_bounds__(256) imad_probe_chain(uint32_t* out, uint32_t seed) {
  uint32_t a[PROBE_LIMBS], t[2 * PROBE_LIMBS];
  uint32_t b = seed * 3u + blockIdx.x + 1u;
#pragma unroll
  for (int i = 0; i < PROBE_LIMBS; i++) a[i] = seed + threadIdx.x * 31u + i * 977u;
#pragma unroll
  for (int i = 0; i < 2 * PROBE_LIMBS; i++) t[i] = i + threadIdx.x;
  for (int it = 0; it < PROBE_ITERS; it++) {
    t[0] = ptx::mad_lo_cc(a[0], b, t[0]);
    t[1] = ptx::madc_hi_cc(a[0], b, t[1]);
#pragma unroll
    for (int j = 1; j < PROBE_LIMBS; j++) {
      t[2 * j] = ptx::madc_lo_cc(a[j], b, t[2 * j]);
      t[2 * j + 1] = ptx::madc_hi_cc(a[j], b, t[2 * j + 1]);
    }
    b += t[3];
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 2 * PROBE_LIMBS; i++) acc ^= t[i];
  if (acc == 0x12345678u) out[0] = acc;  // keep the chains alive
}

__global__ void __launch_bounds__(256) imad_probe_wide(uint32_t* out, uint32_t seed) {
  uint32_t a[PROBE_LIMBS];
  unsigned long long c[PROBE_LIMBS];
  uint32_t b = seed * 3u + blockIdx.x + 1u;
#pragma unroll
  for (int i = 0; i < PROBE_LIMBS; i++) { a[i] = seed + threadIdx.x * 31u + i * 977u; c[i] = i + threadIdx.x; }
  for (int it = 0; it < PROBE_ITERS; it++) {
#pragma unroll
    for (int j = 0; j < PROBE_LIMBS; j++) c[j] += (unsigned long long)a[j] * b;
    b += (uint32_t)c[3];
  }
  unsigned long long x = 0;
#pragma unroll
  for (int i = 0; i < PROBE_LIMBS; i++) x ^= c[i];
  if (x == 0x12345678ull) out[0] = (uint32_t)x;
}

double run_imad_peak(int variant) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0.0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint32_t* d_out = nullptr;
  if (cudaMalloc(&d_out, 4) != cudaSuccess) return 0.0;
  const int blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    for (int k = 0; k < 4; k++) {
      if (variant == 0) imad_probe_chain<<<blocks, threads>>>(d_out, 12345u + rep);
      else imad_probe_wide<<<blocks, threads>>>(d_out, 12345u + rep);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 1 && ms < best) best = ms;
  }
  count_launch(24);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_out);
  if (cudaGetLastError() != cudaSuccess) return 0.0;
  double macs = 4.0 * (double)blocks * threads * (double)PROBE_ITERS * PROBE_LIMBS;
  return macs / (best * 1e-3);
}

}  // namespace lw
