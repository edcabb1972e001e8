// Field-multiplication throughput on one B200: IMAD multiplier (csrc/mont.cuh), its
// Karatsuba variant, the FP64-pipe multiplier (csrc/fpdp.cuh) and the ways of
// running both pipes at once (warp-interleaved and fused pairs).  No memory traffic.
// Also checks on the device that all variants produce identical bits.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../lambdaworks_kzg_b200/csrc/g1.cuh"
#include "fpdp.cuh"
#include "karatsuba.cuh"
using namespace lw;
#define NITER 256

struct Fp2x { Fp a, b; };
__device__ __noinline__ Fp mul_i(Fp a, Fp b) { Fp r; mont_mul<FpCfg>(r.l, a.l, b.l); return r; }
__device__ __noinline__ Fp mul_k(Fp a, Fp b) { Fp r; mont_mul_karatsuba<FpCfg>(r.l, a.l, b.l); return r; }
__device__ __noinline__ Fp mul_d(Fp a, Fp b) { return fp_mul_dp(a, b); }
__device__ __noinline__ Fp sqr_i(Fp a) { return fp_sqr(a); }
__device__ __noinline__ Fp sqr_d(Fp a) { return fp_sqr_dp(a); }
// fused pair: first product on the integer pipe, second on the FP64 pipe; one
// instruction stream, so a single warp feeds both pipes
__device__ __noinline__ Fp2x mul_pair(Fp a1, Fp b1, Fp a2, Fp b2) {
  Fp2x r;
  mont_mul<FpCfg>(r.a.l, a1.l, b1.l);
  dp::mont_mul(r.b.l, a2.l, b2.l);
  return r;
}
__device__ __noinline__ Fp2x mul_pair_k(Fp a1, Fp b1, Fp a2, Fp b2) {
  Fp2x r;
  mont_mul_karatsuba<FpCfg>(r.a.l, a1.l, b1.l);
  dp::mont_mul(r.b.l, a2.l, b2.l);
  return r;
}

// two independent integer-pipe products in one instruction stream (ILP 2 for a lone warp)
__device__ __noinline__ Fp2x mul_pair_ii(Fp a1, Fp b1, Fp a2, Fp b2) {
  Fp2x r;
  mont_mul<FpCfg>(r.a.l, a1.l, b1.l);
  mont_mul<FpCfg>(r.b.l, a2.l, b2.l);
  return r;
}

// MODE 0 imad, 1 karatsuba, 2 dp, 3 warp-mix (odd warps dp), 4 fused pair, 5 fused pair (karatsuba), 6 sqr imad, 7 sqr dp
// 8: warp-mix with 1 of 4 .. generalised: warp w uses dp if (w % MIXDEN) < MIXNUM
template <int MODE, int MINB, int MIXNUM = 1, int MIXDEN = 2>
__global__ void __launch_bounds__(128, MINB) kmul(uint32_t* out, uint32_t seed) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x, y, u, v;
  for (int i = 0; i < 12; i++) { x.l[i] = seed + t * 3 + i; y.l[i] = seed * 5 + i * t; u.l[i] = seed * 7 + t + i; v.l[i] = seed * 11 + i * 3 + t; }
  x.l[11] &= 0xffffff; y.l[11] &= 0xffffff; u.l[11] &= 0xffffff; v.l[11] &= 0xffffff;
  const int warp = (threadIdx.x >> 5) + blockIdx.x * (blockDim.x >> 5);
  // the warp's SM sub-partition is warp_slot % 4; use the warp index inside the block divided by 4
  // plus the block index so that each sub-partition sees both kinds
  const bool use_dp = ((warp / 4 + warp) % MIXDEN) < MIXNUM;
  for (int i = 0; i < NITER; i++) {
    if (MODE == 0) { x = mul_i(x, y); y = mul_i(y, x); u = mul_i(u, v); v = mul_i(v, u); }
    if (MODE == 1) { x = mul_k(x, y); y = mul_k(y, x); u = mul_k(u, v); v = mul_k(v, u); }
    if (MODE == 2) { x = mul_d(x, y); y = mul_d(y, x); u = mul_d(u, v); v = mul_d(v, u); }
    if (MODE == 3) {
      if (use_dp) { x = mul_d(x, y); y = mul_d(y, x); u = mul_d(u, v); v = mul_d(v, u); }
      else { x = mul_i(x, y); y = mul_i(y, x); u = mul_i(u, v); v = mul_i(v, u); }
    }
    if (MODE == 4) { Fp2x r = mul_pair(x, y, u, v); x = r.a; u = r.b; r = mul_pair(y, x, v, u); y = r.a; v = r.b; }
    if (MODE == 5) { Fp2x r = mul_pair_k(x, y, u, v); x = r.a; u = r.b; r = mul_pair_k(y, x, v, u); y = r.a; v = r.b; }
    if (MODE == 8) { Fp2x r = mul_pair_ii(x, y, u, v); x = r.a; u = r.b; r = mul_pair_ii(y, x, v, u); y = r.a; v = r.b; }
    if (MODE == 6) { x = sqr_i(x); y = sqr_i(y); u = sqr_i(u); v = sqr_i(v); }
    if (MODE == 7) { x = sqr_d(x); y = sqr_d(y); u = sqr_d(u); v = sqr_d(v); }
  }
  out[t] = x.l[0] ^ y.l[3] ^ u.l[5] ^ v.l[7];
}

__global__ void kcheck(uint32_t* bad, uint32_t seed) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x, y;
  for (int i = 0; i < 12; i++) { x.l[i] = (seed + t * 0x9e3779b9u) * (i + 1) + i; y.l[i] = (seed * 5 + t) * 0x85ebca6bu + i * 77; }
  x.l[11] &= 0xffffff; y.l[11] &= 0xffffff;
  uint32_t nb = 0;
  for (int it = 0; it < 64; it++) {
    Fp a = mul_i(x, y), b = mul_d(x, y), c = mul_k(x, y);
    Fp2x p = mul_pair(x, y, y, x);
    Fp s1 = sqr_i(x), s2 = sqr_d(x);
    if (!fp_eq(a, b) || !fp_eq(a, c) || !fp_eq(a, p.a) || !fp_eq(a, p.b) || !fp_eq(s1, s2)) nb++;
    x = a; y = fp_add(s1, y);
  }
  if (nb) atomicAdd(bad, nb);
}

static int g_blocks = 148 * 12;
template <int MODE, int MINB, int MIXNUM = 1, int MIXDEN = 2>
void run(const char* name) {
  uint32_t* o; int blocks = g_blocks;
  cudaMalloc(&o, blocks * 128 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kmul<MODE, MINB, MIXNUM, MIXDEN><<<blocks, 128>>>(o, 7); cudaDeviceSynchronize();
  cudaEventRecord(e0); kmul<MODE, MINB, MIXNUM, MIXDEN><<<blocks, 128>>>(o, 9); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double muls = (double)blocks * 128 * NITER * 4;
  printf("%-52s %8.3f ms  %.3e mul/s  (%.0f clk/warp-mul/SM @1.9GHz)\n", name, ms, muls / ms * 1e3, ms * 1e-3 * 1.9e9 / (muls / 32 / 148));
  cudaFree(o);
}
int g_only = -1, g_idx = 0;
#define RUNX(...) do { if (g_only < 0 || g_only == g_idx) run<__VA_ARGS__>; g_idx++; } while (0)
int main(int argc, char** argv) {
  if (argc > 1) g_only = atoi(argv[1]);
  uint32_t* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
  kcheck<<<148, 128>>>(bad, 12345); uint32_t hb = 1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  printf("device cross-check of multiplier variants: %u mismatches (%s)\n", hb, cudaGetErrorString(cudaGetLastError()));
  { if (g_only < 0 || g_only == g_idx) run<0, 3>("imad (minb 3)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<0, 4>("imad (minb 4)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<1, 3>("imad karatsuba (minb 3)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<1, 4>("imad karatsuba (minb 4)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<2, 2>("dfma (minb 2)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<2, 3>("dfma (minb 3)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<2, 4>("dfma (minb 4)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<3, 3, 1, 2>("warp mix 1/2 dfma (minb 3)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<3, 4, 1, 2>("warp mix 1/2 dfma (minb 4)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<3, 3, 2, 3>("warp mix 2/3 dfma (minb 3)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<3, 3, 1, 3>("warp mix 1/3 dfma (minb 3)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<3, 4, 3, 4>("warp mix 3/4 dfma (minb 4)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<3, 4, 1, 4>("warp mix 1/4 dfma (minb 4)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<4, 2>("fused pair imad+dfma (minb 2)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<4, 3>("fused pair imad+dfma (minb 3)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<5, 2>("fused pair karatsuba+dfma (minb 2)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<5, 3>("fused pair karatsuba+dfma (minb 3)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<6, 3>("sqr imad (minb 3)"); g_idx++; }
  { if (g_only < 0 || g_only == g_idx) run<7, 3>("sqr dfma (minb 3)"); g_idx++; }
  // few warps per scheduler: can a lone warp keep the multiplier pipe busy?
  for (int per_sm = 1; per_sm <= 3; per_sm++) {
    g_blocks = 148 * per_sm;
    printf("-- %d block(s) of 4 warps per SM\n", per_sm);
    { if (g_only < 0 || g_only == g_idx) run<0, 3>("imad, one product per call"); g_idx++; }
    { if (g_only < 0 || g_only == g_idx) run<8, 2>("imad, two products per call (ILP 2)"); g_idx++; }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
