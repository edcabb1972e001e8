// clk per warp-level field multiplication on one SM, several implementations
#include <cstdio>
#include <cuda_runtime.h>
#include "../lambdaworks_kzg_b200/csrc/fp29.cuh"
using namespace lw;
#define NITER 512
__device__ __constant__ uint32_t P29V[14];   // not compile-time known
__device__ __constant__ uint32_t PINVV;

__device__ __forceinline__ Fp29 mul29_v(const Fp29& a, const Fp29& b) {
  uint64_t c[15];
#pragma unroll
  for (int j = 0; j < 15; j++) c[j] = 0;
#pragma unroll
  for (int i = 0; i < 14; i++) {
    const uint32_t bi = b.l[i];
#pragma unroll
    for (int j = 0; j < 14; j++) c[j] += (uint64_t)a.l[j] * bi;
    const uint32_t m = ((uint32_t)c[0] * PINVV) & MASK29;
#pragma unroll
    for (int j = 0; j < 14; j++) c[j] += (uint64_t)m * P29V[j];
    const uint64_t carry = c[0] >> W29;
#pragma unroll
    for (int j = 0; j < 14; j++) c[j] = c[j + 1];
    c[14] = 0;
    c[0] += carry;
  }
  Fp29 r;
#pragma unroll
  for (int j = 0; j < 13; j++) { r.l[j] = (uint32_t)c[j] & MASK29; c[j + 1] += c[j] >> W29; }
  r.l[13] = (uint32_t)c[13];
  return r;
}
// product scanning: all 196 products first (27 columns), then reduction
__device__ __forceinline__ Fp29 mul29_ps(const Fp29& a, const Fp29& b) {
  uint64_t c[28];
#pragma unroll
  for (int j = 0; j < 28; j++) c[j] = 0;
#pragma unroll
  for (int i = 0; i < 14; i++)
#pragma unroll
    for (int j = 0; j < 14; j++) c[i + j] += (uint64_t)a.l[j] * b.l[i];
#pragma unroll
  for (int i = 0; i < 14; i++) {
    const uint32_t m = ((uint32_t)c[i] * PINVV) & MASK29;
#pragma unroll
    for (int j = 0; j < 14; j++) c[i + j] += (uint64_t)m * P29V[j];
    c[i + 1] += c[i] >> W29;
  }
  Fp29 r;
#pragma unroll
  for (int j = 0; j < 13; j++) { r.l[j] = (uint32_t)c[14 + j] & MASK29; c[15 + j] += c[14 + j] >> W29; }
  r.l[13] = (uint32_t)c[27];
  return r;
}

template <int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) kmul(uint32_t* out, uint32_t seed) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (MODE == 0) {
    Fp x, y;
    for (int i = 0; i < 12; i++) { x.l[i] = seed + t * 3 + i; y.l[i] = seed * 5 + i; }
    x.l[11] &= 0xffffff; y.l[11] &= 0xffffff;
    for (int i = 0; i < NITER; i++) { x = fp_mul(x, y); y = fp_mul(y, x); }
    out[t] = x.l[0] ^ y.l[3];
  } else {
    Fp29 x, y;
    for (int i = 0; i < 14; i++) { x.l[i] = (seed + t * 3 + i) & MASK29; y.l[i] = (seed * 5 + i) & MASK29; }
    x.l[13] &= 0xf; y.l[13] &= 0xf;
    for (int i = 0; i < NITER; i++) {
      if (MODE == 1) { x = fp29_mul(x, y); y = fp29_mul(y, x); }
      if (MODE == 2) { x = mul29_v(x, y); y = mul29_v(y, x); }
      if (MODE == 3) { x = mul29_ps(x, y); y = mul29_ps(y, x); }
    }
    out[t] = x.l[0] ^ y.l[3];
  }
}
template <int MODE, int MINB>
void run(const char* name) {
  uint32_t* o; int blocks = 148 * 16;
  cudaMalloc(&o, blocks * 128 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kmul<MODE, MINB><<<blocks, 128>>>(o, 7); cudaDeviceSynchronize();
  cudaEventRecord(e0); kmul<MODE, MINB><<<blocks, 128>>>(o, 9); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double muls = (double)blocks * 128 * NITER * 2;
  printf("%-44s %8.3f ms  %.3e mul/s  (%.0f clk/warp-mul/SM @1.9GHz)\n", name, ms, muls / ms * 1e3, ms * 1e-3 * 1.9e9 / (muls / 32 / 148));
  cudaFree(o);
}
int main() {
  cudaMemcpyToSymbol(P29V, k29::P, 56); uint32_t pinv = k29::PINV; cudaMemcpyToSymbol(PINVV, &pinv, 4);
  run<0, 3>("fp32 carry-chain mul (minb 3)");
  run<0, 4>("fp32 carry-chain mul (minb 4)");
  run<1, 3>("fp29 mul, constants known (minb 3)");
  run<2, 3>("fp29 mul, constants in c-mem (minb 3)");
  run<2, 4>("fp29 mul, constants in c-mem (minb 4)");
  run<3, 3>("fp29 product-scanning (minb 3)");
  run<3, 4>("fp29 product-scanning (minb 4)");
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
