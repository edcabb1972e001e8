// IMAD.WIDE issue rate versus operand pattern (register-bank effects)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 1024
__device__ __constant__ uint32_t CM[16];
template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(unsigned long long* cyc, uint32_t* sink, uint32_t seed) {
  uint32_t a[14], b = seed * 7u + blockIdx.x + 1u;
  uint64_t c[15];
  for (int i = 0; i < 14; i++) a[i] = seed + threadIdx.x * 31u + i * 977u;
  for (int i = 0; i < 15; i++) c[i] = i + threadIdx.x;
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {  // in-place accumulate, distinct a[j], shared b
#pragma unroll
      for (int j = 0; j < 14; j++) c[j] += (uint64_t)a[j] * b;
    }
    if (MODE == 1) {  // shifting columns: c[j] = c[j+1] + a[j]*b
#pragma unroll
      for (int j = 0; j < 14; j++) c[j] = c[j + 1] + (uint64_t)a[j] * b;
    }
    if (MODE == 2) {  // constant-bank multiplier
#pragma unroll
      for (int j = 0; j < 14; j++) c[j] += (uint64_t)b * CM[j];
    }
    if (MODE == 3) {  // 32-bit IMAD lo accumulate (for comparison)
#pragma unroll
      for (int j = 0; j < 14; j++) { uint32_t lo = (uint32_t)c[j]; lo = a[j] * b + lo; c[j] = (c[j] & 0xffffffff00000000ull) | lo; }
    }
    if (MODE == 4) {  // carry chain of fused pairs over 12 limbs (reference: 1/clk)
      uint32_t* t = reinterpret_cast<uint32_t*>(c);
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(t[0]), "+r"(t[1]) : "r"(a[0]), "r"(b));
#pragma unroll
      for (int j = 1; j < 14; j++) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(t[2 * j]), "+r"(t[2 * j + 1]) : "r"(a[j]), "r"(b));
    }
    if (MODE == 5) {  // immediate multiplier, the register multiplicand b reused by every MAC: the 64-bit addend is the ONLY register-file read
      c[0] += (uint64_t)b * 0x12345679u;  c[1] += (uint64_t)b * 0x2468acf3u;  c[2] += (uint64_t)b * 0x369d0369u;  c[3] += (uint64_t)b * 0x48d159e7u;
      c[4] += (uint64_t)b * 0x5b05b05bu;  c[5] += (uint64_t)b * 0x6d3a06d3u;  c[6] += (uint64_t)b * 0x7f6e5d4du;  c[7] += (uint64_t)b * 0x91a2b3c5u;
      c[8] += (uint64_t)b * 0xa3d70a3du;  c[9] += (uint64_t)b * 0xb60b60b7u;  c[10] += (uint64_t)b * 0xc83fb72fu; c[11] += (uint64_t)b * 0xda740da7u;
      c[12] += (uint64_t)b * 0xeca8641fu; c[13] += (uint64_t)b * 0xfedcba99u;
    }
    if (MODE == 6) {  // high halves only (what the m[0] * m_i column of the reduction needs)
#pragma unroll
      for (int j = 0; j < 14; j++) { uint32_t lo = (uint32_t)c[j]; lo = __umulhi(a[j], b) + lo; c[j] = (c[j] & 0xffffffff00000000ull) | lo; }
    }
    if (MODE == 7) {  // product without addend (first row of a CIOS product): dest fresh, a[j] distinct, b shared
#pragma unroll
      for (int j = 0; j < 14; j++) c[j] = (uint64_t)(a[j] ^ (uint32_t)c[j]) * b;
    }
    b += (uint32_t)c[3];
  }
  unsigned long long t1 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < 15; i++) acc ^= (uint32_t)c[i] ^ (uint32_t)(c[i] >> 32);
  if (acc == 0x12345u) sink[0] = acc;
  __shared__ unsigned long long smin, smax;
  if (threadIdx.x == 0) { smin = ~0ull; smax = 0; }
  __syncthreads();
  atomicMin(&smin, t0); atomicMax(&smax, t1);
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = smax - smin;
}
template <int MODE>
void run(const char* name) {
  unsigned long long* d_cyc; uint32_t* d_sink;
  cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_sink, 4);
  probe<MODE><<<148, 1024>>>(d_cyc, d_sink, 1u); cudaDeviceSynchronize();
  probe<MODE><<<148, 1024>>>(d_cyc, d_sink, 3u); cudaDeviceSynchronize();
  unsigned long long h[148]; cudaMemcpy(h, d_cyc, 148 * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; i++) avg += (double)h[i]; avg /= 148;
  printf("%-52s %6.3f MAC-instr/clk/SM\n", name, 32.0 * ITERS * 14 / avg);
}
int main() {
  uint32_t cm[16]; for (int i = 0; i < 16; i++) cm[i] = 0x12345 * (i + 3);
  cudaMemcpyToSymbol(CM, cm, 64);
  run<0>("IMAD.WIDE in-place, distinct a[j], shared b");
  run<1>("IMAD.WIDE shifting columns (dest != addend)");
  run<2>("IMAD.WIDE with constant-bank operand");
  run<3>("IMAD (32-bit) accumulate");
  run<4>("carry chain of 14 fused pairs");
  run<5>("IMAD.WIDE, immediate x reused register (only the addend is read)");
  run<6>("IMAD.HI accumulate");
  run<7>("IMAD.WIDE without addend (+1 LOP3 each)");
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
