// Micro-probe of B200 issue rates for the instruction forms a 381-bit modular
// multiplier can be built from.  One block of 1024 threads per SM; every thread
// runs ITERS x UNROLL independent ops; cycles from clock64.  Prints warp-level
// instructions per clock per SM (x32 = lane ops / clk / SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(unsigned long long* cyc, uint32_t* sink, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 7u + blockIdx.x + 1u;
  uint32_t lo[CHAINS], hi[CHAINS];
  unsigned long long w[CHAINS];
  double d[CHAINS];
  float f[CHAINS];
  double da = (double)a * 1.0000001, db = (double)b * 0.9999999;
  for (int c = 0; c < CHAINS; c++) { lo[c] = a + c; hi[c] = b + c; w[c] = a * 3ull + c; d[c] = (double)(a + c); f[c] = (float)(a + c); }
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
      if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a), "r"(b));
      if (MODE == 1) { asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b)); }
      if (MODE == 2) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(a), "r"(b));
      if (MODE == 3) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(a), "r"(b));
      if (MODE == 4) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[c]) : "d"(da), "d"(db));
      if (MODE == 5) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[c]) : "f"((float)da), "f"((float)db));
      if (MODE == 6) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b));
      if (MODE == 7) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a), "r"(b)); asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[c]) : "d"(da), "d"(db)); }
      if (MODE == 8) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a), "r"(b)); asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b)); }
      if (MODE == 9) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[c]) : "d"(da), "d"(db));
    }
    if (MODE == 1) {  // one long carry chain across all CHAINS (like a Montgomery row)
    }
  }
  unsigned long long t1 = clock64();
  uint32_t acc = 0;
  for (int c = 0; c < CHAINS; c++) acc ^= lo[c] ^ hi[c] ^ (uint32_t)w[c] ^ (uint32_t)(w[c] >> 32) ^ (uint32_t)d[c] ^ (uint32_t)f[c];
  if (acc == 0x12345u) sink[0] = acc;
  __shared__ unsigned long long smin, smax;
  if (threadIdx.x == 0) { smin = ~0ull; smax = 0; }
  __syncthreads();
  atomicMin(&smin, t0);
  atomicMax(&smax, t1);
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = smax - smin;
}

// carry chain of 12 wide MACs (6 pairs... exactly one Montgomery row: 12 lo/hi pairs fused -> IMAD.WIDE.X)
__global__ void __launch_bounds__(1024, 1) probe_chain(unsigned long long* cyc, uint32_t* sink, uint32_t seed) {
  uint32_t a[12], t[14];
  uint32_t b = seed * 7u + blockIdx.x + 1u;
  for (int i = 0; i < 12; i++) a[i] = seed + threadIdx.x * 31u + i;
  for (int i = 0; i < 14; i++) t[i] = i;
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(t[0]), "+r"(t[1]) : "r"(a[0]), "r"(b));
#pragma unroll
    for (int j = 2; j < 12; j += 2)
      asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(t[j]), "+r"(t[j + 1]) : "r"(a[j]), "r"(b));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(t[12]));
    b += t[1];
  }
  unsigned long long t1 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < 14; i++) acc ^= t[i];
  if (acc == 0x12345u) sink[0] = acc;
  __shared__ unsigned long long smin, smax;
  if (threadIdx.x == 0) { smin = ~0ull; smax = 0; }
  __syncthreads();
  atomicMin(&smin, t0);
  atomicMax(&smax, t1);
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = smax - smin;
}

template <int MODE>
void run(const char* name, double ops_per_iter_per_thread) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* d_cyc; uint32_t* d_sink;
  cudaMalloc(&d_cyc, sms * 8); cudaMalloc(&d_sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    if (MODE >= 0) probe<(MODE >= 0 ? MODE : 0)><<<sms, 1024>>>(d_cyc, d_sink, 12345u + rep);
    else probe_chain<<<sms, 1024>>>(d_cyc, d_sink, 12345u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long h[256];
  cudaMemcpy(h, d_cyc, sms * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; i++) avg += (double)h[i]; avg /= sms;
  double warp_instr = 32.0 * ITERS * ops_per_iter_per_thread;  // 32 warps per SM
  printf("%-44s %8.3f warp-instr/clk/SM  (%6.1f lane-ops/clk/SM)  %7.3f ms  clk~%.0f MHz\n", name, warp_instr / avg, 32.0 * warp_instr / avg, ms, avg / (ms * 1e3));
  cudaFree(d_cyc); cudaFree(d_sink);
}

int main() {
  run<0>("mad.wide.u32 (no carry)", CHAINS);
  run<1>("mad.lo.cc+madc.hi.cc pair (1 MAC)", CHAINS);
  run<-1>("carry chain of 6 fused pairs (Montgomery row)", 6);
  run<2>("mad.lo.u32", CHAINS);
  run<3>("mad.hi.u32", CHAINS);
  run<4>("fma.rn.f64", CHAINS);
  run<9>("fma.rz.f64", CHAINS);
  run<5>("fma.rn.f32", CHAINS);
  run<6>("add.cc+addc pair (2 instr)", 2 * CHAINS);
  run<7>("mad.wide + fma.f64 interleaved (2 instr)", 2 * CHAINS);
  run<8>("mad.wide + add pair interleaved (3 instr)", 3 * CHAINS);
  return 0;
}
