// XYZZ mixed-addition throughput: fully inlined field multiplications versus an
// out-of-line multiplier (I-cache footprint experiment).  No memory traffic.
#include <cstdio>
#include <cuda_runtime.h>
#include "../lambdaworks_kzg_b200/csrc/g1.cuh"
using namespace lw;
#define NITER 256

__device__ __noinline__ Fp mul_nv(Fp a, Fp b) { Fp r; mont_mul<FpCfg>(r.l, a.l, b.l); return r; }

__device__ __forceinline__ void madd_call(G1Xyzz& acc, const G1Affine& p) {
  if (g1a_is_inf(p)) return;
  if (xyzz_is_inf(acc)) { acc.x = p.x; acc.y = p.y; acc.zz = fp_one(); acc.zzz = fp_one(); return; }
  Fp U2 = mul_nv(p.x, acc.zz);
  Fp S2 = mul_nv(p.y, acc.zzz);
  Fp Pd = fp_sub(U2, acc.x);
  Fp Rd = fp_sub(S2, acc.y);
  if (fp_is_zero(Pd)) { if (fp_is_zero(Rd)) acc = xyzz_dbl_affine(p); else acc = xyzz_inf(); return; }
  Fp PP = mul_nv(Pd, Pd);
  Fp PPP = mul_nv(Pd, PP);
  Fp Q = mul_nv(acc.x, PP);
  Fp X3 = fp_sub(fp_sub(mul_nv(Rd, Rd), PPP), fp_dbl(Q));
  Fp Y3 = fp_sub(mul_nv(Rd, fp_sub(Q, X3)), mul_nv(acc.y, PPP));
  acc.zz = mul_nv(acc.zz, PP);
  acc.zzz = mul_nv(acc.zzz, PPP);
  acc.x = X3; acc.y = Y3;
}

template <int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) kern(G1Xyzz* out, const G1Affine* pts) {
  G1Xyzz acc = xyzz_inf();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < NITER; i++) {
    G1Affine e = pts[(i + t) & 63];
    if (MODE == 0) xyzz_madd(acc, e); else madd_call(acc, e);
  }
  out[t] = acc;
}
__global__ void mkpts(G1Affine* p32) {
  int i = threadIdx.x;
  uint32_t k[8] = {(uint32_t)(i * 2654435761u + 12345u), 7u + i, 0, 0, 0, 0, 0, 1u + i};
  p32[i] = xyzz_to_affine(g1_mul_scalar(g1a_generator(), k, 8));
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 2;
}
int main() {
  G1Affine* p32; G1Xyzz* o32; int blocks = 148 * 12;
  cudaMalloc(&p32, 64 * sizeof(G1Affine)); cudaMalloc(&o32, blocks * 128 * sizeof(G1Xyzz));
  mkpts<<<1, 64>>>(p32); cudaDeviceSynchronize();
  double n = (double)blocks * 128 * NITER; float ms;
  ms = timeit([&] { kern<0, 3><<<blocks, 128>>>(o32, p32); }); printf("inline mul   minb=3: %.3f ms  %.3e madd/s\n", ms, n / ms * 1e3);
  ms = timeit([&] { kern<1, 3><<<blocks, 128>>>(o32, p32); }); printf("call mul     minb=3: %.3f ms  %.3e madd/s\n", ms, n / ms * 1e3);
  ms = timeit([&] { kern<1, 4><<<blocks, 128>>>(o32, p32); }); printf("call mul     minb=4: %.3f ms  %.3e madd/s\n", ms, n / ms * 1e3);
  ms = timeit([&] { kern<1, 5><<<blocks, 128>>>(o32, p32); }); printf("call mul     minb=5: %.3f ms  %.3e madd/s\n", ms, n / ms * 1e3);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
