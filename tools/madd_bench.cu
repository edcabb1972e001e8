// XYZZ mixed-addition throughput: fully inlined field multiplications versus an
// out-of-line multiplier (I-cache footprint experiment).  No memory traffic.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "../lambdaworks_kzg_b200/csrc/g1.cuh"
#include "experiments/fpdp.cuh"
#include "experiments/karatsuba.cuh"
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

// ---- both pipes: each pair of independent products = one integer-pipe product + one FP64-pipe product
struct Fp2x { Fp a, b; };
__device__ __noinline__ Fp mul_k(Fp a, Fp b) { Fp r; mont_mul_karatsuba<FpCfg>(r.l, a.l, b.l); return r; }
__device__ __noinline__ Fp2x mul_pair(Fp a1, Fp b1, Fp a2, Fp b2) {
  Fp2x r; mont_mul<FpCfg>(r.a.l, a1.l, b1.l); dp::mont_mul(r.b.l, a2.l, b2.l); return r;
}
__device__ __noinline__ Fp2x sqr_pair(Fp a1, Fp a2) {
  Fp2x r; mont_sqr<FpCfg>(r.a.l, a1.l); dp::mont_sqr(r.b.l, a2.l); return r;
}
__device__ __noinline__ Fp2x sqr_pair_dd(Fp a1, Fp a2) {
  Fp2x r; dp::mont_sqr(r.a.l, a1.l); dp::mont_sqr(r.b.l, a2.l); return r;
}
__device__ __noinline__ Fp mul_d(Fp a, Fp b) { return fp_mul_dp(a, b); }
__device__ __noinline__ Fp sqr_d(Fp a) { return fp_sqr_dp(a); }

template <int CFG>
__device__ __forceinline__ void madd_pairs(G1Xyzz& acc, const G1Affine& p) {
  if (g1a_is_inf(p)) return;
  if (xyzz_is_inf(acc)) { acc.x = p.x; acc.y = p.y; acc.zz = fp_one(); acc.zzz = fp_one(); return; }
  Fp2x t = mul_pair(p.x, acc.zz, p.y, acc.zzz);
  Fp Pd = fp_sub(t.a, acc.x);
  Fp Rd = fp_sub(t.b, acc.y);
  if (fp_is_zero(Pd)) { xyzz_madd_rare(acc, p); return; }
  t = (CFG == 1) ? sqr_pair_dd(Pd, Rd) : sqr_pair(Pd, Rd);
  Fp PP = t.a, RR = t.b;
  t = mul_pair(Pd, PP, acc.x, PP);
  Fp PPP = t.a, Q = t.b;
  Fp X3 = fp_sub(fp_sub(RR, PPP), fp_dbl(Q));
  t = mul_pair(acc.zz, PP, acc.zzz, PPP);
  acc.zz = t.a; acc.zzz = t.b;
  t = mul_pair(Rd, fp_sub(Q, X3), acc.y, PPP);
  acc.x = X3; acc.y = fp_sub(t.a, t.b);
}
// warp-specialised: the whole mixed addition on one pipe, chosen per warp
template <bool DP>
__device__ __forceinline__ void madd_one_pipe(G1Xyzz& acc, const G1Affine& p) {
  if (g1a_is_inf(p)) return;
  if (xyzz_is_inf(acc)) { acc.x = p.x; acc.y = p.y; acc.zz = fp_one(); acc.zzz = fp_one(); return; }
#define MUL(a, b) (DP ? mul_d(a, b) : fp_mul_nv(a, b))
#define SQR(a) (DP ? sqr_d(a) : fp_sqr_nv(a))
  Fp U2 = MUL(p.x, acc.zz);
  Fp S2 = MUL(p.y, acc.zzz);
  Fp Pd = fp_sub(U2, acc.x);
  Fp Rd = fp_sub(S2, acc.y);
  if (fp_is_zero(Pd)) { xyzz_madd_rare(acc, p); return; }
  Fp PP = SQR(Pd);
  Fp PPP = MUL(Pd, PP);
  Fp Q = MUL(acc.x, PP);
  Fp X3 = fp_sub(fp_sub(SQR(Rd), PPP), fp_dbl(Q));
  Fp Y3 = fp_sub(MUL(Rd, fp_sub(Q, X3)), MUL(acc.y, PPP));
  acc.zz = MUL(acc.zz, PP);
  acc.zzz = MUL(acc.zzz, PPP);
  acc.x = X3; acc.y = Y3;
#undef MUL
#undef SQR
}
__device__ __forceinline__ void madd_kara(G1Xyzz& acc, const G1Affine& p) {
  if (g1a_is_inf(p)) return;
  if (xyzz_is_inf(acc)) { acc.x = p.x; acc.y = p.y; acc.zz = fp_one(); acc.zzz = fp_one(); return; }
  Fp U2 = mul_k(p.x, acc.zz);
  Fp S2 = mul_k(p.y, acc.zzz);
  Fp Pd = fp_sub(U2, acc.x);
  Fp Rd = fp_sub(S2, acc.y);
  if (fp_is_zero(Pd)) { xyzz_madd_rare(acc, p); return; }
  Fp PP = fp_sqr_nv(Pd);
  Fp PPP = mul_k(Pd, PP);
  Fp Q = mul_k(acc.x, PP);
  Fp X3 = fp_sub(fp_sub(fp_sqr_nv(Rd), PPP), fp_dbl(Q));
  Fp Y3 = fp_sub(mul_k(Rd, fp_sub(Q, X3)), mul_k(acc.y, PPP));
  acc.zz = mul_k(acc.zz, PP);
  acc.zzz = mul_k(acc.zzz, PPP);
  acc.x = X3; acc.y = Y3;
}

template <int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) kern(G1Xyzz* out, const G1Affine* pts) {
  G1Xyzz acc = xyzz_inf();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < NITER; i++) {
    G1Affine e = pts[(i + t) & 63];
    if (MODE == 0) xyzz_madd(acc, e);
    else if (MODE == 1) madd_call(acc, e);
    else if (MODE == 2) xyzz_madd_hot(acc, e);
    else if (MODE == 3) madd_kara(acc, e);
    else if (MODE == 4) madd_pairs<0>(acc, e);
    else if (MODE == 5) madd_pairs<1>(acc, e);
    else if (MODE == 6) madd_one_pipe<true>(acc, e);
    else if (MODE == 7 || MODE == 8 || MODE == 9) {
      const int w = threadIdx.x >> 5;  // 4 warps per block, one per SM sub-partition; alternate by block
      const int den = MODE == 7 ? 2 : 3, num = MODE == 9 ? 2 : 1;
      if ((int)((w + blockIdx.x) % den) < num) madd_one_pipe<true>(acc, e); else madd_one_pipe<false>(acc, e);
    }
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
  G1Xyzz* ha = (G1Xyzz*)malloc(sizeof(G1Xyzz) * 256); G1Xyzz* hb = (G1Xyzz*)malloc(sizeof(G1Xyzz) * 256);
  kern<2, 3><<<2, 128>>>(o32, p32); cudaMemcpy(ha, o32, sizeof(G1Xyzz) * 256, cudaMemcpyDeviceToHost);
#define CHECK(M, B) kern<M, B><<<2, 128>>>(o32, p32); cudaMemcpy(hb, o32, sizeof(G1Xyzz) * 256, cudaMemcpyDeviceToHost); \
  printf("mode %d minb %d vs hot: %s (%s)\n", M, B, memcmp(ha, hb, sizeof(G1Xyzz) * 256) ? "MISMATCH" : "identical bits", cudaGetErrorString(cudaGetLastError()));
  CHECK(3, 3) CHECK(4, 2) CHECK(4, 3) CHECK(5, 2) CHECK(6, 3) CHECK(7, 3) CHECK(8, 3)
#define RUN(M, B, NAME) ms = timeit([&] { kern<M, B><<<blocks, 128>>>(o32, p32); }); printf("%-44s minb=%d: %.3f ms  %.3e madd/s  (%.0f clk/warp-madd/SM @1.9GHz)\n", NAME, B, ms, n / ms * 1e3, ms * 1e-3 * 1.9e9 / (n / 32 / 148));
  RUN(2, 3, "hot (imad mul + imad sqr, out of line)")
  RUN(3, 3, "karatsuba mul + imad sqr")
  RUN(3, 4, "karatsuba mul + imad sqr")
  RUN(4, 2, "fused pairs imad|dfma")
  RUN(4, 3, "fused pairs imad|dfma")
  RUN(5, 2, "fused pairs, both squarings dfma")
  RUN(5, 3, "fused pairs, both squarings dfma")
  RUN(6, 3, "all dfma")
  RUN(7, 3, "warp-specialised 1/2 dfma")
  RUN(8, 3, "warp-specialised 1/3 dfma")
  RUN(9, 3, "warp-specialised 2/3 dfma")
  ms = timeit([&] { kern<0, 3><<<blocks, 128>>>(o32, p32); }); printf("inline mul   minb=3: %.3f ms  %.3e madd/s\n", ms, n / ms * 1e3);
  ms = timeit([&] { kern<1, 3><<<blocks, 128>>>(o32, p32); }); printf("call mul     minb=3: %.3f ms  %.3e madd/s\n", ms, n / ms * 1e3);
  ms = timeit([&] { kern<1, 4><<<blocks, 128>>>(o32, p32); }); printf("call mul     minb=4: %.3f ms  %.3e madd/s\n", ms, n / ms * 1e3);
  ms = timeit([&] { kern<1, 5><<<blocks, 128>>>(o32, p32); }); printf("call mul     minb=5: %.3f ms  %.3e madd/s\n", ms, n / ms * 1e3);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
