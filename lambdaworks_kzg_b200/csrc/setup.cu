// Trusted-setup decoding on the device (load_trusted_setup[_file],
// /root/reference/src/lib.rs:709-802, src/srs.rs:25-128): every compressed
// point is validated exactly like the reference's loaders do (G1: curve +
// subgroup; G2: curve only) and returned as canonical integers so the host can
// lay out KZGSettings.g1_values / g2_values the way src/srs.rs:131-213 does.
#include "fp2.cuh"
#include "g1.cuh"
#include "kernels_setup.h"

namespace lw {

__global__ void __launch_bounds__(32) setup_decode_g1_kernel(uint32_t* __restrict__ canon24, int* __restrict__ status, const uint8_t* __restrict__ in48, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t buf[48];
  for (int k = 0; k < 48; k++) buf[k] = in48[(size_t)i * 48 + k];
  G1Affine p;
  bool ok = g1_decompress(p, buf);
  status[i] = ok ? 0 : 2;
  Fp x = fp_zero(), y = fp_zero();
  if (ok && !g1a_is_inf(p)) { x = fp_from_mont(p.x); y = fp_from_mont(p.y); }
  for (int k = 0; k < 12; k++) { canon24[i * 24 + k] = x.l[k]; canon24[i * 24 + 12 + k] = y.l[k]; }
}

// out: x.c0, x.c1, y.c0, y.c1 canonical (48 u32); status 0 ok, 1 = infinity, 2 = rejected
__global__ void __launch_bounds__(32) setup_decode_g2_kernel(uint32_t* __restrict__ canon48, int* __restrict__ status, const uint8_t* __restrict__ in96, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t buf[96];
  for (int k = 0; k < 96; k++) buf[k] = in96[(size_t)i * 96 + k];
  G2Affine q;
  bool inf = false;
  bool ok = g2_decompress(q, inf, buf);
  status[i] = ok ? (inf ? 1 : 0) : 2;
  Fp c[4] = {fp_zero(), fp_zero(), fp_zero(), fp_zero()};
  if (ok && !inf) { c[0] = fp_from_mont(q.x.c0); c[1] = fp_from_mont(q.x.c1); c[2] = fp_from_mont(q.y.c0); c[3] = fp_from_mont(q.y.c1); }
  for (int j = 0; j < 4; j++)
    for (int k = 0; k < 12; k++) canon48[i * 48 + j * 12 + k] = c[j].l[k];
}

void launch_setup_decode_g1(void* d_canon24, int* d_status, const void* d_in48, int n, cudaStream_t st) {
  if (n <= 0) return;
  setup_decode_g1_kernel<<<(n + 31) / 32, 32, 0, st>>>((uint32_t*)d_canon24, d_status, (const uint8_t*)d_in48, n);
  count_launch();
}
void launch_setup_decode_g2(void* d_canon48, int* d_status, const void* d_in96, int n, cudaStream_t st) {
  if (n <= 0) return;
  setup_decode_g2_kernel<<<(n + 31) / 32, 32, 0, st>>>((uint32_t*)d_canon48, d_status, (const uint8_t*)d_in96, n);
  count_launch();
}

}  // namespace lw
