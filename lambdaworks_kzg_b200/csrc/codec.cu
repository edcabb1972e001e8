// Batched G1 point decompression with validation
// (/root/reference/src/compression.rs:62-103).
#include "g1.cuh"
#include "kernels.h"

namespace lw {

__global__ void __launch_bounds__(32) g1_decompress_kernel(G1Affine* __restrict__ aff, uint8_t* __restrict__ recompressed, int* __restrict__ status,
                                                            const uint8_t* __restrict__ in48, int n, int strict, int check_subgroup) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t buf[48];
  for (int k = 0; k < 48; k++) buf[k] = in48[(size_t)i * 48 + k];
  G1Affine p;
  bool ok = strict ? g1_decompress_strict(p, buf, check_subgroup != 0) : g1_decompress(p, buf, check_subgroup != 0);
  if (!ok) p = g1a_inf();
  if (aff) aff[i] = p;
  if (recompressed) {
    // utils.rs:138 hashes compress(decompress(C)), i.e. the canonical encoding
    uint8_t out[48];
    g1_compress(out, p);
    for (int k = 0; k < 48; k++) recompressed[(size_t)i * 48 + k] = out[k];
  }
  status[i] = ok ? 0 : (strict ? 1 : 2);  // C_KZG_ERROR in reference mode, C_KZG_BADARGS in c-kzg mode
}

// the r-torsion test alone, for points decoded with check_subgroup = 0: status[i] = 0 / rejected (as above)
__global__ void __launch_bounds__(32) g1_subgroup_kernel(int* __restrict__ status, const G1Affine* __restrict__ aff, int n, int strict) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  status[i] = g1_in_subgroup(aff[i]) ? 0 : (strict ? 1 : 2);
}

__global__ void status_or_kernel(int* __restrict__ status, const int* __restrict__ other, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && status[i] == 0 && other[i] != 0) status[i] = other[i];
}

// outputs of failed items are zeroed (the host API leaves caller memory untouched for them; a device buffer that
// still held a plausible-looking point would be worse)
__global__ void zero_failed_kernel(uint8_t* __restrict__ out, int bytes_per_item, const int* __restrict__ status, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && status[i] != 0)
    for (int b = 0; b < bytes_per_item; b++) out[(size_t)i * bytes_per_item + b] = 0;
}
void launch_zero_failed(void* d_out, int bytes_per_item, const int* d_status, int n, cudaStream_t st) {
  if (n <= 0 || !d_out || !d_status) return;
  zero_failed_kernel<<<(n + 127) / 128, 128, 0, st>>>((uint8_t*)d_out, bytes_per_item, d_status, n);
  count_launch();
}

void launch_g1_decompress(void* d_aff, void* d_recompressed48, int* d_status, const void* d_in48, int n, cudaStream_t st, bool strict, bool check_subgroup) {
  if (n <= 0) return;
  LW_SAME_CARVEOUT(g1_decompress_kernel);
  g1_decompress_kernel<<<(n + 31) / 32, 32, 0, st>>>((G1Affine*)d_aff, (uint8_t*)d_recompressed48, d_status, (const uint8_t*)d_in48, n, strict ? 1 : 0,
                                                    check_subgroup ? 1 : 0);
  count_launch();
}
void launch_g1_subgroup_check(int* d_status, const void* d_aff, int n, cudaStream_t st, bool strict) {
  if (n <= 0) return;
  LW_SAME_CARVEOUT(g1_subgroup_kernel);
  g1_subgroup_kernel<<<(n + 31) / 32, 32, 0, st>>>(d_status, (const G1Affine*)d_aff, n, strict ? 1 : 0);
  count_launch();
}
void launch_status_or(int* d_status, const int* d_other, int n, cudaStream_t st) {
  if (n <= 0) return;
  LW_SAME_CARVEOUT(status_or_kernel);
  status_or_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_status, d_other, n);
  count_launch();
}

}  // namespace lw
