// batched-affine fixed-base MSM kernel, variant 5: 64 accumulators per thread, 64 threads per blob (128 registers: 8 blocks of 64 threads per SM)
#include "msm_ba.cuh"
namespace lw {
void launch_ba_v5(void* d_partials, const void* d_table, int c, const void* d_scalars, bool be_input, int n_blobs, void* d_scratch, cudaStream_t st) {
  launch_ba<64, -128, 64>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st);
}
}  // namespace lw
