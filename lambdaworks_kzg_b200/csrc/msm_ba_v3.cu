// batched-affine fixed-base MSM kernel, variant 3: 64 accumulators per thread, 32 threads per blob (12 blocks of 32 threads per SM)
#include "msm_ba.cuh"
namespace lw {
void launch_ba_v3(void* d_partials, const void* d_table, int c, const void* d_scalars, bool be_input, int n_blobs, void* d_scratch, cudaStream_t st) {
  launch_ba<64, 12, 32>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st);
}
}  // namespace lw
