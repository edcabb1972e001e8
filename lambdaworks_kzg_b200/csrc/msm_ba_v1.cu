// batched-affine fixed-base MSM kernel, variant 1: 64 accumulators per thread, 64 threads per blob (6 blocks of 64 threads per SM)
#include "msm_ba.cuh"
namespace lw {
void launch_ba_v1(void* d_partials, const void* d_table, int c, const void* d_scalars, bool be_input, int n_blobs, void* d_scratch, cudaStream_t st, int split) {
  launch_ba<64, 6, 64>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st, 0, split);
}
}  // namespace lw
