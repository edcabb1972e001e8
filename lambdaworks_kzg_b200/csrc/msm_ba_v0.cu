// batched-affine fixed-base MSM kernel, variant 0: 64 accumulators per thread, 128 threads per blob (3 blocks of 128 threads per SM)
#include "msm_ba.cuh"
namespace lw {
void launch_ba_v0(void* d_partials, const void* d_table, int c, const void* d_scalars, bool be_input, int n_blobs, void* d_scratch, cudaStream_t st, int split) {
  launch_ba<64, 3, 128>(d_partials, d_table, c, d_scalars, be_input, n_blobs, d_scratch, st, 0, split);
}
// segmented form for the FK20 cell-proof MSMs (cells.cu): 64 results per block, out[seg * seg_stride + blob]
void launch_ba_segmented(void* d_out, const void* d_table, int c, const void* d_scalars, int n_blobs, void* d_scratch, int seg_stride, cudaStream_t st) {
  launch_ba<64, 3, 128>(d_out, d_table, c, d_scalars, false, n_blobs, d_scratch, st, seg_stride);
}
}  // namespace lw
