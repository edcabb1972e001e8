#pragma once
#include "kernels.h"
namespace lw {
void launch_setup_decode_g1(void* d_canon24, int* d_status, const void* d_in48, int n, cudaStream_t st);
void launch_setup_decode_g2(void* d_canon48, int* d_status, const void* d_in96, int n, cudaStream_t st);
}  // namespace lw
