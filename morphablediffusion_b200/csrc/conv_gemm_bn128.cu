// conv_gemm kernels with 128-column tiles (6 pipeline stages): all epilogue variants of this width.
#include "conv_gemm_launch.cuh"

namespace md {

int launch_conv_gemm_bn128(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid, cudaStream_t st) {
  return launch_conv_gemm_variant<128, 6>(tmA, tmB, tmO, p, grid, st);
}

}  // namespace md
