// conv_gemm kernels with 256-column tiles (4 pipeline stages): all epilogue variants of this width.
#include "conv_gemm_launch.cuh"

namespace md {

int launch_conv_gemm_bn256(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid, cudaStream_t st) {
  return launch_conv_gemm_variant<256, 4>(tmA, tmB, tmO, p, grid, st);
}

}  // namespace md
