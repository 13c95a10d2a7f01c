// conv_gemm kernels with 160-column tiles (5 pipeline stages): all epilogue variants of this width.
#include "conv_gemm_launch.cuh"

namespace md {

int launch_conv_gemm_bn160(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid, cudaStream_t st) {
  return launch_conv_gemm_variant<160, 5>(tmA, tmB, tmO, p, grid, st);
}

}  // namespace md
