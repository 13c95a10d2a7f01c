// CTA-pair (cta_group::2) conv_gemm kernels with 160-column tiles (7 pipeline stages): all epilogue variants.
#include "conv_gemm_launch.cuh"

namespace md {

int launch_conv_gemm_cg2_bn160(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid,
                                cudaStream_t st, int* max_pairs) {
  return launch_conv_gemm_cg2_variant<160, 7>(tmA, tmB, tmO, p, grid, st, max_pairs);
}

}  // namespace md
