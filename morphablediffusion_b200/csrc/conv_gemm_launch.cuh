// Host-side launch plumbing of the implicit-GEMM kernel family.  Every tile width lives in its own translation unit
// (conv_gemm_bn*.cu) holding the epilogue variants of that width: residual kind x per-sample vector x fused statistics
// x activation class, one kernel each.
#pragma once
#include "conv_gemm.cuh"
#include "host.h"

namespace md {

int launch_conv_gemm_bn64(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid, cudaStream_t st);
int launch_conv_gemm_bn128(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid, cudaStream_t st);
int launch_conv_gemm_bn160(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid, cudaStream_t st);
int launch_conv_gemm_bn256(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid, cudaStream_t st);
// CTA-pair (cta_group::2) kernels exist for the two widest tiles; max_pairs != nullptr only queries residency
int launch_conv_gemm_cg2_bn160(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid, cudaStream_t st, int* max_pairs);
int launch_conv_gemm_cg2_bn256(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid, cudaStream_t st, int* max_pairs);

template <int BN, int STAGES, int RES, bool RV, bool STATS, int ACTV>
int launch_conv_gemm_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid,
                         cudaStream_t stream) {
  using S = ConvGemmSmem<BN, STAGES>;
  auto kernel = conv_gemm_kernel<BN, STAGES, RES, RV, STATS, ACTV>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(conv_gemm): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  launch_pdl(kernel, dim3(grid), dim3(64 + 32 * kEpiWarps), S::kTotal, stream, tmA, tmB, tmO, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("conv_gemm launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

// CTA-pair kernel: clusters of two CTAs; `grid` is the number of CTAs (even).  *max_pairs (optional) receives the number
// of pairs that can be resident at once (asked once per kernel instance).
template <int BN, int STAGES, int RES, bool RV, bool STATS, int ACTV>
int launch_conv_gemm_cg2_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid,
                             cudaStream_t stream, int* max_pairs) {
  using S = ConvGemmSmem2<BN, STAGES>;
  auto kernel = conv_gemm_cg2_kernel<BN, STAGES, RES, RV, STATS, ACTV>;
  static bool attr_set = false;
  static int resident_pairs = 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.blockDim = dim3(64 + 32 * kEpiWarps); cfg.dynamicSmemBytes = S::kTotal; cfg.stream = stream;
  cfg.attrs = attr; cfg.numAttrs = 2;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(conv_gemm_cg2): %s", cudaGetErrorString(e));
    cfg.gridDim = dim3(2 * (num_sms() / 2));
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
    if (e != cudaSuccess || n < 1) { cudaGetLastError(); n = num_sms() / 2 - 4; }
    resident_pairs = n;
    attr_set = true;
  }
  if (max_pairs) { *max_pairs = resident_pairs; return 0; }
  cfg.gridDim = dim3(grid);
  prefer_max_smem(reinterpret_cast<const void*>(kernel));
  cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmO, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("conv_gemm_cg2 launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

template <int BN, int STAGES, int RES, bool RV, bool STATS>
int launch_conv_gemm_cg2_act(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid,
                             cudaStream_t st, int* max_pairs) {
  if (p.act == ACT_NONE) return launch_conv_gemm_cg2_one<BN, STAGES, RES, RV, STATS, 0>(tmA, tmB, tmO, p, grid, st, max_pairs);
  if (p.act == ACT_GEGLU) {
    if constexpr (BN == kGegluTile && RES == 0 && !RV && !STATS)
      return launch_conv_gemm_cg2_one<BN, STAGES, RES, RV, STATS, 2>(tmA, tmB, tmO, p, grid, st, max_pairs);
    else
      return set_error("conv_gemm: the GEGLU epilogue exists for the widest tile without residual / vector / statistics only");
  }
  return launch_conv_gemm_cg2_one<BN, STAGES, RES, RV, STATS, 1>(tmA, tmB, tmO, p, grid, st, max_pairs);
}

// max_pairs != nullptr: only report how many CTA pairs of this instance fit on the device at once
template <int BN, int STAGES>
int launch_conv_gemm_cg2_variant(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid,
                                 cudaStream_t st, int* max_pairs) {
  const int res = p.res_f32 ? 1 : (p.res_bf16 ? 2 : 0);
  const int key = res * 4 + (p.rowvec ? 2 : 0) + (p.col_stats ? 1 : 0);
  switch (key) {
    case 0: return launch_conv_gemm_cg2_act<BN, STAGES, 0, false, false>(tmA, tmB, tmO, p, grid, st, max_pairs);
    case 1: return launch_conv_gemm_cg2_act<BN, STAGES, 0, false, true>(tmA, tmB, tmO, p, grid, st, max_pairs);
    case 2: return launch_conv_gemm_cg2_act<BN, STAGES, 0, true, false>(tmA, tmB, tmO, p, grid, st, max_pairs);
    case 3: return launch_conv_gemm_cg2_act<BN, STAGES, 0, true, true>(tmA, tmB, tmO, p, grid, st, max_pairs);
    case 4: return launch_conv_gemm_cg2_act<BN, STAGES, 1, false, false>(tmA, tmB, tmO, p, grid, st, max_pairs);
    case 5: return launch_conv_gemm_cg2_act<BN, STAGES, 1, false, true>(tmA, tmB, tmO, p, grid, st, max_pairs);
    case 6: return launch_conv_gemm_cg2_act<BN, STAGES, 1, true, false>(tmA, tmB, tmO, p, grid, st, max_pairs);
    case 7: return launch_conv_gemm_cg2_act<BN, STAGES, 1, true, true>(tmA, tmB, tmO, p, grid, st, max_pairs);
    case 8: return launch_conv_gemm_cg2_act<BN, STAGES, 2, false, false>(tmA, tmB, tmO, p, grid, st, max_pairs);
    case 9: return launch_conv_gemm_cg2_act<BN, STAGES, 2, false, true>(tmA, tmB, tmO, p, grid, st, max_pairs);
    case 10: return launch_conv_gemm_cg2_act<BN, STAGES, 2, true, false>(tmA, tmB, tmO, p, grid, st, max_pairs);
    default: return launch_conv_gemm_cg2_act<BN, STAGES, 2, true, true>(tmA, tmB, tmO, p, grid, st, max_pairs);
  }
}

template <int BN, int STAGES, int RES, bool RV, bool STATS>
int launch_conv_gemm_act(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid,
                         cudaStream_t st) {
  if (p.act == ACT_NONE) return launch_conv_gemm_one<BN, STAGES, RES, RV, STATS, 0>(tmA, tmB, tmO, p, grid, st);
  if (p.act == ACT_GEGLU) {
    if constexpr (BN == kGegluTile && RES == 0 && !RV && !STATS)
      return launch_conv_gemm_one<BN, STAGES, RES, RV, STATS, 2>(tmA, tmB, tmO, p, grid, st);
    else
      return set_error("conv_gemm: the GEGLU epilogue exists for the widest tile without residual / vector / statistics only");
  }
  return launch_conv_gemm_one<BN, STAGES, RES, RV, STATS, 1>(tmA, tmB, tmO, p, grid, st);
}

template <int BN, int STAGES>
int launch_conv_gemm_variant(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmParams& p, int grid,
                             cudaStream_t st) {
  const int res = p.res_f32 ? 1 : (p.res_bf16 ? 2 : 0);
  const int key = res * 4 + (p.rowvec ? 2 : 0) + (p.col_stats ? 1 : 0);
  switch (key) {
    case 0: return launch_conv_gemm_act<BN, STAGES, 0, false, false>(tmA, tmB, tmO, p, grid, st);
    case 1: return launch_conv_gemm_act<BN, STAGES, 0, false, true>(tmA, tmB, tmO, p, grid, st);
    case 2: return launch_conv_gemm_act<BN, STAGES, 0, true, false>(tmA, tmB, tmO, p, grid, st);
    case 3: return launch_conv_gemm_act<BN, STAGES, 0, true, true>(tmA, tmB, tmO, p, grid, st);
    case 4: return launch_conv_gemm_act<BN, STAGES, 1, false, false>(tmA, tmB, tmO, p, grid, st);
    case 5: return launch_conv_gemm_act<BN, STAGES, 1, false, true>(tmA, tmB, tmO, p, grid, st);
    case 6: return launch_conv_gemm_act<BN, STAGES, 1, true, false>(tmA, tmB, tmO, p, grid, st);
    case 7: return launch_conv_gemm_act<BN, STAGES, 1, true, true>(tmA, tmB, tmO, p, grid, st);
    case 8: return launch_conv_gemm_act<BN, STAGES, 2, false, false>(tmA, tmB, tmO, p, grid, st);
    case 9: return launch_conv_gemm_act<BN, STAGES, 2, false, true>(tmA, tmB, tmO, p, grid, st);
    case 10: return launch_conv_gemm_act<BN, STAGES, 2, true, false>(tmA, tmB, tmO, p, grid, st);
    default: return launch_conv_gemm_act<BN, STAGES, 2, true, true>(tmA, tmB, tmO, p, grid, st);
  }
}

}  // namespace md
