// Host-side helpers shared by all translation units of libmdiff.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>

#include "../../include/mdiff.h"

namespace md {

typedef md_conv_gemm_args ConvGemmArgs;

int set_error(const char* fmt, ...);  // stores the message, returns -1
void count_launch(int n = 1);
int num_sms();

int launch_conv_gemm(const ConvGemmArgs& a, cudaStream_t stream);
void* tensor_map_encode_fn();  // cuTensorMapEncodeTiled resolved through the runtime (null without a driver)
// Split-K workspace of the launches issued by the calling host thread.  Every context owns its own (one per stream it
// launches GEMMs on) and selects it around its launches with a SplitScope; nothing here is shared between contexts.
// Without a selection (op-level md_op_conv_gemm) a lazily allocated process-wide workspace is used: such calls must be
// stream-ordered with respect to each other.
struct SplitWorkspace { float* ws = nullptr; size_t bytes = 0; int* cnt = nullptr; size_t ints = 0; };
int alloc_split_workspace(SplitWorkspace& w, size_t bytes, size_t ints);
void free_split_workspace(SplitWorkspace& w);
const SplitWorkspace* select_split_workspace(const SplitWorkspace* w);  // thread-local; returns the previous selection
struct SplitScope {
  const SplitWorkspace* prev;
  explicit SplitScope(const SplitWorkspace* w) : prev(select_split_workspace(w)) {}
  ~SplitScope() { select_split_workspace(prev); }
  SplitScope(const SplitScope&) = delete;
  SplitScope& operator=(const SplitScope&) = delete;
};

#define MD_CHECK(expr)                 \
  do {                                 \
    int _rc = (expr);                  \
    if (_rc != 0) return _rc;          \
  } while (0)

#define MD_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) return md::set_error("%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

bool pdl_enabled();
void prefer_max_smem(const void* kernel);  // once per kernel: pin the L1/shared split so kernel boundaries never reconfigure it

// Launch with programmatic stream serialization: the kernel may start while its predecessor drains; every kernel
// calls pdl_grid_sync() (griddepcontrol.wait) before touching memory.
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  prefer_max_smem(reinterpret_cast<const void*>(kernel));
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("%s launch: %s", what, cudaGetErrorString(e));
  count_launch();
  return 0;
}

}  // namespace md
