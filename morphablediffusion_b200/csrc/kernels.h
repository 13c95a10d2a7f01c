// Internal launcher declarations (host side) for every kernel of libmdiff.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace md {

enum Act : int { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2, ACT_GEGLU = 3, ACT_GELU = 4, ACT_QUICKGELU = 5 };  // QuickGELU: x * sigmoid(1.702 x) (CLIP)
// GEGLU projections are packed per tile of kGegluTile weight rows: kGegluTile/2 value rows, then their kGegluTile/2 gate
// rows, so one accumulator tile holds both halves (the widest tile: fewest re-reads of the activation rows from L2).
constexpr int kGegluTile = 256;

// ---- elementwise.cu
struct GroupNormArgs {
  const void* x0; int C0; int x0_bf16;   // first source  [B][rows][C0]
  const void* x1; int C1; int x1_bf16;   // optional second source (channel concat), C1 = 0 if absent
  int B, rows, groups;
  float eps;
  const float* gamma; const float* beta; // [C0+C1]
  const float* addvec; int addvec_ld;    // optional per-sample vector added before normalising
  float* stats;                          // workspace [B][C][2]
  int stats_prezeroed;                   // 1: buffer is all-zero on entry (the finalize kernel re-zeroes it)
  const float* stats0; const float* stats1;  // optional [B][C0][2] / [B][C1][2] produced by GEMM epilogues (skip the column pass)
  float* scale_shift;                    // workspace [B][C][2]
  void* out;                             // bf16 [B][rows][C]; null = only produce scale_shift
  void* raw_out;                         // optional bf16 copy of the un-normalised input
  int act;
};
int launch_group_norm(const GroupNormArgs& a, cudaStream_t st);
// write_back = 0: the per-sample vector is added for the normalisation only, x itself is left untouched
int launch_layer_norm(float* x, const float* addvec, int addvec_ld, const float* gamma, const float* beta,
                      void* out_bf16, size_t nrows, int rows_per_sample, int C, float eps, cudaStream_t st,
                      int write_back = 1);
int launch_layer_norm_bf16(const void* x_bf16, const float* addvec, int addvec_ld, const float* gamma, const float* beta,
                           void* out_bf16, size_t nrows, int rows_per_sample, int C, float eps, cudaStream_t st);
int launch_small_linear(const float* x, int ldx, const float* W, const float* bias, float* out, int ldo, int B, int K,
                        int N, int act_in, int act_out, int accumulate, cudaStream_t st);
int launch_timestep_embedding(const float* t, float* out, int B, int dim, cudaStream_t st);
int launch_rows_to_nchw(const float* x, int ld, float* out, int B, int C, int HW, cudaStream_t st);
int launch_unet_input(const float* x, const float* xc, int xc_per_sample, float* out, int T, int HW, int cfg,
                      cudaStream_t st);
int launch_cast_bf16(const float* x, void* out, size_t n, cudaStream_t st);
int launch_pad_cast_bf16(const float* x, void* out, size_t rows, int C, int Cpad, cudaStream_t st);
// VAE input: post_quant_conv (1x1, pq = [4][4] weight | [4] bias) of x * in_scale, NCHW fp32 [B][4][HW] -> bf16 [B][HW][64]
// zero-padded to one K block;  row softmax fp32 [rows][n] -> bf16 (AttnBlock, model.py:190-192)
int launch_vae_input(const float* x, const float* pq, float in_scale, void* out_bf16, int B, int HW, cudaStream_t st);
// NCHW fp32 [B][C][HW] (C <= 4) -> bf16 channels-last [B][HW][64], zero-padded (first-stage encoder input)
int launch_nchw_to_cl64(const float* x, void* out_bf16, int B, int C, size_t HW, cudaStream_t st);
// CLIP image tower: preprocess + patchify, token assembly + ln_pre
int launch_clip_patches(const float* image, void* out_bf16, int B, int H, int W, cudaStream_t st);
int launch_clip_tokens(const float* patches, const float* class_emb, const float* pos_emb, const float* g, const float* b,
                       float* x, int B, int ntok, int C, cudaStream_t st);
int launch_softmax_rows(const float* x, void* out_bf16, size_t rows, int n, cudaStream_t st);
int launch_upsample2x(const float* x, void* out, int B, int H, int W, int C, cudaStream_t st);
int launch_ncdhw_to_cl_bf16(const float* x, void* out, int B, int C, size_t S, cudaStream_t st);
int launch_cl_to_ncdhw(const void* x, int x_is_bf16, float* out, int B, int C, size_t S, cudaStream_t st);
int launch_cfg_ddim(const float* eps, float* x, float* eps_out, const float* noise, int T, int n_per_view, int cfg,
                    float cfg_scale, float a_t, float a_prev, float sigma, float sqrt_1m_at, int add_noise,
                    uint64_t seed, uint32_t step, int view0, int do_update, const float* dev_params, cudaStream_t st);

// ---- attention.cu
// Self-attention over tokens: qkv bf16 [B][S][3*C] (q | k | v, head h at columns h*dh), out bf16 [B][S][C].
int launch_self_attention(const void* qkv, void* out, int B, int S, int heads, int dh, cudaStream_t st);
int launch_self_attention_mma(const void* qkv, void* out, int B, int S, int heads, int dh, cudaStream_t st);
// ---- attention_tc.cu: the same contract on tcgen05 tensor cores (S a multiple of 128, dh <= 128)
bool attention_tc_supported(int S, int dh);
int launch_attention_tc(const void* qkv, void* out, int B, int S, int heads, int dh, cudaStream_t st);
// Re-associated depth attention (see attention.cu): qp bf16 [T][HW][4*ctx], c1 bf16 [T][D][HW][ctx] (pre-norm),
// ss fp32 [T][ctx][2] GroupNorm scale/shift, beta fp32 [ctx]; cbar bf16 [B][HW][4*ctx] (samples >= T: zero volume).
int launch_depth_attention(const void* qp, const void* c1, const float* ss, const float* beta, void* cbar, int T, int B,
                           int D, int HW, int ctx, cudaStream_t st);

}  // namespace md
