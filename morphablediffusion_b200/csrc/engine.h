// Engine: model context, packed weights, workspace arena and the host-side orchestration of the denoise step.
#pragma once
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "geometry.h"
#include "host.h"
#include "kernels.h"

namespace md {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ workspace arena (bump allocator, graph friendly)
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool failed = false;
  void* alloc(size_t bytes) {
    const size_t a = (off + 255) & ~size_t(255);
    if (a + bytes > cap) { failed = true; return nullptr; }
    off = a + bytes;
    if (off > peak) peak = off;
    return base + a;
  }
  template <typename T> T* get(size_t n) { return static_cast<T*>(alloc(n * sizeof(T))); }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

// ------------------------------------------------------------------ weights
struct GemmW {           // bf16 [N][taps*K] packed for conv_gemm
  bf16* w = nullptr;
  const float* bias = nullptr;  // fp32 [N] or null
  int N = 0, K = 0, taps = 1;
};
struct NormW { const float* g = nullptr; const float* b = nullptr; int C = 0; };

struct ResW {
  int cin = 0, cout = 0, emb_off = 0;
  NormW n1, n2;
  GemmW c1, c2, skip;
  bool has_skip = false;
};
struct STW {
  int C = 0, heads = 0;
  NormW norm, ln1, ln2, ln3;
  GemmW proj_in, qkv, o1, ff1, ff2, proj_out;
  int v2_off = 0;  // offset of this block's attn2 vector inside UNetW::v2 (single context token => attn2 is affine in ctx)
};
struct DepthW {
  int dim = 0, inner = 0, ctx = 0, dhead = 0;
  GemmW proj_in, proj_ctx, wqk, wov, conv1, conv2;  // wqk / wov: re-associated q.k and out.v products (attention.cu)
  NormW gn_in, gn_ctx, gn_o1, gn_o2;
};
struct UNetLayer {
  int kind = 0;  // 0 conv_in, 1 res, 2 st, 3 down, 4 up
  ResW res; STW st; GemmW conv;
};
struct UNetW {
  int model_channels = 320, in_channels = 8, out_channels = 4, heads = 8, ctx_dim = 768, emb_dim = 1280;
  const float* te0_w = nullptr; const float* te0_b = nullptr; const float* te2_w = nullptr; const float* te2_b = nullptr;
  GemmW emb_g; int emb_total = 0;   // concatenated emb_layers.1 of all ResBlocks: [emb_total][emb_dim]
  GemmW v2_g; int v2_total = 0;     // concatenated (attn2.to_out . attn2.to_v) of all transformer blocks: [v2_total][ctx]
  GemmW conv_in_g;                  // conv_in with Cin zero-padded to one 64-channel K block: bf16 [mch][9][64]
  std::vector<std::vector<UNetLayer>> input_blocks, output_blocks;
  ResW mid0, mid2; STW mid1;
  DepthW mid_cond; std::vector<DepthW> out_cond;
  NormW out_norm; GemmW out_g;  // final conv as a GEMM padded to 8 output columns (rows out_channels..7 are zero)
};

// First-stage decoder (AutoencoderKL.decode = post_quant_conv + Decoder, ldm/models/autoencoder.py:330-333,
// ldm/modules/diffusionmodules/model.py:462-569), attn_resolutions = [] as built by morphable_diffusion.py:399-414.
struct VaeAttnW {
  int C = 0;
  NormW norm;
  GemmW qk;                        // q | k fused along N: [2C][C]
  bf16* wv = nullptr;              // v weight [C][C] (used as the A operand of V^T = Wv . h^T)
  const float* bv = nullptr;       // v bias, added after the P.V product (softmax rows sum to 1)
  GemmW proj;
};
struct VaeW {
  bool loaded = false;
  float* pq = nullptr;             // post_quant_conv: [4][4] weight then [4] bias (fp32)
  int z_channels = 4, out_ch = 3;
  GemmW conv_in, conv_out;         // Cin padded to 64 / N padded to 8
  ResW mid1, mid2;
  VaeAttnW attn;
  std::vector<std::vector<ResW>> up;   // up[i_level][i_block], reference indexing (level 3 runs first)
  GemmW upsample[4];               // up[i_level].upsample.conv for i_level >= 1
  NormW norm_out;
};

// First-stage encoder up to the posterior moments (AutoencoderKL.encode, autoencoder.py:324-328; Encoder, model.py:368-459)
struct VaeEncW {
  bool loaded = false;
  GemmW conv_in;                   // 3 -> 128, Cin padded to 64
  std::vector<std::vector<ResW>> down;   // down[i_level][i_block]
  GemmW downsample[4];             // down[i_level].downsample.conv for i_level < 3 (k3 s2, pad right/bottom)
  ResW mid1, mid2;
  VaeAttnW attn;
  NormW norm_out;
  GemmW conv_out;                  // quant_conv (1x1) folded into conv_out: 512 -> 8 moments
};

// CLIP ViT-L/14 image tower (FrozenCLIPImageEmbedder, ldm/modules/encoders/modules.py:343-382; OpenAI clip VisionTransformer)
struct ClipLayerW { NormW ln1, ln2; GemmW qkv, out, fc, proj; };
struct ClipW {
  bool loaded = false;
  int width = 1024, heads = 16, ntok = 257, out_dim = 768;
  GemmW conv1;                     // patch embedding as a GEMM: [width][640] (K = 3*14*14 = 588 zero-padded)
  const float* class_emb = nullptr; const float* pos_emb = nullptr;
  NormW ln_pre, ln_post;
  std::vector<ClipLayerW> layers;
  GemmW proj;                      // proj^T: [out_dim][width]
};

struct FrBlockW {   // FrustumTVBlock / FrustumTVUpBlock
  int cin = 0, cout = 0, stride = 1; bool up = false;
  const float* t_w = nullptr; const float* t_b = nullptr; const float* v_w = nullptr; const float* v_b = nullptr;
  NormW gn;
  GemmW conv;            // stride 1: 27 taps implicit; stride 2: [N][27*Cin] on gathered patches
  GemmW upc[8];          // transposed conv: one packed weight per output parity class
};
struct FrustumW { GemmW conv0; FrBlockW blk[9]; };  // conv1..conv6, up0..up2

struct SparseLayerW { float* w = nullptr; float* scale = nullptr; float* shift = nullptr; int cin = 0, cout = 0; };
struct VolumeW {
  EncWeightsHost enc;
  const float* te0_w = nullptr; const float* te0_b = nullptr; const float* te2_w = nullptr; const float* te2_b = nullptr;
  const float* smpl_w = nullptr; const float* smpl_b = nullptr;
  SparseLayerW sp[9];   // conv0.0 conv0.3 down0 conv1.0 conv1.3 down1 conv2.0 conv2.3 conv2.6
  FrustumW fr;
};

// ------------------------------------------------------------------ per-sample binding (step invariants)
struct SampleBinding {
  bool bound = false;
  int n_views = 0, view0 = 0, n_local = 0, nv = 0, ortho = 0;
  float* proj = nullptr;        // [N][12] projection rows used by the unproject (diag(r,r,1) K RT | K4 RT4)
  float* cam = nullptr;         // [N][24] frustum back-projection
  float* v_embed = nullptr;     // [N][4]
  float* vertices = nullptr;    // [Nv][3]
  float* pts = nullptr;         // [n_local][D*S*S][3]
  int n0 = 0, n1 = 0, n2 = 0;   // active rows per sparse level
  int32_t* row_vertex = nullptr;  // [n0]
  int32_t* nbr[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // subm0, down0, subm1, down1, subm2
  int32_t* rs_idx = nullptr; float* rs_w = nullptr;                  // [V^3][8]
  std::vector<void*> owned;
};

struct Ctx {
  md_config mcfg;
  UNetW unet;
  VolumeW vol;
  VaeW vae;
  VaeEncW vae_enc;
  ClipW clip;
  bool weights_loaded = false;
  std::vector<void*> weight_allocs;
  Arena arena;
  // conditioning branch (sparse conv + frustum nets) runs on a second stream, concurrently with the UNet's input half
  Arena arena2;
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_levels = nullptr;
  bool levels_pending = false;         // unet_forward must wait on ev_levels before its first depth transformer
  SplitWorkspace split_main, split_side;   // split-K workspaces of the two streams (host.h)
  SampleBinding sb;
  // DDIM schedule (host)
  std::vector<float> alphas, alphas_prev, sigmas, sqrt_1m_alphas;
  std::vector<int> timesteps;
  // step state buffers
  float* d_t = nullptr;  // [max B] timesteps as float
  float* gn_stats = nullptr;  // persistent all-zero GroupNorm statistics scratch (finalize re-zeroes it)
  size_t gn_stats_floats = 0;
  // whole-step CUDA graph (captured on the second call with the same pointers; replayed for every DDIM index)
  cudaStream_t stream = nullptr;       // internal non-blocking stream the step runs on
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  float* d_step = nullptr;             // [16] per-step scalars read by the captured kernels
  cudaGraphExec_t graph = nullptr;     // whole step (single rank) or the part before the cross-rank exchange
  cudaGraphExec_t graph_b = nullptr;   // multi-rank only: the part after the exchange (NCCL runs between the two)
  float* vsum_ptr = nullptr;           // [Nv][16] vertex-feature sums of the current step (all-reduced over ranks)
  struct GraphKey { const void* x; const void* xin; const void* clip; const void* noise; const void* eps; float cfg; int bind_gen; int wgen; } gkey{};
  int graph_warm = 0;                  // calls seen with the current key
  int graph_launches = 0;              // kernels inside the captured graph (launch accounting on replay)
  int bind_gen = 0, weights_gen = 0;
  bool use_graph = true;
  // multi-GPU
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  // NVLink peer exchange of the vertex-feature sums (replaces the NCCL all-reduce inside the step; step.cu):
  // every rank owns one IPC-exported buffer [2 slots][world][slot_floats] + [world] arrival flags.  A rank pushes its
  // partial sums into its region of every peer's buffer and then stores the step's sequence number into the peers'
  // flags; the consumer kernel (view mean + Conv1d + scatter) waits for its own flags and adds the world regions in rank
  // order.  Slots alternate with the sequence number, which is enough: a rank can only reach step k+2's push after it
  // has seen every peer's step-k+1 flag, which a peer raises after its own step-k consumer has finished.
  struct PeerExchange {
    bool on = false;
    size_t slot_floats = 0;
    float* base = nullptr;             // my buffer (device memory, IPC-exported)
    void* peer_base[16] = {};          // peers' buffers mapped into this process (peer_base[rank] == base)
    float** d_peer_data = nullptr;     // device copies of the pointer tables
    unsigned** d_peer_flags = nullptr;
    unsigned* d_seq = nullptr;         // [0] sequence number of the current step, [1] last-block ticket of the push kernel
    unsigned seq = 0;                  // host-side counter (every rank calls md_denoise_step the same number of times)
    int* h_err = nullptr;              // mapped pinned host word: the consumer sets it when a peer's flag never arrives
    int* d_err = nullptr;
  } px;
  std::string err;
};

// activation tensor (channels-last) helpers
struct TF32 { float* p; int B, H, W, C; size_t rows() const { return (size_t)B * H * W; } };

// weights.cu
struct NamedTensor { const float* ptr; std::vector<int64_t> shape; size_t numel; };
typedef std::unordered_map<std::string, NamedTensor> TensorMap;
int load_all_weights(Ctx& c, const TensorMap& tm, cudaStream_t st);
void free_weights(Ctx& c);

// unet.cu
// x_in fp32 NHWC [B][H][W][8]; timesteps fp32 [B]; context fp32 [B][ctx_dim]; levels: bf16 channels-last frustum
// volumes for the B samples {64@D,S,S ; 128 ; 256 ; 512}; eps_out fp32 NCHW [B][4][H][W].
// n_ctx: the first n_ctx samples own a frustum volume in `levels`; samples n_ctx..B-1 are conditioned on an all-zero
// volume (the CFG-unconditional half) and `levels` holds nothing for them.
int unet_forward(Ctx& c, const float* x_in, const float* timesteps, const float* context, const bf16* const levels[4],
                 int B, int n_ctx, int S, int D, float* eps_out, cudaStream_t st);

// unet.cu: decode_first_stage for T latents.  x fp32 NCHW [T][4][S][S] (scaled latents: divided by 0.18215 inside),
// image fp32 NCHW [T][3][8S][8S].
int vae_decode(Ctx& c, const float* x, float* image, int T, int S, cudaStream_t st);
// AutoencoderKL.encode up to the moments: image fp32 NCHW [T][3][8S][8S] in [-1,1] -> moments fp32 NCHW [T][8][S][S]
int vae_encode(Ctx& c, const float* image, float* moments, int T, int S, cudaStream_t st);

// FrozenCLIPImageEmbedder.encode: image fp32 NCHW [n][3][H][W] in [-1,1] -> embedding fp32 [n][768]
int clip_embed(Ctx& c, const float* image, float* out, int n, int H, int W, cudaStream_t st);

// volume.cu
int bind_sample(Ctx& c, const float* K, const float* RT, const float* v_embed, const float* vertices,
                const int32_t* coord, const int32_t* out_sh, const float* bounds, int nv, int n_views, int view0,
                int n_local, int ortho, cudaStream_t st);
void free_binding(Ctx& c);
int embed_time(Ctx& c, const float* t_dev, float* t_embed, cudaStream_t st);  // [1] -> [256]
int vertex_feature_sum(Ctx& c, const float* x_local, const float* t_embed, float* vsum, cudaStream_t st);
int spatial_volume_from_vsum(Ctx& c, const float* vsum, float* vol, cudaStream_t st, bool peer_exchange = false);
int frustum_levels(Ctx& c, const float* vol, int lv0, int T, const float* t_embed, int alloc_samples, bf16* levels[4],
                   cudaStream_t st);

}  // namespace md
