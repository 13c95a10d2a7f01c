/*
 * libmdiff — C ABI of the B200-native Morphable Diffusion multi-view denoise hot path.
 *
 * The reference (xiyichen/morphablediffusion) has no FFI: its boundary is a set of Python classes resolved by
 * dotted name (ldm/util.py:217-232, configs/facescape.yaml:3,27).  This header is the C boundary a binding of
 * those classes calls; each entry point cites the reference function it replaces.  Conventions:
 *   - every function returns 0 on success, a negative code on error; md_last_error() gives the message;
 *   - all tensor pointers are DEVICE pointers owned by the caller unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - nothing here takes or returns a torch type.
 */
#ifndef MDIFF_H_
#define MDIFF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MD_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ library-level */
MD_API int md_version(void);
MD_API const char* md_last_error(void);
/* number of kernels this library launched since the last reset (bench.py "gpu_launches") */
MD_API long long md_launch_count(void);
MD_API void md_reset_launch_count(void);

/* ------------------------------------------------------------------ op level: tensor-core implicit GEMM
 * Dense contraction used by every Conv2d/Conv3d/ConvTranspose3d/Linear on the path
 * (reference: torch.nn.Conv2d in ldm/modules/diffusionmodules/openaimodel.py:199-276, nn.Linear in
 * ldm/modules/attention.py:161-203, nn.Conv3d / ConvTranspose3d in ldm/models/diffusion/network.py:285-347).
 *   A   : bf16 channels-last activations [B][D][H][W][Cpitch], the first Cin channels are contracted
 *   Wt  : bf16 packed weights [N][ntaps*Cin] (K-major; tap-major then channel)
 *   out : row-major [rows][ldo]; row = ((b*OD + z*os+op)*OH + y*os+op)*OW + x*os+op
 */
/* MD_ACT_GEGLU (GEGLU, ldm/modules/attention.py:42-44): Wt / bias rows are packed per tile of 256 rows = 128 value rows
 * followed by their 128 gate rows (N = 2*inner, a multiple of 256); the output [rows][N/2] is value * gelu(gate), bf16. */
enum { MD_ACT_NONE = 0, MD_ACT_SILU = 1, MD_ACT_RELU = 2, MD_ACT_GEGLU = 3, MD_ACT_GELU = 4, MD_ACT_QUICKGELU = 5 };

typedef struct md_conv_gemm_args {
  const void* A;      /* bf16 */
  int B, D, H, W;     /* input dims (a plain GEMM uses D=H=1, W=rows per sample, B=samples) */
  int Cin, Cpitch;    /* Cpitch = 0 -> Cin */
  const void* Wt;     /* bf16 [N][ntaps*Cin] */
  int N;
  int ntaps;
  int tap[27][3];     /* (dx,dy,dz) added to the output coordinate to address the input */
  int OD, OH, OW;     /* output dims; 0 -> same as input */
  int os[3], op[3];   /* output coordinate = input coordinate * os + op (x,y,z); os 0 -> 1 */
  const float* bias;        /* [N] or NULL */
  const float* rowvec;      /* [B][rowvec_ld] per-sample additive vector or NULL */
  int rowvec_ld;
  const float* res_f32;     /* residual added after activation, [rows][ldo] or NULL */
  const void* res_bf16;
  float* out_f32;           /* either or both outputs */
  void* out_bf16;
  int ldo;                  /* 0 -> N (N/2 for GEGLU) */
  int act;
  float out_scale;          /* 0 -> 1 */
  int BN;                   /* tile N override (64/128/160/256), 0 = auto */
  float* col_stats;         /* optional [B][stats_ld][2] sum / sum-of-squares accumulation (atomics; pre-zeroed) */
  int stats_ld;             /* 0 -> N */
  int ksplit;               /* split-K factor: 0 = auto, 1 / -1 = off, >1 = forced */
  int in_stride[3];         /* strided conv: input coord = tile coord * in_stride + tap (x,y,z); 0 -> 1.  B,D,H,W then
                               describe the INPUT tensor and the tile grid covers ceil(dim / in_stride) positions */
  int cta_pair;             /* CTA-pair (cta_group::2, 256-row tiles over two SMs) kernel: 0 = library default
                               (from 24 K blocks up; MD_CG2 overrides), 1 = use it whenever the problem is eligible (tile
                               width 160 or 256, no split-K, at least one full wave of pairs), -1 = never */
  int Wpitch;               /* row pitch of Wt in elements (a weight matrix that is a column slice of a wider one,
                               e.g. the keys inside a fused q|k activation); 0 -> ntaps*Cin; multiple of 8 */
  /* GroupNorm tail (needs col_stats, out_bf16 and the plain output geometry): after the last tile the whole grid waits on
     gn_barrier (an int that is zero on entry), then normalises + activates the rows it produced into gn_out (bf16
     [rows][N]) with the statistics of this launch: GroupNorm(groups, eps, gamma, beta) + gn_act (MD_ACT_NONE/SILU/RELU) */
  void* gn_out;
  const float* gn_gamma;
  const float* gn_beta;
  int gn_groups, gn_act;
  float gn_eps;
  int* gn_barrier;
  int tail_split;           /* last partial wave of tiles split along K: 0 = library default (on; MD_HYBRID=0 switches it
                               off), 1 = on, -1 = off */
} md_conv_gemm_args;

MD_API int md_op_conv_gemm(const md_conv_gemm_args* args, void* stream);


/* ------------------------------------------------------------------ model context
 * One context per GPU / rank.  It owns the packed weights, the workspace arena and the per-sample binding.
 * Replaces the Python objects SyncMultiviewDiffusion / SpatialVolumeNet / UNetWrapper / DepthWiseAttention
 * (ldm/models/diffusion/morphable_diffusion.py:67-151,322-359; ldm/models/diffusion/attention.py:87-138). */
typedef struct md_ctx md_ctx;

typedef struct md_config {
  /* DepthWiseAttention / UNetModel arguments (configs/facescape.yaml:26-42) */
  int model_channels, in_channels, out_channels, num_res_blocks, num_heads, context_dim;
  int channel_mult[4];
  int attn_ds[4];       /* attention at downsample rate 1,2,4,8 (attention_resolutions) */
  int volume_dims[4];
  /* SpatialVolumeNet constants (morphable_diffusion.py:152-180) */
  int latent_size, image_size, spatial_volume_size, frustum_depth, time_embed_dim, view_dim;
  float spatial_volume_length, frustum_volume_length;
  int smpl_num_views;   /* SMPLFeatureExtractor.num_views; 0 = number of bound views (reference hard-codes 16) */
  /* DDIM sampler (SyncDDIMSampler.__init__, morphable_diffusion.py:649-672) */
  int ddim_steps;
  float ddim_eta;
  /* runtime */
  int max_views_per_call;          /* UNet batch = 2 x this with CFG; 0 = 16 */
  unsigned long long workspace_bytes; /* 0 = sized from max_views_per_call */
} md_config;

MD_API void md_default_config(md_config* cfg);
MD_API int md_create(md_ctx** out, const md_config* cfg);
MD_API void md_destroy(md_ctx* ctx);
MD_API unsigned long long md_workspace_peak(md_ctx* ctx);

/* load_state_dict (generate_face.py:75-76): n named fp32 DEVICE tensors keyed by the reference's nn.Module paths
 * ("model.diffusion_model.input_blocks.1.0.in_layers.2.weight", "spatial_volume.xyzc_net.conv0.0.weight", ...).
 * Torch layouts (Conv [O,I,k..], ConvTranspose3d [I,O,k,k,k], Linear [O,I], spconv [O,kd,kh,kw,I]).  The tensors
 * are only read during the call; keys the hot path does not use are ignored (strict=False). */
MD_API int md_load_weights(md_ctx* ctx, int n, const char* const* names, const void* const* ptrs,
                           const long long* numels, void* stream);

/* Step invariants of one sample: the `batch` dict of generate_face.py:227-241.  ALL POINTERS HERE ARE HOST
 * pointers (camera/mesh metadata; copied).  K [n_views][4][4], RT [n_views][3][4] world->cam,
 * v_embed [n_views][4] (get_viewpoint_embedding, morphable_diffusion.py:383-397), vertices [nv][3],
 * coord [nv][3] int32 (d,h,w), out_sh [3], bounds [2][3].  This rank owns views [view0, view0+n_local).
 * projection: 0 perspective, 1 orthographic (utils.py:20-69; anything else -> error like NotImplementedError). */
MD_API int md_bind_sample(md_ctx* ctx, const float* K, const float* RT, const float* v_embed, const float* vertices,
                          const int32_t* coord, const int32_t* out_sh, const float* bounds, int nv, int n_views,
                          int view0, int n_local, int projection, void* stream);

/* Voxelisation rule of generate_face.py:214-225 on the GPU (device pointers): coord [nv][3] i32, out_sh [3] i32,
 * bounds [2][3] f32.  Bit-exact with torch.round((v[:, [2,1,0]] - min) / 0.005).int(). */
MD_API int md_voxelize(const float* vertices, int nv, int32_t* coord, int32_t* out_sh, float* bounds, void* stream);

/* Batch construction either side of the path (generate_face.py:203-249), so that the batch dict can be produced on the
 * device: md_affine_points = the rigid alignment of the fitted mesh (scale, so3 rotation + translation, scale, axis swap
 * composed into one map v' = A v + b; A row-major [9] and b [3] are HOST arrays), md_voxelize = coord / out_sh / bounds
 * (rule a1), md_images_to_u8 = clamp(x,-1,1) -> (x+1)/2*255 -> uint8, NCHW fp32 [n][3][H][W] -> NHWC uint8 [n][H][W][3]. */
MD_API int md_affine_points(const float* v, int n, const float* A9_host, const float* b3_host, float* out, void* stream);
MD_API int md_images_to_u8(const float* img, unsigned char* out, int n, int H, int W, void* stream);

/* SpatialVolumeNet.construct_spatial_volume (morphable_diffusion.py:182-263) for the bound sample.
 * x_local [n_local][4][S][S] fp32; t_embed [time_embed_dim] (md_embed_time); volume_out [64][V][V][V] fp32
 * (NCDHW, B = 1). With a communicator set
 * (md_comm_init) the per-view vertex features are all-reduced over ranks inside this call. */
MD_API int md_spatial_volume(md_ctx* ctx, const float* x_local, const float* t_embed, float* volume_out, void* stream);

/* SyncMultiviewDiffusion.embed_time (morphable_diffusion.py:491-494): sinusoid(256) -> Linear -> SiLU -> Linear.
 * t_embed_out: DEVICE [time_embed_dim]. */
MD_API int md_embed_time(md_ctx* ctx, float timestep, float* t_embed_out, void* stream);

/* SpatialVolumeNet.construct_view_frustum_volume (morphable_diffusion.py:265-320) for T local views starting at
 * local index lv0.  volume [64][V][V][V]; out_levels[i] NCDHW fp32 [T][C_i][D/2^i][S/2^i][S/2^i]. */
MD_API int md_frustum_feats(md_ctx* ctx, const float* volume, int lv0, int T, const float* t_embed,
                            float* const out_levels[4], void* stream);

/* DepthWiseAttention.forward (ldm/models/diffusion/attention.py:117-138): x [B][8][S][S], timesteps [B] (HOST
 * floats), context [B][context_dim] (one token per sample), source[i] NCDHW fp32 frustum volumes, out [B][4][S][S]. */
MD_API int md_unet_forward(md_ctx* ctx, const float* x, const float* timesteps_host, const float* context,
                           const float* const source[4], int B, float* out, void* stream);

/* SyncDDIMSampler.denoise_apply (morphable_diffusion.py:701-739) for the local views:
 * x_local [n_local][4][S][S] is replaced by x_{t-1}.  x_input [4][S][S], clip_embed [context_dim].
 * noise: optional [n_local][4][S][S] standard-normal draws; NULL -> Philox keyed by (seed, index, global view).
 * eps_out: optional [n_local][4][S][S] CFG-combined epsilon. index = DDIM index (49 .. 0); index 0 adds no noise. */
MD_API int md_denoise_step(md_ctx* ctx, float* x_local, const float* x_input, const float* clip_embed, int index,
                           float cfg_scale, const float* noise, unsigned long long seed, float* eps_out,
                           void* stream);
/* SyncMultiviewDiffusion.decode_first_stage (morphable_diffusion.py:468-471) for n latents: image = Decoder(
 * post_quant_conv(x / 0.18215)) (ldm/models/autoencoder.py:330-333, ldm/modules/diffusionmodules/model.py:462-569).
 * x [n][4][S][S] (the sampler's output for n views, device, fp32), image [n][3][8S][8S] (device, fp32, NCHW).
 * latent_size S: 0 -> the context's.  Needs the first_stage_model.decoder.* / post_quant_conv.* tensors among the weights
 * given to md_load_weights (md_has_vae tells); views are independent, so with several ranks each decodes its own. */
MD_API int md_vae_decode(md_ctx* ctx, const float* x, float* image, int n, int latent_size, void* stream);
MD_API int md_has_vae(md_ctx* ctx);
/* The VAE half of SyncMultiviewDiffusion.prepare (morphable_diffusion.py:473-486, encode_first_stage :460-466):
 * moments = quant_conv(Encoder(image)) (ldm/models/autoencoder.py:324-328, ldm/modules/diffusionmodules/model.py:368-459).
 * image [n][3][8S][8S] in [-1, 1] (device, fp32, NCHW), moments [n][8][S][S] = mean | logvar; the caller draws the
 * posterior sample (mean + exp(0.5 logvar) * noise) and applies the 0.18215 scale, as DiagonalGaussianDistribution does.
 * Needs first_stage_model.encoder.* / quant_conv.* among the loaded weights (md_has_vae_encoder tells). */
MD_API int md_vae_encode(md_ctx* ctx, const float* image, float* moments, int n, int latent_size, void* stream);
MD_API int md_has_vae_encoder(md_ctx* ctx);
/* The CLIP half of SyncMultiviewDiffusion.prepare (morphable_diffusion.py:487-488): FrozenCLIPImageEmbedder.encode
 * (ldm/modules/encoders/modules.py:363-382) = bicubic resize to 224 (align_corners), CLIP normalisation, ViT-L/14 image
 * tower, projection.  image [n][3][H][W] in [-1, 1] (device, fp32, NCHW) -> embed [n][768] (the caller adds the token axis).
 * Needs clip_image_encoder.model.visual.* among the loaded weights (md_has_clip tells). */
MD_API int md_clip_embed(md_ctx* ctx, const float* image, float* embed, int n, int H, int W, void* stream);
MD_API int md_has_clip(md_ctx* ctx);
MD_API int md_ddim_timestep(md_ctx* ctx, int index);  /* 1 .. 981 */
/* SyncDDIMSampler(model, ddim_num_steps, "uniform", ddim_eta) (morphable_diffusion.py:649-672): rebuilds the DDIM
 * schedule (timesteps range(0,1000,1000/steps)+1, alphas, alphas_prev, sigmas) of the context.  Cheap; may be called
 * between steps (the per-step scalars are passed to the captured step through a device buffer). */
MD_API int md_set_ddim(md_ctx* ctx, int ddim_steps, float ddim_eta);
MD_API int md_ddim_steps(md_ctx* ctx);


/* ------------------------------------------------------------------ op level: the other kernels of the step
 * (exposed so each kernel can be parity-tested in isolation through the C ABI; all pointers are device pointers) */

/* torch.nn.GroupNorm + optional SiLU/ReLU over channels-last x [B][rows][C] (fp32 or bf16) -> bf16 [B][rows][C].
 * Reference: GroupNorm32 (ldm/modules/diffusionmodules/util.py:199-216), Normalize (ldm/modules/attention.py:85-86),
 * nn.GroupNorm(8, c) in network.py:166,190,288 and attention.py:56-69.  addvec [B][C] (optional) is added to x
 * before normalising (FrustumTVBlock: x + t_conv(t) + v_conv(v), network.py:295). */
MD_API int md_op_group_norm(const void* x, int x_is_bf16, int B, int rows, int C, int groups, float eps,
                            const float* gamma, const float* beta, const float* addvec, int act, void* out_bf16,
                            void* stream);
/* Same GroupNorm with the per-(sample, channel) statistics supplied by the caller: stats fp32 [B][C][2] = (sum, sum of
 * squares) over the rows, exactly what md_op_conv_gemm's col_stats epilogue accumulates.  This is the path the step
 * uses after every tensor-core GEMM (finalize + apply in one kernel). */
MD_API int md_op_group_norm_stats(const void* x, int x_is_bf16, int B, int rows, int C, int groups, float eps,
                                  const float* gamma, const float* beta, const float* addvec, int act,
                                  const float* stats, void* out_bf16, void* stream);
/* nn.LayerNorm over the last dim of x fp32 [rows][C] -> bf16 (ldm/modules/attention.py:257-259) */
/* row softmax fp32 [rows][n] -> bf16 (AttnBlock of the first-stage decoder, model.py:190-192); n a multiple of 4 */
MD_API int md_op_softmax_rows(const float* x, void* out_bf16, long long rows, int n, void* stream);
MD_API int md_op_layer_norm(float* x, const float* gamma, const float* beta, void* out_bf16, long long rows, int C,
                            float eps, void* stream);
/* CrossAttention.forward as self-attention (ldm/modules/attention.py:179-203): qkv bf16 [B][S][3*heads*dh] ->
 * out bf16 [B][S][heads*dh] */
MD_API int md_op_self_attention(const void* qkv, void* out, int B, int S, int heads, int dh, void* stream);
/* Same contract with the kernel chosen explicitly (parity / timing of the two implementations against each other):
 * impl 0 = the library's own choice, 1 = mma.sync flash kernel, 2 = tcgen05/TMEM kernel (S % 128 == 0, dh <= 128). */
MD_API int md_op_self_attention_impl(const void* qkv, void* out, int B, int S, int heads, int dh, int impl,
                                     void* stream);
/* DepthAttention.forward core (ldm/models/diffusion/attention.py:36-46), re-associated: K = W_k c and V = W_v c are
 * never materialised.  qp bf16 [T][HW][4*ctx] = per-head W_k^T q (softmax scale folded in); c1 bf16 [T][D][HW][ctx]
 * = proj_context conv output BEFORE GroupNorm; ss fp32 [T][ctx][2] = its GroupNorm scale/shift; beta fp32 [ctx];
 * out cbar bf16 [B][HW][4*ctx] = per-head sum_d softmax_d(qp . c_d) c_d with c = ReLU(c1*scale+shift).
 * Samples T..B-1 have an all-zero frustum volume (CFG-unconditional half) and receive ReLU(beta). */
MD_API int md_op_depth_attention(const void* qp, const void* c1, const float* ss, const float* beta, void* cbar, int T,
                                 int B, int D, int HW, int ctx, void* stream);
/* CFG combine + DDIM update (morphable_diffusion.py:147-148,675-698) with the schedule of `ctx`.
 * eps [2T or T][n], x [T][n] in place; noise NULL -> Philox(seed, index, view0 + t). */
MD_API int md_op_cfg_ddim(md_ctx* ctx, const float* eps, float* x, float* eps_out, const float* noise, int T,
                          int n_per_view, int index, float cfg_scale, unsigned long long seed, int view0, void* stream);

/* Multi-GPU: one process per GPU; NCCL communicator created from a 128-byte unique id distributed by the host
 * side (torch.distributed).  The only exchange per step is an all-reduce(sum) of the [nv][16] vertex features. */
MD_API int md_comm_unique_id(void* id128);
MD_API int md_comm_init(md_ctx* ctx, int rank, int world, const void* id128);
/* NVLink peer exchange (after md_comm_init): with it attached, md_denoise_step no longer calls NCCL.  Every rank pushes
 * its partial vertex-feature sums into its region of every peer's exchange buffer with plain stores over NVLink and
 * raises a per-rank arrival flag; the kernel that consumes the sums (view mean + Conv1d + voxel scatter) waits for the
 * flags in its own memory and adds the regions in rank order, so the exchange is fused into its producer / consumer
 * kernels and the whole step is one CUDA graph on every rank.
 *   md_peer_buffer : allocates this rank's buffer and writes its 64-byte cudaIpcMemHandle_t to ipc_handle64
 *   md_peer_attach : ipc_handles = world x 64 bytes in rank order (gathered by the host side, e.g. all_gather_object);
 *                    the caller runs a barrier over all ranks before the next md_denoise_step
 * Every rank must then call md_denoise_step the same number of times (a missing rank is reported after ~2 s by the next
 * call instead of hanging the device).  Meshes above 16 384 vertices keep the NCCL path. */
MD_API int md_peer_buffer(md_ctx* ctx, int world, void* ipc_handle64);
MD_API int md_peer_attach(md_ctx* ctx, int rank, int world, const void* ipc_handles);
MD_API int md_peer_attached(md_ctx* ctx);
/* Back to the NCCL all-reduce (every rank must call it before its next md_denoise_step; used by the host side when
 * md_peer_attach failed on some rank, e.g. a launcher that hides the peers' devices from each process). */
MD_API int md_peer_detach(md_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MDIFF_H_ */
