// Context lifecycle, stage-level C ABI and the fused denoise step (SyncDDIMSampler.denoise_apply,
// morphable_diffusion.py:701-739).
#include "engine.h"
#include "ptx.cuh"

#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>

struct md_ctx { md::Ctx c; };

namespace md {

// ------------------------------------------------------------------ tiny kernels local to the step
__global__ void fill_kernel(float* p, float v, int n) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void fill_from_kernel(float* p, const float* src, int n) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = src[0];
}
__global__ void set_step_params_kernel(float* d, float tval, float a_t, float a_prev, float sigma, float s1m, int add_noise,
                                       uint32_t index, uint64_t seed, unsigned* xchg_seq, unsigned seq) {
  pdl_grid_sync();
  if (threadIdx.x == 0) {
    if (xchg_seq) xchg_seq[0] = seq;  // sequence number of this step's peer exchange (read by the captured kernels)
    d[0] = tval; d[1] = a_t; d[2] = a_prev; d[3] = sigma; d[4] = s1m; d[5] = add_noise ? 1.f : 0.f;
    d[6] = __uint_as_float(index);
    d[7] = __uint_as_float(static_cast<uint32_t>(seed)); d[8] = __uint_as_float(static_cast<uint32_t>(seed >> 32));
  }
}
// context rows: first T = clip embedding, remaining (uncond) = 0   (morphable_diffusion.py:135)
__global__ void make_context_kernel(const float* __restrict__ clip, float* __restrict__ out, int T, int B, int dim) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * dim) return;
  out[i] = (i / dim < T) ? clip[i % dim] : 0.f;
}
__global__ void ncdhw_to_cl_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int C, size_t S) {
  pdl_grid_sync();
  const size_t total = static_cast<size_t>(C) * S;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t s = i / C;
    out[i] = x[static_cast<size_t>(c) * S + s];
  }
}

// ------------------------------------------------------------------ NCCL through dlopen (torch already loaded it)
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ struct Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
struct Id128 { char b[128]; };
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return set_error("NCCL library not found (dlopen libnccl.so.2): %s", dlerror());
  *reinterpret_cast<void**>(&g_nccl.GetUniqueId) = dlsym(g_nccl.h, "ncclGetUniqueId");
  *reinterpret_cast<void**>(&g_nccl.CommInitRank) = dlsym(g_nccl.h, "ncclCommInitRank");
  *reinterpret_cast<void**>(&g_nccl.AllReduce) = dlsym(g_nccl.h, "ncclAllReduce");
  *reinterpret_cast<void**>(&g_nccl.AllGather) = dlsym(g_nccl.h, "ncclAllGather");
  *reinterpret_cast<void**>(&g_nccl.CommDestroy) = dlsym(g_nccl.h, "ncclCommDestroy");
  *reinterpret_cast<void**>(&g_nccl.GetErrorString) = dlsym(g_nccl.h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) return set_error("NCCL symbols missing");
  return 0;
}

static void make_schedule(Ctx& c) {
  // _init_schedule (morphable_diffusion.py:428-438) + make_ddim_timesteps + _make_schedule (:658-672)
  const int T = 1000;
  std::vector<float> acp(T);
  const float s0 = sqrtf(0.00085f), s1 = sqrtf(0.0120f);
  float prod = 1.f;
  const float step = (s1 - s0) / static_cast<float>(T - 1);
  for (int i = 0; i < T; ++i) {
    const float b = (i < T / 2) ? (s0 + step * i) : (s1 - step * (T - 1 - i));
    const float beta = b * b;
    prod *= (1.f - beta);
    acp[i] = prod;
  }
  const int n = c.mcfg.ddim_steps;
  const int stride = T / n;
  c.timesteps.clear(); c.alphas.clear(); c.alphas_prev.clear(); c.sigmas.clear(); c.sqrt_1m_alphas.clear();
  for (int i = 0; i * stride < T; ++i) c.timesteps.push_back(i * stride + 1);
  for (size_t i = 0; i < c.timesteps.size(); ++i) {
    const double a = acp[c.timesteps[i]];
    const double ap = (i == 0) ? acp[0] : acp[c.timesteps[i - 1]];
    const double sig = c.mcfg.ddim_eta * sqrt((1.0 - ap) / (1.0 - a) * (1.0 - a / ap));
    c.alphas.push_back(static_cast<float>(a));
    c.alphas_prev.push_back(static_cast<float>(ap));
    c.sigmas.push_back(static_cast<float>(sig));
    c.sqrt_1m_alphas.push_back(sqrtf(1.f - static_cast<float>(a)));
  }
}

static int ensure_ready(md_ctx* ctx, bool need_binding) {
  if (!ctx) return set_error("null context");
  if (!ctx->c.weights_loaded) return set_error("weights not loaded (md_load_weights)");
  if (need_binding && !ctx->c.sb.bound) return set_error("no sample bound (md_bind_sample)");
  ctx->c.arena.failed = false;
  return 0;
}

// the single cross-rank exchange of the step: sum over ranks of the per-vertex features
static int allreduce_vsum(Ctx& c, cudaStream_t st) {
  if (c.world <= 1) return 0;
  if (!c.nccl_comm) return set_error("denoise_step: world=%d but no communicator (md_comm_init)", c.world);
  const int r = g_nccl.AllReduce(c.vsum_ptr, c.vsum_ptr, static_cast<size_t>(c.sb.nv) * 16, /*ncclFloat32*/ 7,
                                 /*ncclSum*/ 0, c.nccl_comm, st);
  if (r != 0) return set_error("ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return 0;
}

// The step's cross-rank exchange runs over NVLink peer memory (md_peer_attach) instead of NCCL when the exchange buffers
// are attached and the mesh fits a region; MD_PEER=0 keeps the NCCL all-reduce between two graphs (A/B switch).
static bool peer_exchange_active(const Ctx& c) {
  static const bool env_on = !(getenv("MD_PEER") != nullptr && atoi(getenv("MD_PEER")) == 0);
  return env_on && c.world > 1 && c.px.on && static_cast<size_t>(c.sb.nv) * 16 <= c.px.slot_floats;
}

// one denoise step for the local views (all chunks)
static int denoise_step_impl(Ctx& c, float* x_local, const float* x_input, const float* clip, int index,
                             float cfg_scale, const float* noise, unsigned long long seed, float* eps_out,
                             int do_update, cudaStream_t st, int phase = 0) {
  // phase 0: whole step (the all-reduce is issued inline); phase 1: only the part before the cross-rank exchange;
  // phase 2: only the part after it.  Phases 1/2 are captured as two CUDA graphs with the NCCL call between them.
  const md_config& mc = c.mcfg;
  const SampleBinding& sb = c.sb;
  if (index < 0 || index >= static_cast<int>(c.timesteps.size())) return set_error("denoise_step: bad DDIM index %d", index);
  const int S = mc.latent_size, V = mc.spatial_volume_size, D = mc.frustum_depth;
  const int HW = S * S;
  const float tval = static_cast<float>(c.timesteps[index]);
  const int cfg = cfg_scale != 1.0f ? 1 : 0;
  Arena& A = c.arena;
  A.off = 0;
  const int chunk = std::min(sb.n_local, mc.max_views_per_call > 0 ? mc.max_views_per_call : 16);
  const int maxB = (cfg ? 2 : 1) * chunk;

  float* d_t = A.get<float>(maxB);
  float* t_embed = A.get<float>(mc.time_embed_dim);
  float* vsum = A.get<float>(static_cast<size_t>(sb.nv) * 16);
  float* vol = A.get<float>(static_cast<size_t>(V) * V * V * 64);
  float* eps_all = A.get<float>(static_cast<size_t>(maxB) * 4 * HW);
  float* ctxv = A.get<float>(static_cast<size_t>(maxB) * mc.context_dim);
  float* x_in = A.get<float>(static_cast<size_t>(maxB) * HW * 8);
  if (A.failed) return set_error("workspace exhausted (step)");
  (void)tval;
  c.vsum_ptr = vsum;
  // Conditioning branch (target-view encoder -> vertex features -> sparse conv -> spatial volume -> frustum nets) and
  // the UNet's input half are independent: with a single view chunk the former runs on a second stream (own arena, own
  // split-K workspace) and the UNet joins it right before its first depth transformer.  On one rank the whole branch
  // moves over (the encoder is one CTA per view: 0.25 ms during which the UNet would otherwise wait); with several
  // ranks the part before the cross-rank exchange stays on the main stream (it is captured as its own graph).
  const bool overlap = c.stream2 != nullptr && sb.n_local <= chunk && getenv("MD_NO_OVERLAP") == nullptr;
  const bool peer = peer_exchange_active(c);  // then the step is one graph (phase 0) on every rank
  const bool vf_on_side = overlap && phase == 0 && (c.world <= 1 || peer);
  if (phase != 2) {
    launch_pdl(fill_from_kernel, dim3((maxB + 63) / 64), dim3(64), 0, st, d_t, c.d_step, maxB);  // timestep of this index (d_step[0])
    MD_CHECK(check_launch("fill"));
    MD_CHECK(embed_time(c, d_t, t_embed, st));
    if (!vf_on_side) MD_CHECK(vertex_feature_sum(c, x_local, t_embed, vsum, st));
  }
  if (phase == 0 && !peer) MD_CHECK(allreduce_vsum(c, st));
  if (phase == 1) return 0;

  const size_t m = A.mark();
  for (int lv0 = 0; lv0 < sb.n_local; lv0 += chunk) {
    const int T = std::min(chunk, sb.n_local - lv0);
    const int B = (cfg ? 2 : 1) * T;
    bf16* levels[4];
    if (overlap) {
      MD_CUDA(cudaEventRecord(c.ev_fork, st));
      MD_CUDA(cudaStreamWaitEvent(c.stream2, c.ev_fork, 0));
      std::swap(c.arena, c.arena2);
      c.arena.off = 0;
      c.arena.failed = false;
      int rc;
      {
        SplitScope side(&c.split_side);
        rc = vf_on_side ? vertex_feature_sum(c, x_local, t_embed, vsum, c.stream2) : 0;
        if (rc == 0) rc = spatial_volume_from_vsum(c, vsum, vol, c.stream2, peer);
        if (rc == 0) rc = frustum_levels(c, vol, lv0, T, t_embed, T, levels, c.stream2);
      }
      std::swap(c.arena, c.arena2);
      if (rc != 0) return rc;
      MD_CUDA(cudaEventRecord(c.ev_levels, c.stream2));
      c.levels_pending = true;
    } else {
      if (lv0 == 0) MD_CHECK(spatial_volume_from_vsum(c, vsum, vol, st, peer));
      MD_CHECK(frustum_levels(c, vol, lv0, T, t_embed, T, levels, st));
    }
    MD_CHECK(launch_unet_input(x_local + static_cast<size_t>(lv0) * 4 * HW, x_input, 0, x_in, T, HW, cfg, st));
    launch_pdl(make_context_kernel, dim3((B * mc.context_dim + 255) / 256), dim3(256), 0, st, clip, ctxv, T, B, mc.context_dim);
    MD_CHECK(check_launch("make_context"));
    MD_CHECK(unet_forward(c, x_in, d_t, ctxv, levels, B, T, S, D, eps_all, st));
    if (c.levels_pending) {  // defensive: a UNet without depth transformers would never have joined
      MD_CUDA(cudaStreamWaitEvent(st, c.ev_levels, 0));
      c.levels_pending = false;
    }
    const int add_noise = (index != 0) ? 1 : 0;
    MD_CHECK(launch_cfg_ddim(eps_all, x_local + static_cast<size_t>(lv0) * 4 * HW,
                             eps_out ? eps_out + static_cast<size_t>(lv0) * 4 * HW : nullptr,
                             noise ? noise + static_cast<size_t>(lv0) * 4 * HW : nullptr, T, 4 * HW, cfg, cfg_scale,
                             c.alphas[index], c.alphas_prev[index], c.sigmas[index], c.sqrt_1m_alphas[index],
                             add_noise, seed, static_cast<uint32_t>(index), sb.view0 + lv0, do_update, c.d_step, st));
    A.release(m);
  }
  return 0;
}

}  // namespace md

using namespace md;

extern "C" {

void md_default_config(md_config* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->model_channels = 320; cfg->in_channels = 8; cfg->out_channels = 4; cfg->num_res_blocks = 2;
  cfg->num_heads = 8; cfg->context_dim = 768;
  const int cm[4] = {1, 2, 4, 4}, at[4] = {1, 1, 1, 0}, vd[4] = {64, 128, 256, 512};
  for (int i = 0; i < 4; ++i) { cfg->channel_mult[i] = cm[i]; cfg->attn_ds[i] = at[i]; cfg->volume_dims[i] = vd[i]; }
  cfg->latent_size = 32; cfg->image_size = 256; cfg->spatial_volume_size = 32; cfg->frustum_depth = 48;
  cfg->time_embed_dim = 256; cfg->view_dim = 4;
  cfg->spatial_volume_length = 0.5f; cfg->frustum_volume_length = 0.86603f;
  cfg->smpl_num_views = 0;
  cfg->ddim_steps = 50; cfg->ddim_eta = 1.0f;
  cfg->max_views_per_call = 16; cfg->workspace_bytes = 0;
}

int md_create(md_ctx** out, const md_config* cfg) {
  if (!out) return set_error("md_create: null out");
  md_config c;
  if (cfg) c = *cfg; else md_default_config(&c);
  if (c.volume_dims[0] != 64) return set_error("md_create: volume_dims[0] must be 64 (spatial volume channels)");
  if (c.latent_size % 8) return set_error("md_create: latent_size must be a multiple of 8");
  md_ctx* ctx = new md_ctx();
  ctx->c.mcfg = c;
  make_schedule(ctx->c);
  const int mv = c.max_views_per_call > 0 ? c.max_views_per_call : 16;
  const double scale = (c.latent_size / 32.0) * (c.latent_size / 32.0);
  size_t bytes = c.workspace_bytes ? static_cast<size_t>(c.workspace_bytes)
                                   : static_cast<size_t>((1.0 + 0.45 * mv * scale) * (1ull << 30));
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ctx->c.arena.base), bytes);
  if (e != cudaSuccess) {
    delete ctx;
    return set_error("md_create: cudaMalloc(%zu) for the workspace failed: %s", bytes, cudaGetErrorString(e));
  }
  ctx->c.arena.cap = bytes;
  ctx->c.gn_stats_floats = static_cast<size_t>(2) * (2 * mv) * 4096;
  e = cudaMalloc(reinterpret_cast<void**>(&ctx->c.gn_stats), ctx->c.gn_stats_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(ctx->c.gn_stats, 0, ctx->c.gn_stats_floats * sizeof(float));
  if (e != cudaSuccess) {
    cudaFree(ctx->c.arena.base);
    delete ctx;
    return set_error("md_create: GroupNorm scratch allocation failed: %s", cudaGetErrorString(e));
  }
  e = cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->c.ev_in, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->c.ev_out, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->c.d_step), 16 * sizeof(float));
  if (e != cudaSuccess) {
    cudaFree(ctx->c.arena.base);
    cudaFree(ctx->c.gn_stats);
    delete ctx;
    return set_error("md_create: stream / event setup failed: %s", cudaGetErrorString(e));
  }
  {  // second stream for the conditioning branch
    const size_t bytes2 = (static_cast<size_t>(384) << 20) + static_cast<size_t>(mv * scale * 48.0) * (1ull << 20);
    cudaError_t e2 = cudaMalloc(reinterpret_cast<void**>(&ctx->c.arena2.base), bytes2);
    if (e2 == cudaSuccess) {
      ctx->c.arena2.cap = bytes2;
      e2 = cudaStreamCreateWithFlags(&ctx->c.stream2, cudaStreamNonBlocking);
    }
    if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&ctx->c.ev_fork, cudaEventDisableTiming);
    if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&ctx->c.ev_levels, cudaEventDisableTiming);
    if (e2 == cudaSuccess && alloc_split_workspace(ctx->c.split_side, static_cast<size_t>(24) << 20, 1 << 14) != 0)
      e2 = cudaErrorMemoryAllocation;
    if (e2 != cudaSuccess) { cudaGetLastError(); ctx->c.stream2 = nullptr; }  // overlap simply stays off
  }
  if (alloc_split_workspace(ctx->c.split_main, static_cast<size_t>(32) << 20, 1 << 14) != 0) {
    md_destroy(ctx);
    return -1;
  }
  ctx->c.use_graph = getenv("MD_NO_GRAPH") == nullptr;
  *out = ctx;
  return 0;
}

void md_destroy(md_ctx* ctx) {
  if (!ctx) return;
  cudaDeviceSynchronize();
  free_binding(ctx->c);
  free_weights(ctx->c);
  if (ctx->c.nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->c.nccl_comm);
  {
    Ctx::PeerExchange& px = ctx->c.px;
    for (int p = 0; p < 16; ++p)
      if (px.peer_base[p] && px.peer_base[p] != px.base) cudaIpcCloseMemHandle(px.peer_base[p]);
    cudaFree(px.d_peer_data); cudaFree(px.d_peer_flags); cudaFree(px.d_seq);
    if (px.h_err) cudaFreeHost(px.h_err);
    cudaFree(px.base);
  }
  if (ctx->c.graph) cudaGraphExecDestroy(ctx->c.graph);
  if (ctx->c.graph_b) cudaGraphExecDestroy(ctx->c.graph_b);
  if (ctx->c.stream) cudaStreamDestroy(ctx->c.stream);
  if (ctx->c.ev_in) cudaEventDestroy(ctx->c.ev_in);
  if (ctx->c.ev_out) cudaEventDestroy(ctx->c.ev_out);
  cudaFree(ctx->c.d_step);
  if (ctx->c.stream2) cudaStreamDestroy(ctx->c.stream2);
  if (ctx->c.ev_fork) cudaEventDestroy(ctx->c.ev_fork);
  if (ctx->c.ev_levels) cudaEventDestroy(ctx->c.ev_levels);
  cudaFree(ctx->c.arena2.base);
  free_split_workspace(ctx->c.split_main);
  free_split_workspace(ctx->c.split_side);
  cudaFree(ctx->c.arena.base);
  cudaFree(ctx->c.gn_stats);
  delete ctx;
}

unsigned long long md_workspace_peak(md_ctx* ctx) { return ctx ? ctx->c.arena.peak : 0; }

int md_load_weights(md_ctx* ctx, int n, const char* const* names, const void* const* ptrs, const long long* numels,
                    void* stream) {
  if (!ctx) return set_error("null context");
  TensorMap tm;
  tm.reserve(static_cast<size_t>(n) * 2);
  for (int i = 0; i < n; ++i) {
    NamedTensor t;
    t.ptr = static_cast<const float*>(ptrs[i]);
    t.numel = static_cast<size_t>(numels[i]);
    tm[names[i]] = t;
  }
  ctx->c.weights_gen++;
  return load_all_weights(ctx->c, tm, static_cast<cudaStream_t>(stream));
}

int md_bind_sample(md_ctx* ctx, const float* K, const float* RT, const float* v_embed, const float* vertices,
                   const int32_t* coord, const int32_t* out_sh, const float* bounds, int nv, int n_views, int view0,
                   int n_local, int projection, void* stream) {
  if (!ctx) return set_error("null context");
  if (projection != 0 && projection != 1) return set_error("NotImplementedError: projection %d", projection);
  ctx->c.bind_gen++;
  return bind_sample(ctx->c, K, RT, v_embed, vertices, coord, out_sh, bounds, nv, n_views, view0, n_local, projection,
                     static_cast<cudaStream_t>(stream));
}

int md_voxelize(const float* vertices, int nv, int32_t* coord, int32_t* out_sh, float* bounds, void* stream) {
  return launch_voxelize(vertices, nv, coord, out_sh, bounds, static_cast<cudaStream_t>(stream));
}

int md_affine_points(const float* v, int n, const float* A9_host, const float* b3_host, float* out, void* stream) {
  if (!v || !out || !A9_host || !b3_host || n < 0) return set_error("md_affine_points: bad arguments");
  return launch_affine_points(v, n, A9_host, b3_host, out, static_cast<cudaStream_t>(stream));
}

int md_images_to_u8(const float* img, unsigned char* out, int n, int H, int W, void* stream) {
  if (!img || !out || n < 0 || H < 1 || W < 1) return set_error("md_images_to_u8: bad arguments");
  return launch_images_to_u8(img, out, n, H * W, static_cast<cudaStream_t>(stream));
}

int md_embed_time(md_ctx* ctx, float timestep, float* t_embed_out, void* stream) {
  MD_CHECK(ensure_ready(ctx, false));
  Ctx& c = ctx->c;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  c.arena.off = 0;
  float* d_t = c.arena.get<float>(1);
  if (c.arena.failed) return set_error("workspace exhausted");
  launch_pdl(fill_kernel, dim3(1), dim3(32), 0, st, d_t, timestep, 1);
  MD_CHECK(check_launch("fill"));
  return embed_time(c, d_t, t_embed_out, st);
}

int md_spatial_volume(md_ctx* ctx, const float* x_local, const float* t_embed, float* volume_out, void* stream) {
  MD_CHECK(ensure_ready(ctx, true));
  Ctx& c = ctx->c;
  SplitScope split_scope(&c.split_main);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const md_config& mc = c.mcfg;
  const int V = mc.spatial_volume_size;
  Arena& A = c.arena;
  A.off = 0;
  float* vsum = A.get<float>(static_cast<size_t>(c.sb.nv) * 16);
  float* vol = A.get<float>(static_cast<size_t>(V) * V * V * 64);
  if (A.failed) return set_error("workspace exhausted");
  MD_CHECK(vertex_feature_sum(c, x_local, t_embed, vsum, st));
  c.vsum_ptr = vsum;
  if (c.world > 1 && c.nccl_comm) MD_CHECK(allreduce_vsum(c, st));
  MD_CHECK(spatial_volume_from_vsum(c, vsum, vol, st));
  return launch_cl_to_ncdhw(vol, 0, volume_out, 1, 64, static_cast<size_t>(V) * V * V, st);
}

int md_frustum_feats(md_ctx* ctx, const float* volume, int lv0, int T, const float* t_embed,
                     float* const out_levels[4], void* stream) {
  MD_CHECK(ensure_ready(ctx, true));
  Ctx& c = ctx->c;
  SplitScope split_scope(&c.split_main);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const md_config& mc = c.mcfg;
  const int V = mc.spatial_volume_size, S = mc.latent_size, D = mc.frustum_depth;
  Arena& A = c.arena;
  A.off = 0;
  float* vol = A.get<float>(static_cast<size_t>(V) * V * V * 64);
  if (A.failed) return set_error("workspace exhausted");
  launch_pdl(ncdhw_to_cl_f32_kernel, dim3(148 * 8), dim3(256), 0, st, volume, vol, 64, static_cast<size_t>(V) * V * V);
  MD_CHECK(check_launch("ncdhw_to_cl_f32"));
  bf16* levels[4];
  MD_CHECK(frustum_levels(c, vol, lv0, T, t_embed, T, levels, st));
  for (int i = 0; i < 4; ++i) {
    const size_t sp = static_cast<size_t>(D >> i) * (S >> i) * (S >> i);
    MD_CHECK(launch_cl_to_ncdhw(levels[i], 1, out_levels[i], T, mc.volume_dims[i], sp, st));
  }
  return 0;
}

int md_unet_forward(md_ctx* ctx, const float* x, const float* timesteps_host, const float* context,
                    const float* const source[4], int B, float* out, void* stream) {
  MD_CHECK(ensure_ready(ctx, false));
  Ctx& c = ctx->c;
  SplitScope split_scope(&c.split_main);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const md_config& mc = c.mcfg;
  const int S = mc.latent_size, D = mc.frustum_depth;
  Arena& A = c.arena;
  A.off = 0;
  float* d_t = A.get<float>(B);
  float* x_in = A.get<float>(static_cast<size_t>(B) * S * S * mc.in_channels);
  bf16* levels[4];
  for (int i = 0; i < 4; ++i) {
    const size_t sp = static_cast<size_t>(D >> i) * (S >> i) * (S >> i);
    levels[i] = A.get<bf16>(static_cast<size_t>(B) * sp * mc.volume_dims[i]);
  }
  if (A.failed) return set_error("workspace exhausted");
  MD_CUDA(cudaMemcpyAsync(d_t, timesteps_host, sizeof(float) * B, cudaMemcpyHostToDevice, st));
  for (int i = 0; i < 4; ++i) {
    const size_t sp = static_cast<size_t>(D >> i) * (S >> i) * (S >> i);
    MD_CHECK(launch_ncdhw_to_cl_bf16(source[i], levels[i], B, mc.volume_dims[i], sp, st));
  }
  // NCHW -> NHWC for the 8-channel input (same transposition kernel, fp32 variant per sample)
  for (int b = 0; b < B; ++b) {
    launch_pdl(ncdhw_to_cl_f32_kernel, dim3(32), dim3(256), 0, st, x + static_cast<size_t>(b) * mc.in_channels * S * S,
                                               x_in + static_cast<size_t>(b) * mc.in_channels * S * S, mc.in_channels,
                                               static_cast<size_t>(S) * S);
    MD_CHECK(check_launch("nchw_to_nhwc"));
  }
  return unet_forward(c, x_in, d_t, context, levels, B, B, S, D, out, st);
}

int md_denoise_step(md_ctx* ctx, float* x_local, const float* x_input, const float* clip_embed, int index,
                    float cfg_scale, const float* noise, unsigned long long seed, float* eps_out, void* stream) {
  MD_CHECK(ensure_ready(ctx, true));
  Ctx& c = ctx->c;
  SplitScope split_scope(&c.split_main);
  if (index < 0 || index >= static_cast<int>(c.timesteps.size())) return set_error("denoise_step: bad DDIM index %d", index);
  cudaStream_t caller = static_cast<cudaStream_t>(stream);
  cudaStream_t st = c.stream;
  // fence the caller's stream into the internal one (the legacy default stream cannot be captured)
  MD_CUDA(cudaEventRecord(c.ev_in, caller));
  MD_CUDA(cudaStreamWaitEvent(st, c.ev_in, 0));
  const bool peer = peer_exchange_active(c);
  if (peer) {
    if (*c.px.h_err != 0)
      return set_error("denoise_step: peer exchange timed out waiting for rank %d (a rank skipped a step, or died)", *c.px.h_err - 1);
    ++c.px.seq;
  }
  launch_pdl(set_step_params_kernel, dim3(1), dim3(32), 0, st, c.d_step, static_cast<float>(c.timesteps[index]), c.alphas[index],
                                           c.alphas_prev[index], c.sigmas[index], c.sqrt_1m_alphas[index], index != 0,
                                           static_cast<uint32_t>(index), seed, peer ? c.px.d_seq : static_cast<unsigned*>(nullptr),
                                           c.px.seq);
  MD_CHECK(check_launch("set_step_params"));
  const Ctx::GraphKey key{x_local, x_input, clip_embed, noise, eps_out, cfg_scale, c.bind_gen, c.weights_gen};
  const bool same = memcmp(&key, &c.gkey, sizeof(key)) == 0;
  int rc = 0;
  if (!c.use_graph) {
    rc = denoise_step_impl(c, x_local, x_input, clip_embed, index, cfg_scale, noise, seed, eps_out, 1, st);
  } else if (same && c.graph) {
    cudaError_t e = cudaGraphLaunch(c.graph, st);
    if (e == cudaSuccess && c.graph_b) {  // multi-rank: [graph A] -> NCCL all-reduce (not captured) -> [graph B]
      rc = allreduce_vsum(c, st);
      if (rc == 0) e = cudaGraphLaunch(c.graph_b, st);
    }
    if (e != cudaSuccess) rc = set_error("cudaGraphLaunch: %s", cudaGetErrorString(e));
    else if (rc == 0) count_launch(c.graph_launches);
  } else {
    if (!same) {
      if (c.graph) { cudaGraphExecDestroy(c.graph); c.graph = nullptr; }
      if (c.graph_b) { cudaGraphExecDestroy(c.graph_b); c.graph_b = nullptr; }
      c.gkey = key;
      c.graph_warm = 0;
    }
    if (c.graph_warm == 0) {  // first call with these pointers: plain launches (also sets function attributes)
      rc = denoise_step_impl(c, x_local, x_input, clip_embed, index, cfg_scale, noise, seed, eps_out, 1, st);
      c.graph_warm = 1;
    } else {                  // second call: capture, instantiate, launch
      const long long before = md_launch_count();
      const bool split = c.world > 1 && !peer;
      auto capture = [&](int phase, cudaGraphExec_t* out) -> int {
        cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
        if (e != cudaSuccess) return set_error("cudaStreamBeginCapture: %s", cudaGetErrorString(e));
        int r = denoise_step_impl(c, x_local, x_input, clip_embed, index, cfg_scale, noise, seed, eps_out, 1, st, phase);
        cudaGraph_t g = nullptr;
        e = cudaStreamEndCapture(st, &g);
        if (r == 0 && e != cudaSuccess) r = set_error("cudaStreamEndCapture: %s", cudaGetErrorString(e));
        if (r == 0) {
          e = cudaGraphInstantiate(out, g, 0);
          if (e != cudaSuccess) r = set_error("cudaGraphInstantiate: %s", cudaGetErrorString(e));
        }
        if (g) cudaGraphDestroy(g);
        return r;
      };
      rc = capture(split ? 1 : 0, &c.graph);
      if (rc == 0 && split) rc = capture(2, &c.graph_b);
      c.graph_launches = static_cast<int>(md_launch_count() - before);
      if (rc == 0) {
        cudaError_t e = cudaGraphLaunch(c.graph, st);
        if (e == cudaSuccess && split) {
          rc = allreduce_vsum(c, st);
          if (rc == 0) e = cudaGraphLaunch(c.graph_b, st);
        }
        if (e != cudaSuccess) rc = set_error("cudaGraphLaunch: %s", cudaGetErrorString(e));
      }
      if (rc != 0) {
        if (c.graph) { cudaGraphExecDestroy(c.graph); c.graph = nullptr; }
        if (c.graph_b) { cudaGraphExecDestroy(c.graph_b); c.graph_b = nullptr; }
      }
    }
  }
  MD_CUDA(cudaEventRecord(c.ev_out, st));
  MD_CUDA(cudaStreamWaitEvent(caller, c.ev_out, 0));
  return rc;
}

int md_has_vae(md_ctx* ctx) { return ctx && ctx->c.vae.loaded ? 1 : 0; }

int md_vae_decode(md_ctx* ctx, const float* x, float* image, int n, int latent_size, void* stream) {
  MD_CHECK(ensure_ready(ctx, false));
  Ctx& c = ctx->c;
  SplitScope split_scope(&c.split_main);
  if (latent_size <= 0) latent_size = c.mcfg.latent_size;
  // views are independent: chunks bound the workspace (and the 32-bit row offsets of the GEMM epilogue)
  const int chunk = std::max(1, c.mcfg.max_views_per_call > 0 ? std::min(c.mcfg.max_views_per_call, 16) : 16);
  const size_t in_per = static_cast<size_t>(4) * latent_size * latent_size;
  const size_t out_per = static_cast<size_t>(3) * 64 * latent_size * latent_size;
  for (int i0 = 0; i0 < n; i0 += chunk) {
    const int T = std::min(chunk, n - i0);
    MD_CHECK(vae_decode(c, x + i0 * in_per, image + i0 * out_per, T, latent_size, static_cast<cudaStream_t>(stream)));
  }
  return 0;
}

int md_has_vae_encoder(md_ctx* ctx) { return ctx && ctx->c.vae_enc.loaded ? 1 : 0; }

int md_vae_encode(md_ctx* ctx, const float* image, float* moments, int n, int latent_size, void* stream) {
  MD_CHECK(ensure_ready(ctx, false));
  Ctx& c = ctx->c;
  SplitScope split_scope(&c.split_main);
  if (latent_size <= 0) latent_size = c.mcfg.latent_size;
  const int chunk = std::max(1, c.mcfg.max_views_per_call > 0 ? std::min(c.mcfg.max_views_per_call, 16) : 16);
  const size_t in_per = static_cast<size_t>(3) * 64 * latent_size * latent_size;
  const size_t out_per = static_cast<size_t>(8) * latent_size * latent_size;
  for (int i0 = 0; i0 < n; i0 += chunk) {
    const int T = std::min(chunk, n - i0);
    MD_CHECK(vae_encode(c, image + i0 * in_per, moments + i0 * out_per, T, latent_size, static_cast<cudaStream_t>(stream)));
  }
  return 0;
}

int md_has_clip(md_ctx* ctx) { return ctx && ctx->c.clip.loaded ? 1 : 0; }

int md_clip_embed(md_ctx* ctx, const float* image, float* embed, int n, int H, int W, void* stream) {
  MD_CHECK(ensure_ready(ctx, false));
  Ctx& c = ctx->c;
  SplitScope split_scope(&c.split_main);
  const int chunk = 16;
  for (int i0 = 0; i0 < n; i0 += chunk) {
    const int T = std::min(chunk, n - i0);
    MD_CHECK(clip_embed(c, image + static_cast<size_t>(i0) * 3 * H * W, embed + static_cast<size_t>(i0) * c.clip.out_dim, T, H, W,
                        static_cast<cudaStream_t>(stream)));
  }
  return 0;
}

int md_set_ddim(md_ctx* ctx, int ddim_steps, float ddim_eta) {
  if (!ctx) return set_error("null context");
  if (ddim_steps < 1 || ddim_steps > 1000) return set_error("md_set_ddim: ddim_steps=%d out of range 1..1000", ddim_steps);
  if (ddim_eta < 0.f) return set_error("md_set_ddim: negative eta");
  ctx->c.mcfg.ddim_steps = ddim_steps;
  ctx->c.mcfg.ddim_eta = ddim_eta;
  make_schedule(ctx->c);  // per-step scalars reach the captured graph through d_step: no re-capture needed
  return 0;
}

int md_ddim_steps(md_ctx* ctx) { return ctx ? static_cast<int>(ctx->c.timesteps.size()) : -1; }

int md_ddim_timestep(md_ctx* ctx, int index) {
  if (!ctx || index < 0 || index >= static_cast<int>(ctx->c.timesteps.size())) return -1;
  return ctx->c.timesteps[index];
}

int md_op_group_norm(const void* x, int x_is_bf16, int B, int rows, int C, int groups, float eps, const float* gamma,
                     const float* beta, const float* addvec, int act, void* out_bf16, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = nullptr;
  MD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), sizeof(float) * 4 * B * C, st));
  GroupNormArgs g;
  memset(&g, 0, sizeof(g));
  g.x0 = x; g.C0 = C; g.x0_bf16 = x_is_bf16; g.B = B; g.rows = rows; g.groups = groups; g.eps = eps;
  g.gamma = gamma; g.beta = beta; g.addvec = addvec; g.addvec_ld = C; g.stats = ws; g.scale_shift = ws + 2 * B * C;
  g.out = out_bf16; g.act = act;
  const int rc = launch_group_norm(g, st);
  cudaFreeAsync(ws, st);
  return rc;
}

int md_op_group_norm_stats(const void* x, int x_is_bf16, int B, int rows, int C, int groups, float eps, const float* gamma,
                           const float* beta, const float* addvec, int act, const float* stats, void* out_bf16,
                           void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!stats || !out_bf16) return set_error("md_op_group_norm_stats: stats and out are required");
  // scale/shift scratch for the (rare) unfused fallback: process-wide, grow-only, so that the call itself launches
  // nothing but the GroupNorm kernels (bench.py times it)
  static float* ws = nullptr;
  static size_t ws_floats = 0;
  const size_t need = static_cast<size_t>(2) * B * C;
  if (need > ws_floats) {
    if (ws) cudaFree(ws);
    ws = nullptr; ws_floats = 0;
    MD_CUDA(cudaMalloc(reinterpret_cast<void**>(&ws), need * sizeof(float)));
    ws_floats = need;
  }
  GroupNormArgs g;
  memset(&g, 0, sizeof(g));
  g.x0 = x; g.C0 = C; g.x0_bf16 = x_is_bf16; g.B = B; g.rows = rows; g.groups = groups; g.eps = eps;
  g.gamma = gamma; g.beta = beta; g.addvec = addvec; g.addvec_ld = C; g.stats0 = stats; g.scale_shift = ws;
  g.out = out_bf16; g.act = act;
  return launch_group_norm(g, st);
}

int md_op_softmax_rows(const float* x, void* out_bf16, long long rows, int n, void* stream) {
  return launch_softmax_rows(x, out_bf16, static_cast<size_t>(rows), n, static_cast<cudaStream_t>(stream));
}

int md_op_layer_norm(float* x, const float* gamma, const float* beta, void* out_bf16, long long rows, int C, float eps,
                     void* stream) {
  return launch_layer_norm(x, nullptr, 0, gamma, beta, out_bf16, static_cast<size_t>(rows), 1, C, eps,
                           static_cast<cudaStream_t>(stream));
}

int md_op_self_attention(const void* qkv, void* out, int B, int S, int heads, int dh, void* stream) {
  return launch_self_attention(qkv, out, B, S, heads, dh, static_cast<cudaStream_t>(stream));
}

int md_op_self_attention_impl(const void* qkv, void* out, int B, int S, int heads, int dh, int impl, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (impl == 1) return launch_self_attention_mma(qkv, out, B, S, heads, dh, st);
  if (impl == 2) return launch_attention_tc(qkv, out, B, S, heads, dh, st);
  return launch_self_attention(qkv, out, B, S, heads, dh, st);
}

int md_op_depth_attention(const void* qp, const void* c1, const float* ss, const float* beta, void* cbar, int T, int B,
                          int D, int HW, int ctx, void* stream) {
  return launch_depth_attention(qp, c1, ss, beta, cbar, T, B, D, HW, ctx, static_cast<cudaStream_t>(stream));
}

int md_op_cfg_ddim(md_ctx* ctx, const float* eps, float* x, float* eps_out, const float* noise, int T, int n_per_view,
                   int index, float cfg_scale, unsigned long long seed, int view0, void* stream) {
  if (!ctx) return set_error("null context");
  Ctx& c = ctx->c;
  if (index < 0 || index >= static_cast<int>(c.timesteps.size())) return set_error("cfg_ddim: bad index %d", index);
  return launch_cfg_ddim(eps, x, eps_out, noise, T, n_per_view, cfg_scale != 1.0f, cfg_scale, c.alphas[index],
                         c.alphas_prev[index], c.sigmas[index], c.sqrt_1m_alphas[index], index != 0, seed,
                         static_cast<uint32_t>(index), view0, 1, nullptr, static_cast<cudaStream_t>(stream));
}

// ---- NVLink peer exchange (Ctx::PeerExchange): md_peer_buffer on every rank, exchange the 64-byte handles out of band
// (torch.distributed all_gather_object in the Python binding), md_peer_attach with all of them, then a barrier.
static const size_t kPeerSlotFloats = static_cast<size_t>(16384) * 16;  // vertices per region (FLAME 5 023, SMPL-X 10 475)

int md_peer_buffer(md_ctx* ctx, int world, void* ipc_handle64) {
  if (!ctx || !ipc_handle64) return set_error("md_peer_buffer: null argument");
  if (world < 2 || world > 16) return set_error("md_peer_buffer: world=%d outside [2, 16]", world);
  Ctx::PeerExchange& px = ctx->c.px;
  if (px.base) return set_error("md_peer_buffer: already allocated");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the C ABI");
  px.slot_floats = kPeerSlotFloats;
  const size_t bytes = static_cast<size_t>(2) * world * px.slot_floats * sizeof(float) + 256;
  MD_CUDA(cudaMalloc(reinterpret_cast<void**>(&px.base), bytes));
  MD_CUDA(cudaMemset(px.base, 0, bytes));
  MD_CUDA(cudaDeviceSynchronize());  // the flags are zero before any peer can learn the handle
  cudaIpcMemHandle_t h;
  MD_CUDA(cudaIpcGetMemHandle(&h, px.base));
  memcpy(ipc_handle64, &h, sizeof(h));
  return 0;
}

int md_peer_attach(md_ctx* ctx, int rank, int world, const void* ipc_handles) {
  if (!ctx || !ipc_handles) return set_error("md_peer_attach: null argument");
  Ctx& c = ctx->c;
  Ctx::PeerExchange& px = c.px;
  if (!px.base) return set_error("md_peer_attach: call md_peer_buffer first");
  if (px.on) return set_error("md_peer_attach: already attached");
  if (world < 2 || world > 16 || rank < 0 || rank >= world) return set_error("md_peer_attach: bad rank %d / world %d", rank, world);
  if (c.world != world || c.rank != rank) return set_error("md_peer_attach: rank/world differ from md_comm_init (%d/%d)", c.rank, c.world);
  const size_t flags_off = static_cast<size_t>(2) * world * px.slot_floats;
  float* data[16];
  unsigned* flags[16];
  for (int p = 0; p < world; ++p) {
    if (p == rank) {
      px.peer_base[p] = px.base;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, static_cast<const char*>(ipc_handles) + static_cast<size_t>(p) * sizeof(h), sizeof(h));
      void* ptr = nullptr;
      const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return set_error("md_peer_attach: cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
      px.peer_base[p] = ptr;
    }
    data[p] = static_cast<float*>(px.peer_base[p]);
    flags[p] = reinterpret_cast<unsigned*>(data[p] + flags_off);
  }
  MD_CUDA(cudaMalloc(reinterpret_cast<void**>(&px.d_peer_data), sizeof(float*) * 16));
  MD_CUDA(cudaMalloc(reinterpret_cast<void**>(&px.d_peer_flags), sizeof(unsigned*) * 16));
  MD_CUDA(cudaMalloc(reinterpret_cast<void**>(&px.d_seq), 4 * sizeof(unsigned)));
  MD_CUDA(cudaMemcpy(px.d_peer_data, data, sizeof(float*) * world, cudaMemcpyHostToDevice));
  MD_CUDA(cudaMemcpy(px.d_peer_flags, flags, sizeof(unsigned*) * world, cudaMemcpyHostToDevice));
  MD_CUDA(cudaMemset(px.d_seq, 0, 4 * sizeof(unsigned)));
  MD_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&px.h_err), sizeof(int), cudaHostAllocMapped));
  *px.h_err = 0;
  MD_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&px.d_err), px.h_err, 0));
  px.seq = 0;
  px.on = true;
  // the step's graph layout changes (one graph instead of two around the NCCL call): drop anything captured before
  if (c.graph) { cudaGraphExecDestroy(c.graph); c.graph = nullptr; }
  if (c.graph_b) { cudaGraphExecDestroy(c.graph_b); c.graph_b = nullptr; }
  c.graph_warm = 0;
  memset(&c.gkey, 0, sizeof(c.gkey));
  return 0;
}

int md_peer_attached(md_ctx* ctx) { return ctx && ctx->c.px.on ? 1 : 0; }

int md_peer_detach(md_ctx* ctx) {
  if (!ctx) return set_error("md_peer_detach: null context");
  Ctx& c = ctx->c;
  if (c.px.on) {  // back to the NCCL all-reduce between two graphs: drop the one-graph capture
    MD_CUDA(cudaDeviceSynchronize());
    c.px.on = false;
    if (c.graph) { cudaGraphExecDestroy(c.graph); c.graph = nullptr; }
    if (c.graph_b) { cudaGraphExecDestroy(c.graph_b); c.graph_b = nullptr; }
    c.graph_warm = 0;
    memset(&c.gkey, 0, sizeof(c.gkey));
  }
  return 0;
}

int md_comm_unique_id(void* id128) {
  MD_CHECK(nccl_load());
  const int r = g_nccl.GetUniqueId(id128);
  return r == 0 ? 0 : set_error("ncclGetUniqueId failed (%d)", r);
}

int md_comm_init(md_ctx* ctx, int rank, int world, const void* id128) {
  if (!ctx) return set_error("null context");
  ctx->c.rank = rank;
  ctx->c.world = world;
  if (world <= 1) return 0;
  MD_CHECK(nccl_load());
  Id128 id;
  memcpy(&id, id128, sizeof(id));
  void* comm = nullptr;
  const int r = g_nccl.CommInitRank(&comm, world, id, rank);
  if (r != 0) return set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  ctx->c.nccl_comm = comm;
  return 0;
}

}  // extern "C"
