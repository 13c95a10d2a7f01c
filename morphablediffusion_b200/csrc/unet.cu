// Host orchestration of DepthWiseAttention.forward (ldm/models/diffusion/attention.py:117-138) on the kernels of
// this library.  Layout: every activation is channels-last; the UNet residual stream `h` is fp32, GEMM operands are
// bf16 produced by the fused GroupNorm/LayerNorm kernels.  GroupNorm statistics are accumulated by the epilogue of
// the GEMM that produces the tensor (per-(sample, channel) sum / sum of squares), so a GroupNorm costs one tiny
// finalize kernel plus one normalise+activate pass.
#include "engine.h"

#include <math.h>
#include <stdlib.h>

namespace md {

namespace {

struct Fwd {
  Ctx& c;
  cudaStream_t st;
  int B;                 // samples in this UNet call
  int n_ctx;             // leading samples that own a frustum volume
  const float* emb_all;  // [B][emb_total] per-ResBlock time-embedding projections
  const float* v2_all;   // [B][v2_total] attn2 output vectors of every transformer block
  float res_eps = 1e-5f; // GroupNorm eps of the ResBlocks (UNet GroupNorm32: 1e-5; first-stage Normalize: 1e-6)
  // statistics pool: one [B][C][2] slab per tensor that feeds a GroupNorm, zeroed once per forward
  float* spool = nullptr;
  size_t spool_cap = 0, spool_off = 0;
  std::unordered_map<const void*, float*> stats_of;

  Arena& A() { return c.arena; }

  float* new_stats(const void* tensor, int nb, int C) {
    const size_t n = static_cast<size_t>(nb) * C * 2;
    if (spool_off + n > spool_cap) return nullptr;
    float* p = spool + spool_off;
    spool_off += n;
    stats_of[tensor] = p;
    return p;
  }
  const float* find_stats(const void* tensor) const {
    auto it = stats_of.find(tensor);
    return it == stats_of.end() ? nullptr : it->second;
  }

  // GroupNorm applied by the producing GEMM itself (conv_gemm's tail phase): norm parameters + destination
  struct GnTail { const NormW* n; int groups; float eps; int act; bf16* out; };
  static bool gn_tail_enabled() {
    static const bool on = getenv("MD_GN_TAIL") == nullptr || atoi(getenv("MD_GN_TAIL")) != 0;
    return on;
  }
  // one grid-barrier counter per launch, taken from the statistics pool (zeroed once per forward)
  int* new_barrier() {
    if (spool_off + 4 > spool_cap) return nullptr;
    int* b = reinterpret_cast<int*>(spool + spool_off);
    spool_off += 4;   // keeps the statistics slabs behind it 16-byte aligned
    return b;
  }
  int attach_tail(md_conv_gemm_args& a, const GnTail* t) {
    if (!t) return 0;
    if (!a.col_stats || !a.out_bf16) return set_error("GroupNorm tail without statistics / bf16 output");
    a.gn_out = t->out; a.gn_gamma = t->n->g; a.gn_beta = t->n->b; a.gn_groups = t->groups; a.gn_eps = t->eps;
    a.gn_act = t->act;
    a.gn_barrier = new_barrier();
    if (!a.gn_barrier) return set_error("statistics pool exhausted");
    return 0;
  }

  static void taps2d(md_conv_gemm_args& a) {
    a.ntaps = 9;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        a.tap[ky * 3 + kx][0] = kx - 1; a.tap[ky * 3 + kx][1] = ky - 1; a.tap[ky * 3 + kx][2] = 0;
      }
  }

  // conv3x3 (pad 1) or 1x1 on a bf16 NHWC tensor of `nb` samples; `stats` = accumulate GroupNorm statistics of the
  // result (only possible when a warp's 32 output rows stay inside one sample)
  int conv(const bf16* a_in, int nb, int H, int W, const GemmW& w, const float* rowvec, int rowvec_ld,
           const float* res_f32, float* out_f32, bf16* out_bf16, bool stats, int act = ACT_NONE,
           const GnTail* tail = nullptr) {
    md_conv_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = a_in; a.B = nb; a.D = 1; a.H = H; a.W = W; a.Cin = w.K; a.Wt = w.w; a.N = w.N;
    if (w.taps == 9) taps2d(a);
    else { a.ntaps = 1; }
    a.bias = w.bias; a.rowvec = rowvec; a.rowvec_ld = rowvec_ld; a.res_f32 = res_f32;
    a.out_f32 = out_f32; a.out_bf16 = out_bf16; a.act = act;
    const void* key = out_f32 ? static_cast<const void*>(out_f32) : static_cast<const void*>(out_bf16);
    if (stats && H * W >= 32) {
      a.col_stats = new_stats(key, nb, w.N);
      if (!a.col_stats) return set_error("statistics pool exhausted");
    } else {
      stats_of.erase(key);  // the arena recycles addresses: statistics of an earlier tensor must not outlive it
    }
    MD_CHECK(attach_tail(a, tail));
    return launch_conv_gemm(a, st);
  }
  // the GroupNorm of a GEMM's own output can ride on the producing launch when that launch carries statistics
  static bool can_tail(int rows_per_sample) { return gn_tail_enabled() && rows_per_sample >= 32 && rows_per_sample % 32 == 0; }
  // plain GEMM over rows = nb * rows_per_sample tokens
  int gemm(const bf16* a_in, int nb, size_t rows_per_sample, const GemmW& w, const float* res_f32, float* out_f32,
           bf16* out_bf16, bool stats, int act = ACT_NONE, const float* rowvec = nullptr, int rowvec_ld = 0,
           const bf16* res_bf16 = nullptr, const GnTail* tail = nullptr) {
    md_conv_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = a_in; a.B = nb; a.D = 1; a.H = 1; a.W = static_cast<int>(rows_per_sample); a.Cin = w.K; a.Wt = w.w; a.N = w.N;
    a.ntaps = 1;
    a.bias = w.bias; a.res_f32 = res_f32; a.out_f32 = out_f32; a.out_bf16 = out_bf16; a.act = act;
    a.rowvec = rowvec; a.rowvec_ld = rowvec_ld; a.res_bf16 = res_bf16;
    const void* key = out_f32 ? static_cast<const void*>(out_f32) : static_cast<const void*>(out_bf16);
    if (stats && rows_per_sample >= 32 && rows_per_sample % 32 == 0) {
      a.col_stats = new_stats(key, nb, w.N);
      if (!a.col_stats) return set_error("statistics pool exhausted");
    } else {
      stats_of.erase(key);
    }
    MD_CHECK(attach_tail(a, tail));
    return launch_conv_gemm(a, st);
  }

  // GroupNorm (+activation) of (x0 | x1) -> bf16.  out == nullptr: only the per-(sample, channel) scale/shift.
  int gn(const void* x0, int C0, bool bf0, const void* x1, int C1, int nb, int rows, int groups, float eps,
         const NormW& n, int act, bf16* out, bf16* raw, float** ss_out = nullptr) {
    const int C = C0 + C1;
    GroupNormArgs g;
    memset(&g, 0, sizeof(g));
    g.x0 = x0; g.C0 = C0; g.x0_bf16 = bf0; g.x1 = x1; g.C1 = C1; g.x1_bf16 = bf0;
    g.B = nb; g.rows = rows; g.groups = groups; g.eps = eps; g.gamma = n.g; g.beta = n.b;
    const float* s0 = find_stats(x0);
    const float* s1 = C1 ? find_stats(x1) : nullptr;
    if (s0 && (!C1 || s1)) {
      g.stats0 = s0; g.stats1 = s1;
    } else {
      if (static_cast<size_t>(nb) * C * 2 > c.gn_stats_floats) return set_error("group norm statistics scratch too small");
      g.stats = c.gn_stats; g.stats_prezeroed = 1;
    }
    g.scale_shift = A().get<float>(static_cast<size_t>(nb) * C * 2);
    if (!g.scale_shift) return set_error("workspace exhausted (group norm)");
    if (ss_out) *ss_out = g.scale_shift;
    g.out = out; g.raw_out = raw; g.act = act;
    return launch_group_norm(g, st);
  }

  // ResBlock._forward (openaimodel.py:256-276).  Input is (x0 | x1) channel-concatenated (x1 may be null).
  // out_b: optional bf16 copy of the result for a consumer that needs a GEMM operand (downsample, depth transformer)
  int res_block(const ResW& r, const float* x0, int C0, const float* x1, int C1, int H, int W, float* out,
                bf16* out_b = nullptr) {
    const size_t rows = static_cast<size_t>(B) * H * W;
    const size_t m = A().mark();
    bf16* a1 = A().get<bf16>(rows * r.cin);
    bf16* raw = r.has_skip ? A().get<bf16>(rows * r.cin) : nullptr;
    bf16* h1 = A().get<bf16>(rows * r.cout);   // consumed by GroupNorm only: bf16 (statistics come from the fp32 accumulators)
    bf16* a2 = A().get<bf16>(rows * r.cout);
    float* skip = r.has_skip ? A().get<float>(rows * r.cout) : nullptr;
    if (A().failed) return set_error("workspace exhausted (res block)");
    MD_CHECK(gn(x0, C0, false, x1, C1, B, H * W, 32, res_eps, r.n1, ACT_SILU, a1, raw));
    // per-sample time-embedding vector (UNet ResBlock); the first-stage ResnetBlock has none (temb is None)
    const float* rv = emb_all ? emb_all + r.emb_off : nullptr;
    if (can_tail(H * W)) {   // GroupNorm + SiLU between the two convolutions rides on conv1's launch
      const GnTail t{&r.n2, 32, res_eps, ACT_SILU, a2};
      MD_CHECK(conv(a1, B, H, W, r.c1, rv, rv ? c.unet.emb_total : 0, nullptr, nullptr, h1, true, ACT_NONE, &t));
    } else {
      MD_CHECK(conv(a1, B, H, W, r.c1, rv, rv ? c.unet.emb_total : 0, nullptr, nullptr, h1, true));
      MD_CHECK(gn(h1, r.cout, true, nullptr, 0, B, H * W, 32, res_eps, r.n2, ACT_SILU, a2, nullptr));
    }
    const float* resid = x0;
    if (r.has_skip) {
      MD_CHECK(conv(raw, B, H, W, r.skip, nullptr, 0, nullptr, skip, nullptr, false));
      resid = skip;
    } else if (C1 != 0) {
      return set_error("res block: identity skip with concatenated input");
    }
    MD_CHECK(conv(a2, B, H, W, r.c2, nullptr, 0, resid, out, out_b, true));
    A().release(m);
    return 0;
  }

  // SpatialTransformer.forward (ldm/modules/attention.py:325-336), depth 1, single-token context.
  int spatial_transformer(const STW& s, const float* x_in, int H, int W, float* out, bf16* out_b = nullptr) {
    const int C = s.C;
    const size_t S = static_cast<size_t>(H) * W;
    const size_t rows = static_cast<size_t>(B) * S;
    const size_t m = A().mark();
    // The block's internal stream x (proj_in output, += attn1, += feed-forward) is bf16: every consumer normalises or
    // rounds it to bf16 anyway, and at level 0 each fp32 pass over it moved 42 MB (the 320-wide linears of the block
    // are bound by exactly that traffic).  MD_ST_FP32=1 keeps it in fp32 (A/B switch).
    static const bool x_fp32 = getenv("MD_ST_FP32") != nullptr;
    bf16* a = A().get<bf16>(rows * C);
    float* x = x_fp32 ? A().get<float>(rows * C) : nullptr;
    bf16* xh = x_fp32 ? nullptr : A().get<bf16>(rows * C);
    bf16* ln = A().get<bf16>(rows * C);
    bf16* qkv = A().get<bf16>(rows * 3 * C);
    bf16* att = A().get<bf16>(rows * C);
    bf16* ff = A().get<bf16>(rows * 4 * C);
    bf16* xb = A().get<bf16>(rows * C);
    if (A().failed) return set_error("workspace exhausted (spatial transformer)");
    const int Si = static_cast<int>(S);
    const float* v2 = v2_all + s.v2_off;
    const int v2_ld = c.unet.v2_total;
    MD_CHECK(gn(x_in, C, false, nullptr, 0, B, Si, 32, 1e-6f, s.norm, ACT_NONE, a, nullptr));
    MD_CHECK(gemm(a, B, S, s.proj_in, nullptr, x, xh, false));
    // attn1 (self-attention)
    if (x_fp32) MD_CHECK(launch_layer_norm(x, nullptr, 0, s.ln1.g, s.ln1.b, ln, rows, Si, C, 1e-5f, st));
    else MD_CHECK(launch_layer_norm_bf16(xh, nullptr, 0, s.ln1.g, s.ln1.b, ln, rows, Si, C, 1e-5f, st));
    MD_CHECK(gemm(ln, B, S, s.qkv, nullptr, nullptr, qkv, false));
    MD_CHECK(launch_self_attention(qkv, att, B, Si, s.heads, C / s.heads, st));
    MD_CHECK(gemm(att, B, S, s.o1, x, x, xh, false, ACT_NONE, nullptr, 0, xh));
    // attn2: one context token => softmax == 1 => attn2(x) = to_out(to_v(ctx)) for every query (precomputed per
    // forward in v2_all): x2 = x + v2[b].  norm3 sees x2 without writing it back; the feed-forward's output GEMM adds
    // the vector again as its per-sample epilogue vector: xb = ff2(geglu(ff1(norm3(x2)))) + v2[b] + x
    if (x_fp32) MD_CHECK(launch_layer_norm(x, v2, v2_ld, s.ln3.g, s.ln3.b, ln, rows, Si, C, 1e-5f, st, /*write_back=*/0));
    else MD_CHECK(launch_layer_norm_bf16(xh, v2, v2_ld, s.ln3.g, s.ln3.b, ln, rows, Si, C, 1e-5f, st));
    MD_CHECK(gemm(ln, B, S, s.ff1, nullptr, nullptr, ff, false, ACT_GEGLU));
    MD_CHECK(gemm(ff, B, S, s.ff2, x, nullptr, xb, false, ACT_NONE, v2, v2_ld, xh));
    MD_CHECK(gemm(xb, B, S, s.proj_out, x_in, out, out_b, true));
    A().release(m);
    return 0;
  }

  // DepthTransformer._forward + DepthAttention.forward (ldm/models/diffusion/attention.py:78-84,26-47), with the
  // attention re-associated (attention.cu) and the zero-volume samples (n_ctx..B-1) short-circuited.
  int depth_transformer(const DepthW& d, const float* x_in, const bf16* x_in_b, int H, int W, const bf16* ctx, int D,
                        float* out) {
    const size_t S = static_cast<size_t>(H) * W;
    const size_t rows = static_cast<size_t>(B) * S;
    const size_t crows = static_cast<size_t>(n_ctx) * S * D;
    const size_t m = A().mark();
    bf16* xb = x_in_b ? nullptr : A().get<bf16>(rows * d.dim);
    bf16* y = A().get<bf16>(rows * d.inner);    // y, y2, y3 feed GroupNorms only: bf16
    bf16* xq = A().get<bf16>(rows * d.inner);
    bf16* qp = A().get<bf16>(static_cast<size_t>(n_ctx) * S * 4 * d.ctx);
    bf16* c1 = A().get<bf16>(crows * d.ctx);
    bf16* cbar = A().get<bf16>(rows * 4 * d.ctx);
    bf16* y2 = A().get<bf16>(rows * d.inner);
    bf16* a1 = A().get<bf16>(rows * d.inner);
    bf16* y3 = A().get<bf16>(rows * d.inner);
    bf16* a2 = A().get<bf16>(rows * d.inner);
    if (A().failed) return set_error("workspace exhausted (depth transformer)");
    if (!x_in_b) MD_CHECK(launch_cast_bf16(x_in, xb, rows * d.dim, st));
    const bool tail = can_tail(static_cast<int>(S));
    if (tail) {
      const GnTail t{&d.gn_in, 8, 1e-5f, ACT_SILU, xq};
      MD_CHECK(gemm(x_in_b ? x_in_b : xb, B, S, d.proj_in, nullptr, nullptr, y, true, ACT_NONE, nullptr, 0, nullptr, &t));
    } else {
      MD_CHECK(gemm(x_in_b ? x_in_b : xb, B, S, d.proj_in, nullptr, nullptr, y, true));
      MD_CHECK(gn(y, d.inner, true, nullptr, 0, B, static_cast<int>(S), 8, 1e-5f, d.gn_in, ACT_SILU, xq, nullptr));
    }
    // queries mapped into context space (only the samples that own a volume)
    MD_CHECK(gemm(xq, n_ctx, S, d.wqk, nullptr, nullptr, qp, false));
    // context branch: proj_context conv -> GroupNorm statistics; the normalisation + ReLU is applied on read
    float* ss_ctx = nullptr;
    MD_CHECK(gemm(ctx, n_ctx, S * D, d.proj_ctx, nullptr, nullptr, c1, true));
    MD_CHECK(gn(c1, d.ctx, true, nullptr, 0, n_ctx, static_cast<int>(S * D), 8, 1e-5f, d.gn_ctx, ACT_RELU, nullptr, nullptr,
                &ss_ctx));
    MD_CHECK(launch_depth_attention(qp, c1, ss_ctx, d.gn_ctx.b, cbar, n_ctx, B, D, static_cast<int>(S), d.ctx, st));
    if (tail) {
      const GnTail t1{&d.gn_o1, 8, 1e-5f, ACT_RELU, a1}, t2{&d.gn_o2, 8, 1e-5f, ACT_RELU, a2};
      MD_CHECK(gemm(cbar, B, S, d.wov, nullptr, nullptr, y2, true, ACT_NONE, nullptr, 0, nullptr, &t1));
      MD_CHECK(conv(a1, B, H, W, d.conv1, nullptr, 0, nullptr, nullptr, y3, true, ACT_NONE, &t2));
    } else {
      MD_CHECK(gemm(cbar, B, S, d.wov, nullptr, nullptr, y2, true));
      MD_CHECK(gn(y2, d.inner, true, nullptr, 0, B, static_cast<int>(S), 8, 1e-5f, d.gn_o1, ACT_RELU, a1, nullptr));
      MD_CHECK(conv(a1, B, H, W, d.conv1, nullptr, 0, nullptr, nullptr, y3, true));
      MD_CHECK(gn(y3, d.inner, true, nullptr, 0, B, static_cast<int>(S), 8, 1e-5f, d.gn_o2, ACT_RELU, a2, nullptr));
    }
    MD_CHECK(conv(a2, B, H, W, d.conv2, nullptr, 0, x_in, out, nullptr, true));
    A().release(m);
    return 0;
  }

  // Downsample: Conv2d 3x3 stride 2 pad 1 (openaimodel.py:159-161): implicit GEMM whose TMA boxes step by 2 pixels
  // tap0 = -1: Conv2d k3 s2 p1 (UNet Downsample); tap0 = 0: the first-stage Downsample (model.py:70-78: zero pad on the
  // right / bottom only, conv k3 s2 p0), i.e. taps 0..2 with the out-of-range column / row zero-filled by TMA
  int downsample(const GemmW& w, const float* x, const bf16* x_b, int H, int W, int C, float* out, int tap0 = -1) {
    const size_t rows = static_cast<size_t>(B) * H * W;
    const size_t m = A().mark();
    bf16* xb = x_b ? nullptr : A().get<bf16>(rows * C);
    if (A().failed) return set_error("workspace exhausted (downsample)");
    if (!x_b) MD_CHECK(launch_cast_bf16(x, xb, rows * C, st));
    md_conv_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = x_b ? x_b : xb; a.B = B; a.D = 1; a.H = H; a.W = W; a.Cin = C; a.Wt = w.w; a.N = w.N;
    taps2d(a);
    for (int t = 0; t < 9; ++t) { a.tap[t][0] += tap0 + 1; a.tap[t][1] += tap0 + 1; }
    a.in_stride[0] = 2; a.in_stride[1] = 2; a.in_stride[2] = 1;
    a.bias = w.bias; a.out_f32 = out;
    if ((H / 2) * (W / 2) >= 32) {
      a.col_stats = new_stats(out, B, w.N);
      if (!a.col_stats) return set_error("statistics pool exhausted");
    } else {
      stats_of.erase(out);
    }
    MD_CHECK(launch_conv_gemm(a, st));
    A().release(m);
    return 0;
  }
  // Upsample: nearest x2 + Conv2d 3x3 (openaimodel.py:110-120)
  int upsample(const GemmW& w, const float* x, int H, int W, int C, float* out, bf16* out_b = nullptr) {
    const size_t orows = static_cast<size_t>(B) * 4 * H * W;
    const size_t m = A().mark();
    bf16* up = A().get<bf16>(orows * C);
    if (A().failed) return set_error("workspace exhausted (upsample)");
    MD_CHECK(launch_upsample2x(x, up, B, H, W, C, st));
    MD_CHECK(conv(up, B, 2 * H, 2 * W, w, nullptr, 0, nullptr, out, out_b, true));
    A().release(m);
    return 0;
  }
};

}  // namespace

int unet_forward(Ctx& c, const float* x_in, const float* timesteps, const float* context, const bf16* const levels[4],
                 int B, int n_ctx, int S, int D, float* eps_out, cudaStream_t st) {
  if (!c.weights_loaded) return set_error("unet_forward: weights not loaded");
  if (n_ctx < 1 || n_ctx > B) return set_error("unet_forward: bad n_ctx=%d (B=%d)", n_ctx, B);
  const UNetW& u = c.unet;
  Arena& A = c.arena;
  const size_t m0 = A.mark();
  const int mch = u.model_channels;

  // time embedding MLP, then every ResBlock's emb projection and every transformer block's attn2 vector as two
  // tensor-core GEMMs over the batch
  float* temb = A.get<float>(static_cast<size_t>(B) * mch);
  float* e1 = A.get<float>(static_cast<size_t>(B) * u.emb_dim);
  float* emb = A.get<float>(static_cast<size_t>(B) * u.emb_dim);
  bf16* emb_b = A.get<bf16>(static_cast<size_t>(B) * u.emb_dim);
  bf16* ctx_b = A.get<bf16>(static_cast<size_t>(B) * u.ctx_dim);
  float* emb_all = A.get<float>(static_cast<size_t>(B) * u.emb_total);
  float* v2_all = A.get<float>(static_cast<size_t>(B) * u.v2_total);
  // statistics pool: ~110 GroupNorm inputs of at most [B][1280][2]
  const size_t spool_floats = static_cast<size_t>(B) * 2 * 96 * 1024;
  float* spool = A.get<float>(spool_floats);
  if (A.failed) return set_error("workspace exhausted (unet embeddings)");
  MD_CUDA(cudaMemsetAsync(spool, 0, spool_floats * sizeof(float), st));
  MD_CHECK(launch_timestep_embedding(timesteps, temb, B, mch, st));
  MD_CHECK(launch_small_linear(temb, mch, u.te0_w, u.te0_b, e1, u.emb_dim, B, mch, u.emb_dim, ACT_NONE, ACT_SILU, 0, st));
  MD_CHECK(launch_small_linear(e1, u.emb_dim, u.te2_w, u.te2_b, emb, u.emb_dim, B, u.emb_dim, u.emb_dim, ACT_NONE, ACT_SILU, 0, st));
  MD_CHECK(launch_cast_bf16(emb, emb_b, static_cast<size_t>(B) * u.emb_dim, st));
  MD_CHECK(launch_cast_bf16(context, ctx_b, static_cast<size_t>(B) * u.ctx_dim, st));
  {
    md_conv_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = emb_b; a.B = 1; a.D = 1; a.H = 1; a.W = B; a.Cin = u.emb_dim; a.Wt = u.emb_g.w; a.N = u.emb_total; a.ntaps = 1;
    a.bias = u.emb_g.bias; a.out_f32 = emb_all;
    MD_CHECK(launch_conv_gemm(a, st));
    a.A = ctx_b; a.Cin = u.ctx_dim; a.Wt = u.v2_g.w; a.N = u.v2_total; a.bias = u.v2_g.bias; a.out_f32 = v2_all;
    MD_CHECK(launch_conv_gemm(a, st));
  }

  Fwd f{c, st, B, n_ctx, emb_all, v2_all};
  f.spool = spool; f.spool_cap = spool_floats;

  struct Skip { float* p; int C, H; };
  std::vector<Skip> hs;
  int H = S, ch = mch;
  // level key -> frustum level index by spatial width (attention.py:128,135)
  auto level_of = [&](int width, const bf16*& ptr, int& depth) {
    int li = 0, w = S, d = D;
    while (w != width && li < 3) { w /= 2; d /= 2; ++li; }
    ptr = levels[li];
    depth = d;
    return w == width ? 0 : -1;
  };

  float* h = A.get<float>(static_cast<size_t>(B) * H * H * mch);
  if (A.failed) return set_error("workspace exhausted (unet)");
  {  // conv_in on the tensor-core path: the 8 input channels are padded to one 64-channel K block (zero weights
     // there), which also gives the first GroupNorm its statistics from the GEMM epilogue
    const size_t m1 = A.mark();
    bf16* xin_b = A.get<bf16>(static_cast<size_t>(B) * H * H * 64);
    if (A.failed) return set_error("workspace exhausted (unet)");
    MD_CHECK(launch_pad_cast_bf16(x_in, xin_b, static_cast<size_t>(B) * H * H, u.in_channels, 64, st));
    MD_CHECK(f.conv(xin_b, B, H, H, u.conv_in_g, nullptr, 0, nullptr, h, nullptr, true));
    A.release(m1);
  }
  hs.push_back({h, mch, H});

  // hb: bf16 copy of h, written by the producing GEMM's epilogue when the next consumer takes h as a GEMM operand
  // (stride-2 downsample convolution, depth transformer proj_in); saves a separate cast pass
  bf16* hb = nullptr;
  for (size_t bi = 1; bi < u.input_blocks.size(); ++bi) {
    const std::vector<UNetLayer>& layers = u.input_blocks[bi];
    const bool next_down = bi + 1 < u.input_blocks.size() && !u.input_blocks[bi + 1].empty() &&
                           u.input_blocks[bi + 1][0].kind == 3;
    for (size_t li = 0; li < layers.size(); ++li) {
      const UNetLayer& L = layers[li];
      const bool want_b = next_down && li + 1 == layers.size();
      if (L.kind == 1) {
        float* o = A.get<float>(static_cast<size_t>(B) * H * H * L.res.cout);
        bf16* ob = want_b ? A.get<bf16>(static_cast<size_t>(B) * H * H * L.res.cout) : nullptr;
        if (A.failed) return set_error("workspace exhausted (unet)");
        MD_CHECK(f.res_block(L.res, h, ch, nullptr, 0, H, H, o, ob));
        h = o; hb = ob; ch = L.res.cout;
      } else if (L.kind == 2) {
        float* o = A.get<float>(static_cast<size_t>(B) * H * H * ch);
        bf16* ob = want_b ? A.get<bf16>(static_cast<size_t>(B) * H * H * ch) : nullptr;
        if (A.failed) return set_error("workspace exhausted (unet)");
        MD_CHECK(f.spatial_transformer(L.st, h, H, H, o, ob));
        h = o; hb = ob;
      } else if (L.kind == 3) {
        float* o = A.get<float>(static_cast<size_t>(B) * (H / 2) * (H / 2) * ch);
        if (A.failed) return set_error("workspace exhausted (unet)");
        MD_CHECK(f.downsample(L.conv, h, hb, H, H, ch, o));
        h = o; hb = nullptr; H /= 2;
      }
    }
    hs.push_back({h, ch, H});
  }
  {  // middle block + middle_conditions
    float* o0 = A.get<float>(static_cast<size_t>(B) * H * H * ch);
    float* o1 = A.get<float>(static_cast<size_t>(B) * H * H * ch);
    float* o2 = A.get<float>(static_cast<size_t>(B) * H * H * ch);
    bf16* o2b = A.get<bf16>(static_cast<size_t>(B) * H * H * ch);
    float* o3 = A.get<float>(static_cast<size_t>(B) * H * H * ch);
    if (A.failed) return set_error("workspace exhausted (unet)");
    MD_CHECK(f.res_block(u.mid0, h, ch, nullptr, 0, H, H, o0));
    MD_CHECK(f.spatial_transformer(u.mid1, o0, H, H, o1));
    MD_CHECK(f.res_block(u.mid2, o1, ch, nullptr, 0, H, H, o2, o2b));
    const bf16* lv; int dd;
    if (level_of(H, lv, dd) != 0) return set_error("unet: no frustum level of width %d", H);
    if (c.levels_pending) {  // the frustum pyramids are produced on the second stream: join here
      MD_CUDA(cudaStreamWaitEvent(st, c.ev_levels, 0));
      c.levels_pending = false;
    }
    MD_CHECK(f.depth_transformer(u.mid_cond, o2, o2b, H, H, lv, dd, o3));
    h = o3;
  }
  for (size_t bi = 0; bi < u.output_blocks.size(); ++bi) {
    const Skip sk = hs.back();
    hs.pop_back();
    const std::vector<UNetLayer>& layers = u.output_blocks[bi];
    const bool cond = bi >= 3 && bi - 3 < u.out_cond.size();  // output_b2c = {3:0, ..., 11:8}
    hb = nullptr;
    for (size_t li = 0; li < layers.size(); ++li) {
      const UNetLayer& L = layers[li];
      const bool want_b = cond && li + 1 == layers.size();
      if (L.kind == 1) {
        float* o = A.get<float>(static_cast<size_t>(B) * H * H * L.res.cout);
        bf16* ob = want_b ? A.get<bf16>(static_cast<size_t>(B) * H * H * L.res.cout) : nullptr;
        if (A.failed) return set_error("workspace exhausted (unet)");
        if (li != 0) return set_error("unet: unexpected ResBlock position");
        MD_CHECK(f.res_block(L.res, h, ch, sk.p, sk.C, H, H, o, ob));
        h = o; hb = ob; ch = L.res.cout;
      } else if (L.kind == 2) {
        float* o = A.get<float>(static_cast<size_t>(B) * H * H * ch);
        bf16* ob = want_b ? A.get<bf16>(static_cast<size_t>(B) * H * H * ch) : nullptr;
        if (A.failed) return set_error("workspace exhausted (unet)");
        MD_CHECK(f.spatial_transformer(L.st, h, H, H, o, ob));
        h = o; hb = ob;
      } else if (L.kind == 4) {
        float* o = A.get<float>(static_cast<size_t>(B) * 4 * H * H * ch);
        bf16* ob = want_b ? A.get<bf16>(static_cast<size_t>(B) * 4 * H * H * ch) : nullptr;
        if (A.failed) return set_error("workspace exhausted (unet)");
        MD_CHECK(f.upsample(L.conv, h, H, H, ch, o, ob));
        h = o; hb = ob; H *= 2;
      }
    }
    if (cond) {
      const bf16* lv; int dd;
      if (level_of(H, lv, dd) != 0) return set_error("unet: no frustum level of width %d", H);
      float* o = A.get<float>(static_cast<size_t>(B) * H * H * ch);
      if (A.failed) return set_error("workspace exhausted (unet)");
      MD_CHECK(f.depth_transformer(u.out_cond[bi - 3], h, hb, H, H, lv, dd, o));
      h = o; hb = nullptr;
    }
  }
  // out: GN32 + SiLU + conv3x3 320 -> 4
  bf16* ao = A.get<bf16>(static_cast<size_t>(B) * H * H * ch);
  if (A.failed) return set_error("workspace exhausted (unet)");
  MD_CHECK(f.gn(h, ch, false, nullptr, 0, B, H * H, 32, 1e-5f, u.out_norm, ACT_SILU, ao, nullptr));
  float* eps8 = A.get<float>(static_cast<size_t>(B) * H * H * 8);
  if (A.failed) return set_error("workspace exhausted (unet)");
  MD_CHECK(f.conv(ao, B, H, H, u.out_g, nullptr, 0, nullptr, eps8, nullptr, false));
  MD_CHECK(launch_rows_to_nchw(eps8, 8, eps_out, B, u.out_channels, H * H, st));
  A.release(m0);
  return 0;
}

// AttnBlock of the first-stage models (model.py:177-203): h fp32 [T][Sx][C] with GEMM-epilogue statistics -> o.
static int vae_attn_block(Fwd& f, Arena& A, const VaeAttnW& a, const float* h, float* o, int T, int Sx, cudaStream_t st) {
  const int C = a.C;
  if (Sx % 64) return set_error("first-stage attention over %d positions (must be a multiple of 64)", Sx);
  const size_t m = A.mark();
  bf16* hn = A.get<bf16>(static_cast<size_t>(T) * Sx * C);
  bf16* qk = A.get<bf16>(static_cast<size_t>(T) * Sx * 2 * C);
  bf16* vT = A.get<bf16>(static_cast<size_t>(C) * Sx);
  float* sc = A.get<float>(static_cast<size_t>(Sx) * Sx);
  bf16* pr = A.get<bf16>(static_cast<size_t>(Sx) * Sx);
  bf16* att = A.get<bf16>(static_cast<size_t>(T) * Sx * C);
  if (A.failed) return set_error("workspace exhausted (first-stage attention)");
  MD_CHECK(f.gn(h, C, false, nullptr, 0, T, Sx, 32, 1e-6f, a.norm, ACT_NONE, hn, nullptr));
  MD_CHECK(f.gemm(hn, T, Sx, a.qk, nullptr, nullptr, qk, false));
  for (int b = 0; b < T; ++b) {
    md_conv_gemm_args g;
    const bf16* hb = hn + static_cast<size_t>(b) * Sx * C;
    const bf16* qb = qk + static_cast<size_t>(b) * Sx * 2 * C;
    // V^T [C][Sx] = Wv [C][C] . h_b^T: the normalised activations of sample b are the K-major "weight" operand
    memset(&g, 0, sizeof(g));
    g.A = a.wv; g.B = 1; g.D = 1; g.H = 1; g.W = C; g.Cin = C; g.Wt = hb; g.N = Sx; g.ntaps = 1; g.out_bf16 = vT;
    MD_CHECK(launch_conv_gemm(g, st));
    // scores [Sx][Sx] = q_b k_b^T * C^-0.5 (q at columns 0..C-1, k at C..2C-1 of the fused activation)
    memset(&g, 0, sizeof(g));
    g.A = qb; g.B = 1; g.D = 1; g.H = 1; g.W = Sx; g.Cin = C; g.Cpitch = 2 * C; g.Wt = qb + C; g.Wpitch = 2 * C;
    g.N = Sx; g.ntaps = 1; g.out_f32 = sc; g.out_scale = 1.f / sqrtf(static_cast<float>(C));
    MD_CHECK(launch_conv_gemm(g, st));
    MD_CHECK(launch_softmax_rows(sc, pr, static_cast<size_t>(Sx), Sx, st));
    // O [Sx][C] = P V + b_v
    memset(&g, 0, sizeof(g));
    g.A = pr; g.B = 1; g.D = 1; g.H = 1; g.W = Sx; g.Cin = Sx; g.Wt = vT; g.N = C; g.ntaps = 1; g.bias = a.bv;
    g.out_bf16 = att + static_cast<size_t>(b) * Sx * C;
    MD_CHECK(launch_conv_gemm(g, st));
  }
  MD_CHECK(f.gemm(att, T, Sx, a.proj, h, o, nullptr, true));
  A.release(m);
  return 0;
}

// decode_first_stage (morphable_diffusion.py:468-471) = AutoencoderKL.decode (ldm/models/autoencoder.py:330-333) =
// post_quant_conv + Decoder.forward (ldm/modules/diffusionmodules/model.py:535-569) on the UNet's kernels: ResnetBlock =
// GroupNorm(eps 1e-6)+SiLU -> conv3x3 -> GroupNorm+SiLU -> conv3x3 (+ 1x1 nin_shortcut), statistics from the GEMM
// epilogues; Upsample = nearest x2 + conv3x3; the single AttnBlock (one head of 512 channels over h*w positions) as
// three tensor-core GEMMs per sample around a row softmax: S = q k^T (the keys are the "weight" operand, addressed
// inside the fused q|k activation through its row pitch), V^T = Wv h^T (so that V arrives K-major for the last
// product), O = softmax(S) V + b_v.
int vae_decode(Ctx& c, const float* x, float* image, int T, int S, cudaStream_t st) {
  if (!c.vae.loaded) return set_error("vae_decode: no first-stage decoder weights were loaded (first_stage_model.*)");
  if (T < 1 || S < 8 || S % 8) return set_error("vae_decode: bad shape T=%d S=%d", T, S);
  const VaeW& v = c.vae;
  Arena& A = c.arena;
  A.off = 0;
  A.failed = false;
  Fwd f{c, st, T, T, nullptr, nullptr};
  f.res_eps = 1e-6f;
  const size_t spool_floats = static_cast<size_t>(T) * 2 * 48 * 512;   // ~35 GroupNorm inputs of at most [T][512][2]
  f.spool = A.get<float>(spool_floats);
  f.spool_cap = spool_floats;
  if (A.failed) return set_error("workspace exhausted (vae statistics)");
  MD_CUDA(cudaMemsetAsync(f.spool, 0, spool_floats * sizeof(float), st));

  int H = S;
  int ch = v.conv_in.N;
  const size_t HW = static_cast<size_t>(H) * H;
  // the decoder has no skip connections: every layer consumes h and produces the next h, so two buffers of the largest
  // activation (the 256-channel tensor at full resolution behind the last Upsample) alternate
  const size_t max_act = static_cast<size_t>(T) * HW * 64 * 256;
  float* bufs[2] = {A.get<float>(max_act), A.get<float>(max_act)};
  if (A.failed) return set_error("workspace exhausted (vae activations: %zu MB)", (2 * max_act * sizeof(float)) >> 20);
  int cur = 0;
  float* h = bufs[0];
  {
    const size_t m = A.mark();
    bf16* zin = A.get<bf16>(static_cast<size_t>(T) * HW * 64);
    if (A.failed) return set_error("workspace exhausted (vae)");
    MD_CHECK(launch_vae_input(x, v.pq, 1.f / 0.18215f, zin, T, static_cast<int>(HW), st));
    MD_CHECK(f.conv(zin, T, H, H, v.conv_in, nullptr, 0, nullptr, h, nullptr, true));
    A.release(m);
  }
  auto next_buf = [&]() { cur ^= 1; return bufs[cur]; };
  auto block = [&](const ResW& r) -> int {
    float* o = next_buf();
    MD_CHECK(f.res_block(r, h, ch, nullptr, 0, H, H, o));
    h = o; ch = r.cout;
    return 0;
  };
  MD_CHECK(block(v.mid1));
  {  // AttnBlock (model.py:177-203)
    float* o = next_buf();
    MD_CHECK(vae_attn_block(f, A, v.attn, h, o, T, H * H, st));
    h = o;
  }
  MD_CHECK(block(v.mid2));
  for (int lev = static_cast<int>(v.up.size()) - 1; lev >= 0; --lev) {
    for (const ResW& r : v.up[lev]) MD_CHECK(block(r));
    if (lev != 0) {
      float* o = next_buf();
      MD_CHECK(f.upsample(v.upsample[lev], h, H, H, ch, o));
      h = o; H *= 2;
    }
  }
  bf16* ao = A.get<bf16>(static_cast<size_t>(T) * H * H * ch);
  float* rgb8 = A.get<float>(static_cast<size_t>(T) * H * H * 8);
  if (A.failed) return set_error("workspace exhausted (vae)");
  MD_CHECK(f.gn(h, ch, false, nullptr, 0, T, H * H, 32, 1e-6f, v.norm_out, ACT_SILU, ao, nullptr));
  MD_CHECK(f.conv(ao, T, H, H, v.conv_out, nullptr, 0, nullptr, rgb8, nullptr, false));
  MD_CHECK(launch_rows_to_nchw(rgb8, 8, image, T, v.out_ch, H * H, st));
  return 0;
}

// AutoencoderKL.encode up to the posterior moments (ldm/models/autoencoder.py:324-328): Encoder.forward
// (ldm/modules/diffusionmodules/model.py:432-459) + quant_conv (folded into conv_out at load time).  Same kernels as the
// decoder; Downsample (:70-78) is the stride-2 implicit GEMM with taps 0..2 (right / bottom zero padding = TMA zero fill).
int vae_encode(Ctx& c, const float* image, float* moments, int T, int S, cudaStream_t st) {
  if (!c.vae_enc.loaded) return set_error("vae_encode: no first-stage encoder weights were loaded (first_stage_model.encoder.*)");
  if (T < 1 || S < 8 || S % 8) return set_error("vae_encode: bad shape T=%d S=%d", T, S);
  const VaeEncW& v = c.vae_enc;
  Arena& A = c.arena;
  A.off = 0;
  A.failed = false;
  Fwd f{c, st, T, T, nullptr, nullptr};
  f.res_eps = 1e-6f;
  const size_t spool_floats = static_cast<size_t>(T) * 2 * 48 * 512;
  f.spool = A.get<float>(spool_floats);
  f.spool_cap = spool_floats;
  if (A.failed) return set_error("workspace exhausted (vae statistics)");
  MD_CUDA(cudaMemsetAsync(f.spool, 0, spool_floats * sizeof(float), st));

  int H = 8 * S;
  int ch = v.conv_in.N;
  const size_t max_act = static_cast<size_t>(T) * H * H * ch;   // the 128-channel tensors at full resolution
  float* bufs[2] = {A.get<float>(max_act), A.get<float>(max_act)};
  if (A.failed) return set_error("workspace exhausted (vae encoder activations: %zu MB)", (2 * max_act * sizeof(float)) >> 20);
  int cur = 0;
  float* h = bufs[0];
  auto next_buf = [&]() { cur ^= 1; return bufs[cur]; };
  {
    const size_t m = A.mark();
    bf16* xin = A.get<bf16>(static_cast<size_t>(T) * H * H * 64);
    if (A.failed) return set_error("workspace exhausted (vae encoder input)");
    MD_CHECK(launch_nchw_to_cl64(image, xin, T, 3, static_cast<size_t>(H) * H, st));
    MD_CHECK(f.conv(xin, T, H, H, v.conv_in, nullptr, 0, nullptr, h, nullptr, true));
    A.release(m);
  }
  auto block = [&](const ResW& r) -> int {
    float* o = next_buf();
    MD_CHECK(f.res_block(r, h, ch, nullptr, 0, H, H, o));
    h = o; ch = r.cout;
    return 0;
  };
  for (size_t lev = 0; lev < v.down.size(); ++lev) {
    for (const ResW& r : v.down[lev]) MD_CHECK(block(r));
    if (lev + 1 != v.down.size()) {
      float* o = next_buf();
      MD_CHECK(f.downsample(v.downsample[lev], h, nullptr, H, H, ch, o, /*tap0=*/0));
      h = o; H /= 2;
    }
  }
  MD_CHECK(block(v.mid1));
  {
    float* o = next_buf();
    MD_CHECK(vae_attn_block(f, A, v.attn, h, o, T, H * H, st));
    h = o;
  }
  MD_CHECK(block(v.mid2));
  bf16* ao = A.get<bf16>(static_cast<size_t>(T) * H * H * ch);
  float* mom8 = A.get<float>(static_cast<size_t>(T) * H * H * 8);
  if (A.failed) return set_error("workspace exhausted (vae encoder)");
  MD_CHECK(f.gn(h, ch, false, nullptr, 0, T, H * H, 32, 1e-6f, v.norm_out, ACT_SILU, ao, nullptr));
  MD_CHECK(f.conv(ao, T, H, H, v.conv_out, nullptr, 0, nullptr, mom8, nullptr, false));
  MD_CHECK(launch_rows_to_nchw(mom8, 8, moments, T, 8, H * H, st));
  return 0;
}

// FrozenCLIPImageEmbedder.encode (ldm/modules/encoders/modules.py:363-382) = preprocess + clip encode_image (OpenAI clip
// VisionTransformer.forward): bicubic resize / normalise / patchify (one kernel) -> patch-embedding GEMM -> class token +
// positions + ln_pre -> 24 x {LayerNorm -> qkv GEMM -> attention (257 tokens, 16 heads of 64: mma.sync kernel) -> out_proj
// GEMM adding into the fp32 residual stream -> LayerNorm -> c_fc GEMM with QuickGELU epilogue -> c_proj GEMM adding into
// the stream} -> ln_post of the class tokens -> projection GEMM.
int clip_embed(Ctx& c, const float* image, float* out, int n, int H, int W, cudaStream_t st) {
  if (!c.clip.loaded) return set_error("clip_embed: no CLIP image-tower weights were loaded (clip_image_encoder.model.visual.*)");
  if (n < 1 || H < 2 || W < 2) return set_error("clip_embed: bad shape n=%d H=%d W=%d", n, H, W);
  const ClipW& q = c.clip;
  const int C = q.width, S = q.ntok, heads = q.heads;
  Arena& A = c.arena;
  A.off = 0;
  A.failed = false;
  Fwd f{c, st, n, n, nullptr, nullptr};
  const size_t rows = static_cast<size_t>(n) * S;
  bf16* patches = A.get<bf16>(static_cast<size_t>(n) * (S - 1) * 640);
  float* pe = A.get<float>(static_cast<size_t>(n) * (S - 1) * C);
  float* x = A.get<float>(rows * C);
  bf16* ln = A.get<bf16>(rows * C);
  bf16* qkv = A.get<bf16>(rows * 3 * C);
  bf16* att = A.get<bf16>(rows * C);
  bf16* hid = A.get<bf16>(rows * 4 * C);
  float* cls = A.get<float>(static_cast<size_t>(n) * C);
  bf16* cls_b = A.get<bf16>(static_cast<size_t>(n) * C);
  if (A.failed) return set_error("workspace exhausted (clip)");
  MD_CHECK(launch_clip_patches(image, patches, n, H, W, st));
  MD_CHECK(f.gemm(patches, n, S - 1, q.conv1, nullptr, pe, nullptr, false));
  MD_CHECK(launch_clip_tokens(pe, q.class_emb, q.pos_emb, q.ln_pre.g, q.ln_pre.b, x, n, S, C, st));
  for (const ClipLayerW& l : q.layers) {
    MD_CHECK(launch_layer_norm(x, nullptr, 0, l.ln1.g, l.ln1.b, ln, rows, S, C, 1e-5f, st));
    MD_CHECK(f.gemm(ln, n, S, l.qkv, nullptr, nullptr, qkv, false));
    MD_CHECK(launch_self_attention(qkv, att, n, S, heads, C / heads, st));
    MD_CHECK(f.gemm(att, n, S, l.out, x, x, nullptr, false));
    MD_CHECK(launch_layer_norm(x, nullptr, 0, l.ln2.g, l.ln2.b, ln, rows, S, C, 1e-5f, st));
    MD_CHECK(f.gemm(ln, n, S, l.fc, nullptr, nullptr, hid, false, ACT_QUICKGELU));
    MD_CHECK(f.gemm(hid, n, S, l.proj, x, x, nullptr, false));
  }
  MD_CUDA(cudaMemcpy2DAsync(cls, sizeof(float) * C, x, sizeof(float) * S * C, sizeof(float) * C, n, cudaMemcpyDeviceToDevice, st));
  MD_CHECK(launch_layer_norm(cls, nullptr, 0, q.ln_post.g, q.ln_post.b, cls_b, n, 1, C, 1e-5f, st));
  MD_CHECK(f.gemm(cls_b, 1, n, q.proj, nullptr, out, nullptr, false));
  return 0;
}

}  // namespace md
