// Weight ingestion: the caller hands over the reference state dict (fp32 device tensors addressed by the
// reference's nn.Module paths) and this file re-lays it out for the kernels:
//   conv / linear weights -> bf16 [N][tap][Cin] (K-major operands of conv_gemm), q|k|v and k|v fused along N,
//   GEGLU projection interleaved per 256-row tile (128 value | 128 gate), ConvTranspose3d split into 8 output-parity classes,
//   BatchNorm1d (eval) folded into scale/shift, all ResBlock emb_layers concatenated into one matrix.
#include "engine.h"

namespace md {

struct TapList { int t[27]; };

template <typename OutT>
__global__ void pack_kernel(const float* __restrict__ src, OutT* __restrict__ dst, int N, int T, int I, long sn,
                            long st, long si, TapList taps, long dn, long dt, long di, int geglu_inner) {
  const size_t total = static_cast<size_t>(N) * T * I;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx % I);
    const int t = static_cast<int>((idx / I) % T);
    const int n = static_cast<int>(idx / (static_cast<size_t>(I) * T));
    int ns = n;
    if (geglu_inner > 0) {
      constexpr int H = kGegluTile / 2;
      const int tile = n / kGegluTile, r = n % kGegluTile;
      ns = (r < H) ? (tile * H + r) : (geglu_inner + tile * H + (r - H));
    }
    const float v = src[ns * sn + taps.t[t] * st + i * si];
    dst[n * dn + t * dt + i * di] = static_cast<OutT>(v);
  }
}

__global__ void permute_rows_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int geglu_inner) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  constexpr int H = kGegluTile / 2;
  const int tile = n / kGegluTile, r = n % kGegluTile;
  const int ns = (r < H) ? (tile * H + r) : (geglu_inner + tile * H + (r - H));
  dst[n] = src[ns];
}

// C[m][n] (bf16, row stride ldc) = scale * sum_k A(m,k) * B(k,n), A(m,k) = transA ? A[k*lda+m] : A[m*lda+k],
// B(k,n) = B[k*ldb+n]   (fp32 inputs; load-time weight products only)
__global__ void matmul_f32_to_bf16_kernel(const float* __restrict__ A, int lda, int transA, const float* __restrict__ Bm,
                                          int ldb, __nv_bfloat16* __restrict__ Cm, int ldc, int M, int N, int K,
                                          float scale) {
  __shared__ float sa[16][17], sb[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int ka = k0 + tx;
    sa[ty][tx] = (row < M && ka < K) ? (transA ? A[static_cast<size_t>(ka) * lda + row] : A[static_cast<size_t>(row) * lda + ka]) : 0.f;
    sb[ty][tx] = (col < N && k0 + ty < K) ? Bm[static_cast<size_t>(k0 + ty) * ldb + col] : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += sa[ty][k] * sb[k][tx];
    __syncthreads();
  }
  if (row < M && col < N) Cm[static_cast<size_t>(row) * ldc + col] = __float2bfloat16(acc * scale);
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

__global__ void bn_fold_kernel(const float* g, const float* b, const float* mean, const float* var, float eps,
                               float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = g[c] / sqrtf(var[c] + eps);
  scale[c] = s;
  shift[c] = b[c] - mean[c] * s;
}

// out[o] = b2[o] + sum_j W[o][j] * b1[j]   (bias of two composed linear maps, n <= 32)
__global__ void fold_bias_kernel(const float* __restrict__ W, const float* __restrict__ b1, const float* __restrict__ b2,
                                 float* __restrict__ out, int n) {
  const int o = threadIdx.x;
  if (o >= n) return;
  float a = b2[o];
  for (int j = 0; j < n; ++j) a += W[o * n + j] * b1[j];
  out[o] = a;
}

struct Loader {
  Ctx& c;
  const TensorMap& tm;
  cudaStream_t st;
  int rc = 0;

  template <typename T> T* dalloc(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) {
      rc = set_error("load_weights: cudaMalloc of %zu bytes failed", n * sizeof(T));
      return nullptr;
    }
    c.weight_allocs.push_back(p);
    return static_cast<T*>(p);
  }
  const NamedTensor* find(const std::string& name, size_t numel) {
    auto it = tm.find(name);
    if (it == tm.end()) { rc = set_error("load_weights: missing tensor '%s'", name.c_str()); return nullptr; }
    if (it->second.numel != numel) {
      rc = set_error("load_weights: tensor '%s' has %zu elements, expected %zu", name.c_str(), it->second.numel, numel);
      return nullptr;
    }
    return &it->second;
  }
  bool has(const std::string& name) const { return tm.find(name) != tm.end(); }

  // fp32 copy owned by the library
  const float* copy_f32(const std::string& name, size_t numel) {
    const NamedTensor* t = find(name, numel);
    if (!t) return nullptr;
    float* d = dalloc<float>(numel);
    if (!d) return nullptr;
    if (cudaMemcpyAsync(d, t->ptr, numel * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      rc = set_error("load_weights: copy of '%s' failed", name.c_str());
    return d;
  }
  NormW norm(const std::string& p, int C) {
    NormW n;
    n.C = C;
    n.g = copy_f32(p + ".weight", C);
    n.b = copy_f32(p + ".bias", C);
    return n;
  }
  template <typename OutT>
  void pack(const float* src, OutT* dst, int N, int T, int I, long sn, long stt, long si, const TapList& taps, long dn,
            long dt, long di, int geglu_inner = 0) {
    const size_t total = static_cast<size_t>(N) * T * I;
    const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 32));
    pack_kernel<OutT><<<blocks, 256, 0, st>>>(src, dst, N, T, I, sn, stt, si, taps, dn, dt, di, geglu_inner);
    if (cudaGetLastError() != cudaSuccess) rc = set_error("pack kernel launch failed");
  }
  static TapList iota() {
    TapList t;
    for (int i = 0; i < 27; ++i) t.t[i] = i;
    return t;
  }
  // Conv (torch layout [O][I][taps]) or Linear ([O][I], taps=1) -> bf16 [O][tap][I], optionally into a slice of a
  // wider fused matrix (dst_rows_off).
  void pack_conv_into(bf16* dst, const std::string& wname, int O, int I, int taps, int geglu_inner = 0) {
    const NamedTensor* t = find(wname, static_cast<size_t>(O) * I * taps);
    if (!t) return;
    pack<bf16>(t->ptr, dst, O, taps, I, static_cast<long>(I) * taps, 1, taps, iota(), static_cast<long>(taps) * I, I, 1,
               geglu_inner);
  }
  GemmW gemm(const std::string& p, int O, int I, int taps, bool bias, int geglu_inner = 0) {
    GemmW g;
    g.N = O; g.K = I; g.taps = taps;
    g.w = dalloc<bf16>(static_cast<size_t>(O) * I * taps);
    if (!g.w) return g;
    pack_conv_into(g.w, p + ".weight", O, I, taps, geglu_inner);
    if (bias) {
      if (geglu_inner > 0) {
        const NamedTensor* t = find(p + ".bias", O);
        float* d = dalloc<float>(O);
        if (t && d) permute_rows_f32_kernel<<<(O + 127) / 128, 128, 0, st>>>(t->ptr, d, O, geglu_inner);
        g.bias = d;
      } else {
        g.bias = copy_f32(p + ".bias", O);
      }
    }
    return g;
  }
  // fused along N: several [Oi][I] matrices stacked
  GemmW gemm_fused(const std::vector<std::string>& names, int O_each, int I) {
    GemmW g;
    g.N = O_each * static_cast<int>(names.size()); g.K = I; g.taps = 1;
    g.w = dalloc<bf16>(static_cast<size_t>(g.N) * I);
    if (!g.w) return g;
    for (size_t k = 0; k < names.size(); ++k) pack_conv_into(g.w + k * static_cast<size_t>(O_each) * I, names[k], O_each, I, 1);
    return g;
  }

  // first-stage ResnetBlock (model.py:82-141, temb_channels = 0) and AttnBlock (:150-203)
  ResW vae_res(const std::string& q, int cin, int cout) {
    ResW r;
    r.cin = cin; r.cout = cout; r.emb_off = -1;
    r.n1 = norm(q + "norm1", cin);
    r.c1 = gemm(q + "conv1", cout, cin, 9, true);
    r.n2 = norm(q + "norm2", cout);
    r.c2 = gemm(q + "conv2", cout, cout, 9, true);
    r.has_skip = cin != cout;
    if (r.has_skip) r.skip = gemm(q + "nin_shortcut", cout, cin, 1, true);
    return r;
  }
  VaeAttnW vae_attn(const std::string& q, int C) {
    // q | k fused; v kept as a [C][C] matrix (it becomes the A operand of V^T = Wv . h^T)
    VaeAttnW t;
    t.C = C;
    t.norm = norm(q + "norm", C);
    t.qk = gemm_fused({q + "q.weight", q + "k.weight"}, C, C);
    float* qkb = dalloc<float>(2 * C);
    const NamedTensor* bq = find(q + "q.bias", C);
    const NamedTensor* bk = find(q + "k.bias", C);
    if (qkb && bq && bk) {
      cudaMemcpyAsync(qkb, bq->ptr, sizeof(float) * C, cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(qkb + C, bk->ptr, sizeof(float) * C, cudaMemcpyDeviceToDevice, st);
    }
    t.qk.bias = qkb;
    t.wv = dalloc<bf16>(static_cast<size_t>(C) * C);
    if (t.wv) pack_conv_into(t.wv, q + "v.weight", C, C, 1);
    t.bv = copy_f32(q + "v.bias", C);
    t.proj = gemm(q + "proj_out", C, C, 1, true);
    return t;
  }
  // conv with Cin <= 64 zero-padded to one K block: bf16 [O][9][64]
  GemmW conv_padded_in(const std::string& p, int O, int I) {
    GemmW g;
    g.N = O; g.K = 64; g.taps = 9;
    const NamedTensor* t = find(p + ".weight", static_cast<size_t>(O) * I * 9);
    g.w = dalloc<bf16>(static_cast<size_t>(O) * 9 * 64);
    if (t && g.w) {
      cudaMemsetAsync(g.w, 0, sizeof(bf16) * O * 9 * 64, st);
      pack<bf16>(t->ptr, g.w, O, 9, I, static_cast<long>(I) * 9, 1, 9, iota(), static_cast<long>(9) * 64, 64, 1);
    }
    g.bias = copy_f32(p + ".bias", O);
    return g;
  }

  ResW res(const std::string& p, int cin, int cout, int emb_dim, int& emb_off, std::vector<std::pair<std::string, int>>& emb_list) {
    ResW r;
    r.cin = cin; r.cout = cout;
    r.n1 = norm(p + "in_layers.0", cin);
    r.c1 = gemm(p + "in_layers.2", cout, cin, 9, true);
    r.emb_off = emb_off;
    emb_list.push_back({p + "emb_layers.1", cout});
    emb_off += cout;
    r.n2 = norm(p + "out_layers.0", cout);
    r.c2 = gemm(p + "out_layers.3", cout, cout, 9, true);
    r.has_skip = cin != cout;
    if (r.has_skip) r.skip = gemm(p + "skip_connection", cout, cin, 1, true);
    (void)emb_dim;
    return r;
  }
  int v2_off = 0;
  std::vector<std::pair<std::string, int>> v2_list;
  STW st_block(const std::string& p, int C, int heads, int ctx) {
    STW s;
    s.C = C; s.heads = heads;
    s.norm = norm(p + "norm", C);
    s.proj_in = gemm(p + "proj_in", C, C, 1, true);
    const std::string t = p + "transformer_blocks.0.";
    s.ln1 = norm(t + "norm1", C); s.ln2 = norm(t + "norm2", C); s.ln3 = norm(t + "norm3", C);
    s.qkv = gemm_fused({t + "attn1.to_q.weight", t + "attn1.to_k.weight", t + "attn1.to_v.weight"}, C, C);
    s.o1 = gemm(t + "attn1.to_out.0", C, C, 1, true);
    s.v2_off = v2_off;
    v2_list.push_back({t + "attn2.", C});
    v2_off += C;
    s.ff1 = gemm(t + "ff.net.0.proj", 8 * C, C, 1, true, /*geglu_inner=*/4 * C);
    s.ff2 = gemm(t + "ff.net.2", C, 4 * C, 1, true);
    s.proj_out = gemm(p + "proj_out", C, C, 1, true);
    return s;
  }
  DepthW depth(const std::string& p, int dim, int dhead, int ctx) {
    DepthW d;
    d.dim = dim; d.dhead = dhead; d.inner = 4 * dhead; d.ctx = ctx;
    d.proj_in = gemm(p + "proj_in.0", d.inner, dim, 1, true);
    d.gn_in = norm(p + "proj_in.1", d.inner);
    d.proj_ctx = gemm(p + "proj_context.0", ctx, ctx, 1, false);
    d.gn_ctx = norm(p + "proj_context.1", ctx);
    // Re-associated depth attention (attention.cu): per head h
    //   W_qk[h] = scale * W_k[h]^T W_q[h]  ([ctx][inner]),  W_ov[:, h] = W_out[:, h] W_v[h]  ([inner][ctx])
    {
      const NamedTensor* wq = find(p + "depth_attn.to_q.weight", static_cast<size_t>(d.inner) * d.inner);
      const NamedTensor* wk = find(p + "depth_attn.to_k.weight", static_cast<size_t>(d.inner) * ctx);
      const NamedTensor* wv = find(p + "depth_attn.to_v.weight", static_cast<size_t>(d.inner) * ctx);
      const NamedTensor* wo = find(p + "depth_attn.to_out.weight", static_cast<size_t>(d.inner) * d.inner);
      d.wqk.N = 4 * ctx; d.wqk.K = d.inner; d.wqk.taps = 1;
      d.wqk.w = dalloc<bf16>(static_cast<size_t>(4) * ctx * d.inner);
      d.wov.N = d.inner; d.wov.K = 4 * ctx; d.wov.taps = 1;
      d.wov.w = dalloc<bf16>(static_cast<size_t>(4) * ctx * d.inner);
      if (wq && wk && wv && wo && d.wqk.w && d.wov.w) {
        const float scale = 1.f / sqrtf(static_cast<float>(dhead));
        dim3 blk(16, 16);
        for (int h = 0; h < 4; ++h) {
          dim3 g1((d.inner + 15) / 16, (ctx + 15) / 16);
          matmul_f32_to_bf16_kernel<<<g1, blk, 0, st>>>(wk->ptr + static_cast<size_t>(h) * dhead * ctx, ctx, 1,
                                                        wq->ptr + static_cast<size_t>(h) * dhead * d.inner, d.inner,
                                                        d.wqk.w + static_cast<size_t>(h) * ctx * d.inner, d.inner, ctx,
                                                        d.inner, dhead, scale);
          dim3 g2((ctx + 15) / 16, (d.inner + 15) / 16);
          matmul_f32_to_bf16_kernel<<<g2, blk, 0, st>>>(wo->ptr + static_cast<size_t>(h) * dhead, d.inner, 0,
                                                        wv->ptr + static_cast<size_t>(h) * dhead * ctx, ctx,
                                                        d.wov.w + static_cast<size_t>(h) * ctx, 4 * ctx, d.inner, ctx,
                                                        dhead, 1.f);
        }
      }
    }
    d.gn_o1 = norm(p + "proj_out.0", d.inner);
    d.conv1 = gemm(p + "proj_out.2", d.inner, d.inner, 9, false);
    d.gn_o2 = norm(p + "proj_out.3", d.inner);
    d.conv2 = gemm(p + "proj_out.5", dim, d.inner, 9, false);
    return d;
  }
};

int load_all_weights(Ctx& c, const TensorMap& tm, cudaStream_t st) {
  free_weights(c);
  Loader L{c, tm, st};
  const md_config& mc = c.mcfg;

  // ------------------------------------------------ UNet
  UNetW& u = c.unet;
  u = UNetW();
  const std::string P = "model.diffusion_model.";
  const int mch = mc.model_channels;
  u.model_channels = mch; u.in_channels = mc.in_channels; u.out_channels = mc.out_channels;
  u.heads = mc.num_heads; u.ctx_dim = mc.context_dim; u.emb_dim = 4 * mch;
  u.te0_w = L.copy_f32(P + "time_embed.0.weight", static_cast<size_t>(u.emb_dim) * mch);
  u.te0_b = L.copy_f32(P + "time_embed.0.bias", u.emb_dim);
  u.te2_w = L.copy_f32(P + "time_embed.2.weight", static_cast<size_t>(u.emb_dim) * u.emb_dim);
  u.te2_b = L.copy_f32(P + "time_embed.2.bias", u.emb_dim);
  int emb_off = 0;
  std::vector<std::pair<std::string, int>> emb_list;

  {  // conv_in (8 -> 320): bf16 [mch][9][64], input channels zero-padded to one K block of the implicit GEMM
    if (mc.in_channels > 64 || mc.in_channels % 4) return set_error("load_weights: in_channels=%d unsupported", mc.in_channels);
    const NamedTensor* t = L.find(P + "input_blocks.0.0.weight", static_cast<size_t>(mch) * mc.in_channels * 9);
    u.conv_in_g.N = mch; u.conv_in_g.K = 64; u.conv_in_g.taps = 9;
    u.conv_in_g.w = L.dalloc<bf16>(static_cast<size_t>(mch) * 9 * 64);
    if (t && u.conv_in_g.w) {
      cudaMemsetAsync(u.conv_in_g.w, 0, sizeof(bf16) * mch * 9 * 64, st);
      L.pack<bf16>(t->ptr, u.conv_in_g.w, mch, 9, mc.in_channels, static_cast<long>(mc.in_channels) * 9, 1, 9,
                   Loader::iota(), static_cast<long>(9) * 64, 64, 1);
    }
    u.conv_in_g.bias = L.copy_f32(P + "input_blocks.0.0.bias", mch);
  }
  std::vector<int> chans{mch};
  int ch = mch, ds = 1;
  u.input_blocks.clear();
  u.input_blocks.push_back({});  // block 0 = conv_in (handled separately)
  auto has_attn = [&](int d) { return (d == 1 && mc.attn_ds[0]) || (d == 2 && mc.attn_ds[1]) || (d == 4 && mc.attn_ds[2]) || (d == 8 && mc.attn_ds[3]); };
  for (int level = 0; level < 4; ++level) {
    const int mult = mc.channel_mult[level];
    for (int nr = 0; nr < mc.num_res_blocks; ++nr) {
      const int bi = static_cast<int>(u.input_blocks.size());
      const std::string bp = P + "input_blocks." + std::to_string(bi) + ".";
      std::vector<UNetLayer> layers;
      UNetLayer r; r.kind = 1;
      r.res = L.res(bp + "0.", ch, mult * mch, u.emb_dim, emb_off, emb_list);
      layers.push_back(r);
      ch = mult * mch;
      if (has_attn(ds)) {
        UNetLayer s; s.kind = 2;
        s.st = L.st_block(bp + "1.", ch, u.heads, u.ctx_dim);
        layers.push_back(s);
      }
      u.input_blocks.push_back(layers);
      chans.push_back(ch);
    }
    if (level != 3) {
      const int bi = static_cast<int>(u.input_blocks.size());
      UNetLayer d; d.kind = 3;
      d.conv = L.gemm(P + "input_blocks." + std::to_string(bi) + ".0.op", ch, ch, 9, true);
      u.input_blocks.push_back({d});
      chans.push_back(ch);
      ds *= 2;
    }
  }
  u.mid0 = L.res(P + "middle_block.0.", ch, ch, u.emb_dim, emb_off, emb_list);
  u.mid1 = L.st_block(P + "middle_block.1.", ch, u.heads, u.ctx_dim);
  u.mid2 = L.res(P + "middle_block.2.", ch, ch, u.emb_dim, emb_off, emb_list);
  u.output_blocks.clear();
  for (int level = 3; level >= 0; --level) {
    const int mult = mc.channel_mult[level];
    for (int i = 0; i <= mc.num_res_blocks; ++i) {
      const int ich = chans.back();
      chans.pop_back();
      const int bi = static_cast<int>(u.output_blocks.size());
      const std::string bp = P + "output_blocks." + std::to_string(bi) + ".";
      std::vector<UNetLayer> layers;
      UNetLayer r; r.kind = 1;
      r.res = L.res(bp + "0.", ch + ich, mch * mult, u.emb_dim, emb_off, emb_list);
      layers.push_back(r);
      ch = mch * mult;
      if (has_attn(ds)) {
        UNetLayer s; s.kind = 2;
        s.st = L.st_block(bp + "1.", ch, u.heads, u.ctx_dim);
        layers.push_back(s);
      }
      if (level && i == mc.num_res_blocks) {
        UNetLayer up; up.kind = 4;
        up.conv = L.gemm(bp + std::to_string(layers.size()) + ".conv", ch, ch, 9, true);
        layers.push_back(up);
        ds /= 2;
      }
      u.output_blocks.push_back(layers);
    }
  }
  u.out_norm = L.norm(P + "out.0", mch);
  {  // final conv 320 -> 4: bf16 [8][9*320] (zero rows 4..7) so it runs on the tensor-core implicit GEMM
    const NamedTensor* t = L.find(P + "out.2.weight", static_cast<size_t>(mc.out_channels) * mch * 9);
    const NamedTensor* b = L.find(P + "out.2.bias", mc.out_channels);
    if (mc.out_channels > 8) return set_error("load_weights: out_channels > 8 unsupported");
    u.out_g.N = 8; u.out_g.K = mch; u.out_g.taps = 9;
    u.out_g.w = L.dalloc<bf16>(static_cast<size_t>(8) * mch * 9);
    float* ob = L.dalloc<float>(8);
    u.out_g.bias = ob;
    if (t && b && u.out_g.w && ob) {
      cudaMemsetAsync(u.out_g.w, 0, sizeof(bf16) * 8 * mch * 9, st);
      cudaMemsetAsync(ob, 0, sizeof(float) * 8, st);
      L.pack<bf16>(t->ptr, u.out_g.w, mc.out_channels, 9, mch, static_cast<long>(mch) * 9, 1, 9, Loader::iota(),
                   static_cast<long>(9) * mch, mch, 1);
      cudaMemcpyAsync(ob, b->ptr, sizeof(float) * mc.out_channels, cudaMemcpyDeviceToDevice, st);
    }
  }
  // depth transformers (attention.py:96-115)
  const int cm2 = mc.channel_mult[2], cm1 = mc.channel_mult[1], cm0 = mc.channel_mult[0];
  const int* vd = mc.volume_dims;
  u.mid_cond = L.depth(P + "middle_conditions.", mch * cm2, vd[3] / 2, vd[3]);
  const int dims[9][2] = {{mch * cm2, vd[2]}, {mch * cm2, vd[2]}, {mch * cm2, vd[1]}, {mch * cm1, vd[1]},
                          {mch * cm1, vd[1]}, {mch * cm1, vd[0]}, {mch * cm0, vd[0]}, {mch * cm0, vd[0]},
                          {mch * cm0, vd[0]}};
  u.out_cond.clear();
  for (int i = 0; i < 9; ++i)
    u.out_cond.push_back(L.depth(P + "output_conditions." + std::to_string(i) + ".", dims[i][0], dims[i][1] / 2, dims[i][1]));
  // concatenated emb_layers: one [emb_total][emb_dim] bf16 GEMM operand
  u.emb_total = emb_off;
  {
    u.emb_g.N = emb_off; u.emb_g.K = u.emb_dim; u.emb_g.taps = 1;
    u.emb_g.w = L.dalloc<bf16>(static_cast<size_t>(emb_off) * u.emb_dim);
    float* eb = L.dalloc<float>(emb_off);
    u.emb_g.bias = eb;
    if (u.emb_g.w && eb) {
      int off = 0;
      for (auto& e : emb_list) {
        const NamedTensor* w = L.find(e.first + ".weight", static_cast<size_t>(e.second) * u.emb_dim);
        const NamedTensor* b = L.find(e.first + ".bias", e.second);
        if (w && b) {
          f32_to_bf16_kernel<<<256, 256, 0, st>>>(w->ptr, u.emb_g.w + static_cast<size_t>(off) * u.emb_dim,
                                                  static_cast<size_t>(e.second) * u.emb_dim);
          cudaMemcpyAsync(eb + off, b->ptr, sizeof(float) * e.second, cudaMemcpyDeviceToDevice, st);
        }
        off += e.second;
      }
    }
  }
  // attn2 with ONE context token: softmax == 1, so attn2(x) = to_out(to_v(ctx)) = (W_out W_v) ctx + b_out for every
  // query.  The products of all transformer blocks are concatenated into one [v2_total][ctx_dim] operand.
  u.v2_total = L.v2_off;
  {
    u.v2_g.N = L.v2_off; u.v2_g.K = u.ctx_dim; u.v2_g.taps = 1;
    u.v2_g.w = L.dalloc<bf16>(static_cast<size_t>(L.v2_off) * u.ctx_dim);
    float* vb = L.dalloc<float>(L.v2_off);
    u.v2_g.bias = vb;
    if (u.v2_g.w && vb) {
      int off = 0;
      for (auto& e : L.v2_list) {
        const int C = e.second;
        const NamedTensor* wv = L.find(e.first + "to_v.weight", static_cast<size_t>(C) * u.ctx_dim);
        const NamedTensor* wo = L.find(e.first + "to_out.0.weight", static_cast<size_t>(C) * C);
        const NamedTensor* bo = L.find(e.first + "to_out.0.bias", C);
        if (wv && wo && bo) {
          dim3 grid((u.ctx_dim + 15) / 16, (C + 15) / 16), blk(16, 16);
          matmul_f32_to_bf16_kernel<<<grid, blk, 0, st>>>(wo->ptr, C, 0, wv->ptr, u.ctx_dim,
                                                          u.v2_g.w + static_cast<size_t>(off) * u.ctx_dim, u.ctx_dim, C,
                                                          u.ctx_dim, C, 1.f);
          cudaMemcpyAsync(vb + off, bo->ptr, sizeof(float) * C, cudaMemcpyDeviceToDevice, st);
        }
        off += C;
      }
    }
  }

  // ------------------------------------------------ SpatialVolumeNet
  VolumeW& v = c.vol;
  v = VolumeW();
  const int td = mc.time_embed_dim, vdm = mc.view_dim;
  v.te0_w = L.copy_f32("time_embed.0.weight", static_cast<size_t>(td) * td);
  v.te0_b = L.copy_f32("time_embed.0.bias", td);
  v.te2_w = L.copy_f32("time_embed.2.weight", static_cast<size_t>(td) * td);
  v.te2_b = L.copy_f32("time_embed.2.bias", td);
  const std::string E = "spatial_volume.target_encoder.";
  v.enc.init_w = L.copy_f32(E + "init_conv.weight", 16 * 4 * 9);
  v.enc.init_b = L.copy_f32(E + "init_conv.bias", 16);
  for (int i = 0; i < 3; ++i) {
    const std::string q = E + "out_conv" + std::to_string(i) + ".";
    auto& R = v.enc.res[i];
    R.te_w = L.copy_f32(q + "time_embed.weight", 16 * td); R.te_b = L.copy_f32(q + "time_embed.bias", 16);
    R.ve_w = L.copy_f32(q + "view_embed.weight", 16 * vdm); R.ve_b = L.copy_f32(q + "view_embed.bias", 16);
    R.gn0_w = L.copy_f32(q + "conv.0.weight", 16); R.gn0_b = L.copy_f32(q + "conv.0.bias", 16);
    R.c0_w = L.copy_f32(q + "conv.2.weight", 16 * 16 * 9); R.c0_b = L.copy_f32(q + "conv.2.bias", 16);
    R.gn1_w = L.copy_f32(q + "conv.3.weight", 16); R.gn1_b = L.copy_f32(q + "conv.3.bias", 16);
    R.c1_w = L.copy_f32(q + "conv.5.weight", 16 * 16 * 9); R.c1_b = L.copy_f32(q + "conv.5.bias", 16);
  }
  v.enc.fgn_w = L.copy_f32(E + "final_out.0.weight", 16); v.enc.fgn_b = L.copy_f32(E + "final_out.0.bias", 16);
  v.enc.fc_w = L.copy_f32(E + "final_out.2.weight", 16 * 16 * 9); v.enc.fc_b = L.copy_f32(E + "final_out.2.bias", 16);
  v.smpl_w = L.copy_f32("spatial_volume.smpl_feature_extractor.conv0.weight", 16 * 16);
  v.smpl_b = L.copy_f32("spatial_volume.smpl_feature_extractor.conv0.bias", 16);
  {  // sparse conv net: spconv layout [O][27][I] -> fp32 [27][I][O], BN folded (eps 1e-3)
    const char* names[9] = {"conv0.0", "conv0.3", "down0.0", "conv1.0", "conv1.3", "down1.0", "conv2.0", "conv2.3", "conv2.6"};
    const int cio[9][2] = {{16, 16}, {16, 16}, {16, 32}, {32, 32}, {32, 32}, {32, 64}, {64, 64}, {64, 64}, {64, 64}};
    for (int i = 0; i < 9; ++i) {
      const std::string base = "spatial_volume.xyzc_net.";
      std::string n = names[i];
      const int ci = cio[i][0], co = cio[i][1];
      SparseLayerW& s = v.sp[i];
      s.cin = ci; s.cout = co;
      const NamedTensor* t = L.find(base + n + ".weight", static_cast<size_t>(co) * 27 * ci);
      s.w = L.dalloc<float>(static_cast<size_t>(co) * 27 * ci);
      if (t && s.w)
        L.pack<float>(t->ptr, s.w, co, 27, ci, static_cast<long>(27) * ci, ci, 1, Loader::iota(), 1,
                      static_cast<long>(ci) * co, co);
      // BatchNorm sits right after the conv in the SparseSequential: index + 1
      const size_t dot = n.rfind('.');
      const std::string bn = base + n.substr(0, dot + 1) + std::to_string(std::stoi(n.substr(dot + 1)) + 1);
      const NamedTensor* g = L.find(bn + ".weight", co);
      const NamedTensor* b = L.find(bn + ".bias", co);
      const NamedTensor* m = L.find(bn + ".running_mean", co);
      const NamedTensor* var = L.find(bn + ".running_var", co);
      s.scale = L.dalloc<float>(co); s.shift = L.dalloc<float>(co);
      if (g && b && m && var && s.scale && s.shift)
        bn_fold_kernel<<<1, 64, 0, st>>>(g->ptr, b->ptr, m->ptr, var->ptr, 1e-3f, s.scale, s.shift, co);
    }
  }
  {  // FrustumTV3DNet
    const std::string F = "spatial_volume.frustum_volume_feats.";
    const int d0 = vd[0], d1 = vd[1], d2 = vd[2], d3 = vd[3];
    v.fr.conv0 = L.gemm(F + "conv0", d0, 64, 27, true);
    const char* bn[9] = {"conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "up0", "up1", "up2"};
    const int io[9][3] = {{d0, d1, 2}, {d1, d1, 1}, {d1, d2, 2}, {d2, d2, 1}, {d2, d3, 2}, {d3, d3, 1},
                          {d3, d2, 0}, {d2, d1, 0}, {d1, d0, 0}};
    for (int i = 0; i < 9; ++i) {
      FrBlockW& b = v.fr.blk[i];
      const std::string q = F + bn[i] + ".";
      b.cin = io[i][0]; b.cout = io[i][1]; b.stride = io[i][2] ? io[i][2] : 2; b.up = io[i][2] == 0;
      b.t_w = L.copy_f32(q + "t_conv.weight", static_cast<size_t>(b.cin) * td); b.t_b = L.copy_f32(q + "t_conv.bias", b.cin);
      b.v_w = L.copy_f32(q + "v_conv.weight", static_cast<size_t>(b.cin) * vdm); b.v_b = L.copy_f32(q + "v_conv.bias", b.cin);
      b.gn = L.norm(q + (b.up ? "norm" : "bn"), b.cin);
      if (!b.up) {
        b.conv = L.gemm(q + "conv", b.cout, b.cin, 27, true);
      } else {
        // ConvTranspose3d weight [I][O][kz][ky][kx]; out[2j+p] gets, per dim: p=0 -> (k=1, in j); p=1 -> (k=0, in j+1), (k=2, in j)
        const NamedTensor* t = L.find(q + "conv.weight", static_cast<size_t>(b.cin) * b.cout * 27);
        const float* bias = L.copy_f32(q + "conv.bias", b.cout);
        for (int cls = 0; cls < 8; ++cls) {
          const int pz = (cls >> 2) & 1, py = (cls >> 1) & 1, px = cls & 1;
          TapList tl{};
          int nt = 0;
          for (int az = 0; az <= pz; ++az)
            for (int ay = 0; ay <= py; ++ay)
              for (int ax = 0; ax <= px; ++ax) {
                // a = 0 -> first tap of this parity, a = 1 -> second tap (only when p = 1)
                const int kz = pz ? (az ? 2 : 0) : 1, ky = py ? (ay ? 2 : 0) : 1, kx = px ? (ax ? 2 : 0) : 1;
                tl.t[nt++] = (kz * 3 + ky) * 3 + kx;
              }
          GemmW& g = b.upc[cls];
          g.N = b.cout; g.K = b.cin; g.taps = nt; g.bias = bias;
          g.w = L.dalloc<bf16>(static_cast<size_t>(b.cout) * nt * b.cin);
          if (t && g.w)
            L.pack<bf16>(t->ptr, g.w, b.cout, nt, b.cin, 27, 1, static_cast<long>(b.cout) * 27, tl,
                         static_cast<long>(nt) * b.cin, b.cin, 1);
        }
      }
    }
  }
  // ------------------------------------------------ first-stage decoder (optional: present when the caller's state
  // dict carries first_stage_model.*, generate_face.py:75-76 loads it with the rest of the checkpoint)
  c.vae = VaeW();
  const std::string FS = "first_stage_model.";
  if (L.has(FS + "decoder.conv_in.weight")) {
    VaeW& a = c.vae;
    const std::string Dp = FS + "decoder.";
    const int ch = 128, nlev = 4, nres = 2;                 // morphable_diffusion.py:399-414
    const int ch_mult[4] = {1, 2, 4, 4};
    int block_in = ch * ch_mult[nlev - 1];
    {  // post_quant_conv: fp32 [4][4] | [4]
      const NamedTensor* w = L.find(FS + "post_quant_conv.weight", 16);
      const NamedTensor* b = L.find(FS + "post_quant_conv.bias", 4);
      a.pq = L.dalloc<float>(20);
      if (w && b && a.pq) {
        cudaMemcpyAsync(a.pq, w->ptr, 16 * sizeof(float), cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(a.pq + 16, b->ptr, 4 * sizeof(float), cudaMemcpyDeviceToDevice, st);
      }
    }
    a.conv_in = L.conv_padded_in(Dp + "conv_in", block_in, 4);   // 4 -> 512
    a.mid1 = L.vae_res(Dp + "mid.block_1.", block_in, block_in);
    a.mid2 = L.vae_res(Dp + "mid.block_2.", block_in, block_in);
    a.attn = L.vae_attn(Dp + "mid.attn_1.", block_in);
    a.up.assign(nlev, std::vector<ResW>());
    for (int lev = nlev - 1; lev >= 0; --lev) {
      const int block_out = ch * ch_mult[lev];
      for (int ib = 0; ib < nres + 1; ++ib) {
        a.up[lev].push_back(L.vae_res(Dp + "up." + std::to_string(lev) + ".block." + std::to_string(ib) + ".", block_in, block_out));
        block_in = block_out;
      }
      if (lev != 0) a.upsample[lev] = L.gemm(Dp + "up." + std::to_string(lev) + ".upsample.conv", block_in, block_in, 9, true);
    }
    a.norm_out = L.norm(Dp + "norm_out", block_in);
    {  // conv_out 128 -> 3, padded to 8 output columns
      const NamedTensor* t = L.find(Dp + "conv_out.weight", static_cast<size_t>(3) * block_in * 9);
      const NamedTensor* b = L.find(Dp + "conv_out.bias", 3);
      a.conv_out.N = 8; a.conv_out.K = block_in; a.conv_out.taps = 9;
      a.conv_out.w = L.dalloc<bf16>(static_cast<size_t>(8) * block_in * 9);
      float* ob = L.dalloc<float>(8);
      a.conv_out.bias = ob;
      if (t && b && a.conv_out.w && ob) {
        cudaMemsetAsync(a.conv_out.w, 0, sizeof(bf16) * 8 * block_in * 9, st);
        cudaMemsetAsync(ob, 0, sizeof(float) * 8, st);
        L.pack<bf16>(t->ptr, a.conv_out.w, 3, 9, block_in, static_cast<long>(block_in) * 9, 1, 9, Loader::iota(),
                     static_cast<long>(9) * block_in, block_in, 1);
        cudaMemcpyAsync(ob, b->ptr, sizeof(float) * 3, cudaMemcpyDeviceToDevice, st);
      }
    }
    a.loaded = L.rc == 0;
  }
  // ------------------------------------------------ first-stage encoder (optional, SURVEY §8f rank 2: the VAE half of prepare())
  c.vae_enc = VaeEncW();
  if (L.has(FS + "encoder.conv_in.weight")) {
    VaeEncW& e = c.vae_enc;
    const std::string Ep = FS + "encoder.";
    const int ch = 128, nlev = 4, nres = 2;
    const int ch_mult[4] = {1, 2, 4, 4}, in_mult[5] = {1, 1, 2, 4, 4};
    e.conv_in = L.conv_padded_in(Ep + "conv_in", ch, 3);
    int block_in = ch;
    e.down.assign(nlev, std::vector<ResW>());
    for (int lev = 0; lev < nlev; ++lev) {
      block_in = ch * in_mult[lev];
      const int block_out = ch * ch_mult[lev];
      for (int ib = 0; ib < nres; ++ib) {
        e.down[lev].push_back(L.vae_res(Ep + "down." + std::to_string(lev) + ".block." + std::to_string(ib) + ".", block_in, block_out));
        block_in = block_out;
      }
      if (lev != nlev - 1) e.downsample[lev] = L.gemm(Ep + "down." + std::to_string(lev) + ".downsample.conv", block_in, block_in, 9, true);
    }
    e.mid1 = L.vae_res(Ep + "mid.block_1.", block_in, block_in);
    e.attn = L.vae_attn(Ep + "mid.attn_1.", block_in);
    e.mid2 = L.vae_res(Ep + "mid.block_2.", block_in, block_in);
    e.norm_out = L.norm(Ep + "norm_out", block_in);
    {  // quant_conv (1x1, 8 -> 8) folded into conv_out (3x3, 512 -> 8): W' = Wq . Wc, b' = Wq bc + bq
      const NamedTensor* wc = L.find(Ep + "conv_out.weight", static_cast<size_t>(8) * block_in * 9);
      const NamedTensor* bc = L.find(Ep + "conv_out.bias", 8);
      const NamedTensor* wq = L.find(FS + "quant_conv.weight", 64);
      const NamedTensor* bq = L.find(FS + "quant_conv.bias", 8);
      const size_t kk = static_cast<size_t>(9) * block_in;
      float* wc_packed = L.dalloc<float>(8 * kk);   // [j][tap][ci]
      e.conv_out.N = 8; e.conv_out.K = block_in; e.conv_out.taps = 9;
      e.conv_out.w = L.dalloc<bf16>(8 * kk);
      float* ob = L.dalloc<float>(8);
      e.conv_out.bias = ob;
      if (wc && bc && wq && bq && wc_packed && e.conv_out.w && ob) {
        L.pack<float>(wc->ptr, wc_packed, 8, 9, block_in, static_cast<long>(block_in) * 9, 1, 9, Loader::iota(),
                      static_cast<long>(kk), block_in, 1);
        dim3 grid(static_cast<unsigned>((kk + 15) / 16), 1), blk(16, 16);
        matmul_f32_to_bf16_kernel<<<grid, blk, 0, st>>>(wq->ptr, 8, 0, wc_packed, static_cast<int>(kk), e.conv_out.w,
                                                        static_cast<int>(kk), 8, static_cast<int>(kk), 8, 1.f);
        fold_bias_kernel<<<1, 32, 0, st>>>(wq->ptr, bc->ptr, bq->ptr, ob, 8);
      }
    }
    e.loaded = L.rc == 0;
  }
  // ------------------------------------------------ CLIP image tower (optional, SURVEY §8f rank 2: the CLIP half of prepare())
  c.clip = ClipW();
  const std::string CV = "clip_image_encoder.model.visual.";
  if (L.has(CV + "conv1.weight")) {
    ClipW& q = c.clip;
    const int Wd = q.width;
    {  // conv1 [width][3][14][14] -> bf16 [width][640]: the torch flattening (c, ky, kx) is already the K order; pad 588 -> 640
      const NamedTensor* t = L.find(CV + "conv1.weight", static_cast<size_t>(Wd) * 588);
      q.conv1.N = Wd; q.conv1.K = 640; q.conv1.taps = 1;
      q.conv1.w = L.dalloc<bf16>(static_cast<size_t>(Wd) * 640);
      if (t && q.conv1.w) {
        cudaMemsetAsync(q.conv1.w, 0, sizeof(bf16) * Wd * 640, st);
        L.pack<bf16>(t->ptr, q.conv1.w, Wd, 1, 588, 588, 0, 1, Loader::iota(), 640, 0, 1);
      }
    }
    q.class_emb = L.copy_f32(CV + "class_embedding", Wd);
    q.pos_emb = L.copy_f32(CV + "positional_embedding", static_cast<size_t>(q.ntok) * Wd);
    q.ln_pre = L.norm(CV + "ln_pre", Wd);
    q.ln_post = L.norm(CV + "ln_post", Wd);
    int n_layers = 0;
    while (L.has(CV + "transformer.resblocks." + std::to_string(n_layers) + ".ln_1.weight")) ++n_layers;
    q.layers.resize(n_layers);
    for (int i = 0; i < n_layers; ++i) {
      const std::string b = CV + "transformer.resblocks." + std::to_string(i) + ".";
      ClipLayerW& l = q.layers[i];
      l.ln1 = L.norm(b + "ln_1", Wd);
      l.ln2 = L.norm(b + "ln_2", Wd);
      l.qkv.N = 3 * Wd; l.qkv.K = Wd; l.qkv.taps = 1;
      l.qkv.w = L.dalloc<bf16>(static_cast<size_t>(3) * Wd * Wd);
      if (l.qkv.w) L.pack_conv_into(l.qkv.w, b + "attn.in_proj_weight", 3 * Wd, Wd, 1);
      l.qkv.bias = L.copy_f32(b + "attn.in_proj_bias", 3 * Wd);
      l.out = L.gemm(b + "attn.out_proj", Wd, Wd, 1, true);
      l.fc = L.gemm(b + "mlp.c_fc", 4 * Wd, Wd, 1, true);
      l.proj = L.gemm(b + "mlp.c_proj", Wd, 4 * Wd, 1, true);
    }
    {  // proj [width][out_dim] -> GEMM weight [out_dim][width] (transposed)
      const NamedTensor* t = L.find(CV + "proj", static_cast<size_t>(Wd) * q.out_dim);
      q.proj.N = q.out_dim; q.proj.K = Wd; q.proj.taps = 1;
      q.proj.w = L.dalloc<bf16>(static_cast<size_t>(q.out_dim) * Wd);
      if (t && q.proj.w) L.pack<bf16>(t->ptr, q.proj.w, q.out_dim, 1, Wd, 1, 0, q.out_dim, Loader::iota(), Wd, 0, 1);
    }
    q.loaded = L.rc == 0 && n_layers > 0;
  }
  if (L.rc != 0) { free_weights(c); return L.rc; }
  MD_CUDA(cudaStreamSynchronize(st));
  c.weights_loaded = true;
  return 0;
}

void free_weights(Ctx& c) {
  for (void* p : c.weight_allocs) cudaFree(p);
  c.weight_allocs.clear();
  c.weights_loaded = false;
  c.vae.loaded = false;
  c.vae_enc.loaded = false;
  c.clip.loaded = false;
}

}  // namespace md
