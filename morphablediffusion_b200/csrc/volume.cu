// SpatialVolumeNet on the GPU: per-sample binding of step invariants (camera matrices, frustum sample points,
// sparse-conv rulebook, resample table) and the per-step construction of the spatial volume and of the per-view
// frustum feature pyramids (morphable_diffusion.py:182-320).
#include "engine.h"

#include <math.h>

namespace md {

namespace {

void mat3x3_mul(const float* a, const float* b, float* o) {  // o = a(3x3) * b(3x3)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s += a[i * 3 + k] * b[k * 3 + j];
      o[i * 3 + j] = s;
    }
}

bool invert4(const double* m, double* inv) {
  double a[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { a[i][j] = m[i * 4 + j]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    for (int r = col + 1; r < 4; ++r) if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
    if (fabs(a[piv][col]) < 1e-300) return false;
    if (piv != col) for (int j = 0; j < 8; ++j) std::swap(a[piv][j], a[col][j]);
    const double d = a[col][col];
    for (int j = 0; j < 8; ++j) a[col][j] /= d;
    for (int r = 0; r < 4; ++r) if (r != col) {
      const double f = a[r][col];
      for (int j = 0; j < 8; ++j) a[r][j] -= f * a[col][j];
    }
  }
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) inv[i * 4 + j] = a[i][4 + j];
  return true;
}

float linspace_host(float length, int V, int i) {
  const float step = (2.f * length) / static_cast<float>(V - 1);
  return (i < V / 2) ? (-length + step * i) : (length - step * (V - 1 - i));
}

template <typename T>
int upload(SampleBinding& sb, const std::vector<T>& h, T** d) {
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(h.size(), 1) * sizeof(T)) != cudaSuccess) return set_error("bind_sample: cudaMalloc failed");
  sb.owned.push_back(p);
  if (!h.empty() && cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
    return set_error("bind_sample: upload failed");
  *d = static_cast<T*>(p);
  return 0;
}

struct Level {
  int D, H, W;
  std::vector<int64_t> lin;                      // sorted linear indices of active voxels
  std::unordered_map<int64_t, int> row;          // lin -> row
  int find(int d, int h, int w) const {
    if (d < 0 || d >= D || h < 0 || h >= H || w < 0 || w >= W) return -1;
    auto it = row.find((static_cast<int64_t>(d) * H + h) * W + w);
    return it == row.end() ? -1 : it->second;
  }
  void finalize() {
    std::sort(lin.begin(), lin.end());
    lin.erase(std::unique(lin.begin(), lin.end()), lin.end());
    row.clear();
    row.reserve(lin.size() * 2);
    for (size_t i = 0; i < lin.size(); ++i) row[lin[i]] = static_cast<int>(i);
  }
  void coords(size_t r, int& d, int& h, int& w) const {
    int64_t l = lin[r];
    w = static_cast<int>(l % W); l /= W;
    h = static_cast<int>(l % H);
    d = static_cast<int>(l / H);
  }
};

// SubMConv3d rulebook: nbr[r][k] = row of (p + k - 1) at the same level
std::vector<int32_t> subm_rules(const Level& L) {
  std::vector<int32_t> nbr(L.lin.size() * 27);
  for (size_t r = 0; r < L.lin.size(); ++r) {
    int d, h, w;
    L.coords(r, d, h, w);
    for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw)
      nbr[r * 27 + (kd * 3 + kh) * 3 + kw] = L.find(d + kd - 1, h + kh - 1, w + kw - 1);
  }
  return nbr;
}

// SparseConv3d(k3,s2,p1): output o is active iff some input 2o-1+k is active; nbr[r_out][k] = input row
Level down_level(const Level& in) {
  Level out;
  out.D = (in.D + 2 - 3) / 2 + 1; out.H = (in.H + 2 - 3) / 2 + 1; out.W = (in.W + 2 - 3) / 2 + 1;
  for (size_t r = 0; r < in.lin.size(); ++r) {
    int d, h, w;
    in.coords(r, d, h, w);
    for (int kd = 0; kd < 3; ++kd) {
      const int nd = d + 1 - kd;
      if (nd < 0 || (nd & 1) || nd / 2 >= out.D) continue;
      for (int kh = 0; kh < 3; ++kh) {
        const int nh = h + 1 - kh;
        if (nh < 0 || (nh & 1) || nh / 2 >= out.H) continue;
        for (int kw = 0; kw < 3; ++kw) {
          const int nw = w + 1 - kw;
          if (nw < 0 || (nw & 1) || nw / 2 >= out.W) continue;
          out.lin.push_back((static_cast<int64_t>(nd / 2) * out.H + nh / 2) * out.W + nw / 2);
        }
      }
    }
  }
  out.finalize();
  return out;
}
std::vector<int32_t> down_rules(const Level& in, const Level& out) {
  std::vector<int32_t> nbr(out.lin.size() * 27);
  for (size_t r = 0; r < out.lin.size(); ++r) {
    int d, h, w;
    out.coords(r, d, h, w);
    for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw)
      nbr[r * 27 + (kd * 3 + kh) * 3 + kw] = in.find(2 * d - 1 + kd, 2 * h - 1 + kh, 2 * w - 1 + kw);
  }
  return nbr;
}

}  // namespace

void free_binding(Ctx& c) {
  for (void* p : c.sb.owned) cudaFree(p);
  c.sb = SampleBinding();
}

int bind_sample(Ctx& c, const float* K, const float* RT, const float* v_embed, const float* vertices,
                const int32_t* coord, const int32_t* out_sh, const float* bounds, int nv, int n_views, int view0,
                int n_local, int ortho, cudaStream_t st) {
  free_binding(c);
  SampleBinding& sb = c.sb;
  const md_config& mc = c.mcfg;
  sb.n_views = n_views; sb.view0 = view0; sb.n_local = n_local; sb.nv = nv; sb.ortho = ortho;
  if (view0 < 0 || n_local < 1 || view0 + n_local > n_views) return set_error("bind_sample: bad view range");
  const int S = mc.latent_size;
  const float ratio = static_cast<float>(S) / static_cast<float>(mc.image_size);

  std::vector<float> proj(static_cast<size_t>(n_views) * 12), cam(static_cast<size_t>(n_views) * 24, 0.f);
  for (int n = 0; n < n_views; ++n) {
    const float* Kn = K + n * 16;
    const float* Rn = RT + n * 12;
    double P4[16], inv[16];
    if (!ortho) {
      // construct_project_matrix (utils.py:46-69): diag(r,r,1) @ K[:3,:3] @ RT
      const float sc[9] = {ratio, 0, 0, 0, ratio, 0, 0, 0, 1};
      const float K3[9] = {Kn[0], Kn[1], Kn[2], Kn[4], Kn[5], Kn[6], Kn[8], Kn[9], Kn[10]};
      float sk[9];
      mat3x3_mul(sc, K3, sk);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
          float s = 0.f;
          for (int k = 0; k < 3; ++k) s += sk[i * 3 + k] * Rn[k * 4 + j];
          proj[n * 12 + i * 4 + j] = s;
          P4[i * 4 + j] = s;
        }
      P4[12] = 0; P4[13] = 0; P4[14] = 0; P4[15] = 1;
      if (!invert4(P4, inv)) return set_error("bind_sample: singular projection for view %d", n);
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) cam[n * 24 + i * 3 + j] = static_cast<float>(inv[i * 4 + j]);
        cam[n * 24 + 9 + i] = static_cast<float>(inv[i * 4 + 3]);
      }
    } else {
      // K(4x4) @ [RT; 0 0 0 1]
      float RT4[16];
      for (int i = 0; i < 12; ++i) RT4[i] = Rn[i];
      RT4[12] = 0; RT4[13] = 0; RT4[14] = 0; RT4[15] = 1;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
          float s = 0.f;
          for (int k = 0; k < 4; ++k) s += Kn[i * 4 + k] * RT4[k * 4 + j];
          proj[n * 12 + i * 4 + j] = s;
        }
      double Kd[16], Kinv[16], Rd[16];
      for (int i = 0; i < 16; ++i) { Kd[i] = Kn[i]; Rd[i] = RT4[i]; }
      if (!invert4(Kd, Kinv) || !invert4(Rd, inv)) return set_error("bind_sample: singular camera for view %d", n);
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) cam[n * 24 + i * 3 + j] = static_cast<float>(inv[i * 4 + j]);
        cam[n * 24 + 9 + i] = static_cast<float>(inv[i * 4 + 3]);
      }
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) cam[n * 24 + 13 + i * 3 + j] = static_cast<float>(Kinv[i * 4 + j]);
    }
    // camera distance: || -R^T t ||  (morphable_diffusion.py:281-284)
    float cp[3];
    for (int i = 0; i < 3; ++i) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s += -Rn[k * 4 + i] * Rn[k * 4 + 3];
      cp[i] = s;
    }
    cam[n * 24 + 12] = sqrtf(cp[0] * cp[0] + cp[1] * cp[1] + cp[2] * cp[2]);
  }
  MD_CHECK(upload(sb, proj, &sb.proj));
  MD_CHECK(upload(sb, cam, &sb.cam));
  MD_CHECK(upload(sb, std::vector<float>(v_embed, v_embed + static_cast<size_t>(n_views) * mc.view_dim), &sb.v_embed));
  MD_CHECK(upload(sb, std::vector<float>(vertices, vertices + static_cast<size_t>(nv) * 3), &sb.vertices));

  // ---- frustum sample points for the local views
  {
    const size_t per_view = static_cast<size_t>(mc.frustum_depth) * S * S;
    void* p = nullptr;
    if (cudaMalloc(&p, per_view * n_local * 3 * sizeof(float)) != cudaSuccess) return set_error("bind_sample: cudaMalloc(pts)");
    sb.owned.push_back(p);
    sb.pts = static_cast<float*>(p);
    MD_CHECK(launch_frustum_points(sb.cam + static_cast<size_t>(view0) * 24, ortho, mc.frustum_depth, S,
                                   mc.spatial_volume_length, mc.frustum_volume_length, sb.pts, n_local, st));
  }

  // ---- sparse-conv rulebook (SparseConvTensor + SubMConv3d/SparseConv3d index generation, done once per mesh)
  Level L0;
  L0.D = out_sh[0]; L0.H = out_sh[1]; L0.W = out_sh[2];
  std::unordered_map<int64_t, int> first_vertex;
  first_vertex.reserve(static_cast<size_t>(nv) * 2);
  for (int i = 0; i < nv; ++i) {
    const int d = coord[i * 3], h = coord[i * 3 + 1], w = coord[i * 3 + 2];
    if (d < 0 || d >= L0.D || h < 0 || h >= L0.H || w < 0 || w >= L0.W)
      return set_error("bind_sample: coord[%d]=(%d,%d,%d) outside out_sh=(%d,%d,%d)", i, d, h, w, L0.D, L0.H, L0.W);
    const int64_t lin = (static_cast<int64_t>(d) * L0.H + h) * L0.W + w;
    if (first_vertex.find(lin) == first_vertex.end()) first_vertex[lin] = i;  // lowest vertex index wins
    L0.lin.push_back(lin);
  }
  L0.finalize();
  std::vector<int32_t> row_vertex(L0.lin.size());
  for (size_t r = 0; r < L0.lin.size(); ++r) row_vertex[r] = first_vertex[L0.lin[r]];
  const Level L1 = down_level(L0);
  const Level L2 = down_level(L1);
  sb.n0 = static_cast<int>(L0.lin.size()); sb.n1 = static_cast<int>(L1.lin.size()); sb.n2 = static_cast<int>(L2.lin.size());
  MD_CHECK(upload(sb, row_vertex, &sb.row_vertex));
  MD_CHECK(upload(sb, subm_rules(L0), &sb.nbr[0]));
  MD_CHECK(upload(sb, down_rules(L0, L1), &sb.nbr[1]));
  MD_CHECK(upload(sb, subm_rules(L1), &sb.nbr[2]));
  MD_CHECK(upload(sb, down_rules(L1, L2), &sb.nbr[3]));
  MD_CHECK(upload(sb, subm_rules(L2), &sb.nbr[4]));

  // ---- resample table: world grid -> dense sparse-conv output (morphable_diffusion.py:234-243,255)
  {
    const int V = mc.spatial_volume_size;
    std::vector<int32_t> idx(static_cast<size_t>(V) * V * V * 8);
    std::vector<float> wgt(idx.size());
    const float mn[3] = {bounds[2], bounds[1], bounds[0]};  // min_dhw = bounds[0][[2,1,0]]
    const int dims[3] = {L2.D, L2.H, L2.W};
    const float voxel = 0.005f;
    for (int i = 0; i < V; ++i) for (int j = 0; j < V; ++j) for (int k = 0; k < V; ++k) {
      const float world_dhw[3] = {linspace_host(mc.spatial_volume_length, V, i), linspace_host(mc.spatial_volume_length, V, j),
                                  linspace_host(mc.spatial_volume_length, V, k)};
      float pos[3], frac[3];
      int base[3];
      for (int a = 0; a < 3; ++a) {
        float g = (world_dhw[a] - mn[a]) / voxel;
        g = g / static_cast<float>(out_sh[a]) * 2.f - 1.f;
        pos[a] = ((g + 1.f) / 2.f) * static_cast<float>(dims[a] - 1);
        const float fl = floorf(pos[a]);
        frac[a] = pos[a] - fl;
        base[a] = static_cast<int>(fminf(fmaxf(fl, -2.f), static_cast<float>(dims[a])));
      }
      const size_t p = (static_cast<size_t>(i) * V + j) * V + k;
      for (int cdx = 0; cdx < 8; ++cdx) {
        const int dz = cdx >> 2, dy = (cdx >> 1) & 1, dx = cdx & 1;
        const int d = base[0] + dz, h = base[1] + dy, w = base[2] + dx;
        idx[p * 8 + cdx] = L2.find(d, h, w);
        wgt[p * 8 + cdx] = (dz ? frac[0] : 1.f - frac[0]) * (dy ? frac[1] : 1.f - frac[1]) * (dx ? frac[2] : 1.f - frac[2]);
      }
    }
    MD_CHECK(upload(sb, idx, &sb.rs_idx));
    MD_CHECK(upload(sb, wgt, &sb.rs_w));
  }
  MD_CUDA(cudaStreamSynchronize(st));
  sb.bound = true;
  return 0;
}

// embed_time (morphable_diffusion.py:491-494): sinusoid(256) -> Linear -> SiLU -> Linear, one timestep
int embed_time(Ctx& c, const float* t_dev, float* t_embed, cudaStream_t st) {
  const int td = c.mcfg.time_embed_dim;
  Arena& A = c.arena;
  const size_t m = A.mark();
  float* s = A.get<float>(td);
  float* h = A.get<float>(td);
  if (A.failed) return set_error("workspace exhausted (embed_time)");
  MD_CHECK(launch_timestep_embedding(t_dev, s, 1, td, st));
  MD_CHECK(launch_small_linear(s, td, c.vol.te0_w, c.vol.te0_b, h, td, 1, td, td, ACT_NONE, ACT_SILU, 0, st));
  MD_CHECK(launch_small_linear(h, td, c.vol.te2_w, c.vol.te2_b, t_embed, td, 1, td, td, ACT_NONE, ACT_NONE, 0, st));
  A.release(m);
  return 0;
}

// K1 + fused K2/K3 for the local views: x_local fp32 NCHW [n_local][4][S][S] -> vsum [Nv][16] (sum over local views)
int vertex_feature_sum(Ctx& c, const float* x_local, const float* t_embed, float* vsum, cudaStream_t st) {
  const SampleBinding& sb = c.sb;
  const md_config& mc = c.mcfg;
  Arena& A = c.arena;
  const size_t m = A.mark();
  float* feats = A.get<float>(static_cast<size_t>(sb.n_local) * mc.latent_size * mc.latent_size * 16);
  if (A.failed) return set_error("workspace exhausted (encoder)");
  MD_CHECK(launch_target_encoder(x_local, t_embed, sb.v_embed + static_cast<size_t>(sb.view0) * mc.view_dim, c.vol.enc,
                                 feats, sb.n_local, mc.time_embed_dim, mc.view_dim, mc.latent_size, st));
  MD_CHECK(launch_vertex_features(feats, sb.proj + static_cast<size_t>(sb.view0) * 12, sb.ortho, mc.latent_size,
                                  mc.spatial_volume_size, mc.spatial_volume_length, sb.vertices, sb.nv, sb.n_local,
                                  vsum, st));
  A.release(m);
  return 0;
}

// K4-K6: view-mean + Conv1d, sparse conv net, resample -> vol fp32 [V][V][V][64]
// peer_exchange: vsum holds this rank's partial sums only; they are pushed to every rank over NVLink and the scatter
// kernel adds the world's contributions (otherwise vsum is already complete: single rank, or all-reduced by the caller)
int spatial_volume_from_vsum(Ctx& c, const float* vsum, float* vol, cudaStream_t st, bool peer_exchange) {
  const SampleBinding& sb = c.sb;
  const md_config& mc = c.mcfg;
  Arena& A = c.arena;
  const size_t m = A.mark();
  const int nmax = std::max(sb.n0, std::max(sb.n1, sb.n2));
  float* bufA = A.get<float>(static_cast<size_t>(nmax) * 64);
  float* bufB = A.get<float>(static_cast<size_t>(nmax) * 64);
  if (A.failed) return set_error("workspace exhausted (sparse conv)");
  const int ninv = mc.smpl_num_views > 0 ? mc.smpl_num_views : sb.n_views;
  if (peer_exchange) {
    const Ctx::PeerExchange& px = c.px;
    const size_t flags_off = static_cast<size_t>(2) * c.world * px.slot_floats;
    MD_CHECK(launch_peer_push(vsum, sb.nv * 16, px.d_peer_data, px.d_peer_flags, px.d_seq, c.rank, c.world, px.slot_floats, st));
    MD_CHECK(launch_smpl_scatter_peer(px.base, reinterpret_cast<const unsigned*>(px.base + flags_off), px.d_seq, c.world,
                                      px.slot_floats, 1.f / static_cast<float>(ninv), c.vol.smpl_w, c.vol.smpl_b,
                                      sb.row_vertex, sb.n0, bufA, px.d_err, st));
  } else {
    MD_CHECK(launch_smpl_scatter(vsum, 1.f / static_cast<float>(ninv), c.vol.smpl_w, c.vol.smpl_b, sb.row_vertex, sb.n0,
                                 bufA, st));
  }
  const int rows[9] = {sb.n0, sb.n0, sb.n1, sb.n1, sb.n1, sb.n2, sb.n2, sb.n2, sb.n2};
  const int rule[9] = {0, 0, 1, 2, 2, 3, 4, 4, 4};
  float* in = bufA;
  float* out = bufB;
  for (int i = 0; i < 9; ++i) {
    const SparseLayerW& s = c.vol.sp[i];
    MD_CHECK(launch_sparse_conv(in, sb.nbr[rule[i]], s.w, s.scale, s.shift, out, rows[i], s.cin, s.cout, st));
    std::swap(in, out);
  }
  const int V = mc.spatial_volume_size;
  MD_CHECK(launch_volume_resample(in, sb.rs_idx, sb.rs_w, vol, V * V * V, st));
  A.release(m);
  return 0;
}

// K7-K9 for T local views starting at local index lv0.  levels[i] (bf16, channels-last) are allocated from the arena
// with room for `alloc_samples` >= T samples; the samples beyond T are zero-filled (the CFG-unconditional half).
int frustum_levels(Ctx& c, const float* vol, int lv0, int T, const float* t_embed, int alloc_samples, bf16* levels[4],
                   cudaStream_t st) {
  const SampleBinding& sb = c.sb;
  const md_config& mc = c.mcfg;
  const FrustumW& F = c.vol.fr;
  Arena& A = c.arena;
  const int S = mc.latent_size, D = mc.frustum_depth, V = mc.spatial_volume_size;
  const int td = mc.time_embed_dim, vdm = mc.view_dim;
  if (lv0 < 0 || lv0 + T > sb.n_local) return set_error("frustum_levels: view range outside the binding");
  const int* vd = mc.volume_dims;
  size_t lrows[4];
  int lD[4], lS[4];
  for (int i = 0; i < 4; ++i) {
    lD[i] = D >> i; lS[i] = S >> i;
    lrows[i] = static_cast<size_t>(lD[i]) * lS[i] * lS[i];
    levels[i] = A.get<bf16>(static_cast<size_t>(alloc_samples) * lrows[i] * vd[i]);
  }
  if (A.failed) return set_error("workspace exhausted (frustum levels)");
  for (int i = 0; i < 4; ++i)
    if (alloc_samples > T)
      MD_CUDA(cudaMemsetAsync(levels[i] + static_cast<size_t>(T) * lrows[i] * vd[i], 0,
                              static_cast<size_t>(alloc_samples - T) * lrows[i] * vd[i] * sizeof(bf16), st));
  const size_t m = A.mark();
  const float* v_emb = sb.v_embed + static_cast<size_t>(sb.view0 + lv0) * vdm;

  auto taps3d = [](md_conv_gemm_args& a) {
    a.ntaps = 27;
    for (int kz = 0; kz < 3; ++kz) for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) {
      const int t = (kz * 3 + ky) * 3 + kx;
      a.tap[t][0] = kx - 1; a.tap[t][1] = ky - 1; a.tap[t][2] = kz - 1;
    }
  };
  // GroupNorm statistics of every tensor that feeds a FrustumTV(Up)Block are accumulated by the producing GEMM
  const int stat_ch = vd[0] + 2 * vd[1] + 2 * vd[2] + 2 * vd[3] + vd[2] + vd[1];
  float* fpool = A.get<float>(static_cast<size_t>(T) * stat_ch * 2);
  if (A.failed) return set_error("workspace exhausted (frustum statistics)");
  MD_CUDA(cudaMemsetAsync(fpool, 0, static_cast<size_t>(T) * stat_ch * 2 * sizeof(float), st));
  size_t fpool_off = 0;
  auto new_stats = [&](int C) {
    float* p = fpool + fpool_off;
    fpool_off += static_cast<size_t>(T) * C * 2;
    return p;
  };
  auto conv3d = [&](const bf16* in, int d, int s, const GemmW& w, bf16* out, float* stats) {
    md_conv_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = in; a.B = T; a.D = d; a.H = s; a.W = s; a.Cin = w.K; a.Wt = w.w; a.N = w.N;
    taps3d(a);
    a.bias = w.bias; a.out_bf16 = out; a.col_stats = stats;
    return launch_conv_gemm(a, st);
  };
  // x + t_conv(t) + v_conv(v) -> GN(8) -> SiLU  (network.py:285-311)
  auto norm_act = [&](const FrBlockW& b, const bf16* x, const float* xstats, size_t rows, bf16* out) {
    float* tv = A.get<float>(static_cast<size_t>(T) * b.cin);
    float* ss = A.get<float>(static_cast<size_t>(T) * b.cin * 2);
    if (A.failed) return set_error("workspace exhausted (frustum norm)");
    MD_CHECK(launch_small_linear(v_emb, vdm, b.v_w, b.v_b, tv, b.cin, T, vdm, b.cin, ACT_NONE, ACT_NONE, 0, st));
    MD_CHECK(launch_small_linear(t_embed, 0, b.t_w, b.t_b, tv, b.cin, T, td, b.cin, ACT_NONE, ACT_NONE, 1, st));
    GroupNormArgs g;
    memset(&g, 0, sizeof(g));
    g.x0 = x; g.C0 = b.cin; g.x0_bf16 = 1; g.B = T; g.rows = static_cast<int>(rows); g.groups = 8; g.eps = 1e-5f;
    g.gamma = b.gn.g; g.beta = b.gn.b; g.addvec = tv; g.addvec_ld = b.cin; g.stats0 = xstats;
    g.scale_shift = ss;
    g.out = out; g.act = ACT_SILU;
    return launch_group_norm(g, st);
  };
  auto block = [&](const FrBlockW& b, const bf16* x, const float* xstats, int d, int s, bf16* out, float* ostats) {
    const size_t rows = static_cast<size_t>(d) * s * s;
    const size_t mm = A.mark();
    bf16* a = A.get<bf16>(static_cast<size_t>(T) * rows * b.cin);
    if (A.failed) return set_error("workspace exhausted (frustum block)");
    MD_CHECK(norm_act(b, x, xstats, rows, a));
    if (b.stride == 1) {
      MD_CHECK(conv3d(a, d, s, b.conv, out, ostats));
    } else {
      // Conv3d 3x3x3 stride 2 pad 1: implicit GEMM whose TMA boxes step by 2 voxels (no patch materialisation)
      md_conv_gemm_args g;
      memset(&g, 0, sizeof(g));
      g.A = a; g.B = T; g.D = d; g.H = s; g.W = s; g.Cin = b.conv.K; g.Wt = b.conv.w; g.N = b.conv.N;
      taps3d(g);
      g.in_stride[0] = g.in_stride[1] = g.in_stride[2] = 2;
      g.bias = b.conv.bias; g.out_bf16 = out; g.col_stats = ostats;
      MD_CHECK(launch_conv_gemm(g, st));
    }
    A.release(mm);
    return 0;
  };
  // ConvTranspose3d(k3,s2,p1,op1) as 8 output-parity classes + skip add
  auto up = [&](const FrBlockW& b, const bf16* x, const float* xstats, int d, int s, const bf16* skip, bf16* out,
                float* ostats) {
    const size_t rows = static_cast<size_t>(d) * s * s;
    const size_t mm = A.mark();
    bf16* a = A.get<bf16>(static_cast<size_t>(T) * rows * b.cin);
    if (A.failed) return set_error("workspace exhausted (frustum up)");
    MD_CHECK(norm_act(b, x, xstats, rows, a));
    for (int cls = 0; cls < 8; ++cls) {
      const int pz = (cls >> 2) & 1, py = (cls >> 1) & 1, px = cls & 1;
      const GemmW& w = b.upc[cls];
      md_conv_gemm_args g;
      memset(&g, 0, sizeof(g));
      g.A = a; g.B = T; g.D = d; g.H = s; g.W = s; g.Cin = w.K; g.Wt = w.w; g.N = w.N;
      int nt = 0;
      for (int az = 0; az <= pz; ++az) for (int ay = 0; ay <= py; ++ay) for (int ax = 0; ax <= px; ++ax) {
        // first tap of an odd output reads input j+1 (kernel index 0), second reads j (kernel index 2)
        g.tap[nt][0] = px ? (ax ? 0 : 1) : 0;
        g.tap[nt][1] = py ? (ay ? 0 : 1) : 0;
        g.tap[nt][2] = pz ? (az ? 0 : 1) : 0;
        ++nt;
      }
      g.ntaps = nt;
      g.OD = 2 * d; g.OH = 2 * s; g.OW = 2 * s;
      g.os[0] = g.os[1] = g.os[2] = 2;
      g.op[0] = px; g.op[1] = py; g.op[2] = pz;
      g.bias = w.bias; g.res_bf16 = skip; g.out_bf16 = out; g.col_stats = ostats;
      MD_CHECK(launch_conv_gemm(g, st));
    }
    A.release(mm);
    return 0;
  };

  const size_t per_view = lrows[0];
  bf16* fr_in = A.get<bf16>(static_cast<size_t>(T) * per_view * 64);
  bf16* x0 = A.get<bf16>(static_cast<size_t>(T) * lrows[0] * vd[0]);
  bf16* t1 = A.get<bf16>(static_cast<size_t>(T) * lrows[1] * vd[1]);
  bf16* x1 = A.get<bf16>(static_cast<size_t>(T) * lrows[1] * vd[1]);
  bf16* t2 = A.get<bf16>(static_cast<size_t>(T) * lrows[2] * vd[2]);
  bf16* x2 = A.get<bf16>(static_cast<size_t>(T) * lrows[2] * vd[2]);
  bf16* t3 = A.get<bf16>(static_cast<size_t>(T) * lrows[3] * vd[3]);
  if (A.failed) return set_error("workspace exhausted (frustum net)");
  MD_CHECK(launch_frustum_gather(vol, sb.pts + static_cast<size_t>(lv0) * per_view * 3, V, fr_in,
                                 static_cast<size_t>(T) * per_view, st));
  float* s_x0 = new_stats(vd[0]); float* s_t1 = new_stats(vd[1]); float* s_x1 = new_stats(vd[1]);
  float* s_t2 = new_stats(vd[2]); float* s_x2 = new_stats(vd[2]); float* s_t3 = new_stats(vd[3]);
  float* s_l3 = new_stats(vd[3]); float* s_l2 = new_stats(vd[2]); float* s_l1 = new_stats(vd[1]);
  MD_CHECK(conv3d(fr_in, lD[0], lS[0], F.conv0, x0, s_x0));
  MD_CHECK(block(F.blk[0], x0, s_x0, lD[0], lS[0], t1, s_t1));          // conv1 (s2)
  MD_CHECK(block(F.blk[1], t1, s_t1, lD[1], lS[1], x1, s_x1));          // conv2
  MD_CHECK(block(F.blk[2], x1, s_x1, lD[1], lS[1], t2, s_t2));          // conv3 (s2)
  MD_CHECK(block(F.blk[3], t2, s_t2, lD[2], lS[2], x2, s_x2));          // conv4
  MD_CHECK(block(F.blk[4], x2, s_x2, lD[2], lS[2], t3, s_t3));          // conv5 (s2)
  MD_CHECK(block(F.blk[5], t3, s_t3, lD[3], lS[3], levels[3], s_l3));   // conv6 -> x3
  MD_CHECK(up(F.blk[6], levels[3], s_l3, lD[3], lS[3], x2, levels[2], s_l2));  // up0 + x2
  MD_CHECK(up(F.blk[7], levels[2], s_l2, lD[2], lS[2], x1, levels[1], s_l1));  // up1 + x1
  MD_CHECK(up(F.blk[8], levels[1], s_l1, lD[1], lS[1], x0, levels[0], nullptr));  // up2 + x0
  A.release(m);
  return 0;
}

}  // namespace md
