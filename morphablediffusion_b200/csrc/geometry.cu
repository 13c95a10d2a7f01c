// Geometry group of the denoise step (HBM / latency bound):
//   voxelize            mesh vertices -> voxel indices (generate_face.py:214-225), bit-exact integer rule
//   target_encoder      NoisyTargetViewEncoder (network.py:163-207), one CTA per view, everything in shared memory
//   vertex_features     fused unproject + trilinear gather at the mesh vertices, summed over views
//   sparse_conv         rulebook gather-conv with folded BatchNorm + ReLU (network.py:74-161)
//   volume_resample     sparse-conv output -> 32^3 world grid (morphable_diffusion.py:234-255)
//   frustum_points      per-view ray sample points (utils.py:79-153)
//   frustum_gather      trilinear gather of the spatial volume along the rays (morphable_diffusion.py:312-315)
#include "host.h"
#include "ptx.cuh"
#include "kernels.h"
#include "geometry.h"

namespace md {

// ------------------------------------------------------------------------------------------------ voxelize (a1)
__global__ void minmax_kernel(const float* __restrict__ v, int nv, float* __restrict__ bounds) {
  pdl_grid_sync();
  __shared__ float smin[3][32], smax[3][32];
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float a = v[i * 3 + k];
      mn[k] = fminf(mn[k], a);
      mx[k] = fmaxf(mx[k], a);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o; o >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffff, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffff, mx[k], o));
    }
    if (lane == 0) { smin[k][warp] = mn[k]; smax[k][warp] = mx[k]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int k = threadIdx.x;
    float a = INFINITY, b = -INFINITY;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { a = fminf(a, smin[k][w]); b = fmaxf(b, smax[k][w]); }
    bounds[k] = a;
    bounds[3 + k] = b;
  }
}

__global__ void voxel_coord_kernel(const float* __restrict__ v, int nv, const float* __restrict__ bounds,
                                   int32_t* __restrict__ coord, int32_t* __restrict__ out_sh) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float voxel = 0.005f;
  if (i < nv) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {  // output order d,h,w = z,y,x
      const int src = 2 - k;
      const float q = __fdiv_rn(__fsub_rn(v[i * 3 + src], bounds[src]), voxel);
      coord[i * 3 + k] = static_cast<int32_t>(rintf(q));  // torch.round: half to even
    }
  }
  if (i < 3) {
    const int src = 2 - i;
    const float q = __fdiv_rn(__fsub_rn(bounds[3 + src], bounds[src]), voxel);
    out_sh[i] = (static_cast<int32_t>(ceilf(q)) | 3) + 1;
  }
}

int launch_voxelize(const float* vertices, int nv, int32_t* coord, int32_t* out_sh, float* bounds, cudaStream_t st) {
  launch_pdl(minmax_kernel, dim3(1), dim3(1024), 0, st, vertices, nv, bounds);
  MD_CHECK(check_launch("minmax"));
  launch_pdl(voxel_coord_kernel, dim3((std::max(nv, 3) + 255) / 256), dim3(256), 0, st, vertices, nv, bounds, coord, out_sh);
  return check_launch("voxel_coord");
}

// ------------------------------------------------------------------------------------------------ batch construction
// Rigid alignment of the fitted mesh (generate_face.py:203-213: scale, so3 rotation + translation, scale, axis swap), one
// affine map v' = A v + b per vertex (A row-major, composed on the host in double precision).
struct Affine3 { float a[9]; float b[3]; };
__global__ void affine_points_kernel(const float* __restrict__ v, int n, Affine3 m, float* __restrict__ out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
#pragma unroll
  for (int k = 0; k < 3; ++k) out[3 * i + k] = fmaf(m.a[3 * k], x, fmaf(m.a[3 * k + 1], y, fmaf(m.a[3 * k + 2], z, m.b[k])));
}

int launch_affine_points(const float* v, int n, const float* A9, const float* b3, float* out, cudaStream_t st) {
  Affine3 m;
  for (int i = 0; i < 9; ++i) m.a[i] = A9[i];
  for (int i = 0; i < 3; ++i) m.b[i] = b3[i];
  launch_pdl(affine_points_kernel, dim3((std::max(n, 1) + 255) / 256), dim3(256), 0, st, v, n, m, out);
  return check_launch("affine_points");
}

// Decoded images -> 8-bit pixels (generate_face.py:246-249): clamp to [-1, 1], (x + 1) / 2 * 255, truncate;
// NCHW fp32 [n][3][HW] -> NHWC uint8 [n][HW][3].
__global__ void images_to_u8_kernel(const float* __restrict__ img, uint8_t* __restrict__ out, int n, int HW) {
  pdl_grid_sync();
  const size_t total = static_cast<size_t>(n) * HW;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / HW, p = i - b * HW;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = fminf(fmaxf(img[(b * 3 + c) * HW + p], -1.f), 1.f);
      out[i * 3 + c] = static_cast<uint8_t>(__fmul_rn(__fmul_rn(__fadd_rn(x, 1.f), 0.5f), 255.f));
    }
  }
}

int launch_images_to_u8(const float* img, uint8_t* out, int n, int HW, cudaStream_t st) {
  const size_t total = static_cast<size_t>(n) * HW;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(148) * 16));
  launch_pdl(images_to_u8_kernel, dim3(std::max(blocks, 1)), dim3(256), 0, st, img, out, n, HW);
  return check_launch("images_to_u8");
}

// ------------------------------------------------------------------------------------------------ target encoder (K1)
// One CTA (1024 threads = one thread per latent pixel) runs the whole 8-conv encoder of one view out of shared memory.
struct EncSmem {
  float tmp[16][32 * 32];    // conv input after GN+SiLU
  float w[9 * 16 * 16];      // current conv weights [tap][ci][co]
  float red[32][16];         // per-warp GN partials
  float gstat[16];           // group mean (8) + rstd (8)
  float tv[16];
};

__device__ void enc_group_stats(EncSmem& s, const float (&v)[16]) {
  // 8 groups of 2 channels over 1024 pixels
  float part[16];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    part[g] = v[2 * g] + v[2 * g + 1];
    part[8 + g] = v[2 * g] * v[2 * g] + v[2 * g + 1] * v[2 * g + 1];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    float a = part[e];
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffff, a, o);
    if (lane == 0) s.red[warp][e] = a;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double s1 = 0.0, s2 = 0.0;
    for (int w = 0; w < 32; ++w) { s1 += s.red[w][threadIdx.x]; s2 += s.red[w][8 + threadIdx.x]; }
    const double n = 2048.0;
    const double mean = s1 / n;
    double var = s2 / n - mean * mean;
    if (var < 0) var = 0;
    s.gstat[threadIdx.x] = static_cast<float>(mean);
    s.gstat[8 + threadIdx.x] = static_cast<float>(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
}

// out[co] = bias[co] + sum_{tap,ci} w[tap][ci][co] * in[ci][pixel+tap]   (in = s.tmp, CIN input channels)
template <int CIN>
__device__ void enc_conv3x3(EncSmem& s, const float* __restrict__ gw, const float* __restrict__ gb, float (&acc)[16]) {
  // stage weights: global layout is torch [co][ci][ky][kx]
  for (int i = threadIdx.x; i < 9 * CIN * 16; i += blockDim.x) {
    const int co = i % 16, ci = (i / 16) % CIN, tap = i / (16 * CIN);
    s.w[i] = gw[(co * CIN + ci) * 9 + tap];
  }
  __syncthreads();
  const int px = threadIdx.x & 31, py = threadIdx.x >> 5;
#pragma unroll
  for (int co = 0; co < 16; ++co) acc[co] = gb[co];
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    if (yy < 0 || yy >= 32) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = px + kx - 1;
      if (xx < 0 || xx >= 32) continue;
      const float* wp = s.w + (ky * 3 + kx) * CIN * 16;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float a = s.tmp[ci][yy * 32 + xx];
        const float4* w4 = reinterpret_cast<const float4*>(wp + ci * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w = w4[q];
          acc[4 * q + 0] += a * w.x; acc[4 * q + 1] += a * w.y; acc[4 * q + 2] += a * w.z; acc[4 * q + 3] += a * w.w;
        }
      }
    }
  }
  __syncthreads();  // everyone done reading s.tmp / s.w
}

struct EncWeights {
  const float* init_w; const float* init_b;
  struct Res {
    const float* te_w; const float* te_b; const float* ve_w; const float* ve_b;
    const float* gn0_w; const float* gn0_b; const float* c0_w; const float* c0_b;
    const float* gn1_w; const float* gn1_b; const float* c1_w; const float* c1_b;
  } res[3];
  const float* fgn_w; const float* fgn_b; const float* fc_w; const float* fc_b;
};

__device__ __forceinline__ float silu_g(float x) { return silu_fast(x); }

__global__ void __launch_bounds__(1024, 1)
target_encoder_kernel(const float* __restrict__ x, const float* __restrict__ t_embed, const float* __restrict__ v_embed,
                      EncWeights W, float* __restrict__ out, int tdim, int vdim) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t enc_raw[];
  EncSmem& s = *reinterpret_cast<EncSmem*>(enc_raw);
  const int view = blockIdx.x;
  const int p = threadIdx.x;
  float f[16], acc[16];

  // init conv 4 -> 16
  for (int c = 0; c < 4; ++c) s.tmp[c][p] = x[(static_cast<size_t>(view) * 4 + c) * 1024 + p];
  __syncthreads();
  enc_conv3x3<4>(s, W.init_w, W.init_b, f);

  for (int rb = 0; rb < 3; ++rb) {
    const EncWeights::Res& R = W.res[rb];
    if (p < 16) {  // per-channel time/view bias (1x1 convs on a 1x1 map)
      float a = R.te_b[p] + R.ve_b[p];
      for (int k = 0; k < tdim; ++k) a += R.te_w[p * tdim + k] * t_embed[k];
      for (int k = 0; k < vdim; ++k) a += R.ve_w[p * vdim + k] * v_embed[view * vdim + k];
      s.tv[p] = a;
    }
    __syncthreads();
    float y[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) y[c] = f[c] + s.tv[c];
    enc_group_stats(s, y);
#pragma unroll
    for (int c = 0; c < 16; ++c)
      s.tmp[c][p] = silu_g((y[c] - s.gstat[c >> 1]) * s.gstat[8 + (c >> 1)] * R.gn0_w[c] + R.gn0_b[c]);
    __syncthreads();
    enc_conv3x3<16>(s, R.c0_w, R.c0_b, acc);
    enc_group_stats(s, acc);
#pragma unroll
    for (int c = 0; c < 16; ++c)
      s.tmp[c][p] = silu_g((acc[c] - s.gstat[c >> 1]) * s.gstat[8 + (c >> 1)] * R.gn1_w[c] + R.gn1_b[c]);
    __syncthreads();
    enc_conv3x3<16>(s, R.c1_w, R.c1_b, acc);
#pragma unroll
    for (int c = 0; c < 16; ++c) f[c] += acc[c];
  }
  enc_group_stats(s, f);
#pragma unroll
  for (int c = 0; c < 16; ++c)
    s.tmp[c][p] = silu_g((f[c] - s.gstat[c >> 1]) * s.gstat[8 + (c >> 1)] * W.fgn_w[c] + W.fgn_b[c]);
  __syncthreads();
  enc_conv3x3<16>(s, W.fc_w, W.fc_b, acc);
  // channels-last output [view][pixel][16]
  float4* o = reinterpret_cast<float4*>(out + (static_cast<size_t>(view) * 1024 + p) * 16);
#pragma unroll
  for (int q = 0; q < 4; ++q) o[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
}

// ---- cluster version: a cluster of 8 CTAs runs the encoder of one view, each CTA owning S/8 latent rows (one thread
// per pixel).  Halo rows of every conv input are copied out of the neighbour CTAs' shared memory (DSMEM) and the
// GroupNorm statistics are reduced over the cluster, so a view's 0.25 ms single-SM latency chain becomes ~8x shorter
// and any latent size with S % 8 == 0 and S*S/8 <= 512 fits (S = 64: BASELINE config 1).
constexpr int kEncCluster = 8;
struct EncClusterLayout {   // offsets in floats into dynamic shared memory
  int tmp, w, red, part, gstat, tv, total;
  __host__ __device__ EncClusterLayout(int S) {
    const int RP = S / kEncCluster;
    tmp = 0;                                  // [2][16][(RP + 2) * S]: ping-pong conv input incl. one halo row either side
    w = tmp + 2 * 16 * (RP + 2) * S;          // [9 * 16 * 16]
    red = w + 9 * 16 * 16;                    // [32 warps][16]
    part = red + 32 * 16;                     // [2][16] this CTA's GroupNorm partial sums (ping-pong)
    gstat = part + 32;                        // [16]
    tv = gstat + 16;                          // [16]
    total = tv + 16;
  }
};

__device__ __forceinline__ float ld_dsmem_f32(const float* local_ptr, uint32_t rank) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(mapa_shared(smem_u32(local_ptr), rank)) : "memory");
  return v;
}

// GroupNorm(8 groups of 2 channels) statistics of this view: CTA partials -> cluster reduction through DSMEM
__device__ void encc_group_stats(float* sm, const EncClusterLayout& L, const float (&v)[16], int pp, float n_elems) {
  float part[16];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    part[g] = v[2 * g] + v[2 * g + 1];
    part[8 + g] = v[2 * g] * v[2 * g] + v[2 * g + 1] * v[2 * g + 1];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    float a = part[e];
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffff, a, o);
    if (lane == 0) sm[L.red + warp * 16 + e] = a;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float a = 0.f;
    for (int w = 0; w < nwarp; ++w) a += sm[L.red + w * 16 + threadIdx.x];
    sm[L.part + pp * 16 + threadIdx.x] = a;
  }
  cluster_sync_all();
  if (threadIdx.x < 8) {
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t r = 0; r < kEncCluster; ++r) {
      s1 += ld_dsmem_f32(sm + L.part + pp * 16 + threadIdx.x, r);
      s2 += ld_dsmem_f32(sm + L.part + pp * 16 + 8 + threadIdx.x, r);
    }
    const double mean = s1 / n_elems;
    double var = s2 / n_elems - mean * mean;
    if (var < 0) var = 0;
    sm[L.gstat + threadIdx.x] = static_cast<float>(mean);
    sm[L.gstat + 8 + threadIdx.x] = static_cast<float>(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
}

// halo rows of buffer `b`: row 0 <- last interior row of the CTA above, row RP+1 <- first interior row of the CTA below
__device__ void encc_halo(float* sm, const EncClusterLayout& L, int b, int S, int RP, uint32_t rank, int nch) {
  const int plane = (RP + 2) * S;
  float* buf = sm + L.tmp + b * 16 * plane;
  cluster_sync_all();  // every CTA's interior rows of buffer b are written
  for (int i = threadIdx.x; i < 2 * nch * S; i += blockDim.x) {
    const int x = i % S, c = (i / S) % nch, side = i / (S * nch);
    float v = 0.f;
    if (side == 0) {
      if (rank > 0) v = ld_dsmem_f32(buf + c * plane + RP * S + x, rank - 1);
      buf[c * plane + x] = v;
    } else {
      if (rank + 1 < kEncCluster) v = ld_dsmem_f32(buf + c * plane + S + x, rank + 1);
      buf[c * plane + (RP + 1) * S + x] = v;
    }
  }
  __syncthreads();
}

template <int CIN>
__device__ void encc_conv3x3(float* sm, const EncClusterLayout& L, int b, int S, int RP, const float* __restrict__ gw,
                             const float* __restrict__ gb, float (&acc)[16]) {
  for (int i = threadIdx.x; i < 9 * CIN * 16; i += blockDim.x) {
    const int co = i % 16, ci = (i / 16) % CIN, tap = i / (16 * CIN);
    sm[L.w + i] = gw[(co * CIN + ci) * 9 + tap];
  }
  __syncthreads();
  const int plane = (RP + 2) * S;
  const float* buf = sm + L.tmp + b * 16 * plane;
  const int px = threadIdx.x % S, py = threadIdx.x / S;
#pragma unroll
  for (int co = 0; co < 16; ++co) acc[co] = gb[co];
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky;  // halo-inclusive row index; rows outside the image hold zeros
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = px + kx - 1;
      if (xx < 0 || xx >= S) continue;
      const float* wp = sm + L.w + (ky * 3 + kx) * CIN * 16;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float a = buf[ci * plane + yy * S + xx];
        const float4* w4 = reinterpret_cast<const float4*>(wp + ci * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w = w4[q];
          acc[4 * q + 0] += a * w.x; acc[4 * q + 1] += a * w.y; acc[4 * q + 2] += a * w.z; acc[4 * q + 3] += a * w.w;
        }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512, 1)
target_encoder_cluster_kernel(const float* __restrict__ x, const float* __restrict__ t_embed,
                              const float* __restrict__ v_embed, EncWeights W, float* __restrict__ out, int tdim, int vdim,
                              int S) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float encc_sm[];
  float* sm = encc_sm;
  const EncClusterLayout L(S);
  const int RP = S / kEncCluster;
  const uint32_t rank = cluster_ctarank();
  const int view = blockIdx.x / kEncCluster;
  const int p = threadIdx.x;
  const int px = p % S, py = p / S;
  const int gy = static_cast<int>(rank) * RP + py;
  const int plane = (RP + 2) * S;
  const float n_elems = 2.f * S * S;
  const size_t HW = static_cast<size_t>(S) * S;
  float f[16], acc[16];
  int b = 0, pp = 0;

  // init conv 4 -> 16: the input comes from global memory, halo rows included
  for (int i = p; i < 4 * plane; i += blockDim.x) {
    const int xx = i % S, r = (i / S) % (RP + 2), c = i / plane;
    const int yy = static_cast<int>(rank) * RP + r - 1;
    sm[L.tmp + c * plane + r * S + xx] =
        (yy >= 0 && yy < S) ? x[(static_cast<size_t>(view) * 4 + c) * HW + static_cast<size_t>(yy) * S + xx] : 0.f;
  }
  __syncthreads();
  encc_conv3x3<4>(sm, L, 0, S, RP, W.init_w, W.init_b, f);

  auto activate_into = [&](const float (&y)[16], const float* gw, const float* gbv) {
    b ^= 1;
    float* buf = sm + L.tmp + b * 16 * plane;
#pragma unroll
    for (int c = 0; c < 16; ++c)
      buf[c * plane + (py + 1) * S + px] =
          silu_g((y[c] - sm[L.gstat + (c >> 1)]) * sm[L.gstat + 8 + (c >> 1)] * gw[c] + gbv[c]);
    encc_halo(sm, L, b, S, RP, rank, 16);
  };

  for (int rb = 0; rb < 3; ++rb) {
    const EncWeights::Res& R = W.res[rb];
    if (p < 16) {  // per-channel time/view bias (1x1 convs on a 1x1 map)
      float a = R.te_b[p] + R.ve_b[p];
      for (int k = 0; k < tdim; ++k) a += R.te_w[p * tdim + k] * t_embed[k];
      for (int k = 0; k < vdim; ++k) a += R.ve_w[p * vdim + k] * v_embed[view * vdim + k];
      sm[L.tv + p] = a;
    }
    __syncthreads();
    float y[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) y[c] = f[c] + sm[L.tv + c];
    encc_group_stats(sm, L, y, pp, n_elems); pp ^= 1;
    activate_into(y, R.gn0_w, R.gn0_b);
    encc_conv3x3<16>(sm, L, b, S, RP, R.c0_w, R.c0_b, acc);
    encc_group_stats(sm, L, acc, pp, n_elems); pp ^= 1;
    activate_into(acc, R.gn1_w, R.gn1_b);
    encc_conv3x3<16>(sm, L, b, S, RP, R.c1_w, R.c1_b, acc);
#pragma unroll
    for (int c = 0; c < 16; ++c) f[c] += acc[c];
  }
  encc_group_stats(sm, L, f, pp, n_elems); pp ^= 1;
  activate_into(f, W.fgn_w, W.fgn_b);
  encc_conv3x3<16>(sm, L, b, S, RP, W.fc_w, W.fc_b, acc);
  // channels-last output [view][pixel][16]
  float4* o = reinterpret_cast<float4*>(out + (static_cast<size_t>(view) * HW + static_cast<size_t>(gy) * S + px) * 16);
#pragma unroll
  for (int q = 0; q < 4; ++q) o[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
  cluster_sync_all();  // no CTA may exit while a neighbour can still read its shared memory
}

int launch_target_encoder(const float* x, const float* t_embed, const float* v_embed, const EncWeightsHost& w,
                          float* out, int n_views, int tdim, int vdim, int S, cudaStream_t st) {
  EncWeights W;
  static_assert(sizeof(EncWeights) == sizeof(EncWeightsHost), "layout mismatch");
  memcpy(&W, &w, sizeof(W));
  static const bool single = getenv("MD_ENC_SINGLE") != nullptr;  // the one-CTA-per-view kernel (S = 32 only), for A/B
  if (single && S == 32) {
    static bool attr = false;
    const int smem = sizeof(EncSmem);
    if (!attr) {
      MD_CUDA(cudaFuncSetAttribute(target_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr = true;
    }
    launch_pdl(target_encoder_kernel, dim3(n_views), dim3(1024), smem, st, x, t_embed, v_embed, W, out, tdim, vdim);
    return check_launch("target_encoder");
  }
  if (S % kEncCluster != 0 || S * S / kEncCluster > 512 || S * S / kEncCluster < 32)
    return set_error("target_encoder: latent size %d unsupported (needs S %% 8 == 0 and 32 <= S*S/8 <= 512)", S);
  const EncClusterLayout L(S);
  const int smem = L.total * static_cast<int>(sizeof(float));
  static int smem_set = 0;
  if (smem > smem_set) {
    MD_CUDA(cudaFuncSetAttribute(target_encoder_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kEncCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.gridDim = dim3(n_views * kEncCluster); cfg.blockDim = dim3(S * S / kEncCluster);
  cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 2;
  cudaLaunchKernelEx(&cfg, target_encoder_cluster_kernel, x, t_embed, v_embed, W, out, tdim, vdim, S);
  return check_launch("target_encoder_cluster");
}

// ------------------------------------------------------------------------------------------------ projection helpers
// Projects a world point with the 3x4 matrix P (row-major) and returns pixel coordinates in the 32x32 feature map,
// following utils.py:20-43 and grid_sample's align_corners=True un-normalisation.
__device__ __forceinline__ void project_point(const float* __restrict__ P, int ortho, float wx, float wy, float wz,
                                              int size, float& px, float& py) {
  const float u = P[0] * wx + P[1] * wy + P[2] * wz + P[3];
  const float v = P[4] * wx + P[5] * wy + P[6] * wz + P[7];
  float gx, gy;
  if (!ortho) {
    float w = P[8] * wx + P[9] * wy + P[10] * wz + P[11];
    if (w < 1e-4f) w = 1e-4f;
    const float half = (size - 1) / 2.f;
    gx = (u / w) / half - 1.f;
    gy = (v / w) / half - 1.f;
  } else {
    gx = u;
    gy = v;
  }
  px = ((gx + 1.f) / 2.f) * (size - 1);
  py = ((gy + 1.f) / 2.f) * (size - 1);
}

// bilinear sample (zeros padding) of a channels-last [size][size][16] map; accumulates wgt * value into acc[4]
// for channel quad cq.
__device__ __forceinline__ void bilinear16_acc(const float* __restrict__ fmap, int size, float px, float py, int cq,
                                               float wgt, float4& acc) {
  const float fx = floorf(px), fy = floorf(py);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const float ax = px - fx, ay = py - fy;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int xx = x0 + dx, yy = y0 + dy;
      if (xx < 0 || xx >= size || yy < 0 || yy >= size) continue;
      const float w = wgt * (dx ? ax : 1.f - ax) * (dy ? ay : 1.f - ay);
      const float4 v = *reinterpret_cast<const float4*>(fmap + (static_cast<size_t>(yy) * size + xx) * 16 + cq * 4);
      acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
  }
}

// linspace(-length, length, V)[i] exactly as torch computes it (start + i*step for the first half, end - (V-1-i)*step
// for the second half).
__device__ __forceinline__ float linspace_at(float length, int V, int i) {
  const float step = (2.f * length) / static_cast<float>(V - 1);
  return (i < V / 2) ? (-length + step * i) : (length - step * (V - 1 - i));
}

// ------------------------------------------------------------------------------------------------ vertex features (K2+K3)
// sum over the views [view0, view0+n_views) of the trilinear sample (at vertex/length) of each view's unprojected
// volume.  The volume is never materialised: each of the 8 corner voxels is projected and bilinearly sampled on the fly.
// out [Nv][16] (sum over views, not yet divided by N).
__global__ void vertex_features_kernel(const float* __restrict__ feats, const float* __restrict__ proj, int ortho,
                                       int size, int V, float length, const float* __restrict__ vertices, int nv,
                                       int n_views, float* __restrict__ out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv * 4) return;
  const int cq = i & 3, vi = i >> 2;
  const float gx = vertices[vi * 3 + 0] / length, gy = vertices[vi * 3 + 1] / length, gz = vertices[vi * 3 + 2] / length;
  const float ix = ((gx + 1.f) / 2.f) * (V - 1), iy = ((gy + 1.f) / 2.f) * (V - 1), iz = ((gz + 1.f) / 2.f) * (V - 1);
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy), z0 = static_cast<int>(fz);
  const float ax = ix - fx, ay = iy - fy, az = iz - fz;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int n = 0; n < n_views; ++n) {
    const float* P = proj + n * 12;
    const float* fm = feats + static_cast<size_t>(n) * size * size * 16;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
      const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
      if (xx < 0 || xx >= V || yy < 0 || yy >= V || zz < 0 || zz >= V) continue;
      const float w = (dx ? ax : 1.f - ax) * (dy ? ay : 1.f - ay) * (dz ? az : 1.f - az);
      float px, py;
      project_point(P, ortho, linspace_at(length, V, xx), linspace_at(length, V, yy), linspace_at(length, V, zz), size,
                    px, py);
      bilinear16_acc(fm, size, px, py, cq, w, acc);
    }
  }
  *reinterpret_cast<float4*>(out + static_cast<size_t>(vi) * 16 + cq * 4) = acc;
}

int launch_vertex_features(const float* feats, const float* proj, int ortho, int size, int V, float length,
                           const float* vertices, int nv, int n_views, float* out, cudaStream_t st) {
  launch_pdl(vertex_features_kernel, dim3((nv * 4 + 127) / 128), dim3(128), 0, st, feats, proj, ortho, size, V, length, vertices, nv,
                                                                n_views, out);
  return check_launch("vertex_features");
}

// ------------------------------------------------------------------------------------------------ sparse conv (K4+K5)
// SMPLFeatureExtractor (Conv1d 1x1 on the view-mean, network.py:41-72) fused with the scatter into voxel rows:
// row r takes the features of its lowest-index vertex.
__global__ void smpl_scatter_kernel(const float* __restrict__ vsum, float inv_views, const float* __restrict__ W,
                                    const float* __restrict__ bias, const int32_t* __restrict__ row_vertex, int n_rows,
                                    float* __restrict__ out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * 16) return;
  const int co = i & 15, r = i >> 4;
  const float* f = vsum + static_cast<size_t>(row_vertex[r]) * 16;
  float a = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) a += W[co * 16 + k] * (f[k] * inv_views);
  out[i] = a + bias[co];
}

int launch_smpl_scatter(const float* vsum, float inv_views, const float* W, const float* bias,
                        const int32_t* row_vertex, int n_rows, float* out, cudaStream_t st) {
  launch_pdl(smpl_scatter_kernel, dim3((n_rows * 16 + 127) / 128), dim3(128), 0, st, vsum, inv_views, W, bias, row_vertex, n_rows, out);
  return check_launch("smpl_scatter");
}

// ---- cross-rank exchange of the vertex-feature sums over NVLink peer memory (Ctx::PeerExchange, engine.h) -----------------
// Push: this rank's [n4] float4 partial sums go into region (seq & 1, rank) of EVERY rank's exchange buffer (plain remote
// stores: fire and forget), each CTA fences at system scope, and the last CTA to finish raises this rank's flag in every
// peer's buffer to the step's sequence number.
__global__ void peer_push_kernel(const float4* __restrict__ vsum, int n4, float* const* __restrict__ peer_data,
                                 unsigned* const* __restrict__ peer_flags, unsigned* __restrict__ seq_ticket, int rank,
                                 int world, size_t slot_floats) {
  pdl_grid_sync();
  const unsigned seq = seq_ticket[0];
  const size_t off4 = ((seq & 1u) * world + rank) * (slot_floats / 4);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const float4 v = vsum[i];
    for (int p = 0; p < world; ++p) reinterpret_cast<float4*>(peer_data[p])[off4 + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(seq_ticket + 1, 1u);
    if (t == gridDim.x - 1) {
      seq_ticket[1] = 0u;  // every other CTA has taken its ticket: reset for the next step
      __threadfence_system();
      for (int p = 0; p < world; ++p) st_release_sys_u32(peer_flags[p] + rank, seq);
    }
  }
}

// Consume: SMPLFeatureExtractor + scatter as smpl_scatter_kernel, on the sum over ranks (added in rank order, so that
// every rank builds bit-identical volumes).  Each CTA first waits until every rank's flag in THIS GPU's buffer has
// reached the step's sequence number; the regions are read past L1 (__ldcg: peers write them through this GPU's L2).
__global__ void smpl_scatter_peer_kernel(const float* __restrict__ slots, const unsigned* __restrict__ flags,
                                         const unsigned* __restrict__ seq_ticket, int world, size_t slot_floats,
                                         float inv_views, const float* __restrict__ W, const float* __restrict__ bias,
                                         const int32_t* __restrict__ row_vertex, int n_rows, float* __restrict__ out,
                                         int* __restrict__ err) {
  pdl_grid_sync();
  const unsigned seq = seq_ticket[0];
  if (threadIdx.x < world) {
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_sys_u32(flags + threadIdx.x) - seq) < 0) {
      if (clock64() - t0 > (1LL << 32)) { *err = 1 + static_cast<int>(threadIdx.x); break; }  // ~2 s: report, never hang the device
      __nanosleep(100);
    }
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * 16) return;
  const int co = i & 15, r = i >> 4;
  const float* base = slots + (seq & 1u) * world * slot_floats + static_cast<size_t>(row_vertex[r]) * 16;
  float f[16];
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) {
    float4 s4 = __ldcg(reinterpret_cast<const float4*>(base) + k4);
    for (int p = 1; p < world; ++p) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(base + p * slot_floats) + k4);
      s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
    }
    f[k4 * 4] = s4.x; f[k4 * 4 + 1] = s4.y; f[k4 * 4 + 2] = s4.z; f[k4 * 4 + 3] = s4.w;
  }
  float a = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) a += W[co * 16 + k] * (f[k] * inv_views);
  out[i] = a + bias[co];
}

int launch_peer_push(const float* vsum, int n_floats, float* const* peer_data, unsigned* const* peer_flags,
                     unsigned* seq_ticket, int rank, int world, size_t slot_floats, cudaStream_t st) {
  const int n4 = n_floats / 4;
  launch_pdl(peer_push_kernel, dim3(std::min((n4 + 255) / 256, 64)), dim3(256), 0, st, reinterpret_cast<const float4*>(vsum), n4,
             peer_data, peer_flags, seq_ticket, rank, world, slot_floats);
  return check_launch("peer_push");
}

int launch_smpl_scatter_peer(const float* slots, const unsigned* flags, const unsigned* seq_ticket, int world,
                             size_t slot_floats, float inv_views, const float* W, const float* bias,
                             const int32_t* row_vertex, int n_rows, float* out, int* err, cudaStream_t st) {
  launch_pdl(smpl_scatter_peer_kernel, dim3((n_rows * 16 + 127) / 128), dim3(128), 0, st, slots, flags, seq_ticket, world,
             slot_floats, inv_views, W, bias, row_vertex, n_rows, out, err);
  return check_launch("smpl_scatter_peer");
}

// out[r][co] = relu( scale[co] * sum_{k,ci} W[k][ci][co] * in[nbr[r][k]][ci] + shift[co] ),  nbr < 0 = inactive.
// One warp per output row: lane k < 27 owns kernel tap k (gathers its neighbour row and multiplies by W[k]), then the
// 27 partial vectors are summed with a butterfly reduction.
template <int CIN, int COUT>
__global__ void sparse_conv_kernel(const float* __restrict__ in, const int32_t* __restrict__ nbr,
                                   const float* __restrict__ W, const float* __restrict__ scale,
                                   const float* __restrict__ shift, float* __restrict__ out, int n_rows) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  float acc[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
  const int j = (lane < 27) ? __ldg(nbr + r * 27 + lane) : -1;
  if (j >= 0) {
    const float4* ip = reinterpret_cast<const float4*>(in + static_cast<size_t>(j) * CIN);
    const float* wp = W + static_cast<size_t>(lane) * CIN * COUT;
#pragma unroll 2
    for (int c4 = 0; c4 < CIN / 4; ++c4) {
      const float4 a = __ldg(ip + c4);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4* w4 = reinterpret_cast<const float4*>(wp + static_cast<size_t>(c4 * 4 + e) * COUT);
#pragma unroll
        for (int q = 0; q < COUT / 4; ++q) {
          const float4 w = __ldg(w4 + q);
          acc[4 * q] += av[e] * w.x; acc[4 * q + 1] += av[e] * w.y; acc[4 * q + 2] += av[e] * w.z; acc[4 * q + 3] += av[e] * w.w;
        }
      }
    }
  }
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[co] += __shfl_xor_sync(0xffffffff, acc[co], o);
  }
  // lane co writes output channel co (COUT <= 64: two channels per lane at most)
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    if ((co & 31) == lane) out[static_cast<size_t>(r) * COUT + co] = fmaxf(acc[co] * scale[co] + shift[co], 0.f);
  }
}

// Wide layers (Cin >= 32): one thread per (row, 4 output channels) — float4 weight loads coalesced over co.
__global__ void sparse_conv_quad_kernel(const float* __restrict__ in, const int32_t* __restrict__ nbr,
                                        const float* __restrict__ W, const float* __restrict__ scale,
                                        const float* __restrict__ shift, float* __restrict__ out, int n_rows, int Cin,
                                        int Cout) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int CQ = Cout >> 2;
  if (i >= n_rows * CQ) return;
  const int cq = i % CQ, r = i / CQ;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < 27; ++k) {
    const int j = __ldg(nbr + r * 27 + k);
    if (j < 0) continue;
    const float4* ip = reinterpret_cast<const float4*>(in + static_cast<size_t>(j) * Cin);
    const float* wp = W + static_cast<size_t>(k) * Cin * Cout + cq * 4;
#pragma unroll 4
    for (int c4 = 0; c4 < (Cin >> 2); ++c4) {
      const float4 a = __ldg(ip + c4);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + static_cast<size_t>(c4 * 4 + 0) * Cout));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + static_cast<size_t>(c4 * 4 + 1) * Cout));
      const float4 w2 = __ldg(reinterpret_cast<const float4*>(wp + static_cast<size_t>(c4 * 4 + 2) * Cout));
      const float4 w3 = __ldg(reinterpret_cast<const float4*>(wp + static_cast<size_t>(c4 * 4 + 3) * Cout));
      acc.x += a.x * w0.x + a.y * w1.x + a.z * w2.x + a.w * w3.x;
      acc.y += a.x * w0.y + a.y * w1.y + a.z * w2.y + a.w * w3.y;
      acc.z += a.x * w0.z + a.y * w1.z + a.z * w2.z + a.w * w3.z;
      acc.w += a.x * w0.w + a.y * w1.w + a.z * w2.w + a.w * w3.w;
    }
  }
  const float4 sc = *reinterpret_cast<const float4*>(scale + cq * 4);
  const float4 sh = *reinterpret_cast<const float4*>(shift + cq * 4);
  float4 o;
  o.x = fmaxf(acc.x * sc.x + sh.x, 0.f); o.y = fmaxf(acc.y * sc.y + sh.y, 0.f);
  o.z = fmaxf(acc.z * sc.z + sh.z, 0.f); o.w = fmaxf(acc.w * sc.w + sh.w, 0.f);
  *reinterpret_cast<float4*>(out + static_cast<size_t>(r) * Cout + cq * 4) = o;
}

// Wide layers as a register-tiled gather GEMM: one CTA owns TM output rows x COUT channels.  Per kernel tap that is active
// for at least one row of the tile, the gathered input rows ([row][ci], rows of inactive neighbours zero-filled) and the
// tap's CIN x COUT weight block are copied into shared memory with cp.async through an NST-stage ring, so the gathers of
// the next NST - 1 taps are in flight while the current one is multiplied.  Every thread accumulates an RPT x 4
// micro-tile: per four input channels, RPT + 4 16-byte shared-memory loads feed 16 * RPT FMAs.  NGRP groups of 256
// threads share the tile: the active taps are dealt round-robin to the groups, each with its own ring and named
// barrier, and the partial sums meet in shared memory before the epilogue.
// History (64 -> 64 layer of the FLAME mesh, 7 051 rows, ncu): one thread per (row, 4 channels) re-reading every weight
// from L1 for 4 FMAs: 154 us (L1 bandwidth); tiles with weights through L1: 170 us (L2 latency of each tap's first
// touch); weights staged one tap ahead, register-staged transposed gather: 76 us (one gather in flight per group,
// issue slots 50 % busy); this version: see profiles/r02_sparse_conv.md.
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void group_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int CIN, int COUT, int TM>
constexpr int sparse_stage_floats() { return TM * (CIN + 4) + CIN * COUT; }
template <int CIN, int COUT, int TM, int NGRP, int NST>
constexpr int sparse_tile_smem() { return 27 * TM * 4 + NGRP * NST * sparse_stage_floats<CIN, COUT, TM>() * 4; }

template <int CIN, int COUT, int TM, int NGRP, int NST>
__global__ void __launch_bounds__(256 * NGRP) sparse_conv_tile_kernel(const float* __restrict__ in, const int32_t* __restrict__ nbr,
                                                                      const float* __restrict__ W, const float* __restrict__ scale,
                                                                      const float* __restrict__ shift, float* __restrict__ out,
                                                                      int n_rows) {
  constexpr int CQ = COUT / 4;       // column quads
  constexpr int RG = 256 / CQ;       // row groups
  constexpr int RPT = TM / RG;       // rows per thread
  constexpr int LDA = CIN + 4;       // padded row pitch: the row groups of one warp land on different banks
  constexpr int NA = TM * CIN / 4;   // activation float4 per tap
  constexpr int NWT = CIN * COUT / 4;  // weight float4 per tap
  constexpr int STAGE = sparse_stage_floats<CIN, COUT, TM>();
  constexpr int GRP_FLOATS = NST * STAGE;
  static_assert(RPT == 2 || RPT == 4, "micro-tile is 2 or 4 rows");
  static_assert(NA % 256 == 0, "gather mapping");
  static_assert(NGRP == 1 || TM * COUT <= GRP_FLOATS, "partial sums reuse a group's ring");
  extern __shared__ __align__(16) uint8_t sparse_smem[];
  const int grp = threadIdx.x >> 8;
  const int t = threadIdx.x & 255;
  float* ring = reinterpret_cast<float*>(sparse_smem) + grp * GRP_FLOATS;
  int (*s_nbr)[TM] = reinterpret_cast<int (*)[TM]>(reinterpret_cast<float*>(sparse_smem) + NGRP * GRP_FLOATS);
  __shared__ unsigned s_mask;
  const int row0 = blockIdx.x * TM;
  if (threadIdx.x == 0) s_mask = 0u;
  __syncthreads();
  unsigned mine = 0u;
  for (int i = threadIdx.x; i < 27 * TM; i += 256 * NGRP) {  // the rulebook is static (uploaded at bind): read it before the grid dependency
    const int r = i / 27, k = i - r * 27;
    const int j = (row0 + r < n_rows) ? __ldg(nbr + static_cast<size_t>(row0) * 27 + i) : -1;
    s_nbr[k][r] = j;
    if (j >= 0) mine |= 1u << k;
  }
  mine = __reduce_or_sync(0xffffffffu, mine);
  if ((t & 31) == 0 && mine) atomicOr(&s_mask, mine);
  __syncthreads();
  unsigned mask = s_mask;
  if (NGRP > 1) {  // keep every NGRP-th active tap, starting at this group's index
    unsigned keep = 0u, rest = mask;
    for (int i = 0; rest; ++i, rest &= rest - 1)
      if (i % NGRP == grp) keep |= rest & (0u - rest);
    mask = keep;
  }
  const int n_taps = __popc(mask);

  const int cq = t % CQ, rg = t / CQ;
  float acc[RPT][4];
#pragma unroll
  for (int r = 0; r < RPT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;

  unsigned to_issue = mask;
  auto issue_next = [&](int stage) {  // copies of the next unissued tap into ring slot `stage`; always commits a group
    if (to_issue) {
      const int k = __ffs(to_issue) - 1;
      to_issue &= to_issue - 1;
      float* sa = ring + stage * STAGE;
      float* sw = sa + TM * LDA;
      const float4* wsrc = reinterpret_cast<const float4*>(W + static_cast<size_t>(k) * CIN * COUT);
#pragma unroll
      for (int q = 0; q < (NWT + 255) / 256; ++q)
        if (NWT % 256 == 0 || t + q * 256 < NWT) cp_async16_zfill(sw + (t + q * 256) * 4, wsrc + t + q * 256, 16);
#pragma unroll
      for (int q = 0; q < NA / 256; ++q) {
        const int idx = t + q * 256;
        const int r = idx / (CIN / 4), c4 = idx - r * (CIN / 4);
        const int j = s_nbr[k][r];
        cp_async16_zfill(sa + r * LDA + c4 * 4, in + (j >= 0 ? static_cast<size_t>(j) * CIN + c4 * 4 : 0), j >= 0 ? 16 : 0);
      }
    }
    cp_async_commit();
  };

  pdl_grid_sync();
#pragma unroll
  for (int sidx = 0; sidx < NST - 1; ++sidx) issue_next(sidx);
  for (int i = 0; i < n_taps; ++i) {
    cp_async_wait<NST - 2>();         // this thread's copies of tap i have landed
    group_barrier(1 + grp, 256);      // ... everyone's have, and everyone is past the multiply of tap i - 1
    issue_next((i + NST - 1) % NST);  // refill the slot tap i - 1 used
    const float* sa = ring + (i % NST) * STAGE + rg * RPT * LDA;
    const float* sw = ring + (i % NST) * STAGE + TM * LDA + cq * 4;
#pragma unroll 2
    for (int c4 = 0; c4 < CIN / 4; ++c4) {
      float4 av[RPT];
#pragma unroll
      for (int r = 0; r < RPT; ++r) av[r] = *reinterpret_cast<const float4*>(sa + r * LDA + c4 * 4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 w = *reinterpret_cast<const float4*>(sw + (c4 * 4 + e) * COUT);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          const float a = e == 0 ? av[r].x : e == 1 ? av[r].y : e == 2 ? av[r].z : av[r].w;
          acc[r][0] += a * w.x; acc[r][1] += a * w.y; acc[r][2] += a * w.z; acc[r][3] += a * w.w;
        }
      }
    }
  }
  cp_async_wait<0>();
  if (NGRP > 1) {  // groups 1.. park their partial tiles in their own (now idle) ring; group 0 adds them
    group_barrier(1 + grp, 256);  // the whole group is past its last multiply before the ring is overwritten
    if (grp > 0) {
      float4* part = reinterpret_cast<float4*>(ring);
#pragma unroll
      for (int r = 0; r < RPT; ++r) part[r * 256 + t] = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
    __syncthreads();
    if (grp > 0) return;
    for (int gi = 1; gi < NGRP; ++gi) {
      const float4* other = reinterpret_cast<const float4*>(reinterpret_cast<float*>(sparse_smem) + gi * GRP_FLOATS);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float4 v = other[r * 256 + t];
        acc[r][0] += v.x; acc[r][1] += v.y; acc[r][2] += v.z; acc[r][3] += v.w;
      }
    }
  }
  const float4 sc = *reinterpret_cast<const float4*>(scale + cq * 4);
  const float4 sh = *reinterpret_cast<const float4*>(shift + cq * 4);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int row = row0 + rg * RPT + r;
    if (row >= n_rows) continue;
    float4 o;
    o.x = fmaxf(acc[r][0] * sc.x + sh.x, 0.f); o.y = fmaxf(acc[r][1] * sc.y + sh.y, 0.f);
    o.z = fmaxf(acc[r][2] * sc.z + sh.z, 0.f); o.w = fmaxf(acc[r][3] * sc.w + sh.w, 0.f);
    *reinterpret_cast<float4*>(out + static_cast<size_t>(row) * COUT + cq * 4) = o;
  }
}

template <int CIN, int COUT, int NGRP, int NST>
static void launch_sparse_tile(const float* in, const int32_t* nbr, const float* W, const float* scale, const float* shift,
                               float* out, int n_rows, cudaStream_t st) {
  constexpr int TM = 64;
  constexpr int smem = sparse_tile_smem<CIN, COUT, TM, NGRP, NST>();
  static_assert(smem <= 227 * 1024, "shared memory per CTA");
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(sparse_conv_tile_kernel<CIN, COUT, TM, NGRP, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    configured = true;
  }
  launch_pdl(sparse_conv_tile_kernel<CIN, COUT, TM, NGRP, NST>, dim3((n_rows + TM - 1) / TM), dim3(256 * NGRP), smem, st, in,
             nbr, W, scale, shift, out, n_rows);
}

int launch_sparse_conv(const float* in, const int32_t* nbr, const float* W, const float* scale, const float* shift,
                       float* out, int n_rows, int Cin, int Cout, cudaStream_t st) {
  if (n_rows == 0) return 0;
  const unsigned blocks = (static_cast<unsigned>(n_rows) * 32 + 127) / 128;
  // MD_SPARSE_TILE: 0 = the one-warp-per-row / one-thread-per-quad kernels for every layer, 1 = tile kernel with one
  // group of 256 threads, 2 (default) = two groups
  static const int tile = getenv("MD_SPARSE_TILE") != nullptr ? atoi(getenv("MD_SPARSE_TILE")) : 2;
  if (Cin == 16 && Cout == 16) launch_pdl(sparse_conv_kernel<16, 16>, dim3(blocks), dim3(128), 0, st, in, nbr, W, scale, shift, out, n_rows);
  else if (tile == 1 && Cin == 16 && Cout == 32) launch_sparse_tile<16, 32, 1, 4>(in, nbr, W, scale, shift, out, n_rows, st);
  else if (tile == 1 && Cin == 32 && Cout == 32) launch_sparse_tile<32, 32, 1, 4>(in, nbr, W, scale, shift, out, n_rows, st);
  else if (tile == 1 && Cin == 32 && Cout == 64) launch_sparse_tile<32, 64, 1, 4>(in, nbr, W, scale, shift, out, n_rows, st);
  else if (tile == 1 && Cin == 64 && Cout == 64) launch_sparse_tile<64, 64, 1, 4>(in, nbr, W, scale, shift, out, n_rows, st);
  else if (tile >= 2 && Cin == 16 && Cout == 32) launch_sparse_tile<16, 32, 2, 4>(in, nbr, W, scale, shift, out, n_rows, st);
  else if (tile >= 2 && Cin == 32 && Cout == 32) launch_sparse_tile<32, 32, 2, 4>(in, nbr, W, scale, shift, out, n_rows, st);
  else if (tile >= 2 && Cin == 32 && Cout == 64) launch_sparse_tile<32, 64, 2, 3>(in, nbr, W, scale, shift, out, n_rows, st);
  else if (tile >= 2 && Cin == 64 && Cout == 64) launch_sparse_tile<64, 64, 2, 3>(in, nbr, W, scale, shift, out, n_rows, st);
  else if (Cin == 16 && Cout == 32) launch_pdl(sparse_conv_kernel<16, 32>, dim3(blocks), dim3(128), 0, st, in, nbr, W, scale, shift, out, n_rows);
  else if (Cin % 4 == 0 && Cout % 4 == 0)
    launch_pdl(sparse_conv_quad_kernel, dim3((n_rows * (Cout / 4) + 63) / 64), dim3(64), 0, st, in, nbr, W, scale, shift, out, n_rows, Cin, Cout);
  else return set_error("sparse_conv: unsupported channels %d -> %d", Cin, Cout);
  return check_launch("sparse_conv");
}

// ------------------------------------------------------------------------------------------------ volume resample (K6)
// vol[p][c] = sum_j w[p][j] * feat[idx[p][j]][c];  idx < 0 = empty voxel.  feat [n2][64], vol [V^3][64] fp32.
__global__ void volume_resample_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx,
                                       const float* __restrict__ wgt, float* __restrict__ vol, int npts) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<size_t>(npts) * 16) return;
  const int cq = static_cast<int>(i & 15);
  const size_t p = i >> 4;
  float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = idx[p * 8 + j];
    if (r < 0) continue;
    const float w = wgt[p * 8 + j];
    const float4 v = *reinterpret_cast<const float4*>(feat + static_cast<size_t>(r) * 64 + cq * 4);
    acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
  }
  *reinterpret_cast<float4*>(vol + p * 64 + cq * 4) = acc;
}

int launch_volume_resample(const float* feat, const int32_t* idx, const float* wgt, float* vol, int npts,
                           cudaStream_t st) {
  const size_t total = static_cast<size_t>(npts) * 16;
  launch_pdl(volume_resample_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, feat, idx, wgt, vol, npts);
  return check_launch("volume_resample");
}

// ------------------------------------------------------------------------------------------------ frustum points (K7)
// pts[view][d][y][x] = world point / length (normalised grid_sample coordinate), utils.py:79-153.
// persp: world = M * (x*dep, y*dep, dep) + t ;  ortho: world = M * (kx, ky, dep) + t with (kx,ky) = Kinv*(gx,gy,1).
__global__ void frustum_points_kernel(const float* __restrict__ cam, int ortho, int D, int size, float length,
                                      float frustum_len, float* __restrict__ pts, int n_views) {
  pdl_grid_sync();
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t per_view = static_cast<size_t>(D) * size * size;
  if (i >= per_view * n_views) return;
  const int view = static_cast<int>(i / per_view);
  size_t r = i % per_view;
  const int x = static_cast<int>(r % size); r /= size;
  const int y = static_cast<int>(r % size);
  const int d = static_cast<int>(r / size);
  const float* c = cam + view * 24;  // M (9), t (3), dist (1), Kinv rows for ortho (6: k00 k01 k02 k10 k11 k12), pad
  const float dist = c[12];
  const float nearv = dist - frustum_len, farv = dist + frustum_len;
  // torch.linspace(0,1,D)
  const float lstep = 1.f / static_cast<float>(D - 1);
  const float lin = (d < D / 2) ? (lstep * d) : (1.f - lstep * (D - 1 - d));
  const float dep = lin * (farv - nearv) + nearv;
  float a, b, cc;
  if (!ortho) {
    a = x * dep; b = y * dep; cc = dep;
  } else {
    const float gx = (2.f * x) / (size - 1) - 1.f, gy = (2.f * y) / (size - 1) - 1.f;
    a = c[13] * gx + c[14] * gy + c[15];
    b = c[16] * gx + c[17] * gy + c[18];
    cc = dep;
  }
  const float wx = c[0] * a + c[1] * b + c[2] * cc + c[9];
  const float wy = c[3] * a + c[4] * b + c[5] * cc + c[10];
  const float wz = c[6] * a + c[7] * b + c[8] * cc + c[11];
  pts[i * 3 + 0] = wx / length;
  pts[i * 3 + 1] = wy / length;
  pts[i * 3 + 2] = wz / length;
}

int launch_frustum_points(const float* cam, int ortho, int D, int size, float length, float frustum_len, float* pts,
                          int n_views, cudaStream_t st) {
  const size_t total = static_cast<size_t>(D) * size * size * n_views;
  launch_pdl(frustum_points_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, cam, ortho, D, size, length,
                                                                                    frustum_len, pts, n_views);
  return check_launch("frustum_points");
}

// ------------------------------------------------------------------------------------------------ frustum gather (K8)
// Trilinear sample of the shared spatial volume at every frustum point (grid_sample, align_corners=True, zeros padding;
// morphable_diffusion.py:301-307).  vol fp32 [V][V][V][64] (8.4 MB: L2 resident); out bf16 [npts][64].
// Eight lanes per point, eight channels per lane: a warp works on four consecutive points, every lane has its sixteen
// 16-byte corner loads in flight together and stores 16 bytes (512 contiguous bytes per warp store).  Consecutive points
// are neighbouring pixels of one depth slice, so the corner reads are L1 hits; the kernel is bound by the 2 KB of corner
// data each point pulls through the L1 (3.2 GB per 16-view step), not by the 101 MB it writes.
__global__ void __launch_bounds__(256) frustum_gather_kernel(const float* __restrict__ vol, const float* __restrict__ pts,
                                                             int V, __nv_bfloat16* __restrict__ out, size_t npts) {
  pdl_grid_sync();
  const int oct = threadIdx.x & 7;           // channels [8*oct, 8*oct + 8)
  const size_t stride = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 3;
  for (size_t p = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 3; p < npts; p += stride) {
    const float gx = pts[p * 3], gy = pts[p * 3 + 1], gz = pts[p * 3 + 2];
    const float ix = ((gx + 1.f) / 2.f) * (V - 1), iy = ((gy + 1.f) / 2.f) * (V - 1), iz = ((gz + 1.f) / 2.f) * (V - 1);
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const float ax = ix - fx, ay = iy - fy, az = iz - fz;
    // clamp before the int conversion so far-away points cannot overflow
    const int x0 = static_cast<int>(fminf(fmaxf(fx, -2.f), static_cast<float>(V)));
    const int y0 = static_cast<int>(fminf(fmaxf(fy, -2.f), static_cast<float>(V)));
    const int z0 = static_cast<int>(fminf(fmaxf(fz, -2.f), static_cast<float>(V)));
    float4 lo[8], hi[8];
    float w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
      const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
      const bool ok = xx >= 0 && xx < V && yy >= 0 && yy < V && zz >= 0 && zz < V;
      w[c] = ok ? (dx ? ax : 1.f - ax) * (dy ? ay : 1.f - ay) * (dz ? az : 1.f - az) : 0.f;
      const float* src = vol + ((static_cast<size_t>(ok ? zz : 0) * V + (ok ? yy : 0)) * V + (ok ? xx : 0)) * 64 + oct * 8;
      lo[c] = __ldg(reinterpret_cast<const float4*>(src));
      hi[c] = __ldg(reinterpret_cast<const float4*>(src) + 1);
    }
    // same accumulation order as the one-warp-per-point kernel it replaces: corners 0..7, a += w * v
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (w[c] != 0.f) {
        a[0] += w[c] * lo[c].x; a[1] += w[c] * lo[c].y; a[2] += w[c] * lo[c].z; a[3] += w[c] * lo[c].w;
        a[4] += w[c] * hi[c].x; a[5] += w[c] * hi[c].y; a[6] += w[c] * hi[c].z; a[7] += w[c] * hi[c].w;
      }
    }
    uint4 o;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(a[0], a[1]); o.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(a[2], a[3]); o.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(a[4], a[5]); o.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(a[6], a[7]); o.w = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint4*>(out + p * 64 + oct * 8) = o;
  }
}

int launch_frustum_gather(const float* vol, const float* pts, int V, void* out_bf16, size_t npts, cudaStream_t st) {
  const size_t threads = npts * 8;
  const unsigned blocks = static_cast<unsigned>(std::min<size_t>((threads + 255) / 256, static_cast<size_t>(num_sms()) * 32));
  launch_pdl(frustum_gather_kernel, dim3(std::max(blocks, 1u)), dim3(256), 0, st, vol, pts, V,
             static_cast<__nv_bfloat16*>(out_bf16), npts);
  return check_launch("frustum_gather");
}

}  // namespace md
