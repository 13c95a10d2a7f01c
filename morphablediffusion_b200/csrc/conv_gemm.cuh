// Implicit-GEMM convolution / GEMM on tcgen05 tensor cores (sm_100a).
//
// One kernel family covers every dense contraction of the denoise step:
//   * Linear / 1x1 conv           : 1 tap, A viewed as [rows, C]
//   * Conv2d 3x3 / Conv3d 3x3x3   : 9 / 27 taps; every tap is a TMA box load of the channels-last activation
//                                   tensor at shifted coordinates (out-of-bounds = zero fill = conv padding)
//   * ConvTranspose3d (k3,s2,p1,op1): 8 launches, one per output parity class, with 1/2/4/8 taps each and a
//                                   strided output scatter in the epilogue.
// Operands are bf16, K-major, 128B-swizzled in shared memory; the accumulator lives in TMEM (fp32).
// Warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps2..5 = epilogue.
#pragma once
#include "ptx.cuh"
#include "kernels.h"

namespace md {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kMaxTaps = 27;


struct ConvGemmParams {
  // input (A) geometry, channels-last [B][D][H][W][C]
  int W, H, D, B;
  int bw, bh, bd, bb;      // TMA box (bw*bh*bd*bb == 128)
  int nxb, nyb, nzb, nbb;  // blocks per dim
  int kblocks_per_tap;     // Cin / 64
  int ntaps;
  int8_t tdx[kMaxTaps], tdy[kMaxTaps], tdz[kMaxTaps];
  int N;        // GEMM N (rows of the packed weight matrix)
  int n_tiles;  // ceil(N / BN)
  int m_tiles;
  // output geometry: out coord = in coord * os + op ; output tensor dims (OW,OH,OD), row stride ldo
  int OW, OH, OD;
  int osx, osy, osz, opx, opy, opz;
  int ldo;  // elements per output row
  // epilogue
  const float* bias;    // [N]
  const float* rowvec;  // [B][rowvec_ld] per-sample additive vector
  int rowvec_ld;
  const float* res_f32;          // [Mout][ldo]
  const __nv_bfloat16* res_bf16; // [Mout][ldo]
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int act;
  float out_scale;  // multiplies the result before residual (1.0 default)
};

__device__ __forceinline__ float act_silu(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float act_gelu(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

template <int BN, int STAGES>
struct ConvGemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvGemmParams p) {
  using S = ConvGemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int kblocks = p.ntaps * p.kblocks_per_tap;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int m = tile / p.n_tiles;
        const int xb = m % p.nxb; m /= p.nxb;
        const int yb = m % p.nyb; m /= p.nyb;
        const int zb = m % p.nzb; m /= p.nzb;
        const int x0 = xb * p.bw, y0 = yb * p.bh, z0 = zb * p.bd, b0 = m * p.bb;
        const int n0 = n_tile * BN;
        int kcol = 0;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int cx = x0 + p.tdx[tap], cy = y0 + p.tdy[tap], cz = z0 + p.tdz[tap];
          for (int kc = 0; kc < p.kblocks_per_tap; ++kc, ++it, kcol += kBlockK) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + s * S::kStageBytes;
            uint8_t* sb = sa + S::kABytes;
            mbar_expect_tx(&full_bar[s], S::kStageBytes);
            tma_load_5d(sa, &tmA, &full_bar[s], kc * kBlockK, cx, cy, cz, b0);
            tma_load_2d(sb, &tmB, &full_bar[s], kcol, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN);
      int it = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int a = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        mbar_wait(&tmem_empty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
          const uint64_t da = make_sw128_kmajor_desc(sa);
          const uint64_t db = make_sw128_kmajor_desc(sb);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128-B swizzle atom: +2 in the (addr >> 4) field
            tc_mma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[s]);
        }
        tc_commit(&tmem_full[a]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int a = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      const int n_tile = tile % p.n_tiles;
      int m = tile / p.n_tiles;
      const int xb = m % p.nxb; m /= p.nxb;
      const int yb = m % p.nyb; m /= p.nyb;
      const int zb = m % p.nzb; m /= p.nzb;
      int rr = r;
      const int ix = rr % p.bw; rr /= p.bw;
      const int iy = rr % p.bh; rr /= p.bh;
      const int iz = rr % p.bd; rr /= p.bd;
      const int x = xb * p.bw + ix, y = yb * p.bh + iy, z = zb * p.bd + iz, b = m * p.bb + rr;
      const bool valid = (x < p.W) && (y < p.H) && (z < p.D) && (b < p.B);
      const long long orow =
          ((static_cast<long long>(b) * p.OD + (z * p.osz + p.opz)) * p.OH + (y * p.osy + p.opy)) * p.OW +
          (x * p.osx + p.opx);

      mbar_wait(&tmem_full[a], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN;

      if (p.act == ACT_GEGLU) {
        constexpr int HALF = BN / 2;
        const int n_out0 = n_tile * HALF;
#pragma unroll 1
        for (int c = 0; c < HALF / 32; ++c) {
          uint32_t v[32], g[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_32x32(taddr + HALF + c * 32, g);
          tc_wait_ld();
          if (valid) {
            const int nb = n_tile * BN + c * 32;  // packed-row index of the value half
            __nv_bfloat16* o = p.out_bf16 + orow * p.ldo + n_out0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (n_out0 + c * 32 + j < p.N / 2) {
                __align__(16) __nv_bfloat16 ob[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  float val = __uint_as_float(v[j + e]) + (p.bias ? p.bias[nb + j + e] : 0.f);
                  float gate = __uint_as_float(g[j + e]) + (p.bias ? p.bias[nb + HALF + j + e] : 0.f);
                  ob[e] = __float2bfloat16(val * act_gelu(gate));
                }
                *reinterpret_cast<uint4*>(o + j) = *reinterpret_cast<const uint4*>(ob);
              }
            }
          }
        }
      } else {
        const int n0 = n_tile * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tc_wait_ld();
          if (valid) {
            const int nb = n0 + c * 32;
            const long long obase = orow * p.ldo + nb;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (nb + j < p.N) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]) * p.out_scale;
                if (p.bias) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] += __ldg(p.bias + nb + j + e);
                }
                if (p.rowvec) {
                  const float* rv = p.rowvec + static_cast<long long>(b) * p.rowvec_ld + nb + j;
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] += __ldg(rv + e);
                }
                if (p.act == ACT_SILU) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] = act_silu(f[e]);
                } else if (p.act == ACT_RELU) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
                } else if (p.act == ACT_GELU) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] = act_gelu(f[e]);
                }
                if (p.res_f32) {
                  const float4 r0 = *reinterpret_cast<const float4*>(p.res_f32 + obase + j);
                  const float4 r1 = *reinterpret_cast<const float4*>(p.res_f32 + obase + j + 4);
                  f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
                  f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
                }
                if (p.res_bf16) {
                  const uint4 rb = *reinterpret_cast<const uint4*>(p.res_bf16 + obase + j);
                  const __nv_bfloat16* rbh = reinterpret_cast<const __nv_bfloat16*>(&rb);
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] += __bfloat162float(rbh[e]);
                }
                if (p.out_f32) {
                  *reinterpret_cast<float4*>(p.out_f32 + obase + j) = make_float4(f[0], f[1], f[2], f[3]);
                  *reinterpret_cast<float4*>(p.out_f32 + obase + j + 4) = make_float4(f[4], f[5], f[6], f[7]);
                }
                if (p.out_bf16) {
                  __align__(16) __nv_bfloat16 ob[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) ob[e] = __float2bfloat16(f[e]);
                  *reinterpret_cast<uint4*>(p.out_bf16 + obase + j) = *reinterpret_cast<const uint4*>(ob);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[a]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace md
