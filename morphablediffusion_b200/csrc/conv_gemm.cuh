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

// Development-only phase stamps (build with -DMD_KPROF): globaltimer per CTA and phase, read back with md_debug_kprof.
#ifdef MD_KPROF
__device__ __forceinline__ void kprof(unsigned long long* buf, int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (blockIdx.x < 160) buf[blockIdx.x * 32 + slot] = t;
}
#define KPROF(slot, cond) do { if (cond) kprof(p.kprof, slot); } while (0)
#else
#define KPROF(slot, cond) do { } while (0)
#endif

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kMaxTaps = 27;
constexpr int kKprofTile = 2;  // tile whose steady-state phases the MD_KPROF build stamps


// Division by a launch-time constant as multiply + shift (exact for n < 2^31): the per-tile coordinate arithmetic runs
// in every one of the 18 warps, so runtime integer division there costs more issue slots than the epilogue's payload.
struct FastDiv {
  uint32_t mul, shift, d;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = static_cast<uint32_t>(d < 1 ? 1 : d);
  uint32_t s = 0;
  while ((1u << s) < f.d) ++s;
  f.shift = 31 + s;
  f.mul = static_cast<uint32_t>(((1ull << f.shift) + f.d - 1) / f.d);
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
  return static_cast<int>((static_cast<unsigned long long>(static_cast<uint32_t>(n)) * f.mul) >> f.shift);
}
__device__ __forceinline__ void fdivmod(int n, const FastDiv& f, int& q, int& r) {
  q = fdiv(n, f);
  r = n - q * static_cast<int>(f.d);
}

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
  float* col_stats; // optional [B][stats_ld][2]: per-(sample, channel) sum / sum of squares of the fp32 result
  int stats_ld;
  int contig;       // contiguous tile run per CTA (see kernel)
  int ksplit;       // split-K factor of the SPLIT tiles: work item = (tile, split); partial accumulators meet in split_ws,
                    // the last arriving warp of every (tile, epilogue-warp) pair reduces them and runs the real epilogue
  int split_tiles;  // the first split_tiles tiles (CTA-pair kernel: pair tiles) are split ksplit ways and their items come
                    // FIRST in the schedule, the remaining tiles run unsplit behind them.  All tiles = tile-starved
                    // problems; 0 = no split; in between = the partial wave of a persistent grid is split along K to
                    // fill the machine, and running it first hides the partial-sum exchange behind the whole waves
  int split_items;  // split_tiles * ksplit
  int total_items;  // split_items + (tiles - split_tiles)
  float* split_ws;  // [split tiles (x2 in the pair kernel)][ksplit][128][BN] fp32
  int* split_cnt;   // [split tiles (x2)][EPI_WARPS] arrival counters (zero on entry, reset by the last arrival)
  int isx, isy, isz; // input coordinate = output-tile coordinate * is + tap offset (strided convolution via TMA element strides)
  FastDiv fd_ksplit, fd_ntiles, fd_nxb, fd_nyb, fd_nzb;
  int epi_tma;  // 1: the output leaves through TMA stores of the per-warp staging tile (no fused statistics, one output,
                // plain output geometry); qorg = position of an epilogue warp's 32 rows inside the tile box, per lane quarter
  int16_t qorg[4][4];
  // GroupNorm tail (optional, needs col_stats and a bf16 output): after its last tile every CTA waits on a grid-wide
  // counter, then the whole grid normalises + activates the rows it just produced into gn_out (bf16 [rows][N]) with the
  // statistics its own epilogues accumulated: the GroupNorm between two convolutions costs no launch of its own
  __nv_bfloat16* gn_out;
  const float* gn_gamma; const float* gn_beta;
  int gn_groups, gn_act, gn_rows;   // gn_rows = output rows per sample
  float gn_eps;
  int* gn_bar;                      // grid barrier counter, zero on entry
  int dbg;      // development only (MD_GEMM_DBG): 1 = skip the epilogue work, 2 = skip the MMAs, 4 = skip the TMA loads;
                // results are garbage, the timings isolate which pipeline stage bounds a shape
  int cg2;      // 1: CTA-pair kernel; a work item is an (M-tile pair, N tile) and this CTA owns M tile 2*pair + rank
  int m_pairs;  // ceil(m_tiles / 2)
#ifdef MD_KPROF
  unsigned long long* kprof;  // [160 CTAs][32 slots] phase stamps
#endif
};

// work item -> (n tile, box coordinates, K split)
struct TileCoord {
  int tile, sp, ns, n_tile, xb, yb, zb, bblk;   // ns: number of K splits of this tile (1 or p.ksplit)
};
// work item -> (tile, split index, split count)
__device__ __forceinline__ void decode_split(const ConvGemmParams& p, int item, int& tile, int& sp, int& ns) {
  if (item < p.split_items) {
    fdivmod(item, p.fd_ksplit, tile, sp);
    ns = p.ksplit;
  } else {
    tile = p.split_tiles + (item - p.split_items); sp = 0; ns = 1;
  }
}
// K-block range [kb0, kb1) of split sp out of ns
__device__ __forceinline__ void split_krange(const ConvGemmParams& p, int kblocks, int sp, int ns, int& kb0, int& kb1) {
  if (ns == 1) { kb0 = 0; kb1 = kblocks; }
  else { kb0 = fdiv(kblocks * sp, p.fd_ksplit); kb1 = fdiv(kblocks * (sp + 1), p.fd_ksplit); }
}
__device__ __forceinline__ TileCoord decode_item(const ConvGemmParams& p, int item) {
  TileCoord t;
  decode_split(p, item, t.tile, t.sp, t.ns);
  int m;
  fdivmod(t.tile, p.fd_ntiles, m, t.n_tile);
  if (p.cg2) m = 2 * m + static_cast<int>(cluster_ctarank());  // an odd tile count leaves the last pair's second
                                                              // tile out of range: zero-filled loads, masked stores
  int m2;
  fdivmod(m, p.fd_nxb, m2, t.xb);
  fdivmod(m2, p.fd_nyb, m, t.yb);
  fdivmod(m, p.fd_nzb, t.bblk, t.zb);
  return t;
}

__device__ __forceinline__ float act_silu(float x) { return silu_fast(x); }
__device__ __forceinline__ float act_gelu(float x) { return gelu_fast(x); }
__device__ __forceinline__ float act_quickgelu(float x) { return x * rcp_approx(1.f + ex2_approx(-1.702f * 1.4426950408889634f * x)); }

#ifndef MD_EPI_WARPS
#define MD_EPI_WARPS 16
#endif
constexpr int kEpiWarps = MD_EPI_WARPS;   // kEpiWarps/4 warps per TMEM lane quarter; each owns every kCStride-th chunk
constexpr int kCStride = kEpiWarps / 4;   // 16-column chunks between two chunks of the same warp
static_assert(kEpiWarps % 4 == 0 && kEpiWarps >= 4 && kEpiWarps <= 16, "epilogue warps come in groups of four");
constexpr int kChunk = 16;      // accumulator columns per epilogue chunk (tcgen05.ld 32x32b.x16)

template <int BN, int STAGES>
struct ConvGemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStageOffset = STAGES * kStageBytes;            // kEpiWarps x 2 KB staging tiles (32 rows x 16 fp32)
  static constexpr int kBarOffset = kStageOffset + kEpiWarps * 2048;
  static constexpr int kTotal = kBarOffset + 256;
};

// 32 lanes x 16 columns of 32-bit accumulators <-> registers
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// Adds a warp's running GroupNorm partial sums to global memory.  Phase-2 layout: lane = (row-in-group-of-8 << 2) |
// 16-byte chunk, so lanes l, l^4, l^8, l^16 own the same 4 columns.
template <int NCH>
__device__ __forceinline__ void flush_col_stats(const ConvGemmParams& p, int lane, int sample, int n_out0, int n_limit,
                                                int c_begin, float (&st1)[NCH][4], float (&st2)[NCH][4]) {
  const int prow = lane >> 2, pchunk = lane & 3;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        st1[k][e] += __shfl_xor_sync(0xffffffff, st1[k][e], o);
        st2[k][e] += __shfl_xor_sync(0xffffffff, st2[k][e], o);
      }
    }
    const int col = n_out0 + (c_begin + k * kCStride) * kChunk + pchunk * 4;
    if (prow == 0 && sample >= 0 && col < n_limit) {
      float* cs = p.col_stats + (static_cast<long long>(sample) * p.stats_ld + col) * 2;
#pragma unroll
      for (int e = 0; e < 4; ++e) { atomicAdd(cs + 2 * e, st1[k][e]); atomicAdd(cs + 2 * e + 1, st2[k][e]); }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { st1[k][e] = 0.f; st2[k][e] = 0.f; }
  }
}

// Epilogue of one output tile for one warp: its 32 accumulator rows x the 16-column chunks c_begin, c_begin+4, ...
// Specialised on the residual kind / per-sample vector / presence of an activation so the inner loops are branch-free.
//   phase 0: issue the chunk's global loads (bias, per-sample vector, residual) in the coalesced phase-2 layout;
//   phase 1 (row owner): TMEM -> registers, scale (+bias, GEGLU), swizzled store into a private 32x16 staging tile;
//   phase 2 (coalesced; one warp instruction = 8 rows x 64 B): + bias, + per-sample vector, activation, + residual,
//            fp32 / bf16 stores, per-(sample, channel) sum / sum-of-squares for the next GroupNorm.
// ri[it] = (output row offset in elements or -1, sample) of phase-2 row it*8 + (lane>>2), fetched once per tile.
template <int BN, int RES, bool RV, int NCH, bool STATS, int ACTV>
__device__ __forceinline__ void epilogue_tile(const ConvGemmParams& p, uint32_t taddr, float* stage, int lane, int n_tile,
                                              const int2 (&ri)[4], int c_begin, float (&st1)[NCH][4],
                                              float (&st2)[NCH][4], bool kp = false) {
  const int prow = lane >> 2;   // phase-2 row within a group of 8
  const int pchunk = lane & 3;  // phase-2 16-byte chunk within the 64-byte row
  constexpr bool geglu = (ACTV == 2);
  constexpr int HALF = BN / 2;
  const int out_cols = geglu ? HALF : BN;
  const int n_limit = geglu ? p.N / 2 : p.N;
  const int n_out0 = n_tile * out_cols;
  const float* const bias = p.bias;
  float* const out_f32 = p.out_f32;
  __nv_bfloat16* const out_bf16 = p.out_bf16;
  const bool scaled = p.out_scale != 1.f;
  const float* const st_rd = stage + prow * 16;  // phase-2 read base; row it*8+prow -> + it*128, chunk swizzled below
  const int sw = (prow >> 1) & 3;                // ((it*8 + prow) >> 1) & 3 == (prow >> 1) & 3
  float* const st_wr = stage + lane * 16;
  const int swl = (lane >> 1) & 3;
  int ci = 0;
#pragma unroll 1
  for (int c = c_begin; c < out_cols / kChunk; c += kCStride, ++ci) {
    const int col = n_out0 + c * kChunk + pchunk * 4;
    const bool col_ok = col < n_limit;
    const int col_safe = col_ok ? col : 0;
    KPROF(21, kp && ci == 0);
    // ---- phase 0: all global loads of this chunk in flight together (invalid rows read a safe address)
    float4 rv4[4];
    float4 rs4[4];
    uint2 rb2[4];
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias && !geglu) b4 = __ldg(reinterpret_cast<const float4*>(bias + col_safe));
    if (RV || RES != 0) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const bool ok = ri[it].x >= 0;
        const long long ro = ok ? ri[it].x : 0;
        if (RV) rv4[it] = __ldg(reinterpret_cast<const float4*>(p.rowvec + static_cast<long long>(ok ? ri[it].y : 0) * p.rowvec_ld + col_safe));
        if (RES == 1) rs4[it] = *reinterpret_cast<const float4*>(p.res_f32 + ro + col_safe);
        if (RES == 2) rb2[it] = *reinterpret_cast<const uint2*>(p.res_bf16 + ro + col_safe);
      }
    }
    // ---- phase 1: accumulator chunk -> staging tile (16-byte chunk j of row `lane` lands at chunk j ^ ((lane>>1)&3))
    {
      uint32_t v[16];
      if (geglu) {
        const int nb = n_tile * BN + c * kChunk;  // packed-row index of the value half
        float4 bv4[4], bg4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bv4[j] = bias ? __ldg(reinterpret_cast<const float4*>(bias + nb) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          bg4[j] = bias ? __ldg(reinterpret_cast<const float4*>(bias + nb + HALF) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_32x16(taddr + c * kChunk, v);
        uint32_t g[16];
        tmem_ld_32x16(taddr + HALF + c * kChunk, g);
        tc_wait_ld();
        KPROF(22, kp && ci == 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 bv = bv4[j], bg = bg4[j];
          float4 o;
          o.x = (__uint_as_float(v[4 * j]) + bv.x) * act_gelu(__uint_as_float(g[4 * j]) + bg.x);
          o.y = (__uint_as_float(v[4 * j + 1]) + bv.y) * act_gelu(__uint_as_float(g[4 * j + 1]) + bg.y);
          o.z = (__uint_as_float(v[4 * j + 2]) + bv.z) * act_gelu(__uint_as_float(g[4 * j + 2]) + bg.z);
          o.w = (__uint_as_float(v[4 * j + 3]) + bv.w) * act_gelu(__uint_as_float(g[4 * j + 3]) + bg.w);
          *reinterpret_cast<float4*>(st_wr + ((j ^ swl) << 2)) = o;
        }
      } else {
        tmem_ld_32x16(taddr + c * kChunk, v);
        tc_wait_ld();
        KPROF(22, kp && ci == 0);
        if (scaled) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * p.out_scale);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(st_wr + ((j ^ swl) << 2)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
    __syncwarp();
    KPROF(23, kp && ci == 0);
    // ---- phase 2: coalesced finish; all four staged rows are fetched before the arithmetic
    float4 t4[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) t4[it] = *reinterpret_cast<const float4*>(st_rd + it * 128 + ((pchunk ^ sw) << 2));
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const bool ok = (ri[it].x >= 0) && col_ok;
      float4 v4 = t4[it];
      v4.x += b4.x; v4.y += b4.y; v4.z += b4.z; v4.w += b4.w;
      if (RV) { v4.x += rv4[it].x; v4.y += rv4[it].y; v4.z += rv4[it].z; v4.w += rv4[it].w; }
      if (ACTV == 1) {
        if (p.act == ACT_SILU) {
          v4.x = act_silu(v4.x); v4.y = act_silu(v4.y); v4.z = act_silu(v4.z); v4.w = act_silu(v4.w);
        } else if (p.act == ACT_RELU) {
          v4.x = fmaxf(v4.x, 0.f); v4.y = fmaxf(v4.y, 0.f); v4.z = fmaxf(v4.z, 0.f); v4.w = fmaxf(v4.w, 0.f);
        } else if (p.act == ACT_GELU) {
          v4.x = act_gelu(v4.x); v4.y = act_gelu(v4.y); v4.z = act_gelu(v4.z); v4.w = act_gelu(v4.w);
        } else if (p.act == ACT_QUICKGELU) {
          v4.x = act_quickgelu(v4.x); v4.y = act_quickgelu(v4.y); v4.z = act_quickgelu(v4.z); v4.w = act_quickgelu(v4.w);
        }
      }
      if (RES == 1) { v4.x += rs4[it].x; v4.y += rs4[it].y; v4.z += rs4[it].z; v4.w += rs4[it].w; }
      if (RES == 2) {
        v4.x += __uint_as_float(rb2[it].x << 16); v4.y += __uint_as_float(rb2[it].x & 0xffff0000u);
        v4.z += __uint_as_float(rb2[it].y << 16); v4.w += __uint_as_float(rb2[it].y & 0xffff0000u);
      }
      const long long off = static_cast<long long>(ri[it].x) + col;
      if (out_f32 && ok) *reinterpret_cast<float4*>(out_f32 + off) = v4;
      if (out_bf16 && ok) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v4.x, v4.y);
        __nv_bfloat162 hi = __floats2bfloat162_rn(v4.z, v4.w);
        uint2 o2;
        o2.x = *reinterpret_cast<uint32_t*>(&lo);
        o2.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(out_bf16 + off) = o2;
      }
      if (STATS && ok) {
        s1[0] += v4.x; s1[1] += v4.y; s1[2] += v4.z; s1[3] += v4.w;
        s2[0] += v4.x * v4.x; s2[1] += v4.y * v4.y; s2[2] += v4.z * v4.z; s2[3] += v4.w * v4.w;
      }
    }
    if (STATS) {
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        if (k == ci) {
#pragma unroll
          for (int e = 0; e < 4; ++e) { st1[k][e] += s1[e]; st2[k][e] += s2[e]; }
        }
      }
    }
    __syncwarp();
    KPROF(24, kp && ci == 0);
    KPROF(25, kp && ci == 1);
  }
}

// TMA-store epilogue of one output tile for one warp (experimental, p.epi_tma): everything is finished in the
// row-owner layout the accumulator comes in (lane = output row): bias, per-sample vector and fp32 residual are read as
// four 16-byte loads per lane, the result is written into the warp's swizzled staging tile (exactly the layout of a
// SWIZZLE_64B / SWIZZLE_32B TMA box of 16 columns x 32 rows) and one lane hands it to the TMA.  No per-element global
// addressing, no second pass over shared memory: ~1/3 of the instructions of the coalesced epilogue.
template <int BN, int RES, bool RV, int ACTV>
__device__ __forceinline__ void epilogue_tile_tma(const ConvGemmParams& p, const CUtensorMap* tmO, uint32_t taddr,
                                                  float* stage, int lane, int n_tile, long long orow, int bv,
                                                  int c_begin, int x0, int y0, int z0, int b0) {
  constexpr bool geglu = (ACTV == 2);
  constexpr int HALF = BN / 2;
  constexpr int out_cols = geglu ? HALF : BN;
  const int n_limit = geglu ? p.N / 2 : p.N;
  const int n_out0 = n_tile * out_cols;
  const bool f32out = p.out_f32 != nullptr;
  const bool valid = bv >= 0;
  const long long ro = valid ? orow : 0;
  const float* rvp = RV ? p.rowvec + static_cast<long long>(valid ? bv : 0) * p.rowvec_ld : nullptr;
  const bool scaled = p.out_scale != 1.f;
#pragma unroll 1
  for (int c = c_begin; c < out_cols / kChunk; c += kCStride) {
    const int col = n_out0 + c * kChunk;
    if (col >= n_limit) break;
    float o[16];
    if constexpr (geglu) {
      // value | gate halves of the accumulator tile; packed bias rows follow the same interleave
      const int nb = n_tile * BN + c * kChunk;
      float4 bv4[4], bg4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bv4[j] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + nb) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        bg4[j] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + nb + HALF) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      uint32_t v[16], g[16];
      tmem_ld_32x16(taddr + c * kChunk, v);
      tmem_ld_32x16(taddr + HALF + c * kChunk, g);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[4 * j + 0] = (__uint_as_float(v[4 * j + 0]) + bv4[j].x) * act_gelu(__uint_as_float(g[4 * j + 0]) + bg4[j].x);
        o[4 * j + 1] = (__uint_as_float(v[4 * j + 1]) + bv4[j].y) * act_gelu(__uint_as_float(g[4 * j + 1]) + bg4[j].y);
        o[4 * j + 2] = (__uint_as_float(v[4 * j + 2]) + bv4[j].z) * act_gelu(__uint_as_float(g[4 * j + 2]) + bg4[j].z);
        o[4 * j + 3] = (__uint_as_float(v[4 * j + 3]) + bv4[j].w) * act_gelu(__uint_as_float(g[4 * j + 3]) + bg4[j].w);
      }
    } else {
    const bool full = col + kChunk <= p.N;  // a ragged last chunk reads bias / residual element-safe below
    float4 b4[4], r4[4], v4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cj = (full || col + 4 * j + 4 <= p.N) ? col + 4 * j : 0;
      b4[j] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + cj)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (RV) v4[j] = __ldg(reinterpret_cast<const float4*>(rvp + cj));
      if (RES == 1) r4[j] = *reinterpret_cast<const float4*>(p.res_f32 + ro + cj);
    }
    uint32_t v[16];
    tmem_ld_32x16(taddr + c * kChunk, v);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float e[4] = {__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                    __uint_as_float(v[4 * j + 3])};
      const float bb[4] = {b4[j].x, b4[j].y, b4[j].z, b4[j].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float y = scaled ? e[t] * p.out_scale : e[t];
        y += bb[t];
        if (RV) y += (t == 0 ? v4[j].x : t == 1 ? v4[j].y : t == 2 ? v4[j].z : v4[j].w);
        if (ACTV == 1) {
          if (p.act == ACT_SILU) y = act_silu(y);
          else if (p.act == ACT_RELU) y = fmaxf(y, 0.f);
          else if (p.act == ACT_GELU) y = act_gelu(y);
          else if (p.act == ACT_QUICKGELU) y = act_quickgelu(y);
        }
        if (RES == 1) y += (t == 0 ? r4[j].x : t == 1 ? r4[j].y : t == 2 ? r4[j].z : r4[j].w);
        o[4 * j + t] = y;
      }
    }
    }  // !geglu
    // the previous chunk's store must have finished reading the staging tile
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
    if (f32out) {
      float* st_wr = stage + lane * 16;
      const int swl = (lane >> 1) & 3;  // SWIZZLE_64B: 16-byte chunk index ^= address bits 7..8
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(st_wr + ((j ^ swl) << 2)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    } else {
      uint8_t* st_wr = reinterpret_cast<uint8_t*>(stage) + lane * 32;
      const int sw1 = (lane >> 2) & 1;  // SWIZZLE_32B: 16-byte chunk index ^= address bit 7
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          __nv_bfloat162 h = __floats2bfloat162_rn(o[8 * j + 2 * t], o[8 * j + 2 * t + 1]);
          w[t] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(st_wr + ((j ^ sw1) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_5d(tmO, stage, col, x0, y0, z0, b0);
      tma_store_commit();
    }
  }
}

// Pulls `bytes` (multiple of 16) at a 16-byte aligned global address into L2 without occupying registers.
__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// Epilogue warps: loop over this CTA's work items (same schedule as the producer / MMA warps).
template <int BN, int STAGES, int RES, bool RV, bool STATS, int ACTV, int STAGE_OFFSET>
__device__ __forceinline__ void epilogue_loop(const ConvGemmParams& p, const CUtensorMap* tmO, uint8_t* smem,
                                              uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t tmem_base, int warp,
                                              int lane, int tile_begin, int tile_end, int tile_step) {
  const int ew = warp - 2;
  const int q = warp & 3;        // TMEM lane quarter this warp may read
  const int c_begin = ew >> 2;   // this warp owns chunks c_begin, c_begin + 4, ...
  const int r = q * 32 + lane;
  float* stage = reinterpret_cast<float*>(smem + STAGE_OFFSET) + ew * 512;
  const int out_cols_t = (ACTV == 2) ? BN / 2 : BN;
  const int n_limit_t = (ACTV == 2) ? p.N / 2 : p.N;
  constexpr int NCH = (BN / kChunk + kCStride - 1) / kCStride;  // chunks one warp owns per tile
  float st1[NCH][4], st2[NCH][4];
#pragma unroll
  for (int k = 0; k < NCH; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) { st1[k][e] = 0.f; st2[k][e] = 0.f; }
  int st_sample = -1, st_ntile = -1;
  // row -> position inside the TMA box (tile independent)
  int rr = r;
  const int ix = rr % p.bw; rr /= p.bw;
  const int iy = rr % p.bh; rr /= p.bh;
  const int iz = rr % p.bd; rr /= p.bd;
  const long long plane = static_cast<long long>(p.OW) * p.OH;
  // The residual rows of a tile are read chunk by chunk in the coalesced phase, each read a full memory round trip
  // on the warp's critical path: one tile ahead, the first warp of every lane quarter pulls its 32 residual rows
  // (this tile's column range) into L2 so those reads become L2 hits.
  const bool do_prefetch = (c_begin == 0) && (p.res_f32 != nullptr || p.res_bf16 != nullptr);
  auto prefetch_residual = [&](int item) {
    if (!do_prefetch || item >= tile_end) return;
    const TileCoord t = decode_item(p, item);
    if (t.ns != 1) return;   // split tiles: only the last arrival reads the residual
    const int x = t.xb * p.bw + ix, y = t.yb * p.bh + iy, z = t.zb * p.bd + iz, b = t.bblk * p.bb + rr;
    const int n0 = t.n_tile * out_cols_t;
    const int cols = min(out_cols_t, n_limit_t - n0);
    if ((x < p.W) && (y < p.H) && (z < p.D) && (b < p.B) && cols > 0) {
      const long long off =
          ((static_cast<long long>(b) * p.OD + (z * p.osz + p.opz)) * plane + (y * p.osy + p.opy) * p.OW +
           (x * p.osx + p.opx)) * p.ldo + n0;
      if (p.res_f32) prefetch_l2_bulk(p.res_f32 + off, static_cast<uint32_t>(cols) * 4u);
      else prefetch_l2_bulk(p.res_bf16 + off, static_cast<uint32_t>(cols) * 2u);
    }
  };
  prefetch_residual(tile_begin);
  int lt = 0;
  for (int item = tile_begin; item < tile_end; item += tile_step, ++lt) {
    prefetch_residual(item + tile_step);
    const TileCoord tc = decode_item(p, item);
    const int tile = tc.tile, sp = tc.sp, n_tile = tc.n_tile;
    const int a = lt & 1;
    const uint32_t aph = (lt >> 1) & 1;
    const int x = tc.xb * p.bw + ix, y = tc.yb * p.bh + iy, z = tc.zb * p.bd + iz, b = tc.bblk * p.bb + rr;
    const bool valid = (x < p.W) && (y < p.H) && (z < p.D) && (b < p.B);
    const long long orow =
        ((static_cast<long long>(b) * p.OD + (z * p.osz + p.opz)) * plane + (y * p.osy + p.opy) * p.OW +
         (x * p.osx + p.opx)) * p.ldo;
    const int bv = valid ? b : -1;
    // sample of the warp's rows (fused GroupNorm statistics: the host guarantees one sample per warp there)
    const int warp_sample = __reduce_max_sync(0xffffffff, bv);
    if (STATS && (warp_sample != st_sample || n_tile != st_ntile)) {
      if (st_sample >= 0) flush_col_stats<NCH>(p, lane, st_sample, st_ntile * out_cols_t, n_limit_t, c_begin, st1, st2);
      st_sample = warp_sample; st_ntile = n_tile;
    }
    __syncwarp();

    KPROF(16, lt == kKprofTile && warp == 2 && lane == 0);
    mbar_wait(&tmem_full[a], aph);
    tc_fence_after();
    KPROF(6, lt == 0 && warp == 2 && lane == 0);
    KPROF(17, lt == kKprofTile && warp == 2 && lane == 0);
    // (output row offset or -1, sample) of the four rows this lane finishes in the coalesced phase: they belong to
    // lanes it*8 + (lane>>2) of this warp
    int2 ri[4];
    {
      const int my_off = valid ? static_cast<int>(orow) : -1;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        ri[it].x = __shfl_sync(0xffffffff, my_off, it * 8 + (lane >> 2));
        ri[it].y = __shfl_sync(0xffffffff, bv, it * 8 + (lane >> 2));
      }
    }
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN;
    bool run_epilogue = true;
    if (tc.ns > 1) {
      // split-K: publish this warp's partial accumulators, take a ticket; the last arrival for this (tile, warp)
      // adds the other splits' partials back into TMEM and then runs the ordinary epilogue on the full sums.
      constexpr int acc_chunks = BN / kChunk;
      // slot of this CTA's tile among the split tiles (both CTAs of a pair work on the same pair tile)
      const int ws_idx = p.cg2 ? 2 * tile + static_cast<int>(cluster_ctarank()) : tile;
      // Exchange through ONE fp32 accumulation tile per split tile: every split adds its partial with vector
      // red.global.add (fire and forget, no round trip; four floats per lane and 512 contiguous bytes per warp
      // instruction: layout [16-column chunk][4-column group][128 rows][4]), the last arrival reads the sums once,
      // restores the zeros and puts them into TMEM.  Cost per CTA is independent of the split count; the tile is all-zero
      // between launches (allocated zeroed, every element that is added to is read and re-zeroed by the last arrival).
      float* ws_row = p.split_ws + static_cast<size_t>(ws_idx) * (128 * BN) + r * 4;
      for (int c = c_begin; c < acc_chunks; c += kCStride) {
        uint32_t v[16];
        tmem_ld_32x16(taddr + c * kChunk, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          red_add_v4_f32(ws_row + (c * 4 + j) * 512, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
      __threadfence();
      __syncwarp();
      int last = 0;
      if (lane == 0) {
        int* cnt = p.split_cnt + ws_idx * kEpiWarps + ew;
        last = (atomicAdd(cnt, 1) == p.ksplit - 1) ? 1 : 0;
        if (last) *cnt = 0;  // ready for the next launch
      }
      last = __shfl_sync(0xffffffff, last, 0);
      run_epilogue = last != 0;
      KPROF(7, lt == 0 && warp == 2 && lane == 0);
      if (run_epilogue) {
        // every split has published before taking its ticket and the loads below depend on the ticket; they bypass L1
        // (ld.cg), so no second fence is needed.
        // Two chunks (32 registers) are requested per memory round trip.
        for (int c0 = c_begin; c0 < acc_chunks; c0 += 2 * kCStride) {
          const int c1 = c0 + kCStride;
          const bool two = c1 < acc_chunks;
          float4 t0[4], t1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) t0[j] = __ldcg(reinterpret_cast<const float4*>(ws_row + (c0 * 4 + j) * 512));
          if (two) {
#pragma unroll
            for (int j = 0; j < 4; ++j) t1[j] = __ldcg(reinterpret_cast<const float4*>(ws_row + (c1 * 4 + j) * 512));
          }
          uint32_t v[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[4 * j] = __float_as_uint(t0[j].x); v[4 * j + 1] = __float_as_uint(t0[j].y);
            v[4 * j + 2] = __float_as_uint(t0[j].z); v[4 * j + 3] = __float_as_uint(t0[j].w);
            __stcg(reinterpret_cast<float4*>(ws_row + (c0 * 4 + j) * 512), make_float4(0.f, 0.f, 0.f, 0.f));
          }
          tmem_st_32x16(taddr + c0 * kChunk, v);
          if (two) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[4 * j] = __float_as_uint(t1[j].x); v[4 * j + 1] = __float_as_uint(t1[j].y);
              v[4 * j + 2] = __float_as_uint(t1[j].z); v[4 * j + 3] = __float_as_uint(t1[j].w);
              __stcg(reinterpret_cast<float4*>(ws_row + (c1 * 4 + j) * 512), make_float4(0.f, 0.f, 0.f, 0.f));
            }
            tmem_st_32x16(taddr + c1 * kChunk, v);
          }
        }
        tc_wait_st();
        KPROF(8, lt == 0 && warp == 2 && lane == 0);
      }
    }
#ifdef MD_KPROF
    const bool kpf = lt == kKprofTile && warp == 2 && lane == 0;
#else
    constexpr bool kpf = false;
#endif
    if (run_epilogue && !(p.dbg & 1)) {
      if constexpr (!STATS && RES != 2) {
        if (p.epi_tma) {
          const int16_t* qo = p.qorg[q];
          epilogue_tile_tma<BN, RES, RV, ACTV>(p, tmO, taddr, stage, lane, n_tile, orow, bv, c_begin,
                                               tc.xb * p.bw + qo[0], tc.yb * p.bh + qo[1], tc.zb * p.bd + qo[2],
                                               tc.bblk * p.bb + qo[3]);
        } else {
          epilogue_tile<BN, RES, RV, NCH, STATS, ACTV>(p, taddr, stage, lane, n_tile, ri, c_begin, st1, st2, kpf);
        }
      } else {
        epilogue_tile<BN, RES, RV, NCH, STATS, ACTV>(p, taddr, stage, lane, n_tile, ri, c_begin, st1, st2, kpf);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (p.cg2) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[a]), 0));  // the pair leader issues the MMAs
      else mbar_arrive(&tmem_empty[a]);
    }
    KPROF(9, lt == 0 && warp == 2 && lane == 0);
    KPROF(18, lt == kKprofTile && warp == 2 && lane == 0);
    KPROF(10, warp == 2 && lane == 0);
  }
  if (STATS && st_sample >= 0)
    flush_col_stats<NCH>(p, lane, st_sample, st_ntile * out_cols_t, n_limit_t, c_begin, st1, st2);
  if (p.epi_tma && lane == 0) tma_store_wait_all();  // the CTA's shared memory must outlive its bulk stores
}

// GroupNorm tail of a statistics-carrying launch (p.gn_out != nullptr), run by all threads of every CTA after the tile loop.
// 1. grid barrier: the CTA's output stores and statistics atomics are fenced, thread 0 arrives on p.gn_bar and spins until
//    every CTA of the (fully resident: grid <= SMs, one CTA per SM) grid has arrived.  Only launches of the UNet's main
//    stream use the tail, so a CTA that is still waiting for an SM is at most delayed by side-stream kernels, which
//    never wait on anything of ours.
// 2. the rows are split evenly over the CTAs; per sample a CTA touches, it reduces the per-channel sums to group
//    mean / rstd (fp32 sums, double only for E[x^2] - mean^2, as gn_apply_fused_kernel), builds scale / shift in shared
//    memory and streams its rows: 8-byte loads of the bf16 pre-norm tensor, normalise, activation, 8-byte stores.
struct GnTailArgs {
  const __nv_bfloat16* out_bf16; __nv_bfloat16* gn_out; const float* col_stats; const float* gn_gamma; const float* gn_beta;
  int* gn_bar; int N, B, ldo, stats_ld, gn_groups, gn_act, gn_rows; float gn_eps;
};
static __device__ __noinline__ void gn_tail(const GnTailArgs p, uint8_t* smem) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(p.gn_bar, 1);
    const unsigned target = gridDim.x;
    const long long t0 = clock64();
    while (ld_acquire_gpu_u32(p.gn_bar) < target) {
      __nanosleep(32);
      if (clock64() - t0 > (1LL << 33)) break;   // ~4 s: never hang the device on a scheduling assumption gone wrong
    }
  }
  __syncthreads();
  const int C = p.N;
  const int CQ = C >> 2;
  const int G = p.gn_groups, cpg = C / G;
  const int nthreads = blockDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = nthreads >> 5;
  float* chs = reinterpret_cast<float*>(smem);          // [C][2] per-channel (sum, sum of squares) -> (scale, shift)
  float* gsm = chs + 2 * C;                             // [G][2] mean, rstd
  const long long rps = p.gn_rows;
  const long long rows_total = static_cast<long long>(p.B) * rps;
  const long long rpc = (rows_total + gridDim.x - 1) / gridDim.x;
  const long long r0 = static_cast<long long>(blockIdx.x) * rpc;
  const long long r1 = min(rows_total, r0 + rpc);
  if (r0 >= r1) return;
  const int R = nthreads / CQ;                          // row phases (C <= 1280 -> CQ <= 320 <= threads)
  const int cq = threadIdx.x % CQ, rsub = threadIdx.x / CQ;
  const double inv_n = 1.0 / (static_cast<double>(rps) * cpg);
  for (long long b = r0 / rps; b * rps < r1; ++b) {
    const long long s0 = max(r0, b * rps), s1 = min(r1, (b + 1) * rps);
    for (int ch = threadIdx.x; ch < C; ch += nthreads) {
      const float2 sq = __ldcg(reinterpret_cast<const float2*>(p.col_stats + (b * p.stats_ld + ch) * 2));
      chs[2 * ch] = sq.x; chs[2 * ch + 1] = sq.y;
    }
    __syncthreads();
    for (int g = warp; g < G; g += nwarps) {
      float a = 0.f, q = 0.f;
      for (int ch = g * cpg + lane; ch < (g + 1) * cpg; ch += 32) { a += chs[2 * ch]; q += chs[2 * ch + 1]; }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffff, a, o);
        q += __shfl_xor_sync(0xffffffff, q, o);
      }
      if (lane == 0) {
        const double mean = static_cast<double>(a) * inv_n;
        double var = static_cast<double>(q) * inv_n - mean * mean;
        if (var < 0.0) var = 0.0;
        gsm[2 * g] = static_cast<float>(mean);
        gsm[2 * g + 1] = rsqrtf(static_cast<float>(var) + p.gn_eps);
      }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += nthreads) {
      const int g = ch / cpg;
      const float sc = __ldg(p.gn_gamma + ch) * gsm[2 * g + 1];
      chs[2 * ch] = sc;
      chs[2 * ch + 1] = __ldg(p.gn_beta + ch) - gsm[2 * g] * sc;
    }
    __syncthreads();
    if (rsub < R) {
      const float4 s01 = *reinterpret_cast<const float4*>(chs + 8 * cq);       // sc0 sh0 sc1 sh1
      const float4 s23 = *reinterpret_cast<const float4*>(chs + 8 * cq + 4);   // sc2 sh2 sc3 sh3
      const __nv_bfloat16* src = p.out_bf16 + cq * 4;
      __nv_bfloat16* dst = p.gn_out + cq * 4;
      const int act = p.gn_act;
      auto one = [&](const uint2 u) {
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        float y0 = __low2float(lo) * s01.x + s01.y, y1 = __high2float(lo) * s01.z + s01.w;
        float y2 = __low2float(hi) * s23.x + s23.y, y3 = __high2float(hi) * s23.z + s23.w;
        if (act == ACT_SILU) { y0 = act_silu(y0); y1 = act_silu(y1); y2 = act_silu(y2); y3 = act_silu(y3); }
        else if (act == ACT_RELU) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
        __nv_bfloat162 a = __floats2bfloat162_rn(y0, y1), c2 = __floats2bfloat162_rn(y2, y3);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&c2);
        return o;
      };
      long long r = s0 + rsub;
      for (; r + 3 * R < s1; r += 4 * R) {   // four rows in flight per thread
        uint2 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = __ldcg(reinterpret_cast<const uint2*>(src + (r + k * R) * p.ldo));
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<uint2*>(dst + (r + k * R) * C) = one(u[k]);
      }
      for (; r < s1; r += R) *reinterpret_cast<uint2*>(dst + r * C) = one(__ldcg(reinterpret_cast<const uint2*>(src + r * p.ldo)));
    }
    __syncthreads();
  }
}

// RES: 0 none / 1 fp32 / 2 bf16 residual; RV: per-sample additive vector; STATS: fused GroupNorm statistics; ACTV: 0 no
// activation / 1 SiLU, ReLU or GELU (p.act) / 2 GEGLU.  One kernel per combination keeps every instance a few thousand instructions, so the
// cold instruction fetches of these short launches stay small.
template <int BN, int STAGES, int RES, bool RV, bool STATS, int ACTV>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const ConvGemmParams p) {
  using S = ConvGemmSmem<BN, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  KPROF(0, threadIdx.x == 0);
  constexpr uint32_t kTmemCols = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.epi_tma) prefetch_tmap(&tmO);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);
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
  KPROF(1, threadIdx.x == 0);
  // barriers, TMEM and descriptors are set up: everything above overlapped the predecessor kernel's tail
  pdl_grid_sync();
  KPROF(2, threadIdx.x == 0);

  const int total_tiles = p.total_items;  // work items: (tile, K split)
  const int kblocks = p.ntaps * p.kblocks_per_tap;
  // Tile schedule: interleaved (tile = cta + i*grid) or, for tall single-N-tile problems, one contiguous run per CTA
  // (consecutive tiles then share a sample, which lets the epilogue keep GroupNorm partial sums in registers).
  const int per_cta = (total_tiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int tile_begin = p.contig ? static_cast<int>(blockIdx.x) * per_cta : static_cast<int>(blockIdx.x);
  const int tile_end = p.contig ? min(total_tiles, tile_begin + per_cta) : total_tiles;
  const int tile_step = p.contig ? 1 : static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp walks the schedule (warp-uniform control flow keeps addresses and coordinates in uniform
    // registers: no per-instruction R2UR election loops around UTMALDG); one elected lane issues.
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    int plt = 0;
    for (int item = tile_begin; item < tile_end; item += tile_step, ++plt) {
      const TileCoord tc = decode_item(p, item);
      int kb0, kb1;
      split_krange(p, kblocks, tc.sp, tc.ns, kb0, kb1);
      const int x0 = tc.xb * p.bw * p.isx, y0 = tc.yb * p.bh * p.isy, z0 = tc.zb * p.bd * p.isz, b0 = tc.bblk * p.bb;
      const int n0 = tc.n_tile * BN;
      int tap = kb0 / p.kblocks_per_tap, kc = kb0 - tap * p.kblocks_per_tap;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        KPROF(19, lane == 0 && plt == kKprofTile && kb == kb0);
        KPROF(20, lane == 0 && plt == kKprofTile && kb == kb1 - 1);
        uint8_t* sa = smem + s * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        if (elect_one()) {
          if (p.dbg & 4) {
            mbar_arrive(&full_bar[s]);
          } else {
            mbar_expect_tx(&full_bar[s], S::kStageBytes);
            tma_load_5d(sa, &tmA, &full_bar[s], kc * kBlockK, x0 + p.tdx[tap], y0 + p.tdy[tap], z0 + p.tdz[tap], b0);
            tma_load_2d(sb, &tmB, &full_bar[s], kb * kBlockK, n0);
          }
        }
        KPROF(3, lane == 0 && it == 0);
        if (++kc == p.kblocks_per_tap) { kc = 0; ++tap; }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform loop as well; the MMAs and their commits come from the same elected lane (tcgen05.commit tracks the
    // issuing thread's MMAs; elect.sync with a full mask always names the same lane).
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    int lt = 0;
    for (int item = tile_begin; item < tile_end; item += tile_step, ++lt) {
      int tile_, sp, ns, kb0, kb1;
      decode_split(p, item, tile_, sp, ns);
      split_krange(p, kblocks, sp, ns, kb0, kb1);
      const int nkb = kb1 - kb0;
      const int a = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      mbar_wait(&tmem_empty[a], aph ^ 1);
      tc_fence_after();
      KPROF(13, lane == 0 && lt == kKprofTile);
      const uint32_t d_tmem = tmem_base + a * BN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        KPROF(4, lane == 0 && it == 0);
        KPROF(5, lane == 0 && lt == 0 && kb == nkb - 1);
        KPROF(14, lane == 0 && lt == kKprofTile && kb == 0);
        const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
        const uint32_t sb = sa + S::kABytes;
        const uint64_t da = make_sw128_kmajor_desc(sa);
        const uint64_t db = make_sw128_kmajor_desc(sb);
        if (elect_one()) {
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // advance 16 bf16 = 32 B inside the 128-B swizzle atom: +2 in the (addr >> 4) field
              tc_mma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(&empty_bar[s]);
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one()) tc_commit(&tmem_full[a]);
      KPROF(15, lane == 0 && lt == kKprofTile);
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2 .. 17) =====================
    epilogue_loop<BN, STAGES, RES, RV, STATS, ACTV, S::kStageOffset>(p, &tmO, smem, tmem_full, tmem_empty, tmem_base, warp, lane,
                                                                    tile_begin, tile_end, tile_step);
  }

  tc_fence_before();
  __syncthreads();
  KPROF(11, threadIdx.x == 0);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
    KPROF(12, lane == 0);
  }
  if constexpr (STATS) {
    if (p.gn_out) {
      const GnTailArgs ga{p.out_bf16, p.gn_out, p.col_stats, p.gn_gamma, p.gn_beta, p.gn_bar, p.N, p.B, p.ldo, p.stats_ld,
                          p.gn_groups, p.gn_act, p.gn_rows, p.gn_eps};
      gn_tail(ga, smem);
    }
  }
}

// ------------------------------------------------------------------------------------------------ CTA-pair variant
// Two CTAs of a cluster compute one 256 x BN tile with cta_group::2 MMAs: each CTA stages its own 128 A rows and HALF
// of the B rows (BN/2 weight rows), so a pair pulls (256 + BN) operand rows from L2 for 256 x BN outputs: twice the
// FLOP per L2 byte of the single-CTA kernel, which is what that kernel is bound by.  The leader (cluster rank 0) issues
// the MMAs; its full barrier counts the TMA bytes of both CTAs; commits are multicast to both CTAs' barriers; both CTAs
// run the ordinary epilogue on their own 128 accumulator rows and release the accumulator on the leader's barrier.
template <int BN, int STAGES>
struct ConvGemmSmem2 {
  static constexpr int kABytes = kBlockM * kBlockK * 2;        // 16 KB: this CTA's 128 rows
  static constexpr int kBBytes = (BN / 2) * kBlockK * 2;       // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStageOffset = STAGES * kStageBytes;
  static constexpr int kBarOffset = kStageOffset + kEpiWarps * 2048;
  static constexpr int kTotal = kBarOffset + 256;
};

template <int BN, int STAGES, int RES, bool RV, bool STATS, int ACTV>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps, 1)
conv_gemm_cg2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const ConvGemmParams p) {
  using S = ConvGemmSmem2<BN, STAGES>;
  static_assert(S::kBBytes % 1024 == 0, "B half tile must keep the 1024-byte swizzle alignment");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  constexpr uint32_t kTmemCols = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.epi_tma) prefetch_tmap(&tmO);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader: one arrive.expect_tx per phase, bytes of both CTAs
      mbar_init(&empty_bar[i], 1);  // one multicast commit per phase
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);  // epilogue warps of both CTAs (used on the leader only)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, kTmemCols);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast commit / pair TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();

  const int npairs = static_cast<int>(gridDim.x) >> 1;
  const int pair = static_cast<int>(blockIdx.x) >> 1;
  const int total_tiles = p.total_items;  // work items: (M-tile pair, N tile, K split)
  const int kblocks = p.ntaps * p.kblocks_per_tap;
  const int per_pair = (total_tiles + npairs - 1) / npairs;
  const int tile_begin = p.contig ? pair * per_pair : pair;
  const int tile_end = p.contig ? min(total_tiles, tile_begin + per_pair) : total_tiles;
  const int tile_step = p.contig ? 1 : npairs;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    // warp-uniform loop, one elected lane issues (see conv_gemm_kernel)
    int s = 0;
    uint32_t ph = 0;
    for (int item = tile_begin; item < tile_end; item += tile_step) {
      const TileCoord tc = decode_item(p, item);
      const int x0 = tc.xb * p.bw * p.isx, y0 = tc.yb * p.bh * p.isy, z0 = tc.zb * p.bd * p.isz, b0 = tc.bblk * p.bb;
      const int n0 = tc.n_tile * BN + static_cast<int>(rank) * (BN / 2);
      int kb0, kb1;
      split_krange(p, kblocks, tc.sp, tc.ns, kb0, kb1);
      int tap = kb0 / p.kblocks_per_tap, kc = kb0 - tap * p.kblocks_per_tap;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        const uint32_t lead_bar = mapa_shared(smem_u32(&full_bar[s]), 0);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * S::kStageBytes);
          tma_load_5d_cg2(sa, &tmA, lead_bar, kc * kBlockK, x0 + p.tdx[tap], y0 + p.tdy[tap], z0 + p.tdz[tap], b0);
          tma_load_2d_cg2(sb, &tmB, lead_bar, kb * kBlockK, n0);
        }
        if (++kc == p.kblocks_per_tap) { kc = 0; ++tap; }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair leader only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, BN);
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int item = tile_begin; item < tile_end; item += tile_step, ++lt) {
        const int a = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        mbar_wait(&tmem_empty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        int tile_, sp, ns, kb0, kb1;
        decode_split(p, item, tile_, sp, ns);
        split_krange(p, kblocks, sp, ns, kb0, kb1);
        const int nkb = kb1 - kb0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
          const uint64_t da = make_sw128_kmajor_desc(sa);
          const uint64_t db = make_sw128_kmajor_desc(sb);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              tc_mma_f16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            tc_commit_cg2(&empty_bar[s], 3);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) tc_commit_cg2(&tmem_full[a], 3);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (both CTAs, their own 128 rows) =====================
    epilogue_loop<BN, STAGES, RES, RV, STATS, ACTV, S::kStageOffset>(p, &tmO, smem, tmem_full, tmem_empty, tmem_base, warp, lane,
                                                                    tile_begin, tile_end, tile_step);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory and both CTAs' barriers are still in use
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
  if constexpr (STATS) {
    if (p.gn_out) {
      const GnTailArgs ga{p.out_bf16, p.gn_out, p.col_stats, p.gn_gamma, p.gn_beta, p.gn_bar, p.N, p.B, p.ldo, p.stats_ld,
                          p.gn_groups, p.gn_act, p.gn_rows, p.gn_eps};
      gn_tail(ga, smem);
    }
  }
}

}  // namespace md
