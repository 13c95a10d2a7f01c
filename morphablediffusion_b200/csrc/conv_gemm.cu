// Host launcher for the tcgen05 implicit-GEMM kernel: builds the TMA tensor maps and picks the tile shape.
#include "conv_gemm_launch.cuh"

#include <mutex>
#include <stdio.h>
#include <stdlib.h>

namespace md {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

#ifdef MD_KPROF
extern unsigned long long* g_kprof_buf;
#endif
void* tensor_map_encode_fn() { return reinterpret_cast<void*>(get_encode()); }

// ---- split-K workspaces: per context, selected per host thread (host.h); a process-wide default only serves the
// op-level entry point
static thread_local const SplitWorkspace* t_split = nullptr;
const SplitWorkspace* select_split_workspace(const SplitWorkspace* w) {
  const SplitWorkspace* prev = t_split;
  t_split = w;
  return prev;
}
int alloc_split_workspace(SplitWorkspace& w, size_t bytes, size_t ints) {
  void* p = nullptr;
  void* c = nullptr;
  // both all-zero between launches: the last arrival of every split tile restores the zeros it consumed
  if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMalloc(&c, ints * sizeof(int)) != cudaSuccess ||
      cudaMemset(p, 0, bytes) != cudaSuccess || cudaMemset(c, 0, ints * sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    if (p) cudaFree(p);
    if (c) cudaFree(c);
    return set_error("conv_gemm: split-K workspace allocation failed");
  }
  w.ws = static_cast<float*>(p); w.bytes = bytes;
  w.cnt = static_cast<int*>(c); w.ints = ints;
  return 0;
}
void free_split_workspace(SplitWorkspace& w) {
  if (w.ws) cudaFree(w.ws);
  if (w.cnt) cudaFree(w.cnt);
  w = SplitWorkspace();
}
static const SplitWorkspace* default_split_workspace() {
  static SplitWorkspace def;
  static std::once_flag once;
  std::call_once(once, [] { alloc_split_workspace(def, static_cast<size_t>(32) << 20, 1 << 14); });
  return def.ws ? &def : nullptr;
}

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

int launch_conv_gemm(const ConvGemmArgs& a, cudaStream_t stream) {
  const SplitWorkspace* sw = a.ksplit < 0 ? nullptr : (t_split ? t_split : default_split_workspace());
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if (a.Cin % kBlockK != 0) return set_error("conv_gemm: Cin=%d must be a multiple of 64", a.Cin);
  if (a.N % 8 != 0) return set_error("conv_gemm: N=%d must be a multiple of 8", a.N);
  if (a.ntaps < 1 || a.ntaps > kMaxTaps) return set_error("conv_gemm: bad tap count %d", a.ntaps);
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.Wt) & 15))
    return set_error("conv_gemm: operands must be 16-byte aligned");
  const int Cpitch = a.Cpitch ? a.Cpitch : a.Cin;
  if (Cpitch % 8 != 0) return set_error("conv_gemm: channel pitch %d must be a multiple of 8", Cpitch);
  if (a.Wpitch % 8 != 0 || (a.Wpitch && a.Wpitch < a.ntaps * a.Cin))
    return set_error("conv_gemm: weight pitch %d must be a multiple of 8 and cover K=%d", a.Wpitch, a.ntaps * a.Cin);

  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.isx = a.in_stride[0] ? a.in_stride[0] : 1; p.isy = a.in_stride[1] ? a.in_stride[1] : 1;
  p.isz = a.in_stride[2] ? a.in_stride[2] : 1;
  // tile grid = output positions (= input positions / stride)
  p.W = (a.W + p.isx - 1) / p.isx; p.H = (a.H + p.isy - 1) / p.isy; p.D = (a.D + p.isz - 1) / p.isz; p.B = a.B;
  int rem = kBlockM;
  p.bw = std::min(pow2_floor(p.W), rem); rem /= p.bw;
  p.bh = std::min(pow2_floor(p.H), rem); rem /= p.bh;
  p.bd = std::min(pow2_floor(p.D), rem); rem /= p.bd;
  p.bb = rem;
  p.nxb = (p.W + p.bw - 1) / p.bw;
  p.nyb = (p.H + p.bh - 1) / p.bh;
  p.nzb = (p.D + p.bd - 1) / p.bd;
  p.nbb = (a.B + p.bb - 1) / p.bb;
  p.m_tiles = p.nxb * p.nyb * p.nzb * p.nbb;
  p.kblocks_per_tap = a.Cin / kBlockK;
  p.ntaps = a.ntaps;
  for (int t = 0; t < a.ntaps; ++t) {
    p.tdx[t] = static_cast<int8_t>(a.tap[t][0]);
    p.tdy[t] = static_cast<int8_t>(a.tap[t][1]);
    p.tdz[t] = static_cast<int8_t>(a.tap[t][2]);
  }
  p.N = a.N;
  p.OW = a.OW ? a.OW : p.W; p.OH = a.OH ? a.OH : p.H; p.OD = a.OD ? a.OD : p.D;
  p.osx = a.os[0] ? a.os[0] : 1; p.osy = a.os[1] ? a.os[1] : 1; p.osz = a.os[2] ? a.os[2] : 1;
  p.opx = a.op[0]; p.opy = a.op[1]; p.opz = a.op[2];
  p.act = a.act;
  const int n_out = (a.act == ACT_GEGLU) ? a.N / 2 : a.N;
  p.ldo = a.ldo ? a.ldo : n_out;
  p.bias = a.bias; p.rowvec = a.rowvec; p.rowvec_ld = a.rowvec_ld ? a.rowvec_ld : a.N;
  p.res_f32 = a.res_f32; p.res_bf16 = static_cast<const __nv_bfloat16*>(a.res_bf16);
  p.out_f32 = a.out_f32; p.out_bf16 = static_cast<__nv_bfloat16*>(a.out_bf16);
  p.out_scale = a.out_scale == 0.f ? 1.f : a.out_scale;
  p.col_stats = a.col_stats; p.stats_ld = a.stats_ld ? a.stats_ld : n_out;
  if (a.col_stats && ((p.bw * p.bh * p.bd) % 32) != 0)
    return set_error("conv_gemm: fused statistics need >= 32 rows per sample per tile (box %dx%dx%d)", p.bw, p.bh, p.bd);
  if (p.ldo % 8 != 0) return set_error("conv_gemm: ldo=%d must be a multiple of 8", p.ldo);
  if (static_cast<long long>(a.B) * p.OD * p.OH * p.OW * p.ldo >= (1LL << 31))
    return set_error("conv_gemm: output of %lld elements exceeds the 32-bit row offsets of the epilogue",
                     static_cast<long long>(a.B) * p.OD * p.OH * p.OW * p.ldo);
  if (a.act == ACT_GEGLU && !a.out_bf16) return set_error("conv_gemm: GEGLU epilogue writes bf16 only");
  if (a.gn_out) {
    const bool plain = p.osx == 1 && p.osy == 1 && p.osz == 1 && p.opx == 0 && p.opy == 0 && p.opz == 0 && p.OW == p.W &&
                       p.OH == p.H && p.OD == p.D;
    if (!a.col_stats || !a.out_bf16 || !a.gn_gamma || !a.gn_beta || !a.gn_barrier || !plain || a.gn_groups < 1 ||
        a.N % a.gn_groups || a.N % 4 || a.N > 1280 || p.ldo % 4)
      return set_error("conv_gemm: the GroupNorm tail needs col_stats, out_bf16, gamma / beta, a barrier counter, the "
                       "plain output geometry and N <= 1280 divisible by the groups (N=%d groups=%d)", a.N, a.gn_groups);
    p.gn_out = static_cast<__nv_bfloat16*>(a.gn_out);
    p.gn_gamma = a.gn_gamma; p.gn_beta = a.gn_beta; p.gn_groups = a.gn_groups; p.gn_act = a.gn_act; p.gn_eps = a.gn_eps;
    p.gn_bar = a.gn_barrier;
    p.gn_rows = p.W * p.H * p.D;
  }

  // ---- tile N and split-K selection: minimise  waves x (per-tile MMA time ~ BN, divided by the K split)  plus a
  // reduction overhead for split tiles; prefer the wider tile on ties (fewer re-reads of the activation tile)
  const int kblocks_all = p.ntaps * p.kblocks_per_tap;
  auto auto_split = [&](long long tiles) {
    if (a.ksplit != 0 || a.act == ACT_GEGLU || !sw) return 1;
    if (tiles * 2 > num_sms() || kblocks_all < 8) return 1;
    // at least `split_div` K blocks per split (MD_SPLIT_DIV; the sweep in profiles/r02_gemm_autotune.md found 9-way splits
    // of 36-block loops losing to 2-3-way ones)
    static const int split_div = getenv("MD_SPLIT_DIV") ? std::max(1, atoi(getenv("MD_SPLIT_DIV"))) : 4;
    long long ks = std::min<long long>(std::min<long long>(num_sms() / tiles, kblocks_all / split_div), 16);
    // the partial-sum exchange is 5-10 us of pure latency (publish, ticket, read back: profiles/r02 phase stamps), about
    // as long as `split_min_saved` K blocks: a split that shortens the K loop by less than that loses.  MD_SPLIT_MIN
    // overrides the threshold (0 = the round-1 rule).
    static const int split_min_saved = getenv("MD_SPLIT_MIN") ? atoi(getenv("MD_SPLIT_MIN")) : 28;
    if (ks > 1 && kblocks_all - kblocks_all / ks < split_min_saved) ks = 1;
    return static_cast<int>(std::max<long long>(ks, 1));
  };
  int BN = a.BN;
  if (BN == 0) {
    if (a.act == ACT_GEGLU) {
      BN = kGegluTile;
    } else {
      // Two cost models over the candidate widths.  Tile-starved problems (some candidate wants split-K) keep the
      // MMA-time model  waves x (bn + 24) / ksplit  that the split heuristic was tuned with.  Everything else is bound
      // by the bytes a tile pulls from L2 per K block, (128 + bn) rows (the 12 TB/s L2 caps a 128 x bn tile at
      // bn / (128 + bn) of ~1.5 PFLOP/s), so there the model is  waves x (bn + 128): it favours the widest tile that
      // still fills the waves.
      const int cands[4] = {256, 160, 128, 64};
      double best_mma = -1.0, best_l2 = -1.0;
      int bn_mma = 64, bn_l2 = 0, ks_mma = 1;
      for (int ci = 0; ci < 4; ++ci) {
        const int bn = cands[ci];
        if (bn > 64 && a.N <= bn / 2) continue;                       // mostly-empty tile
        const long long tiles = static_cast<long long>(p.m_tiles) * ((a.N + bn - 1) / bn);
        const int ks = auto_split(tiles);
        const long long waves = (tiles * ks + num_sms() - 1) / num_sms();
        const double cost = waves * (bn + 24.0) / ks + (ks > 1 ? 0.15 * (bn + 24.0) : 0.0);
        if (best_mma < 0 || cost < best_mma) { best_mma = cost; bn_mma = bn; ks_mma = ks; }
        if (ks == 1) {
          const double cost_l2 = waves * (bn + 128.0);
          if (best_l2 < 0 || cost_l2 < best_l2) { best_l2 = cost_l2; bn_l2 = bn; }
        }
      }
      BN = (ks_mma == 1 && bn_l2 != 0) ? bn_l2 : bn_mma;
    }
  }
  if (a.act == ACT_GEGLU && (BN != kGegluTile || a.N % kGegluTile != 0))
    return set_error("conv_gemm: GEGLU requires BN=%d and N a multiple of it (N=%d)", kGegluTile, a.N);
  p.n_tiles = (a.N + BN - 1) / BN;

  // ---- tensor maps
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.D, (cuuint64_t)a.B};
    cuuint64_t strides[4] = {(cuuint64_t)Cpitch * 2, (cuuint64_t)Cpitch * 2 * a.W,
                             (cuuint64_t)Cpitch * 2 * a.W * a.H, (cuuint64_t)Cpitch * 2 * a.W * a.H * a.D};
    cuuint32_t box[5] = {(cuuint32_t)kBlockK, (cuuint32_t)(p.bw * p.isx), (cuuint32_t)(p.bh * p.isy),
                         (cuuint32_t)(p.bd * p.isz), (cuuint32_t)p.bb};
    cuuint32_t es[5] = {1, (cuuint32_t)p.isx, (cuuint32_t)p.isy, (cuuint32_t)p.isz, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(a.A), dims, strides, box,
                     es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(A) failed: %d (W=%d H=%d D=%d B=%d C=%d)", (int)r,
                                            a.W, a.H, a.D, a.B, a.Cin);
  }
  // ---- split-K for tile-starved problems (few output tiles, long K loop): fills the machine and multiplies the
  // bytes in flight of what is otherwise a weight-streaming loop on a handful of SMs
  p.ksplit = 1;
  if (a.ksplit > 1) {
    p.ksplit = a.ksplit;
  } else {
    p.ksplit = auto_split(static_cast<long long>(p.m_tiles) * p.n_tiles);
  }
  if (p.ksplit > 1) {
    if (a.act == ACT_GEGLU) return set_error("conv_gemm: split-K is not available with the GEGLU epilogue");
    const size_t need = static_cast<size_t>(p.m_tiles) * p.n_tiles * 128 * BN * sizeof(float);  // one accumulation tile per split tile
    float* ws = sw ? sw->ws : nullptr;
    int* cnt = sw ? sw->cnt : nullptr;
    const size_t ws_bytes = sw ? sw->bytes : 0;
    const size_t cnt_ints = sw ? sw->ints : 0;
    if (!ws || need > ws_bytes || static_cast<size_t>(p.m_tiles) * p.n_tiles * kEpiWarps > cnt_ints) {
      if (a.ksplit > 1) return set_error("conv_gemm: split-K workspace too small (%zu bytes needed)", need);
      p.ksplit = 1;
    }
    p.split_ws = ws;
    p.split_cnt = cnt;
  }
  // ---- CTA pairs (cta_group::2) for tile-rich problems on the two widest tiles: half the weight bytes per FLOP.
  // With the warp-uniform issue loops the long-K convolutions are bound by operand traffic, and pairs win from ~24 K
  // blocks up (r02a, isolated, warm: 640->640 @32x32 207 -> 181 us, 1920->640 @16x16 148 -> 129 us = 1.40 PFLOP/s,
  // 320->320 @32x32 72 -> 63 us); below that the launches are epilogue-bound and pairs lose (no TMA-store epilogue).
  // MD_CG2=0 switches pairs off, MD_CG2=2 uses them whenever eligible; args.cta_pair overrides per call.
  static const int cg2_env = getenv("MD_CG2") ? atoi(getenv("MD_CG2")) : 1;
  const int cg2_mode = a.cta_pair != 0 ? (a.cta_pair > 0 ? 2 : 0) : cg2_env;
  const int cg2_min_kblocks = cg2_mode >= 2 ? 1 : 24;
  p.m_pairs = (p.m_tiles + 1) / 2;
  int cg2_pairs = 0;
  if (cg2_mode > 0 && p.ksplit == 1 && (BN == 160 || BN == 256) && p.m_tiles >= 2 && kblocks_all >= cg2_min_kblocks) {
    int resident = 0;
    const int rc = (BN == 160) ? launch_conv_gemm_cg2_bn160(tmA, tmA, tmA, p, 0, stream, &resident)
                               : launch_conv_gemm_cg2_bn256(tmA, tmA, tmA, p, 0, stream, &resident);
    if (rc == 0 && resident > 0 && p.m_pairs * p.n_tiles >= resident) {
      p.cg2 = 1;
      cg2_pairs = std::min(p.m_pairs * p.n_tiles, resident);
    }
  }
  {
    const cuuint64_t Ktot = (cuuint64_t)a.ntaps * a.Cin;
    cuuint64_t dims[2] = {Ktot, (cuuint64_t)a.N};
    cuuint64_t strides[1] = {(a.Wpitch ? (cuuint64_t)a.Wpitch : Ktot) * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)(p.cg2 ? BN / 2 : BN)};  // a pair CTA stages half the rows
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a.Wt), dims, strides, box,
                     es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(B) failed: %d (K=%llu N=%d)", (int)r,
                                            (unsigned long long)Ktot, a.N);
  }

  // ---- whole waves unsplit, the partial wave split along K.  A persistent grid of G CTAs (pairs) walks T tiles in
  // ceil(T / G) rounds; when one round holds only rem = T mod G tiles, those are split ks = G / rem ways so that the
  // round costs ~1/ks of a tile instead of a whole one, e.g. 320 tiles of 256 columns on 148 SMs: 2.16 -> 3 rounds
  // become 2 + 1/6.  The split items run FIRST: the partial-sum exchange (measured 5-10 us of pure latency: publish,
  // ticket, read back) then overlaps the next tile's main loop through the second TMEM accumulator instead of
  // ending the kernel.  Uses the split-K machinery of the tile-starved launches.  MD_HYBRID=0 switches it off.
  const int tiles_all = p.cg2 ? p.m_pairs * p.n_tiles : p.m_tiles * p.n_tiles;
  const int slots = p.cg2 ? cg2_pairs : std::min(tiles_all, num_sms());
  const bool contig = p.n_tiles == 1 && p.ksplit == 1 && tiles_all > 2 * slots;
  p.split_tiles = p.ksplit > 1 ? tiles_all : 0;
  {
    static const int hybrid_env = getenv("MD_HYBRID") ? atoi(getenv("MD_HYBRID")) : 1;
    const int hybrid = a.tail_split != 0 ? (a.tail_split > 0 ? 1 : 0) : hybrid_env;
    if (hybrid && p.ksplit == 1 && a.ksplit == 0 && !contig && sw && tiles_all > slots && kblocks_all >= 96) {
      const int rem = tiles_all % slots;
      const int ks = rem ? std::min(std::min(slots / rem, kblocks_all / 4), 6) : 1;
      if (ks >= 2) {
        const size_t ws_slots = static_cast<size_t>(rem) * (p.cg2 ? 2 : 1);
        const size_t need = ws_slots * 128 * BN * sizeof(float);
        if (need <= sw->bytes && ws_slots * kEpiWarps <= sw->ints) {
          p.ksplit = ks;
          p.split_tiles = rem;
          p.split_ws = sw->ws;
          p.split_cnt = sw->cnt;
        }
      }
    }
  }
  p.split_items = p.split_tiles * p.ksplit;
  p.total_items = p.split_items + (tiles_all - p.split_tiles);

  // ---- TMA-store epilogue: one output, no fused statistics, no residual, no split-K, plain output geometry; the
  // output leaves as 16-column x 32-row boxes of the per-warp staging tile.  Measured (gpurun r01y/r01z): -10..19 % on
  // the bf16-out GEMMs without residual (qkv, GEGLU); fp32-residual launches get 5-11 % slower this way (their per-lane
  // residual reads touch 32 sectors per request) and keep the coalesced epilogue unless MD_EPI_TMA=2.  MD_EPI_TMA=0
  // switches the path off.
  CUtensorMap tmO = tmA;
  {
    static const int epi_tma_mode = getenv("MD_EPI_TMA") ? atoi(getenv("MD_EPI_TMA")) : 1;
    const bool epi_tma_on = epi_tma_mode != 0, epi_tma_res = epi_tma_mode >= 2;
    const bool one_out = (a.out_f32 != nullptr) != (a.out_bf16 != nullptr);
    const bool plain = p.osx == 1 && p.osy == 1 && p.osz == 1 && p.opx == 0 && p.opy == 0 && p.opz == 0 && p.OW == p.W &&
                       p.OH == p.H && p.OD == p.D;
    void* optr = a.out_f32 ? static_cast<void*>(a.out_f32) : a.out_bf16;
    if (epi_tma_on && !p.cg2 && one_out && plain && !a.col_stats && !a.res_bf16 && (!a.res_f32 || epi_tma_res) &&
        (p.ksplit == 1 || p.split_tiles < tiles_all) &&
        !(reinterpret_cast<uintptr_t>(optr) & 15)) {
      // a lane quarter's 32 rows inside the tile box (x fastest): sub-box dims and the origin of every quarter
      int qd[4], bd4[4] = {p.bw, p.bh, p.bd, p.bb}, rem = 32;
      for (int i = 0; i < 4; ++i) { qd[i] = std::min(bd4[i], rem); rem /= qd[i]; }
      for (int qi = 0; qi < 4; ++qi) {
        int r = qi * 32;
        for (int i = 0; i < 4; ++i) { p.qorg[qi][i] = static_cast<int16_t>(r % bd4[i]); r /= bd4[i]; }
      }
      const size_t es = a.out_f32 ? 4 : 2;
      cuuint64_t dims[5] = {(cuuint64_t)n_out, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)a.B};
      cuuint64_t strides[4] = {(cuuint64_t)p.ldo * es, (cuuint64_t)p.ldo * es * p.W, (cuuint64_t)p.ldo * es * p.W * p.H,
                               (cuuint64_t)p.ldo * es * p.W * p.H * p.D};
      cuuint32_t box[5] = {(cuuint32_t)kChunk, (cuuint32_t)qd[0], (cuuint32_t)qd[1], (cuuint32_t)qd[2], (cuuint32_t)qd[3]};
      cuuint32_t es5[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&tmO, a.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, optr, dims,
                       strides, box, es5, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       a.out_f32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r == CUDA_SUCCESS) p.epi_tma = 1;
    }
  }
#ifdef MD_KPROF
  if (!g_kprof_buf) {
    cudaMalloc(&g_kprof_buf, sizeof(unsigned long long) * 160 * 32);
    cudaMemset(g_kprof_buf, 0, sizeof(unsigned long long) * 160 * 32);
  }
  p.kprof = g_kprof_buf;
#endif
  static const int gemm_dbg = getenv("MD_GEMM_DBG") ? atoi(getenv("MD_GEMM_DBG")) : 0;
  p.dbg = gemm_dbg;
  p.fd_ksplit = make_fastdiv(p.ksplit);
  p.fd_ntiles = make_fastdiv(p.n_tiles);
  p.fd_nxb = make_fastdiv(p.nxb);
  p.fd_nyb = make_fastdiv(p.nyb);
  p.fd_nzb = make_fastdiv(p.nzb);
  const int total = p.total_items;
  const int grid = p.cg2 ? 2 * std::min(total, cg2_pairs) : std::min(total, num_sms());
  p.contig = contig ? 1 : 0;
  static const bool trace = getenv("MD_TRACE") != nullptr;
  if (trace)
    fprintf(stderr, "conv_gemm B=%d D=%d H=%d W=%d Cin=%d taps=%d N=%d BN=%d tiles=%dx%d ks=%d split_tiles=%d cg2=%d act=%d f32=%d bf16=%d res=%d "
            "stats=%d rowvec=%d tail=%d stride=%d\n",
            a.B, a.D, a.H, a.W, a.Cin, a.ntaps, a.N, BN, p.m_tiles, p.n_tiles, p.ksplit, p.split_tiles, p.cg2, a.act, a.out_f32 != nullptr,
            a.out_bf16 != nullptr, a.res_f32 != nullptr ? 1 : a.res_bf16 != nullptr ? 2 : 0, a.col_stats != nullptr, a.rowvec != nullptr,
            a.gn_out != nullptr, a.in_stride[0] > 1 || a.os[0] > 1);
  if (p.cg2)
    return (BN == 160) ? launch_conv_gemm_cg2_bn160(tmA, tmB, tmO, p, grid, stream, nullptr)
                       : launch_conv_gemm_cg2_bn256(tmA, tmB, tmO, p, grid, stream, nullptr);
  switch (BN) {
    case 64:  return launch_conv_gemm_bn64(tmA, tmB, tmO, p, grid, stream);
    case 128: return launch_conv_gemm_bn128(tmA, tmB, tmO, p, grid, stream);
    case 160: return launch_conv_gemm_bn160(tmA, tmB, tmO, p, grid, stream);
    case 256: return launch_conv_gemm_bn256(tmA, tmB, tmO, p, grid, stream);
    default:  return set_error("conv_gemm: unsupported BN=%d", BN);
  }
}

}  // namespace md

#ifdef MD_KPROF
// Development builds only: phase time stamps of the last conv_gemm launch ([160 CTAs][32 slots], globaltimer ns).
namespace md { unsigned long long* g_kprof_buf = nullptr; }
extern "C" __attribute__((visibility("default"))) int md_debug_kprof(unsigned long long* host_out, int clear) {
  if (!md::g_kprof_buf) return -1;
  if (cudaMemcpy(host_out, md::g_kprof_buf, sizeof(unsigned long long) * 160 * 32, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (clear) cudaMemset(md::g_kprof_buf, 0, sizeof(unsigned long long) * 160 * 32);
  return 0;
}
#endif
