// Self-attention of BasicTransformerBlock.attn1 (ldm/modules/attention.py:179-203) on tcgen05 tensor cores.
//
// One CTA owns NWG x 128 queries of one (sample, head) and walks the keys in tiles of 128:
//   S = Q K^T   : tcgen05.mma, Q and the K tile K-major (head dim contiguous) in 128B-swizzled shared memory,
//                 fp32 scores in TMEM (128 lanes = queries, 128 columns = keys);
//   softmax     : one thread per query row reads its 128 scores with tcgen05.ld, keeps the running maximum / sum in
//                 registers and writes P (bf16) into a K-major swizzled shared-memory tile;
//   O += P V    : tcgen05.mma with P as the K-major A operand and the V tile as an MN-major B operand (head dim
//                 contiguous, exactly how TMA lands the [key][dim] rows), fp32 output accumulated in TMEM.
// Q, K and V tiles come straight out of the fused qkv activation [B*S][3*heads][dh] through one 3-D tensor map: a box
// of 64 dims x 128 rows whose dims beyond dh are out of bounds and arrive as zeros, so head dims that are not a
// multiple of 64 (40, 80) need no padded copy.  The output rescale of the online softmax is lazy: the exponent offset
// only moves when the running maximum grew by more than 2^8, and then the warp rescales its TMEM rows in place.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2.. = NWG softmax warpgroups.
#include "host.h"
#include "ptx.cuh"
#include "kernels.h"

#include <cuda.h>
#include <stdlib.h>

namespace md {

namespace {

constexpr int kBox = 128 * 128;  // bytes of one TMA box: 128 rows x 64 bf16

template <int NDB, int NWG, int STAGES>
struct AttSmem {
  static constexpr int kQ = 0;                                   // [NWG][NDB] boxes
  static constexpr int kKV = kQ + NWG * NDB * kBox;              // [STAGES] x (K: NDB boxes | V: NDB boxes)
  static constexpr int kStageBytes = 2 * NDB * kBox;
  static constexpr int kP = kKV + STAGES * kStageBytes;          // [NWG] x 2 boxes (128 keys = 2 x 64)
  static constexpr int kBar = kP + NWG * 2 * kBox;
  static constexpr int kTotal = kBar + 256;
};

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// qkv viewed as [B*S rows][3*heads slots][dh].  PERSISTENT: the grid is one or two CTAs per SM and every CTA walks the
// work items (sample, head, tile of 128*NWG queries) item = cta, cta + grid, ...: barriers, TMEM and the K/V ring live
// across items, the Q tile of the next item is requested as soon as the last Q K^T of the current one has run and its
// first key tiles are already in flight during the current item's last softmax, so the TMA round trip + first-MMA
// latency that opened every non-persistent CTA (14 % of the stall samples, gpurun_out/r02ae) is paid once per SM.
// One-warpgroup instances leave room for two CTAs per SM (112 KB of shared memory, 256 TMEM columns, 168 registers x
// 192 threads each): one CTA's output epilogue and score waits overlap the other one's softmax.
template <int NDB, int NWG, int STAGES>
__global__ void __launch_bounds__(64 + 128 * NWG, (NWG == 1 && NDB == 1 && STAGES == 2) ? 2 : 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, int S, int heads, int dh,
                    float scale_log2e, int n_items) {
  using L = AttSmem<NDB, NWG, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* q_empty = q_full + 1;         // the last Q K^T of an item has run: the Q tile may be replaced
  // K and V tiles of a stage have their own barriers: the K half is free again as soon as Q K^T of its tile has run,
  // long before the P V product that frees the V half, so the next key tile is already in flight while the softmax of
  // the current one runs.  With one barrier per stage the load of tile j+1 could only start after P V of tile j-1, and
  // the softmax warps waited for their scores every tile (14 % of all stall samples at that wait in the ncu source
  // page, gpurun_out/r02ab): 135 -> 121 us on the level-0 launch.
  uint64_t* k_full = q_empty + 1;
  uint64_t* k_empty = k_full + STAGES;
  uint64_t* v_full = k_empty + STAGES;
  uint64_t* v_empty = v_full + STAGES;
  uint64_t* s_full = v_empty + STAGES;    // [NWG] scores of the current key tile are in TMEM
  uint64_t* s_free = s_full + NWG;        // [NWG] the warpgroup has read them into registers
  uint64_t* p_ready = s_free + NWG;       // [NWG] P tile written (and O rescaled if needed)
  uint64_t* pv_done = p_ready + NWG;      // [NWG] O += P V of the previous key tile has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + NWG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = S >> 7;                    // key tiles
  const int nq = S / (128 * NWG);           // query tiles per (sample, head)
  const int ksteps = (dh + 15) >> 4;        // 16-wide K steps of Q K^T
  const int npv = ksteps << 4;              // N of the P V product (head dim rounded up to 16)
  constexpr uint32_t kTmemCols = NWG == 2 ? 512 : 256;  // S[w] at column w*128, O[w] at column NWG*128 + w*128

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int w = 0; w < NWG; ++w) {
      mbar_init(&s_full[w], 1);
      mbar_init(&s_free[w], 128);
      mbar_init(&p_ready[w], 128);
      mbar_init(&pv_done[w], 1);
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
  pdl_grid_sync();

  // item -> (sample b, head h, first query q0); consecutive CTAs work on neighbouring query tiles of one (b, h): their
  // K / V tiles are shared through the L2
  auto item_coords = [&](int item, int& row0, int& h, int& q0) {
    const int qt = item % nq;
    const int bh = item / nq;
    h = bh % heads;
    row0 = (bh / heads) * S;   // first token row of the sample in the [B*S] row space
    q0 = qt * (128 * NWG);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    // warp-uniform control flow (operands stay in uniform registers), one elected lane issues
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      int row0, h, q0;
      item_coords(item, row0, h, q0);
      mbar_wait(q_empty, (it & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(q_full, NWG * NDB * kBox);
        for (int w = 0; w < NWG; ++w)
          for (int db = 0; db < NDB; ++db)
            tma_load_3d(smem + L::kQ + (w * NDB + db) * kBox, &tmQKV, q_full, db * 64, h, row0 + q0 + w * 128);
      }
      for (int j = 0; j < nt; ++j) {
        uint8_t* base = smem + L::kKV + s * L::kStageBytes;
        mbar_wait(&k_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&k_full[s], NDB * kBox);
#pragma unroll
          for (int db = 0; db < NDB; ++db)
            tma_load_3d(base + db * kBox, &tmQKV, &k_full[s], db * 64, heads + h, row0 + j * 128);
        }
        mbar_wait(&v_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&v_full[s], NDB * kBox);
#pragma unroll
          for (int db = 0; db < NDB; ++db)
            tma_load_3d(base + (NDB + db) * kBox, &tmQKV, &v_full[s], db * 64, 2 * heads + h, row0 + j * 128);
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // warp-uniform loop; every MMA and commit comes from the same elected lane.  g counts this CTA's key tiles across
    // items: stage index / parities of the rings and of the per-tile barriers follow it.
    const uint32_t idesc_qk = make_idesc_bf16_ex(128, 128, 0);
    const uint32_t idesc_pv = make_idesc_bf16_ex(128, npv, 1);
    int g = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      mbar_wait(q_full, it & 1);
      tc_fence_after();
      for (int j = 0; j <= nt; ++j) {
        if (j < nt) {
          const int gj = g + j;
          const int s = gj % STAGES;
          mbar_wait(&k_full[s], (gj / STAGES) & 1);
          tc_fence_after();
          const uint32_t kbase = smem_u32(smem + L::kKV + s * L::kStageBytes);
          for (int w = 0; w < NWG; ++w) {
            if (gj > 0) {
              mbar_wait(&s_free[w], (gj - 1) & 1);
              tc_fence_after();
            }
            const uint32_t qbase = smem_u32(smem + L::kQ + w * NDB * kBox);
            if (elect_one()) {
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint32_t off = (ks >> 2) * kBox;
                const uint64_t da = make_sw128_kmajor_desc(qbase + off) + 2 * (ks & 3);
                const uint64_t db = make_sw128_kmajor_desc(kbase + off) + 2 * (ks & 3);
                tc_mma_f16(tmem_base + w * 128, da, db, idesc_qk, ks != 0 ? 1u : 0u);
              }
              tc_commit(&s_full[w]);
            }
          }
          if (elect_one()) {
            tc_commit(&k_empty[s]);                     // the K tile is free once every warpgroup's Q K^T has run
            if (j == nt - 1) tc_commit(q_empty);        // ... and after the item's last Q K^T so is the Q tile
          }
        }
        if (j > 0) {
          const int jp = j - 1;
          const int gp = g + jp;
          const int sp = gp % STAGES;
          const uint32_t vbase = smem_u32(smem + L::kKV + sp * L::kStageBytes + NDB * kBox);
          mbar_wait(&v_full[sp], (gp / STAGES) & 1);
          tc_fence_after();
          for (int w = 0; w < NWG; ++w) {
            mbar_wait(&p_ready[w], gp & 1);
            tc_fence_after();
            const uint32_t pbase = smem_u32(smem + L::kP + w * 2 * kBox);
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {  // 128 keys = 8 K steps of 16
                const uint64_t da = make_sw128_kmajor_desc(pbase + (ks >> 2) * kBox) + 2 * (ks & 3);
                // V tile: [key][64 dims] rows of 128 B; a K step = 16 keys = 2048 B; dim blocks of 64 are kBox apart
                const uint64_t dv = make_sw128_mnmajor_desc(vbase + ks * 2048, kBox);
                tc_mma_f16(tmem_base + NWG * 128 + w * 128, da, dv, idesc_pv, (jp != 0 || ks != 0) ? 1u : 0u);
              }
              tc_commit(&pv_done[w]);
            }
          }
          if (elect_one()) tc_commit(&v_empty[sp]);
        }
      }
      g += nt;
    }
    __syncwarp();
  } else {
    // ===================== softmax warpgroups =====================
    const int w = (warp - 2) >> 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // query row inside the warpgroup's 128
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + w * 128;
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + NWG * 128 + w * 128;
    uint8_t* prow = smem + L::kP + w * 2 * kBox + r * 128;
    const int sw = r & 7;
    int g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int row0, h, q0;
      item_coords(item, row0, h, q0);
      float m_used = 0.f, l = 0.f;
      for (int j = 0; j < nt; ++j) {
        const int gj = g + j;
        mbar_wait(&s_full[w], gj & 1);
        tc_fence_after();
        uint32_t sv[128];
#pragma unroll
        for (int i = 0; i < 4; ++i) tmem_ld_32x32(t_s + 32 * i, *reinterpret_cast<uint32_t(*)[32]>(&sv[32 * i]));
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(&s_free[w]);              // the MMA warp may overwrite S with the next tile's scores
        float mx4[4] = {__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]), __uint_as_float(sv[3])};
#pragma unroll
        for (int i = 4; i < 128; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[i]));
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        float corr = 1.f;
        bool grow = false;
        if (j == 0) {
          m_used = mx;
        } else if ((mx - m_used) * scale_log2e > 8.f) {
          corr = ex2_approx((m_used - mx) * scale_log2e);
          m_used = mx;
          l *= corr;
          grow = true;
        }
        // exponentials first, into packed bf16 registers: the wait for the previous tile's P V product (which frees the P
        // tile and completes O) then sits behind ~1 us of arithmetic instead of in front of it
        const float mb = m_used * scale_log2e;
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        uint4 pk[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {          // 16-byte chunks of 8 keys
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pv[i] = ex2_approx(fmaf(__uint_as_float(sv[8 * c + i]), scale_log2e, -mb));
            l4[i & 3] += pv[i];
          }
          pk[c] = make_uint4(pack2_bf16(pv[0], pv[1]), pack2_bf16(pv[2], pv[3]), pack2_bf16(pv[4], pv[5]),
                             pack2_bf16(pv[6], pv[7]));
        }
        if (j > 0) {                            // (the item's first tile: the output epilogue below already waited)
          mbar_wait(&pv_done[w], (gj - 1) & 1); // P tile free again, O complete up to tile j-1
          if (__any_sync(0xffffffff, grow)) {
            tc_fence_after();
            for (int c = 0; c < npv; c += 16) {
              uint32_t o[16];
              tmem_ld_x16(t_o + c, o);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
              tmem_st_x16(t_o + c, o);
            }
            tc_wait_st();
          }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c)
          *reinterpret_cast<uint4*>(prow + (c >> 3) * kBox + (((c & 7) ^ sw) << 4)) = pk[c];
        l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        fence_proxy_async();                    // generic-proxy writes of P -> visible to the tensor core's async proxy
        tc_fence_before();
        mbar_arrive(&p_ready[w]);
      }
      g += nt;
      // ---- output: O / l -> bf16 [B][S][heads*dh]
      mbar_wait(&pv_done[w], (g - 1) & 1);
      tc_fence_after();
      const float inv = 1.f / l;
      __nv_bfloat16* orow = out + (static_cast<size_t>(row0 + q0 + w * 128 + r) * heads + h) * dh;
      for (int c = 0; c < npv; c += 16) {
        uint32_t o[16];
        tmem_ld_x16(t_o + c, o);
        tc_wait_ld();
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
          if (c + gq * 8 < dh) {
            const uint4 pk = make_uint4(
                pack2_bf16(__uint_as_float(o[gq * 8 + 0]) * inv, __uint_as_float(o[gq * 8 + 1]) * inv),
                pack2_bf16(__uint_as_float(o[gq * 8 + 2]) * inv, __uint_as_float(o[gq * 8 + 3]) * inv),
                pack2_bf16(__uint_as_float(o[gq * 8 + 4]) * inv, __uint_as_float(o[gq * 8 + 5]) * inv),
                pack2_bf16(__uint_as_float(o[gq * 8 + 6]) * inv, __uint_as_float(o[gq * 8 + 7]) * inv));
            *reinterpret_cast<uint4*>(orow + c + gq * 8) = pk;
          }
        }
      }
      tc_fence_before();   // the O reads above are ordered before the next item's first P V (issued after our p_ready arrive)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NDB, int NWG, int STAGES>
int attention_tc_impl(const void* qkv, void* out, int B, int S, int heads, int dh, cudaStream_t st) {
  using L = AttSmem<NDB, NWG, STAGES>;
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(tensor_map_encode_fn());
  if (!enc) return set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)dh, (cuuint64_t)(3 * heads), (cuuint64_t)B * S};
  const cuuint64_t strides[2] = {(cuuint64_t)dh * 2, (cuuint64_t)3 * heads * dh * 2};
  const cuuint32_t box[3] = {64, 1, 128};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(qkv) failed: %d (B=%d S=%d heads=%d dh=%d)", (int)r, B, S, heads, dh);
  static bool attr = false;
  if (!attr) {
    MD_CUDA(cudaFuncSetAttribute(attention_tc_kernel<NDB, NWG, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 L::kTotal));
    attr = true;
  }
  const int n_items = (S / (128 * NWG)) * heads * B;
  constexpr int ctas_per_sm = (NWG == 1 && NDB == 1 && STAGES == 2) ? 2 : 1;
  const dim3 grid(std::min(n_items, ctas_per_sm * num_sms()));
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(dh));
  launch_pdl(attention_tc_kernel<NDB, NWG, STAGES>, grid, dim3(64 + 128 * NWG), L::kTotal, st, tm,
             static_cast<__nv_bfloat16*>(out), S, heads, dh, scale_log2e, n_items);
  return check_launch("attention_tc");
}

}  // namespace

// true when the tensor-core kernel covers the shape (sequence a multiple of 128, head dim <= 128 and a multiple of 8)
bool attention_tc_supported(int S, int dh) { return S >= 128 && (S % 128) == 0 && dh >= 16 && dh <= 128 && (dh % 8) == 0; }

int launch_attention_tc(const void* qkv, void* out, int B, int S, int heads, int dh, cudaStream_t st) {
  if (!attention_tc_supported(S, dh)) return set_error("attention_tc: unsupported shape S=%d dh=%d", S, dh);
  if (reinterpret_cast<uintptr_t>(qkv) & 15) return set_error("attention_tc: qkv must be 16-byte aligned");
  if (dh <= 64) {
    // measured (r02f, 32 x 8 heads x 1024 x 40): two one-warpgroup CTAs per SM 134 us, one two-warpgroup CTA 141 us
    static const int variant = getenv("MD_ATT_VARIANT") ? atoi(getenv("MD_ATT_VARIANT")) : 1;
    if (variant == 1) return attention_tc_impl<1, 1, 2>(qkv, out, B, S, heads, dh, st);
    if (S % 256 == 0) return attention_tc_impl<1, 2, 3>(qkv, out, B, S, heads, dh, st);
    return attention_tc_impl<1, 1, 3>(qkv, out, B, S, heads, dh, st);
  }
  return attention_tc_impl<2, 1, 2>(qkv, out, B, S, heads, dh, st);
}

}  // namespace md
