// Attention kernels of the UNet:
//   * self-attention of BasicTransformerBlock.attn1 (ldm/modules/attention.py:179-203): fused flash-style kernel
//     (QK^T -> online softmax -> PV) on bf16 mma.sync tiles, head dims 40/80/160, sequences 16..4096;
//     the [B*heads, S, S] score tensor the reference materialises never exists.
//   * depth attention of DepthAttention.forward (ldm/models/diffusion/attention.py:26-47): per-pixel softmax over
//     the D depth samples of a view's frustum volume; one HBM pass over K|V.
#include "host.h"
#include "ptx.cuh"
#include "kernels.h"

#include <stdlib.h>

namespace md {

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float fast_exp2(float x) {  // ex2.approx.ftz: no denormal fix-up code around the MUFU
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kAttBK = 64;   // keys per tile

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// DHP: head dim padded to a multiple of 16; NW warps x 16 queries per CTA.  Row stride DHP+8 keeps ldmatrix rows on
// distinct banks.  K/V tiles are double-buffered with cp.async so the next tile streams in while this one is used.
#ifndef MD_ATT_MINBLOCKS
#define MD_ATT_MINBLOCKS 2
#endif
template <int DHP, int NW>
__global__ void __launch_bounds__(NW * 32, (DHP <= 48 && NW == 8) ? MD_ATT_MINBLOCKS : 1)
self_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int S, int heads, int dh,
                      float scale_log2e) {
  pdl_grid_sync();
  constexpr int LD = DHP + 8;
  constexpr int BQ = NW * 16;
  constexpr int NT = NW * 32;
  extern __shared__ __align__(16) uint8_t att_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* sK = sQ + BQ * LD;            // [2][64][LD]
  __nv_bfloat16* sV = sK + 2 * kAttBK * LD;    // [2][64][LD]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int C = heads * dh;
  const size_t row_stride = static_cast<size_t>(3) * C;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * S * row_stride + h * dh;
  const int chunks = dh / 8;          // 16-byte chunks of real data per row
  constexpr int chunksP = DHP / 8;    // including zero padding

  // Q tile (synchronous) + zero the padding columns of the K/V buffers once
  for (int i = threadIdx.x; i < BQ * chunksP; i += NT) {
    const int r = i / chunksP, c = i % chunksP;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q0 + r < S && c < chunks) v = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(q0 + r) * row_stride + c * 8);
    *reinterpret_cast<uint4*>(sQ + r * LD + c * 8) = v;
  }
  if (chunks < chunksP) {
    for (int i = threadIdx.x; i < 4 * kAttBK * (chunksP - chunks); i += NT) {
      const int r = i / (chunksP - chunks), c = chunks + i % (chunksP - chunks);
      *reinterpret_cast<uint4*>(sK + r * LD + c * 8) = make_uint4(0, 0, 0, 0);  // sK and sV are contiguous: 4*64 rows
    }
  }
  // K/V tile copy: every thread owns up to kSlots fixed (row, 16-byte chunk) slots of a tile; only the tile origin
  // changes between tiles, so the index arithmetic is done once
  constexpr int kSlots = (kAttBK * chunksP + NT - 1) / NT;
  int slot[kSlots];  // row | chunk << 8, or -1
#pragma unroll
  for (int k = 0; k < kSlots; ++k) {
    const int i = threadIdx.x + k * NT;
    const int r = i / chunks, c = i - r * chunks;
    slot[k] = (i < kAttBK * chunks) ? (r | (c << 8)) : -1;
  }
  auto prefetch = [&](int tile) {
    const int k0 = tile * kAttBK;
    __nv_bfloat16* dk = sK + (tile & 1) * kAttBK * LD;
    __nv_bfloat16* dv = sV + (tile & 1) * kAttBK * LD;
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
      if (slot[k] >= 0) {
        const int r = slot[k] & 255, c8 = (slot[k] >> 8) * 8;
        const bool ok = k0 + r < S;
        const __nv_bfloat16* src = base + static_cast<size_t>(ok ? k0 + r : 0) * row_stride + c8;
        cp_async16(dk + r * LD + c8, src + C, ok);
        cp_async16(dv + r * LD + c8, src + 2 * C, ok);
      }
    }
    cp_async_commit();
  };
  const int ntiles = (S + kAttBK - 1) / kAttBK;
  prefetch(0);
  __syncthreads();

  // Q fragments stay in registers for the whole kernel
  uint32_t qf[DHP / 16][4];
#pragma unroll
  for (int kk = 0; kk < DHP / 16; ++kk) {
    const int r = warp * 16 + (lane & 15);
    const int c = kk * 16 + (lane >> 4) * 8;
    ldsm_x4(qf[kk], sQ + r * LD + c);
  }

  float o[DHP / 8][4];
#pragma unroll
  for (int i = 0; i < DHP / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int t = 0; t < ntiles; ++t) {
    if (t + 1 < ntiles) { prefetch(t + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const __nv_bfloat16* tK = sK + (t & 1) * kAttBK * LD;
    const __nv_bfloat16* tV = sV + (t & 1) * kAttBK * LD;
    const int k0 = t * kAttBK;

    float s[kAttBK / 8][4];
#pragma unroll
    for (int j = 0; j < kAttBK / 8; j += 2) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      s[j + 1][0] = s[j + 1][1] = s[j + 1][2] = s[j + 1][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < DHP / 16; ++kk) {
        // one ldmatrix.x4 = B fragments of two 8-key groups: lanes 0-15 address group j, lanes 16-31 group j+1
        uint32_t bf[4];
        const int r = (j + (lane >> 4)) * 8 + (lane & 7);
        const int c = kk * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(bf, tK + r * LD + c);
        const uint32_t b0[2] = {bf[0], bf[1]}, b1[2] = {bf[2], bf[3]};
        mma_bf16_16816(s[j], qf[kk], b0);
        mma_bf16_16816(s[j + 1], qf[kk], b1);
      }
    }
    // mask keys beyond the sequence
    if (k0 + kAttBK > S) {
#pragma unroll
      for (int j = 0; j < kAttBK / 8; ++j) {
        const int key = k0 + j * 8 + (lane & 3) * 2;
        if (key >= S) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= S) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < kAttBK / 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffff, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffff, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffff, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffff, mx1, 2));
    const float corr0 = fast_exp2((m0 - mx0) * scale_log2e);
    const float corr1 = fast_exp2((m1 - mx1) * scale_log2e);
    m0 = mx0; m1 = mx1;
    l0 *= corr0; l1 *= corr1;
#pragma unroll
    for (int i = 0; i < DHP / 8; ++i) { o[i][0] *= corr0; o[i][1] *= corr0; o[i][2] *= corr1; o[i][3] *= corr1; }
    const float mb0 = m0 * scale_log2e, mb1 = m1 * scale_log2e;
    uint32_t pf[kAttBK / 16][4];
#pragma unroll
    for (int j = 0; j < kAttBK / 8; ++j) {
      const float p0 = fast_exp2(fmaf(s[j][0], scale_log2e, -mb0));
      const float p1 = fast_exp2(fmaf(s[j][1], scale_log2e, -mb0));
      const float p2 = fast_exp2(fmaf(s[j][2], scale_log2e, -mb1));
      const float p3 = fast_exp2(fmaf(s[j][3], scale_log2e, -mb1));
      l0 += p0 + p1; l1 += p2 + p3;
      pf[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int kk = 0; kk < kAttBK / 16; ++kk) {
#pragma unroll
      for (int i = 0; i < DHP / 8; i += 2) {
        // transposed x4: lanes 0-15 address dims i*8.., lanes 16-31 dims (i+1)*8.. of the same 16 keys
        uint32_t bf[4];
        const int r = kk * 16 + (lane & 15);
        ldsm_x4_trans(bf, tV + r * LD + (i + (lane >> 4)) * 8);
        const uint32_t b0[2] = {bf[0], bf[1]}, b1[2] = {bf[2], bf[3]};
        mma_bf16_16816(o[i], pf[kk], b0);
        mma_bf16_16816(o[i + 1], pf[kk], b1);
      }
    }
    __syncthreads();  // everyone done with buffer t&1 before it is refilled (tile t+2)
  }
  l0 += __shfl_xor_sync(0xffffffff, l0, 1);
  l0 += __shfl_xor_sync(0xffffffff, l0, 2);
  l1 += __shfl_xor_sync(0xffffffff, l1, 1);
  l1 += __shfl_xor_sync(0xffffffff, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + (lane >> 2);
  const int r1 = r0 + 8;
  __nv_bfloat16* ob = out + static_cast<size_t>(b) * S * C + h * dh;
#pragma unroll
  for (int i = 0; i < DHP / 8; ++i) {
    const int d = i * 8 + (lane & 3) * 2;
    if (d < dh) {
      if (r0 < S) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0) * C + d) = pack_bf16(o[i][0] * inv0, o[i][1] * inv0);
      if (r1 < S) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r1) * C + d) = pack_bf16(o[i][2] * inv1, o[i][3] * inv1);
    }
  }
}

template <int DHP, int NW>
static int self_attention_impl(const void* qkv, void* out, int B, int S, int heads, int dh, cudaStream_t st) {
  constexpr int LD = DHP + 8;
  constexpr int BQ = NW * 16;
  const int smem = (BQ + 4 * kAttBK) * LD * 2;
  static bool attr = false;
  if (!attr) {
    MD_CUDA(cudaFuncSetAttribute(self_attention_kernel<DHP, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  dim3 grid((S + BQ - 1) / BQ, heads, B);
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(dh));
  launch_pdl(self_attention_kernel<DHP, NW>, dim3(grid), dim3(NW * 32), smem, st, static_cast<const __nv_bfloat16*>(qkv),
                                                              static_cast<__nv_bfloat16*>(out), S, heads, dh,
                                                              scale_log2e);
  return check_launch("self_attention");
}

int launch_self_attention(const void* qkv, void* out, int B, int S, int heads, int dh, cudaStream_t st) {
  // long sequences run on the tcgen05 kernel (attention_tc.cu); MD_ATT_MMA=1 keeps the mma.sync kernel for A/B timing
  static const bool force_mma = getenv("MD_ATT_MMA") != nullptr;
  if (!force_mma && attention_tc_supported(S, dh)) return launch_attention_tc(qkv, out, B, S, heads, dh, st);
  return launch_self_attention_mma(qkv, out, B, S, heads, dh, st);
}

int launch_self_attention_mma(const void* qkv, void* out, int B, int S, int heads, int dh, cudaStream_t st) {
  if (dh % 8) return set_error("self_attention: head dim %d must be a multiple of 8", dh);
  const bool big = S >= 128;  // 8 warps x 16 queries per CTA when the sequence is long enough to fill them
  if (dh <= 48) return big ? self_attention_impl<48, 8>(qkv, out, B, S, heads, dh, st) : self_attention_impl<48, 4>(qkv, out, B, S, heads, dh, st);
  if (dh <= 80) return big ? self_attention_impl<80, 8>(qkv, out, B, S, heads, dh, st) : self_attention_impl<80, 4>(qkv, out, B, S, heads, dh, st);
  if (dh <= 160) return big ? self_attention_impl<160, 8>(qkv, out, B, S, heads, dh, st) : self_attention_impl<160, 4>(qkv, out, B, S, heads, dh, st);
  return set_error("self_attention: head dim %d unsupported", dh);
}

// ------------------------------------------------------------------------------------------------ depth attention
// DepthAttention.forward (ldm/models/diffusion/attention.py:26-47) re-associated so that K and V are never built:
//   sim[h][d] = q_h . (W_k,h c_d) = (W_k,h^T q_h) . c_d =: qp_h . c_d        (qp comes from one GEMM, scale folded in)
//   out       = W_out concat_h( W_v,h sum_d attn[h][d] c_d ) = W_ov . cbar     (second GEMM on cbar = sum_d attn c_d)
// with c = ReLU(GroupNorm(proj_context(ctx))) applied on the fly from the pre-norm tensor c1 and its per-(sample,
// channel) scale/shift.  A single pass over the D depth samples with an online softmax per head.  Samples b >= T have an all-zero frustum volume
// (the CFG-unconditional half): c == ReLU(beta) for every depth, attention is uniform and cbar = ReLU(beta).
//   qp   bf16 [T][HW][4*ctx]      c1 bf16 [T][D][HW][ctx]      ss fp32 [T][ctx][2]      beta fp32 [ctx]
//   cbar bf16 [B][HW][4*ctx]
// Thread mapping: TPP = ctx/16 lanes per (sample, pixel), each owning 16 contiguous context channels; a warp covers
// 32/TPP consecutive pixels (one 16-byte-vectorised, fully coalesced row segment per depth sample).  Per depth sample a
// lane spends 16 FMAs per head on the score and 16 on the accumulator; the partial scores meet with log2(TPP) shuffle
// steps per head.  The exponent offset is lazy (moves only when the running maximum grew by more than 2^8), so the
// accumulator rescale is off the common path.
template <int TPP>
__global__ void __launch_bounds__(128)
depth_attention_kernel(const __nv_bfloat16* __restrict__ qp, const __nv_bfloat16* __restrict__ c1,
                       const float* __restrict__ ss, const float* __restrict__ beta, __nv_bfloat16* __restrict__ cbar,
                       int T, int B, int D, int HW) {
  pdl_grid_sync();
  constexpr int ctx = TPP * 16;
  constexpr int PPW = 32 / TPP;  // pixels per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane % TPP, pw = lane / TPP;
  const size_t wid = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
  const size_t gp = wid * PPW + pw;  // global pixel index b*HW + pix; HW is a multiple of PPW: one sample per warp
  if (gp >= static_cast<size_t>(B) * HW) return;
  const int b = static_cast<int>(gp / HW);
  const int pix = static_cast<int>(gp % HW);
  const int j0 = sub * 16;
  __nv_bfloat16* op = cbar + gp * (4 * ctx) + j0;
  if (b >= T) {
    uint32_t w[8];
#pragma unroll
    for (int e = 0; e < 16; e += 2) w[e / 2] = pack_bf16(fmaxf(beta[j0 + e], 0.f), fmaxf(beta[j0 + e + 1], 0.f));
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      *reinterpret_cast<uint4*>(op + h * ctx) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(op + h * ctx + 8) = make_uint4(w[4], w[5], w[6], w[7]);
    }
    return;
  }
  float sc[16], sh[16], q[4][16], acc[4][16];
  {
    const float4* sp = reinterpret_cast<const float4*>(ss + (static_cast<size_t>(b) * ctx + j0) * 2);
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
      const float4 v = sp[e / 2];
      sc[e] = v.x; sh[e] = v.y; sc[e + 1] = v.z; sh[e + 1] = v.w;
    }
  }
  const __nv_bfloat16* qb = qp + gp * (4 * ctx) + j0;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const uint4 a = *reinterpret_cast<const uint4*>(qb + h * ctx);
    const uint4 c = *reinterpret_cast<const uint4*>(qb + h * ctx + 8);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {  // scores in the log2 domain: fold log2(e) into the query
      q[h][2 * e] = __uint_as_float(w[e] << 16) * 1.4426950408889634f;
      q[h][2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u) * 1.4426950408889634f;
      acc[h][2 * e] = 0.f; acc[h][2 * e + 1] = 0.f;
    }
  }
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, l[4] = {0.f, 0.f, 0.f, 0.f};
  const __nv_bfloat16* cp = c1 + (static_cast<size_t>(b) * D * HW + pix) * ctx + j0;
  const size_t dstride = static_cast<size_t>(HW) * ctx;
  // depth samples are fetched three at a time, one batch ahead of the arithmetic (D is a multiple of 6 at every level)
  constexpr int DB = 3;
  uint4 nxt[DB][2];
#pragma unroll
  for (int dd = 0; dd < DB; ++dd) {
    const int d = min(dd, D - 1);
    nxt[dd][0] = *reinterpret_cast<const uint4*>(cp + d * dstride);
    nxt[dd][1] = *reinterpret_cast<const uint4*>(cp + d * dstride + 8);
  }
  for (int d0 = 0; d0 < D; d0 += DB) {
    uint4 cur[DB][2];
#pragma unroll
    for (int dd = 0; dd < DB; ++dd) { cur[dd][0] = nxt[dd][0]; cur[dd][1] = nxt[dd][1]; }
    if (d0 + DB < D) {
#pragma unroll
      for (int dd = 0; dd < DB; ++dd) {
        const int d = min(d0 + DB + dd, D - 1);
        nxt[dd][0] = *reinterpret_cast<const uint4*>(cp + d * dstride);
        nxt[dd][1] = *reinterpret_cast<const uint4*>(cp + d * dstride + 8);
      }
    }
#pragma unroll
    for (int dd = 0; dd < DB; ++dd) {
      if (d0 + dd >= D) break;
      const uint32_t w[8] = {cur[dd][0].x, cur[dd][0].y, cur[dd][0].z, cur[dd][0].w,
                             cur[dd][1].x, cur[dd][1].y, cur[dd][1].z, cur[dd][1].w};
      float c[16];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        c[2 * e] = fmaxf(fmaf(__uint_as_float(w[e] << 16), sc[2 * e], sh[2 * e]), 0.f);
        c[2 * e + 1] = fmaxf(fmaf(__uint_as_float(w[e] & 0xffff0000u), sc[2 * e + 1], sh[2 * e + 1]), 0.f);
      }
      float s[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int e = 0; e < 16; e += 2) { a0 = fmaf(q[h][e], c[e], a0); a1 = fmaf(q[h][e + 1], c[e + 1], a1); }
        s[h] = a0 + a1;
      }
#pragma unroll
      for (int o = TPP / 2; o; o >>= 1) {
#pragma unroll
        for (int h = 0; h < 4; ++h) s[h] += __shfl_xor_sync(0xffffffff, s[h], o);
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        if (s[h] > m[h] + 8.f) {  // also the first sample (m = -inf): corr = 0 clears the (already zero) state
          const float corr = ex2_approx(m[h] - s[h]);
          m[h] = s[h];
          l[h] *= corr;
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[h][e] *= corr;
        }
        const float pe = ex2_approx(s[h] - m[h]);
        l[h] += pe;
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[h][e] = fmaf(pe, c[e], acc[h][e]);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float inv = 1.f / l[h];
    uint32_t w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = pack_bf16(acc[h][2 * e] * inv, acc[h][2 * e + 1] * inv);
    *reinterpret_cast<uint4*>(op + h * ctx) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(op + h * ctx + 8) = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

int launch_depth_attention(const void* qp, const void* c1, const float* ss, const float* beta, void* cbar, int T, int B,
                           int D, int HW, int ctx, cudaStream_t st) {
  if (ctx != 64 && ctx != 128 && ctx != 256 && ctx != 512)
    return set_error("depth_attention: context dim %d unsupported (64/128/256/512)", ctx);
  const int ppw = 32 / (ctx / 16);
  if (HW % ppw) return set_error("depth_attention: %d pixels per sample is not a multiple of %d", HW, ppw);
  const size_t warps = (static_cast<size_t>(B) * HW + ppw - 1) / ppw;
  const unsigned blocks = static_cast<unsigned>((warps * 32 + 127) / 128);
  const __nv_bfloat16* qq = static_cast<const __nv_bfloat16*>(qp);
  const __nv_bfloat16* cc = static_cast<const __nv_bfloat16*>(c1);
  __nv_bfloat16* oo = static_cast<__nv_bfloat16*>(cbar);
  switch (ctx) {
    case 64: launch_pdl(depth_attention_kernel<4>, dim3(blocks), dim3(128), 0, st, qq, cc, ss, beta, oo, T, B, D, HW); break;
    case 128: launch_pdl(depth_attention_kernel<8>, dim3(blocks), dim3(128), 0, st, qq, cc, ss, beta, oo, T, B, D, HW); break;
    case 256: launch_pdl(depth_attention_kernel<16>, dim3(blocks), dim3(128), 0, st, qq, cc, ss, beta, oo, T, B, D, HW); break;
    default: launch_pdl(depth_attention_kernel<32>, dim3(blocks), dim3(128), 0, st, qq, cc, ss, beta, oo, T, B, D, HW); break;
  }
  return check_launch("depth_attention");
}

}  // namespace md
