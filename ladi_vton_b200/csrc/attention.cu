// Flash-style fused attention for sm_100a, head_dim 64, bf16 in / bf16 out, fp32 softmax state.
//   S = Q K^T  : tcgen05.mma (M=128 queries, N=128 keys, K=64) into TMEM;
//   softmax    : one thread per query row, S read with tcgen05.ld, online max/sum in the exp2 domain, P written as bf16 into
//                128B-swizzled shared memory;
//   O += P V   : tcgen05.mma with P (K-major, smem) x V (MN-major descriptor straight on the TMA-loaded [keys][64] tile); O
//                accumulates in TMEM and is rescaled in place (tcgen05.ld/st) only when a row maximum moved.
// Q/K/V tiles arrive by TMA through 3-D tensor maps over [batch][tokens][row pitch], so per-head slices of the fused QKV
// projection output are read in place (no head split/transposes); out-of-range tokens are zero-filled and masked.
//
// Two kernels:
//   attention_pair_kernel  (nkv > 128): one CTA owns TWO 128-row query tiles, each with its own softmax warps AND its own MMA issuer
//       warp, so each tile's chain  softmax(j) -> S(j+1), P(j) V(j) -> softmax(j+1)  advances independently and the tensor core works
//       on one tile while the other is in its softmax; S(j+1) is issued before P(j) V(j) (shortest path back to the softmax warps);
//       K/V tiles are loaded once for both query tiles (3-stage TMA ring, released when both issuers are done with a tile).
//   attention_single_kernel<SHORT>: one query tile per CTA; SHORT (nkv <= 128) needs one K/V stage and 256 TMEM columns, so two CTAs share
//       an SM and hide each other's latency chain.
//   attention_short_persistent_kernel (round 2, the default for nkv <= 128 = the 77 text tokens of cross-attention): persistent CTAs looping
//       over (query tile, head, image) items with a two-stage TMA ring and P in tensor memory.
//
// Replaces diffusers CrossAttention (attn1/attn2 of BasicTransformerBlock) inside UNet2DConditionModel.forward, which the
// reference runs through torch SDPA or xformers (/root/reference/src/inference.py:143-147; call-site tryon_pipe.py:732).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace {

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KiB: [128 tokens][64 bf16]

struct AttnParams {
  int nq, nkv, n_kv_tiles;
  bf16* out;
  int out_pitch;
  int64_t out_batch_stride;
  float scale_log2;  // scale * log2(e)
  int b_delay;       // pair kernel: cycles tile B's first S = Q K^T is held back after tile A's (phase offset of the two softmaxes)
  long long* trace;  // optional per-phase clock64() trace of CTA (0,0,0) (tools/attn_trace.py); null in production
};

#ifdef LADI_ATTN_TRACE  // debug builds only (tools/attn_trace.py): per-phase clock64() stamps of CTA (0,0,0)
#define TRACE(slot)                                                                         \
  do {                                                                                      \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && trace_ok) \
      p.trace[trace_base + (slot)] = clock64();                                             \
  } while (0)
#else
#define TRACE(slot) \
  do {              \
  } while (0)
#endif

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// HALVES threads per query row (HALVES = 1: 128 threads per tile; HALVES = 2: 256 threads, the two halves split the key
// columns of every S tile, the O columns and the P K-chunks, and agree on the running row maximum through shared memory).
// Consumes S tiles from TMEM, produces P tiles in shared memory, keeps O normalised.
//   S for KV tile j lives at tS + s_stride * (j & s_mask); barriers: s_full (per S buffer), p_full (count 128*HALVES), o_ready.
template <int HALVES, bool SINGLE_READ, bool PT = false>
__device__ __forceinline__ void softmax_rows(const AttnParams& p, int n_tiles, uint32_t tS, uint32_t s_stride, uint32_t s_mask, uint32_t tO,
                                             uint8_t* sP, uint32_t s_full0, uint32_t p_full, uint32_t o_ready, int ew, int lane, int q0,
                                             int h, int b, int half, float* xm, uint32_t bar_id, uint32_t tP = 0) {
  constexpr int COLS = BKV / HALVES;        // key columns of each S tile owned by this thread
  constexpr int OCOLS = HD / HALVES;        // O columns owned by this thread
  const int r = ew * 32 + lane;
  const uint32_t lane_off = (uint32_t)(ew * 32) << 16;
  const int col0 = half * COLS;
  const bool trace_ok = (ew == 0 && lane == 0);
  int trace_base = 0;
  float m = -INFINITY, l = 0.f;
  uint8_t* prow0 = sP + r * 128 + (HALVES == 2 ? half * TILE_BYTES : 0);
  const int sw = r & 7;
  for (int j = 0; j < n_tiles; ++j) {
    const uint32_t sb = j & s_mask;
    const int valid = min(COLS, p.nkv - j * BKV - col0);  // may be <= 0 for the upper half of a ragged last tile
    trace_base = 1024 * (int)(bar_id * 2 + half) + 8 * j;  // bar_id: 1 = tile A, 2 = tile B
    TRACE(0);
    ptx::mbar_wait(s_full0 + 8 * sb, s_mask ? ((j >> 1) & 1) : (j & 1));
    ptx::tc_fence_after();
    TRACE(1);
    const uint32_t ts = tS + sb * s_stride + lane_off + col0;
    bool moved;
    float alpha;
    if constexpr (SINGLE_READ) {
    // ---- one TMEM read per tile: this thread's COLS scores stay in registers for both the maximum and the exponentials
    uint32_t v[COLS];
#pragma unroll
    for (int c = 0; c < COLS; c += 32) {
      if (c < valid) ptx::tmem_ld32(ts + c, reinterpret_cast<uint32_t(&)[32]>(v[c]));
    }
    ptx::tmem_wait_ld();
    float mx = -INFINITY;
    if (valid >= COLS) {
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;  // independent chains
#pragma unroll
      for (int i = 0; i < COLS; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(v[i + 2])); m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
      }
      mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    } else {
#pragma unroll
      for (int i = 0; i < COLS; ++i)
        if (i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
    }
    if (HALVES == 2) {  // both halves of a row must scale P and O with the same maximum
      float* slot = xm + (j & 1) * 256;
      slot[half * 128 + r] = mx;
      ptx::named_barrier_sync(bar_id, 256);
      mx = fmaxf(mx, slot[(half ^ 1) * 128 + r]);
    }
    const float m_true = mx * p.scale_log2;
    const float m_new = (m_true > m + 8.f) ? m_true : m;  // lazy rescaling, see the two-pass path
    alpha = ex2(m - m_new);  // 0 on the first tile (m = -inf)
    moved = m_new > m;
    m = m_new;
    // P buffer and O are free once P V of the previous tile has completed
    if (j > 0) {
      ptx::mbar_wait(o_ready, (j - 1) & 1);
      ptx::tc_fence_after();
    }
    // ---- p = exp2(s*scale - m_new) -> bf16, streamed to the swizzled P tile 8 keys (16 bytes) at a time
    float sum = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
    for (int q = 0; q < COLS / 8; ++q) {
      float e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = ex2(fmaf(__uint_as_float(v[8 * q + i]), p.scale_log2, -m_new));
        e[i] = (valid >= COLS || 8 * q + i < valid) ? x : 0.f;
      }
      sum += e[0] + e[4]; sum1 += e[1] + e[5]; sum2 += e[2] + e[6]; sum3 += e[3] + e[7];
      uint8_t* dst = prow0 + (q >> 3) * TILE_BYTES + (((q & 7) ^ sw) << 4);
      *reinterpret_cast<uint4*>(dst) =
          make_uint4(ptx::pack_bf16(e[0], e[1]), ptx::pack_bf16(e[2], e[3]), ptx::pack_bf16(e[4], e[5]), ptx::pack_bf16(e[6], e[7]));
    }
    l = l * alpha + ((sum + sum1) + (sum2 + sum3));
    } else {
    // pass 1: row maximum
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < COLS; c += 32) {
      if (c >= valid) break;
      uint32_t v[32];
      ptx::tmem_ld32(ts + c, v);
      ptx::tmem_wait_ld();
      if (c + 32 <= valid) {
        float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;  // independent chains: 4x shorter dependency
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(v[i + 2])); m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
    }
    TRACE(2);
    if (HALVES == 2) {  // both halves of a row must scale P and O with the same maximum
      float* slot = xm + (j & 1) * 256;
      slot[half * 128 + r] = mx;
      ptx::named_barrier_sync(bar_id, 256);
      mx = fmaxf(mx, slot[(half ^ 1) * 128 + r]);
    }
    TRACE(3);
    // lazy rescaling: the reference maximum m only moves when the true row maximum exceeds it by more than 2^8; until then P is
    // scaled with the stale m (values <= 256, exact in bf16/fp32 range) and O / l need no correction -- O/l is invariant to m
    const float m_true = mx * p.scale_log2;
    const float m_new = (m_true > m + 8.f) ? m_true : m;
    alpha = ex2(m - m_new);  // 0 on the first tile (m = -inf), 1 when the reference did not move
    // pass 2: p = exp2(s*scale - m_new), packed to bf16
    uint32_t pk[COLS / 2];
    float sum = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
    for (int c = 0; c < COLS; c += 32) {
      uint32_t v[32];
      if (c < valid) {
        ptx::tmem_ld32(ts + c, v);
        ptx::tmem_wait_ld();
      }
      if (c + 32 <= valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float p0 = ex2(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
          const float p1 = ex2(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -m_new));
          const float p2 = ex2(fmaf(__uint_as_float(v[i + 2]), p.scale_log2, -m_new));
          const float p3 = ex2(fmaf(__uint_as_float(v[i + 3]), p.scale_log2, -m_new));
          sum += p0; sum1 += p1; sum2 += p2; sum3 += p3;
          pk[(c + i) >> 1] = ptx::pack_bf16(p0, p1);
          pk[((c + i) >> 1) + 1] = ptx::pack_bf16(p2, p3);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = 0.f, p1 = 0.f;
          if (c + i < valid) p0 = ex2(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
          if (c + i + 1 < valid) p1 = ex2(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -m_new));
          sum += p0 + p1;
          pk[(c + i) >> 1] = ptx::pack_bf16(p0, p1);
        }
      }
    }
    l = l * alpha + ((sum + sum1) + (sum2 + sum3));
    moved = m_new > m;
    m = m_new;
    TRACE(4);
    // P buffer and O are free once P V of the previous tile has completed
    if (j > 0) {
      ptx::mbar_wait(o_ready, (j - 1) & 1);
      ptx::tc_fence_after();
    }
    TRACE(5);
    if constexpr (PT) {  // P goes to tensor memory (A operand of the PV MMA): lane = row, 2 bf16 per column, COLS/2 columns
      static_assert(COLS == 64 || COLS == 128, "P tile chunks of 32 packed columns");
#pragma unroll
      for (int c = 0; c < COLS / 2; c += 32)
        ptx::tmem_st32(tP + lane_off + half * (COLS / 2) + c, reinterpret_cast<const uint32_t(&)[32]>(pk[c]));
      ptx::tmem_wait_st();
    } else {
#pragma unroll
    for (int q = 0; q < COLS / 8; ++q) {  // 16-byte pieces: 8 per 64-key chunk, XOR-swizzled by (row & 7)
      uint8_t* dst = prow0 + (q >> 3) * TILE_BYTES + (((q & 7) ^ sw) << 4);
      *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
    }
    }
    if (j > 0 && __any_sync(0xffffffffu, moved)) {
#pragma unroll
      for (int c = 0; c < OCOLS; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(tO + lane_off + half * OCOLS + c, v);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
        ptx::tmem_st32(tO + lane_off + half * OCOLS + c, v);
      }
      ptx::tmem_wait_st();
    }
    TRACE(6);
    ptx::tc_fence_before();
    if constexpr (!PT) ptx::fence_proxy_async_smem();  // generic-proxy P writes -> visible to the tensor core (async proxy)
    ptx::mbar_arrive(p_full);
    TRACE(7);
  }
  // ---- output: O / l
  if (HALVES == 2) {  // row sum = sum over both column halves (they used identical maxima throughout)
    float* slot = xm + (n_tiles & 1) * 256;
    slot[half * 128 + r] = l;
    ptx::named_barrier_sync(bar_id, 256);
    l += slot[(half ^ 1) * 128 + r];
  }
  ptx::mbar_wait(o_ready, (n_tiles - 1) & 1);
  ptx::tc_fence_after();
  const int qi = q0 + r;
  const float inv = 1.f / l;
  bf16* orow = p.out + (size_t)b * p.out_batch_stride + (size_t)qi * p.out_pitch + h * HD + half * OCOLS;
#pragma unroll
  for (int c = 0; c < OCOLS; c += 32) {
    uint32_t v[32];
    ptx::tmem_ld32(tO + lane_off + half * OCOLS + c, v);
    ptx::tmem_wait_ld();
    if (qi < p.nq) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 u;
        u.x = ptx::pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
        u.y = ptx::pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
        u.z = ptx::pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
        u.w = ptx::pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + c + i) = u;
      }
    }
  }
}

// Production softmax of the pair kernel: two threads per query row (64 key columns each), P to tensor memory, and a LAZY
// reference maximum: tile 0 takes an exact row maximum (two TMEM passes); every later tile is ONE pass that scales with the
// reference m carried over from earlier tiles while tracking its own maximum; m is only raised (and O, l rescaled) at tile
// boundaries when the running maximum outgrew it by more than 2^8.  O / l is invariant to m, and P <= 2^(growth) stays far inside
// bf16 / fp32 range, so the result equals the exact online softmax up to rounding.
__device__ __forceinline__ void softmax_pair_lazy(const AttnParams& p, int n_tiles, uint32_t tS, uint32_t tO, uint32_t tP, uint32_t s_full,
                                                  uint32_t p_full, uint32_t o_ready, int ew, int lane, int q0, int h, int b, int half,
                                                  float* xm, uint32_t bar_id) {
  constexpr int COLS = BKV / 2, OCOLS = HD / 2;
  const int r = ew * 32 + lane;
  const uint32_t lane_off = (uint32_t)(ew * 32) << 16;
  const int col0 = half * COLS;
  const uint32_t ts = tS + lane_off + col0;
  float m = -INFINITY, l = 0.f, alpha_pending = 1.f;
  bool rescale_pending = false;
  const bool trace_ok = (ew == 0 && lane == 0);
  int trace_base = 0;
  for (int j = 0; j < n_tiles; ++j) {
    const int valid = min(COLS, p.nkv - j * BKV - col0);  // may be <= 0 for the upper half of a ragged last tile
    trace_base = 1024 * (int)(bar_id * 2 + half) + 8 * j;  // bar_id: 1 = tile A, 2 = tile B
    TRACE(0);
    ptx::mbar_wait(s_full, j & 1);
    ptx::tc_fence_after();
    TRACE(1);
    if (j == 0) {  // exact maximum for the first tile
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < COLS; c += 32) {
        if (c >= valid) break;
        uint32_t v[32];
        ptx::tmem_ld32(ts + c, v);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      xm[half * 128 + r] = mx;
      ptx::named_barrier_sync(bar_id, 256);
      m = fmaxf(mx, xm[(half ^ 1) * 128 + r]) * p.scale_log2;
    }
    // S(j) is issued BEFORE P(j-1) V (shorter critical chain), so O and the P region are only free again once o_ready(j-1) has
    // completed; by the time the first 32 keys have been exponentiated that MMA group (256 cycles) is long done.
    if (j > 0 && __any_sync(0xffffffffu, rescale_pending)) {  // the reference moved at the last boundary: bring O to the new scale
      ptx::mbar_wait(o_ready, (j - 1) & 1);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld32(tO + lane_off + half * OCOLS, v);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha_pending);
      ptx::tmem_st32(tO + lane_off + half * OCOLS, v);
    }
    // ---- one pass: p = exp2(s*scale - m) packed to bf16 and streamed to tensor memory 32 keys at a time, plus this tile's own
    // maximum for the next reference
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    float x0 = -INFINITY, x1 = -INFINITY, x2 = -INFINITY, x3 = -INFINITY;
    // (A variant with double-buffered 16-column TMEM reads -- next read in flight under the current exponentials -- measured SLOWER on
    // B200: 351 vs 333 us at 3072 tokens, 2257 vs 2124 us at 12288, profiles/r01_attn_bench.jsonl; the two 32-column reads stay.)
    {
#pragma unroll
    for (int c = 0; c < COLS; c += 32) {
      uint32_t v[32], pk[16];
      if (c < valid) {
        ptx::tmem_ld32(ts + c, v);
        ptx::tmem_wait_ld();
      }
      if (c == 0) TRACE(2);
      else TRACE(4);
      if (c + 32 <= valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float a0 = __uint_as_float(v[i]), a1 = __uint_as_float(v[i + 1]), a2 = __uint_as_float(v[i + 2]), a3 = __uint_as_float(v[i + 3]);
          x0 = fmaxf(x0, a0); x1 = fmaxf(x1, a1); x2 = fmaxf(x2, a2); x3 = fmaxf(x3, a3);
          const float p0 = ex2(fmaf(a0, p.scale_log2, -m)), p1 = ex2(fmaf(a1, p.scale_log2, -m));
          const float p2 = ex2(fmaf(a2, p.scale_log2, -m)), p3 = ex2(fmaf(a3, p.scale_log2, -m));
          s0 += p0; s1 += p1; s2 += p2; s3 += p3;
          pk[i >> 1] = ptx::pack_bf16(p0, p1);
          pk[(i >> 1) + 1] = ptx::pack_bf16(p2, p3);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = 0.f, p1 = 0.f;
          if (c + i < valid) { const float a = __uint_as_float(v[i]); x0 = fmaxf(x0, a); p0 = ex2(fmaf(a, p.scale_log2, -m)); }
          if (c + i + 1 < valid) { const float a = __uint_as_float(v[i + 1]); x1 = fmaxf(x1, a); p1 = ex2(fmaf(a, p.scale_log2, -m)); }
          s0 += p0; s1 += p1;
          pk[i >> 1] = ptx::pack_bf16(p0, p1);
        }
      }
      if (c == 0 && j > 0) {  // P(j-1) must have been consumed before it is overwritten
        ptx::mbar_wait(o_ready, (j - 1) & 1);
        ptx::tc_fence_after();
      }
      ptx::tmem_st16(tP + lane_off + half * (COLS / 2) + (c >> 1), pk);
      if (c == 0) TRACE(3);
      else TRACE(5);
    }
    }
    l += (s0 + s1) + (s2 + s3);
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    ptx::mbar_arrive(p_full);
    TRACE(6);
    // ---- off the MMA critical path: agree on the reference for the next tile
    if (j + 1 < n_tiles) {
      float* slot = xm + ((j + 1) & 1) * 256;
      const float mxl = fmaxf(fmaxf(x0, x1), fmaxf(x2, x3));
      slot[half * 128 + r] = mxl;
      ptx::named_barrier_sync(bar_id, 256);
      const float m_true = fmaxf(mxl, slot[(half ^ 1) * 128 + r]) * p.scale_log2;
      rescale_pending = m_true > m + 8.f;
      alpha_pending = rescale_pending ? ex2(m - m_true) : 1.f;
      if (rescale_pending) { l *= alpha_pending; m = m_true; }
    }
    TRACE(7);
  }
  // ---- output: O / l, row sum = sum over both column halves (identical references throughout)
  {
    float* slot = xm + ((n_tiles + 1) & 1) * 256;
    slot[half * 128 + r] = l;
    ptx::named_barrier_sync(bar_id, 256);
    l += slot[(half ^ 1) * 128 + r];
  }
  ptx::mbar_wait(o_ready, (n_tiles - 1) & 1);
  ptx::tc_fence_after();
  const int qi = q0 + r;
  const float inv = 1.f / l;
  bf16* orow = p.out + (size_t)b * p.out_batch_stride + (size_t)qi * p.out_pitch + h * HD + half * OCOLS;
  uint32_t v[32];
  ptx::tmem_ld32(tO + lane_off + half * OCOLS, v);
  ptx::tmem_wait_ld();
  if (qi < p.nq) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      uint4 u;
      u.x = ptx::pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
      u.y = ptx::pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
      u.z = ptx::pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
      u.w = ptx::pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
      *reinterpret_cast<uint4*>(orow + i) = u;
    }
  }
}

// EXPERIMENT (variant 6, not the default): softmax_pair_lazy with the two query tiles' exponential phases taking turns on the MUFU pipe
// (named-barrier token, ids 3 / 4) and 16-column TMEM reads double-buffered under the exponentials.  Round-2 measurements
// (profiles/r02_attn_experiments.txt): the same 331 us as variant 5 at 16 x 5 x 3072 tokens, and timing-only builds of this function
// showed why -- with the exponentials REMOVED the kernel still takes 281 us (-15 %), without the running-maximum FMNMX 313 us, without the
// in-phase TMEM reads 322 us: the d = 64 kernel is not MUFU-bound but bound by the per-KV-tile dependency chain softmax -> p_full ->
// MMA issue (~350 clk for S(j+1)) -> s_full -> TMEM read, with issue slots (57 %) and MUFU (60 %) both only moderately busy.  The way
// down is fewer instructions per score (packed f16x2 exponentials writing P directly, row sums from a ones column through the tensor
// core), not MUFU scheduling.
__device__ __forceinline__ void softmax_pair_pingpong(const AttnParams& p, int n_tiles, uint32_t tS, uint32_t tO, uint32_t tP, uint32_t s_full,
                                                      uint32_t p_full, uint32_t o_ready, int ew, int lane, int q0, int h, int b, int half,
                                                      float* xm, uint32_t wg, bool pp) {
  constexpr int COLS = BKV / 2, OCOLS = HD / 2;
  const uint32_t bar_id = 1 + wg, tok_mine = 3 + wg, tok_other = 4 - wg;
  const int r = ew * 32 + lane;
  const uint32_t lane_off = (uint32_t)(ew * 32) << 16;
  const int col0 = half * COLS;
  const uint32_t ts = tS + lane_off + col0;
  const uint32_t tp = tP + lane_off + half * (COLS / 2);
  float m = -INFINITY, l = 0.f, alpha_pending = 1.f;
  bool rescale_pending = false;
  const bool trace_ok = (ew == 0 && lane == 0);
  int trace_base = 0;
  if (pp && wg == 1) ptx::named_barrier_arrive(3, 512);  // prime: tile A owns the first turn
  for (int j = 0; j < n_tiles; ++j) {
    const int valid = min(COLS, p.nkv - j * BKV - col0);  // may be <= 0 for the upper half of a ragged last tile
    trace_base = 1024 * (int)(bar_id * 2 + half) + 8 * j;
    TRACE(0);
    ptx::mbar_wait(s_full, j & 1);
    ptx::tc_fence_after();
    TRACE(1);
    uint32_t va[16], vb[16];  // 16-column TMEM reads, double-buffered: the next read is in flight under the current exponentials
    if (j == 0) {  // exact maximum for the first tile (one extra pass over the 64 columns)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < COLS; c += 16) {
        if (c >= valid) break;
        ptx::tmem_ld16(ts + c, va);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c + i < valid) mx = fmaxf(mx, __uint_as_float(va[i]));
      }
      xm[half * 128 + r] = mx;
      ptx::named_barrier_sync(bar_id, 256);
      m = fmaxf(mx, xm[(half ^ 1) * 128 + r]) * p.scale_log2;
    } else {
      ptx::mbar_wait(o_ready, (j - 1) & 1);  // P(j-1) V(j-1) done: the P region and O are free (issued right behind S(j): long complete)
      ptx::tc_fence_after();
      if (__any_sync(0xffffffffu, rescale_pending)) {  // the reference moved at the last boundary: bring O to the new scale
        uint32_t o[32];
        ptx::tmem_ld32(tO + lane_off + half * OCOLS, o);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha_pending);
        ptx::tmem_st32(tO + lane_off + half * OCOLS, o);
      }
    }
    if (0 < valid) ptx::tmem_ld16(ts, va);
    ptx::tmem_wait_ld();
    TRACE(2);
    // ---------------------------------------------------------------- this tile's turn on the MUFU pipe
    if (pp) ptx::named_barrier_sync(tok_mine, 512);
    TRACE(3);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    float x0 = -INFINITY, x1 = -INFINITY, x2 = -INFINITY, x3 = -INFINITY;
    auto sub = [&](const uint32_t (&v)[16], int c) {  // 16 keys -> 8 packed columns of P
      uint32_t pk[8];
      if (c + 16 <= valid) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float a0 = __uint_as_float(v[i]), a1 = __uint_as_float(v[i + 1]), a2 = __uint_as_float(v[i + 2]), a3 = __uint_as_float(v[i + 3]);
          x0 = fmaxf(x0, a0); x1 = fmaxf(x1, a1); x2 = fmaxf(x2, a2); x3 = fmaxf(x3, a3);
          const float p0 = ex2(fmaf(a0, p.scale_log2, -m)), p1 = ex2(fmaf(a1, p.scale_log2, -m));
          const float p2 = ex2(fmaf(a2, p.scale_log2, -m)), p3 = ex2(fmaf(a3, p.scale_log2, -m));
          s0 += p0; s1 += p1; s2 += p2; s3 += p3;
          pk[i >> 1] = ptx::pack_bf16(p0, p1);
          pk[(i >> 1) + 1] = ptx::pack_bf16(p2, p3);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float p0 = 0.f, p1 = 0.f;
          if (c + i < valid) { const float a = __uint_as_float(v[i]); x0 = fmaxf(x0, a); p0 = ex2(fmaf(a, p.scale_log2, -m)); }
          if (c + i + 1 < valid) { const float a = __uint_as_float(v[i + 1]); x1 = fmaxf(x1, a); p1 = ex2(fmaf(a, p.scale_log2, -m)); }
          s0 += p0; s1 += p1;
          pk[i >> 1] = ptx::pack_bf16(p0, p1);
        }
      }
      ptx::tmem_st8(tp + (c >> 1), pk);
    };
    if (16 < valid) ptx::tmem_ld16(ts + 16, vb);
    sub(va, 0);
    ptx::tmem_wait_ld();
    if (32 < valid) ptx::tmem_ld16(ts + 32, va);
    sub(vb, 16);
    TRACE(4);
    ptx::tmem_wait_ld();
    if (48 < valid) ptx::tmem_ld16(ts + 48, vb);
    sub(va, 32);
    ptx::tmem_wait_ld();
    if (pp && !(wg == 1 && j == n_tiles - 1)) {
      // the last 16 exponentials of this tile overlap the other tile's start; the hand-over is a named-barrier arrival (non-blocking)
      ptx::named_barrier_arrive(tok_other, 512);
    }
    sub(vb, 48);
    TRACE(5);
    l += (s0 + s1) + (s2 + s3);
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    ptx::mbar_arrive(p_full);
    TRACE(6);
    // ---- off the MMA critical path, under the OTHER tile's exponentials: agree on the reference for the next tile
    if (j + 1 < n_tiles) {
      float* slot = xm + ((j + 1) & 1) * 256;
      const float mxl = fmaxf(fmaxf(x0, x1), fmaxf(x2, x3));
      slot[half * 128 + r] = mxl;
      ptx::named_barrier_sync(bar_id, 256);
      const float m_true = fmaxf(mxl, slot[(half ^ 1) * 128 + r]) * p.scale_log2;
      rescale_pending = m_true > m + 8.f;
      alpha_pending = rescale_pending ? ex2(m - m_true) : 1.f;
      if (rescale_pending) { l *= alpha_pending; m = m_true; }
    }
    TRACE(7);
  }
  // ---- output: O / l, row sum = sum over both column halves (identical references throughout)
  {
    float* slot = xm + ((n_tiles + 1) & 1) * 256;
    slot[half * 128 + r] = l;
    ptx::named_barrier_sync(bar_id, 256);
    l += slot[(half ^ 1) * 128 + r];
  }
  ptx::mbar_wait(o_ready, (n_tiles - 1) & 1);
  ptx::tc_fence_after();
  const int qi = q0 + r;
  const float inv = 1.f / l;
  bf16* orow = p.out + (size_t)b * p.out_batch_stride + (size_t)qi * p.out_pitch + h * HD + half * OCOLS;
  uint32_t v[32];
  ptx::tmem_ld32(tO + lane_off + half * OCOLS, v);
  ptx::tmem_wait_ld();
  if (qi < p.nq) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      uint4 u;
      u.x = ptx::pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
      u.y = ptx::pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
      u.z = ptx::pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
      u.w = ptx::pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
      *reinterpret_cast<uint4*>(orow + i) = u;
    }
  }
}

__device__ __forceinline__ void issue_qk(uint32_t tS, uint64_t qdesc, uint64_t kdesc) {
  constexpr uint32_t idesc_qk = ptx::idesc_bf16(128, BKV, 0, 0);
#pragma unroll
  for (int k = 0; k < HD / 16; ++k) ptx::mma_ss(tS, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
}
__device__ __forceinline__ void issue_pv(uint32_t tO, uint64_t pdesc0, uint64_t pdesc1, uint64_t vdesc, int first) {
  constexpr uint32_t idesc_pv = ptx::idesc_bf16(128, HD, 0, 1);  // B = V is MN-major
#pragma unroll
  for (int k = 0; k < BKV / 16; ++k)
    ptx::mma_ss(tO, (k < 4 ? pdesc0 + 2 * k : pdesc1 + 2 * (k - 4)), vdesc + 128 * k, idesc_pv, (!first) || (k != 0));
}

__device__ __forceinline__ void issue_pv_ts(uint32_t tO, uint32_t tP, uint64_t vdesc, int first) {
  constexpr uint32_t idesc_pv = ptx::idesc_bf16(128, HD, 0, 1);  // A = P from tensor memory (K-major), B = V MN-major
#pragma unroll
  for (int k = 0; k < BKV / 16; ++k) ptx::mma_ts(tO, tP + 8 * k, vdesc + 128 * k, idesc_pv, (!first) || (k != 0));
}

// ============================================================================================ one query tile per CTA
template <bool SHORT>
__global__ void __launch_bounds__(256, SHORT ? 2 : 1)
attention_single_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  constexpr int KV_STAGES = SHORT ? 1 : 2;
  constexpr uint32_t TMEM_COLS = SHORT ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + KV_STAGES * TILE_BYTES;
  uint8_t* sP = sV + KV_STAGES * TILE_BYTES;  // 2 K-chunks of [128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * TILE_BYTES);
  const uint32_t b0 = ptx::smem_u32(bars);
  const uint32_t q_full = b0, kv_full0 = b0 + 8, kv_empty0 = b0 + 24, s_full0 = b0 + 40, p_full = b0 + 56, o_ready = b0 + 64;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(kv_full0 + 8 * s, 1);
      ptx::mbar_init(kv_empty0 + 8 * s, 1);
      ptx::mbar_init(s_full0 + 8 * s, 1);
    }
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(o_ready, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + (SHORT ? 128 : 256);
  ptx::pdl_wait();

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(q_full, TILE_BYTES);
      ptx::tma_load_3d(&tmQ, ptx::smem_u32(sQ), q_full, h * HD, qt * BQ, b);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int st = SHORT ? 0 : (j & 1);
      ptx::mbar_wait(kv_empty0 + 8 * st, ((j >> 1) & 1) ^ 1);
      const uint32_t fb = kv_full0 + 8 * st;
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(fb, 2 * TILE_BYTES);
        ptx::tma_load_3d(&tmK, ptx::smem_u32(sK + st * TILE_BYTES), fb, h * HD, j * BKV, b);
        ptx::tma_load_3d(&tmV, ptx::smem_u32(sV + st * TILE_BYTES), fb, h * HD, j * BKV, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, one elected lane issues (see ptx::elect_one)
    const uint64_t qdesc = ptx::smem_desc_sw128(ptx::smem_u32(sQ));
    const uint64_t pdesc0 = ptx::smem_desc_sw128(ptx::smem_u32(sP));
    const uint64_t pdesc1 = ptx::smem_desc_sw128(ptx::smem_u32(sP + TILE_BYTES));
    ptx::mbar_wait(q_full, 0);
    for (int j = 0; j <= n_tiles; ++j) {
      if (j < n_tiles) {
        const int st = SHORT ? 0 : (j & 1);
        ptx::mbar_wait(kv_full0 + 8 * st, (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint64_t kd = ptx::smem_desc_sw128(ptx::smem_u32(sK + st * TILE_BYTES));
        if (ptx::elect_one()) {
          issue_qk(tS + st * BKV, qdesc, kd);
          ptx::mma_commit(s_full0 + 8 * st);
        }
        __syncwarp();
      }
      if (j > 0) {
        const int i = j - 1, st = SHORT ? 0 : (i & 1);
        ptx::mbar_wait(p_full, i & 1);
        ptx::tc_fence_after();
        const uint64_t vd = ptx::smem_desc_sw128(ptx::smem_u32(sV + st * TILE_BYTES));
        if (ptx::elect_one()) {
          issue_pv(tO, pdesc0, pdesc1, vd, i == 0);
          ptx::mma_commit(kv_empty0 + 8 * st);
          ptx::mma_commit(o_ready);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // SHORT has a single KV tile, so only S buffer 0 / parity 0 is ever used and the 2-buffer indexing below stays valid
    softmax_rows<1, !SHORT>(p, n_tiles, tS, BKV, SHORT ? 0u : 1u, tO, sP, s_full0, p_full, o_ready, warp & 3, lane, qt * BQ, h, b, 0, nullptr, 0);
  }

  __syncwarp();
  ptx::pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ============================================================================================ short K/V, persistent
// Cross-attention over the 77 text tokens (attn2 of every BasicTransformerBlock; nkv <= 128 = ONE K/V tile).  The one-tile-per-CTA kernel
// above spends most of its ~5 us per CTA on fixed costs -- barrier init, TMEM allocation, tensor-map prefetch, the first TMA round trip --
// for 1920 CTAs per launch at the 64x48 level.  Here a CTA is persistent: it loops over (query tile, head, image) items, two-stage TMA ring for
// Q / K / V so the loads of item i+1 fly under item i, P goes to tensor memory (no smem round trip), and two CTAs share an SM (96 KiB
// smem, 256 TMEM columns each) so one CTA's softmax runs under the other's MMAs and loads.
//   TMEM: S [0,128)  O [128,192)  P [192,256)       barriers: ld_full[2] ld_empty[2] s_full p_full o_full o_free
__global__ void __launch_bounds__(256, 2)
attention_short_persistent_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                                  const __grid_constant__ AttnParams p, int q_tiles, int heads, int items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // 2 stages
  uint8_t* sK = sQ + 2 * TILE_BYTES;        // 2 stages
  uint8_t* sV = sK + 2 * TILE_BYTES;        // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * TILE_BYTES);
  const uint32_t b0 = ptx::smem_u32(bars);
  const uint32_t ld_full0 = b0, ld_empty0 = b0 + 16, s_full = b0 + 32, p_full = b0 + 40, o_full = b0 + 48, o_free = b0 + 56;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ld_full0 + 8 * s, 1);
      ptx::mbar_init(ld_empty0 + 8 * s, 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_free, 128);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128, tP = tmem_base + 192;
  ptx::pdl_wait();

  auto decode = [&](int item, int& qt, int& h, int& b) {  // query tile fastest: consecutive items of a CTA re-read the same K/V from L2
    qt = item % q_tiles;
    const int r = item / q_tiles;
    h = r % heads;
    b = r / heads;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int st = it & 1;
      int qt, h, b;
      decode(item, qt, h, b);
      ptx::mbar_wait(ld_empty0 + 8 * st, ((it >> 1) & 1) ^ 1);
      const uint32_t fb = ld_full0 + 8 * st;
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(fb, 3 * TILE_BYTES);
        ptx::tma_load_3d(&tmQ, ptx::smem_u32(sQ + st * TILE_BYTES), fb, h * HD, qt * BQ, b);
        ptx::tma_load_3d(&tmK, ptx::smem_u32(sK + st * TILE_BYTES), fb, h * HD, 0, b);
        ptx::tma_load_3d(&tmV, ptx::smem_u32(sV + st * TILE_BYTES), fb, h * HD, 0, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int st = it & 1;
      ptx::mbar_wait(ld_full0 + 8 * st, (it >> 1) & 1);
      if (it > 0) ptx::mbar_wait(o_free, (it - 1) & 1);  // the softmax warps have read O(it-1) (hence S / P of it-1 are dead too)
      ptx::tc_fence_after();
      const uint64_t qd = ptx::smem_desc_sw128(ptx::smem_u32(sQ + st * TILE_BYTES));
      const uint64_t kd = ptx::smem_desc_sw128(ptx::smem_u32(sK + st * TILE_BYTES));
      const uint64_t vd = ptx::smem_desc_sw128(ptx::smem_u32(sV + st * TILE_BYTES));
      if (ptx::elect_one()) {
        issue_qk(tS, qd, kd);
        ptx::mma_commit(s_full);
      }
      __syncwarp();
      ptx::mbar_wait(p_full, it & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        issue_pv_ts(tO, tP, vd, 1);
        ptx::mma_commit(o_full);
        ptx::mma_commit(ld_empty0 + 8 * st);  // Q / K / V stage reusable
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output: one thread per query row, single K/V tile
    const int ew = warp & 3;
    const int r = ew * 32 + lane;
    const uint32_t lane_off = (uint32_t)(ew * 32) << 16;
    const int valid = p.nkv;  // <= 128
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      int qt, h, b;
      decode(item, qt, h, b);
      ptx::mbar_wait(s_full, it & 1);
      ptx::tc_fence_after();
      // pass 1: row maximum over the valid keys
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        if (c >= valid) break;
        uint32_t v[32];
        ptx::tmem_ld32(tS + lane_off + c, v);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m = mx * p.scale_log2;
      // pass 2: p = exp2(s * scale - m) packed to bf16 into tensor memory (zeros for the padding keys), row sum in fp32
      float l = 0.f;
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t v[32], pk[16];
        if (c < valid) {
          ptx::tmem_ld32(tS + lane_off + c, v);
          ptx::tmem_wait_ld();
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = 0.f, p1 = 0.f;
          if (c + i < valid) p0 = ex2(fmaf(__uint_as_float(v[i]), p.scale_log2, -m));
          if (c + i + 1 < valid) p1 = ex2(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -m));
          l += p0 + p1;
          pk[i >> 1] = ptx::pack_bf16(p0, p1);
        }
        ptx::tmem_st16(tP + lane_off + (c >> 1), pk);
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
      // output: O / l
      ptx::mbar_wait(o_full, it & 1);
      ptx::tc_fence_after();
      const int qi = qt * BQ + r;
      const float inv = 1.f / l;
      bf16* orow = p.out + (size_t)b * p.out_batch_stride + (size_t)qi * p.out_pitch + h * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(tO + lane_off + c, v);
        ptx::tmem_wait_ld();
        if (c + 32 == HD) {  // every accumulator column is in registers: S, P and O of this item may be overwritten
          ptx::tc_fence_before();
          ptx::mbar_arrive(o_free);
        }
        if (qi < p.nq) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 u;
            u.x = ptx::pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
            u.y = ptx::pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
            u.z = ptx::pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
            u.w = ptx::pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c + i) = u;
          }
        }
      }
    }
  }

  __syncwarp();
  ptx::pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}
constexpr size_t SMEM_SHORT_PERSIST = 6 * TILE_BYTES + 8 * 8 + 16 + 1024;

// ============================================================================================ two query tiles per CTA
// Tile B runs ONE KV tile behind tile A: while warpgroup A is in its softmax the tensor core executes B's PV / QK^T and vice
// versa, so MMA latency is hidden instead of being added to every iteration.  K/V ring of 3 stages (tile t is needed from
// S_A(t), issued in iteration t-1, until PV_B(t), issued in iteration t+1).
constexpr int PAIR_KV_STAGES = 3;
template <bool SR, bool PT, bool LAZY = false, bool PINGPONG = false>
__global__ void __launch_bounds__(640, 1)
attention_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 2 query tiles
  uint8_t* sK = sQ + 2 * TILE_BYTES;                    // 3 stages
  uint8_t* sV = sK + PAIR_KV_STAGES * TILE_BYTES;       // 3 stages
  uint8_t* sP = sV + PAIR_KV_STAGES * TILE_BYTES;       // 2 query tiles x 2 K-chunks
  float* xm = reinterpret_cast<float*>(sP + 4 * TILE_BYTES);  // row-max / row-sum exchange: [tile][parity][half][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xm + 2 * 2 * 256);
  const uint32_t b0 = ptx::smem_u32(bars);
  const uint32_t q_full = b0, kv_full0 = b0 + 8, kv_empty0 = b0 + 32, s_full0 = b0 + 56 /* A: +0, B: +8 */, p_full0 = b0 + 72,
                 o_ready0 = b0 + 88;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qp = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;
  const bool has_b = (qp * 2 + 1) * BQ < p.nq;  // second query tile holds at least one real row

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < PAIR_KV_STAGES; ++s) {
      ptx::mbar_init(kv_full0 + 8 * s, 1);
      ptx::mbar_init(kv_empty0 + 8 * s, has_b ? 2 : 1);  // one commit per query tile's MMA issuer
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(s_full0 + 8 * s, 1);
      ptx::mbar_init(p_full0 + 8 * s, 256);
      ptx::mbar_init(o_ready0 + 8 * s, 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)  P_A [384,448)  P_B [448,512)
  ptx::pdl_wait();
  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(q_full, (has_b ? 2 : 1) * TILE_BYTES);
      ptx::tma_load_3d(&tmQ, ptx::smem_u32(sQ), q_full, h * HD, qp * 2 * BQ, b);
      if (has_b) ptx::tma_load_3d(&tmQ, ptx::smem_u32(sQ + TILE_BYTES), q_full, h * HD, (qp * 2 + 1) * BQ, b);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j % PAIR_KV_STAGES;
      ptx::mbar_wait(kv_empty0 + 8 * st, ((j / PAIR_KV_STAGES) & 1) ^ 1);
      const uint32_t fb = kv_full0 + 8 * st;
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(fb, 2 * TILE_BYTES);
        ptx::tma_load_3d(&tmK, ptx::smem_u32(sK + st * TILE_BYTES), fb, h * HD, j * BKV, b);
        ptx::tma_load_3d(&tmV, ptx::smem_u32(sV + st * TILE_BYTES), fb, h * HD, j * BKV, b);
      }
      __syncwarp();
    }
  } else if (warp == 1 || (warp == 3 && has_b)) {
    // MMA issuers: ONE WARP PER QUERY TILE (warp 1 = tile A, warp 3 = tile B), each a warp-uniform loop with one elected lane
    // issuing (see ptx::elect_one).  Each tile's chain  softmax(j) -> P V(j), S(j+1) -> softmax(j+1)  advances on its own: a single
    // issuer serving A then B in program order forced the two tiles into lockstep (both softmaxes, then both MMA groups, measured
    // with tools/attn_trace.py: the MUFU pipe idled through the MMA phases and the tensor pipe through the softmax phases).
    // A K/V stage is released when BOTH issuers have committed their P V of that tile (kv_empty count 2).
    const int wg = warp == 1 ? 0 : 1;
    const uint64_t qd = ptx::smem_desc_sw128(ptx::smem_u32(sQ + wg * TILE_BYTES));
    const uint64_t p0 = ptx::smem_desc_sw128(ptx::smem_u32(sP + wg * 2 * TILE_BYTES)),
                   p1 = ptx::smem_desc_sw128(ptx::smem_u32(sP + wg * 2 * TILE_BYTES + TILE_BYTES));
    const uint32_t tS = tmem_base + 128 * wg, tO = tmem_base + 256 + 64 * wg, tP = tmem_base + 384 + 64 * wg;  // P tile in TMEM (PT)
    const uint32_t s_full = s_full0 + 8 * wg, p_full = p_full0 + 8 * wg, o_ready = o_ready0 + 8 * wg;
    const uint64_t k0 = ptx::smem_desc_sw128(ptx::smem_u32(sK)), v0 = ptx::smem_desc_sw128(ptx::smem_u32(sV));
    constexpr uint64_t STAGE_DESC = TILE_BYTES >> 4;  // descriptor start-address units per K/V stage
    const bool trace_ok = (lane == 0);
    int trace_base = 1024 * (6 + wg);
    ptx::mbar_wait(q_full, 0);
    if (wg == 1 && p.b_delay > 0) {
      // optional phase offset of tile B's first S = Q K^T behind tile A's (tuning knob, env LADI_ATTN_B_DELAY in cycles; measured
      // neutral on B200 -- profiles/r01_attn_bench.jsonl -- because the two tiles drift back into phase within ~10 KV tiles)
      ptx::mbar_wait(s_full0, 0);
      const long long t_start = clock64();
      while (clock64() - t_start < p.b_delay) {
      }
    }
    ptx::mbar_wait(kv_full0, 0);
    ptx::tc_fence_after();
    if (ptx::elect_one()) {
      issue_qk(tS, qd, k0);
      ptx::mma_commit(s_full);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      trace_base = 1024 * (6 + wg) + 8 * j;
      TRACE(0);
      const uint64_t kd_n = k0 + (uint64_t)((j + 1) % PAIR_KV_STAGES) * STAGE_DESC;
      const uint64_t vd_j = v0 + (uint64_t)(j % PAIR_KV_STAGES) * STAGE_DESC;
      // K/V tile j+1 was requested two tiles ago: observe it BEFORE blocking on the softmax, off the critical chain
      if (j + 1 < n_tiles) ptx::mbar_wait(kv_full0 + 8 * ((j + 1) % PAIR_KV_STAGES), ((j + 1) / PAIR_KV_STAGES) & 1);
      TRACE(1);
      ptx::mbar_wait(p_full, j & 1);  // P(j) is complete, hence S(j) has been consumed
      ptx::tc_fence_after();
      TRACE(2);
      // ---- S(j+1) FIRST (the softmax warps' critical chain is  P(j) -> S(j+1) -> softmax(j+1)), then O += P(j) V(j).  The softmax
      // of tile j+1 waits on o_ready(j) before it overwrites P or rescales O.
      if (j + 1 < n_tiles) {
        if (ptx::elect_one()) {
          issue_qk(tS, qd, kd_n);
          ptx::mma_commit(s_full);
        }
        __syncwarp();
      }
      TRACE(3);
      if (ptx::elect_one()) {
        if constexpr (PT) issue_pv_ts(tO, tP, vd_j, j == 0);
        else issue_pv(tO, p0, p1, vd_j, j == 0);
        ptx::mma_commit(o_ready);
        ptx::mma_commit(kv_empty0 + 8 * (j % PAIR_KV_STAGES));  // this tile is done with K/V tile j
      }
      __syncwarp();
      TRACE(4);
    }
  } else if (warp >= 4) {
    const int wg = (warp - 4) >> 3;         // 0 = tile A (warps 4-11), 1 = tile B (warps 12-19)
    const int half = ((warp - 4) >> 2) & 1;  // which half of the key columns / O columns this warpgroup owns
    if (wg == 0 || has_b) {
      if constexpr (PINGPONG)
        softmax_pair_pingpong(p, n_tiles, tmem_base + 128 * wg, tmem_base + 256 + 64 * wg, tmem_base + 384 + 64 * wg, s_full0 + 8 * wg,
                              p_full0 + 8 * wg, o_ready0 + 8 * wg, warp & 3, lane, (qp * 2 + wg) * BQ, h, b, half, xm + wg * 512, (uint32_t)wg, has_b);
      else if constexpr (LAZY)
        softmax_pair_lazy(p, n_tiles, tmem_base + 128 * wg, tmem_base + 256 + 64 * wg, tmem_base + 384 + 64 * wg, s_full0 + 8 * wg,
                          p_full0 + 8 * wg, o_ready0 + 8 * wg, warp & 3, lane, (qp * 2 + wg) * BQ, h, b, half, xm + wg * 512, 1 + wg);
      else
        softmax_rows<2, SR, PT>(p, n_tiles, tmem_base + 128 * wg, 0, 0u, tmem_base + 256 + 64 * wg, sP + wg * 2 * TILE_BYTES, s_full0 + 8 * wg,
                                p_full0 + 8 * wg, o_ready0 + 8 * wg, warp & 3, lane, (qp * 2 + wg) * BQ, h, b, half, xm + wg * 512, 1 + wg,
                                tmem_base + 384 + 64 * wg);
    }
  }

  __syncwarp();
  ptx::pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}


// ============================================================================================ one head, head_dim 512
// VAE mid-block AttentionBlock (/root/reference/src/models/vae.py:81-90,112 encoder; :142-150,187 decoder; diffusers 0.14
// AttentionBlock: 1 head, d = C = 512, N = h*w tokens -- 3072 at 512x384, 12288 at 1024x768).  Flash-style: the N x N score matrix
// never exists.  One CTA = one 128-row query tile x ONE HALF of the value/output width (blockIdx.y), because the fp32 accumulators of
// a full 128 x 512 output tile would fill all 512 tensor-memory columns by themselves:
//   TMEM  [0,128) S buffer 0   [128,256) S buffer 1   [256,512) O (128 x 256 fp32);  P(j) (bf16) overwrites S(j) in place;
//   SMEM  Q tile resident as 8 K-chunks [128 x 64] (128 KiB) + a ring of 16-KiB stages through which BOTH operands stream:
//         the 8 d-chunks [128 keys x 64] of K tile j (S += Q_c K_c^T), then the 4 column chunks [128 keys x 64] of V tile j
//         (O[:, 64c..] += P V_c, V read MN-major straight from the TMA tile);
//   issue order on the tensor pipe: QK(0), QK(1), PV(0), QK(2), PV(1), ...  so the softmax of tile j (MUFU floor ~1024 clk for
//         128 x 128 exponentials) runs under QK(j+1) (2048 clk) and the pipe never waits for it.
// The half split recomputes S once per half (1.5x the tensor work of an ideal kernel); d = 512 makes the kernel MMA-bound
// (3072 tensor clk vs ~1600 softmax clk per KV tile), unlike the d = 64 kernels above.
// Templated on the head width D in {512, 256}: 256 (the reduced-width test models) needs no value split (O = 256 columns).
constexpr int D5_VHALF = 256, D5_VCHUNKS = D5_VHALF / 64;
constexpr int D5_STAGES = 5;

template <int D5>
__global__ void __launch_bounds__(384, 1)
attention_d512_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int D5_CHUNKS = D5 / 64;
  uint8_t* sQ = smem;                                   // D/64 chunks [128 rows][64 d]
  uint8_t* sR = sQ + D5_CHUNKS * TILE_BYTES;            // operand ring
  float* xm = reinterpret_cast<float*>(sR + D5_STAGES * TILE_BYTES);  // row-max / row-sum exchange [parity][half][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xm + 2 * 256);
  const uint32_t b0 = ptx::smem_u32(bars);
  const uint32_t q_full = b0, full0 = b0 + 8, empty0 = full0 + 8 * D5_STAGES, s_full0 = empty0 + 8 * D5_STAGES, p_full0 = s_full0 + 16,
                 o_ready = p_full0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * D5_STAGES + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, vh = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < D5_STAGES; ++s) {
      ptx::mbar_init(full0 + 8 * s, 1);
      ptx::mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(s_full0 + 8 * s, 1);
      ptx::mbar_init(p_full0 + 8 * s, 256);
    }
    ptx::mbar_init(o_ready, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tO = tmem_base + 256;
  ptx::pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: Q once, then the K / V chunk stream in the
    // exact order the MMA issuer consumes it: K(0), K(1), V(0), K(2), V(1), ..., K(n-1), V(n-2), V(n-1)
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(q_full, D5_CHUNKS * TILE_BYTES);
      for (int c = 0; c < D5_CHUNKS; ++c) ptx::tma_load_3d(&tmQ, ptx::smem_u32(sQ + c * TILE_BYTES), q_full, c * 64, qt * BQ, b);
    }
    __syncwarp();
    uint32_t stage = 0, phase = 0;
    auto load_chunks = [&](const CUtensorMap* tm, int col0, int nchunks, int tile) {
      for (int c = 0; c < nchunks; ++c) {
        ptx::mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t fb = full0 + 8 * stage;
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(fb, TILE_BYTES);
          ptx::tma_load_3d(tm, ptx::smem_u32(sR + stage * TILE_BYTES), fb, col0 + c * 64, tile * BKV, b);
        }
        __syncwarp();
        if (++stage == D5_STAGES) { stage = 0; phase ^= 1; }
      }
    };
    load_chunks(&tmK, 0, D5_CHUNKS, 0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) load_chunks(&tmK, 0, D5_CHUNKS, j + 1);
      load_chunks(&tmV, vh * D5_VHALF, D5_VCHUNKS, j);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues)
    constexpr uint32_t idesc_qk = ptx::idesc_bf16(128, BKV, 0, 0);
    constexpr uint32_t idesc_pv = ptx::idesc_bf16(128, 64, 0, 1);  // A = P from tensor memory, B = V chunk, MN-major
    const uint64_t q0 = ptx::smem_desc_sw128(ptx::smem_u32(sQ));
    constexpr uint64_t CHUNK_DESC = TILE_BYTES >> 4;
    uint32_t stage = 0, phase = 0;
    auto issue_qk_tile = [&](int tile) {   // S(tile) = Q K(tile)^T over the 8 d-chunks
      const uint32_t tS = tmem_base + (tile & 1) * BKV;
      for (int c = 0; c < D5_CHUNKS; ++c) {
        ptx::mbar_wait(full0 + 8 * stage, phase);
        ptx::tc_fence_after();
        const uint64_t kd = ptx::smem_desc_sw128(ptx::smem_u32(sR + stage * TILE_BYTES));
        const uint64_t qd = q0 + (uint64_t)c * CHUNK_DESC;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::mma_ss(tS, qd + 2 * k, kd + 2 * k, idesc_qk, (c | k) != 0);
          ptx::mma_commit(empty0 + 8 * stage);
          if (c == D5_CHUNKS - 1) ptx::mma_commit(s_full0 + 8 * (tile & 1));
        }
        __syncwarp();
        if (++stage == D5_STAGES) { stage = 0; phase ^= 1; }
      }
    };
    ptx::mbar_wait(q_full, 0);
    issue_qk_tile(0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) issue_qk_tile(j + 1);            // overwrites S/P buffer (j+1)&1: P(j-1) V was issued before -> in order
      ptx::mbar_wait(p_full0 + 8 * (j & 1), (j >> 1) & 1);  // P(j) is in tensor memory (and O has been rescaled if needed)
      ptx::tc_fence_after();
      const uint32_t tP = tmem_base + (j & 1) * BKV;
      for (int c = 0; c < D5_VCHUNKS; ++c) {
        ptx::mbar_wait(full0 + 8 * stage, phase);
        ptx::tc_fence_after();
        const uint64_t vd = ptx::smem_desc_sw128(ptx::smem_u32(sR + stage * TILE_BYTES));
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) ptx::mma_ts(tO + 64 * c, tP + 8 * k, vd + 128 * k, idesc_pv, (j != 0) || (k != 0));
          ptx::mma_commit(empty0 + 8 * stage);
          if (c == D5_VCHUNKS - 1) ptx::mma_commit(o_ready);
        }
        __syncwarp();
        if (++stage == D5_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax: two threads per query row (64 key columns each)
    const int ew = warp & 3, half = (warp - 4) >> 2;
    const int r = ew * 32 + lane;
    const uint32_t lane_off = (uint32_t)(ew * 32) << 16;
    constexpr int COLS = BKV / 2, OCOLS = D5_VHALF / 2;
    const int col0 = half * COLS;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int valid = min(COLS, p.nkv - j * BKV - col0);  // may be <= 0 for the upper half of a ragged last tile
      const uint32_t tS = tmem_base + (j & 1) * BKV + lane_off;
      ptx::mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t v[COLS];
      ptx::tmem_ld32(tS + col0, reinterpret_cast<uint32_t(&)[32]>(v[0]));
      ptx::tmem_ld32(tS + col0 + 32, reinterpret_cast<uint32_t(&)[32]>(v[32]));
      ptx::tmem_wait_ld();
      float mx = -INFINITY;
      if (valid >= COLS) {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < COLS; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(v[i + 2])); m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      } else {
#pragma unroll
        for (int i = 0; i < COLS; ++i)
          if (i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      // both halves of a row agree on the maximum; the barrier also orders "both halves have read S(j)" before "P(j) overwrites it"
      float* slot = xm + (j & 1) * 256;
      slot[half * 128 + r] = mx;
      ptx::tc_fence_before();
      ptx::named_barrier_sync(1, 256);
      ptx::tc_fence_after();
      mx = fmaxf(mx, slot[(half ^ 1) * 128 + r]);
      // lazy rescaling: the reference maximum only moves when the true maximum outgrew it by more than 2^8 (P <= 256, exact range)
      const float m_true = mx * p.scale_log2;
      const float m_new = (m_true > m + 8.f) ? m_true : m;
      const float alpha = ex2(m - m_new);  // 0 on the first tile
      const bool moved = m_new > m;
      m = m_new;
      uint32_t pk[COLS / 2];
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      if (valid >= COLS) {
#pragma unroll
        for (int i = 0; i < COLS; i += 4) {
          const float p0 = ex2(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new)), p1 = ex2(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -m_new));
          const float p2 = ex2(fmaf(__uint_as_float(v[i + 2]), p.scale_log2, -m_new)), p3 = ex2(fmaf(__uint_as_float(v[i + 3]), p.scale_log2, -m_new));
          s0 += p0; s1 += p1; s2 += p2; s3 += p3;
          pk[i >> 1] = ptx::pack_bf16(p0, p1);
          pk[(i >> 1) + 1] = ptx::pack_bf16(p2, p3);
        }
      } else {
#pragma unroll
        for (int i = 0; i < COLS; i += 2) {
          float p0 = 0.f, p1 = 0.f;
          if (i < valid) p0 = ex2(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
          if (i + 1 < valid) p1 = ex2(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -m_new));
          s0 += p0 + p1;
          pk[i >> 1] = ptx::pack_bf16(p0, p1);
        }
      }
      l = l * alpha + ((s0 + s1) + (s2 + s3));
      // EVERY phase of o_ready is observed, in order (a parity wait only tells "the phase before the current one": skipping phases made
      // the final wait below pass while only P(0) V(0) had completed whenever the V stream ran late).  P(j-1) V(j-1) runs right behind
      // S(j) on the tensor pipe, so by now it is normally long done.
      if (j > 0) {
        ptx::mbar_wait(o_ready, (j - 1) & 1);
        ptx::tc_fence_after();
      }
      if (j > 0 && __any_sync(0xffffffffu, moved)) {   // rare: bring this thread's 128 accumulator columns to the new scale
#pragma unroll 1
        for (int c = 0; c < OCOLS; c += 32) {
          uint32_t o[32];
          ptx::tmem_ld32(tO + lane_off + half * OCOLS + c, o);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          ptx::tmem_st32(tO + lane_off + half * OCOLS + c, o);
        }
      }
      ptx::tmem_st32(tS + half * (COLS / 2), reinterpret_cast<const uint32_t(&)[32]>(pk[0]));  // P(j) in place of S(j)
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full0 + 8 * (j & 1));
    }
    // ---- output: O / l (row sum over both halves: identical references throughout)
    {
      float* slot = xm + (n_tiles & 1) * 256;
      slot[half * 128 + r] = l;
      ptx::named_barrier_sync(1, 256);
      l += slot[(half ^ 1) * 128 + r];
    }
    ptx::mbar_wait(o_ready, (n_tiles - 1) & 1);
    ptx::tc_fence_after();
    const int qi = qt * BQ + r;
    const float inv = 1.f / l;
    bf16* orow = p.out + (size_t)b * p.out_batch_stride + (size_t)qi * p.out_pitch + vh * D5_VHALF + half * OCOLS;
#pragma unroll 1
    for (int c = 0; c < OCOLS; c += 32) {
      uint32_t o[32];
      ptx::tmem_ld32(tO + lane_off + half * OCOLS + c, o);
      ptx::tmem_wait_ld();
      if (qi < p.nq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = ptx::pack_bf16(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          u.y = ptx::pack_bf16(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          u.z = ptx::pack_bf16(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          u.w = ptx::pack_bf16(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c + i) = u;
        }
      }
    }
  }

  __syncwarp();
  ptx::pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

constexpr size_t smem_d512(int d) { return (size_t)(d / 64 + D5_STAGES) * TILE_BYTES + 2 * 256 * 4 + (1 + 2 * D5_STAGES + 5) * 8 + 16 + 1024; }
static_assert(smem_d512(512) <= 232448, "d=512 attention: shared memory over the 227 KiB per-CTA limit");

constexpr size_t SMEM_SINGLE = 7 * TILE_BYTES + 10 * 8 + 16 + 1024;
constexpr size_t SMEM_SHORT = 5 * TILE_BYTES + 10 * 8 + 16 + 1024;
constexpr size_t SMEM_PAIR = (2 + 2 * PAIR_KV_STAGES + 4) * TILE_BYTES + 4 * 256 * 4 + 14 * 8 + 16 + 1024;

int make_map(CUtensorMap* m, const void* ptr, int cols, int tokens, int batch, int pitch, int64_t batch_stride) {
  const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)tokens, (uint64_t)batch};
  const uint64_t strides[2] = {(uint64_t)pitch * 2, (uint64_t)batch_stride * 2};
  const uint32_t box[3] = {(uint32_t)HD, 128u, 1u};
  return ladi_encode_tmap_bf16(m, ptr, 3, dims, strides, box);
}

}  // namespace

extern "C" int ladi_attention_bf16(const ladi_attn_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LADI_CHECK(d != nullptr && d->q && d->k && d->v && d->out, "null attention operand");
  LADI_CHECK(d->batch > 0 && d->heads > 0 && d->nq > 0 && d->nkv > 0, "bad attention extent");
  LADI_CHECK(d->q_pitch % 8 == 0 && d->k_pitch % 8 == 0 && d->v_pitch % 8 == 0 && d->out_pitch % 8 == 0, "pitches must be multiples of 8");
  LADI_CHECK(d->q_pitch >= d->heads * HD && d->k_pitch >= d->heads * HD && d->v_pitch >= d->heads * HD, "pitch < heads*64");
  CUtensorMap tq, tk, tv;
  if (make_map(&tq, d->q, d->heads * HD, d->nq, d->batch, d->q_pitch, d->q_batch_stride)) return LADI_ERR_CUDA;
  if (make_map(&tk, d->k, d->heads * HD, d->nkv, d->batch, d->k_pitch, d->k_batch_stride)) return LADI_ERR_CUDA;
  if (make_map(&tv, d->v, d->heads * HD, d->nkv, d->batch, d->v_pitch, d->v_batch_stride)) return LADI_ERR_CUDA;
  AttnParams p;
  p.nq = d->nq; p.nkv = d->nkv; p.n_kv_tiles = (d->nkv + BKV - 1) / BKV;
  p.out = reinterpret_cast<bf16*>(d->out); p.out_pitch = d->out_pitch; p.out_batch_stride = d->out_batch_stride;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.trace = reinterpret_cast<long long*>(d->trace);
  {
    static int b_delay = -1;  // env LADI_ATTN_B_DELAY (cycles) for tuning (tools/attn_bench.py); default 0
    if (b_delay < 0) {
      const char* e = getenv("LADI_ATTN_B_DELAY");
      b_delay = e != nullptr ? atoi(e) : 0;
      if (b_delay < 0) b_delay = 0;
    }
    p.b_delay = b_delay;
  }
  static bool attr_set = false;
  if (!attr_set) {
    LADI_CUDA(cudaFuncSetAttribute(attention_single_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SINGLE));
    LADI_CUDA(cudaFuncSetAttribute(attention_single_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SHORT));
    LADI_CUDA(cudaFuncSetAttribute(attention_short_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SHORT_PERSIST));
    LADI_CUDA(cudaFuncSetAttribute(attention_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PAIR));
    LADI_CUDA(cudaFuncSetAttribute(attention_pair_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PAIR));
    LADI_CUDA(cudaFuncSetAttribute(attention_pair_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PAIR));
    LADI_CUDA(cudaFuncSetAttribute(attention_pair_kernel<false, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PAIR));
    attr_set = true;
  }
  // variant: 0 auto, 1 one query tile per CTA, 2 pair (P in smem), 4 pair (P in tensor memory), 5 pair (P in TMEM + lazy single-pass softmax)
  int variant = d->variant;
  // auto: one K/V tile (cross-attention over the text tokens, tiny self-attention) -> the persistent short-K/V kernel (1.14-1.41x over one tile per
  // CTA, profiles/r02_xattn_bench.jsonl); long sequences -> two query tiles per CTA from ~512 queries up
  if (variant == 0) variant = d->nkv <= BKV ? 8 : (d->nq < 512 ? 1 : 5);
  if (variant == 8) {
    LADI_CHECK(d->nkv <= BKV, "variant 8 (persistent short-K/V kernel) needs nkv <= 128");
    const int q_tiles = (d->nq + BQ - 1) / BQ, items = q_tiles * d->heads * d->batch;
    const int grid = items < 2 * ladi_num_sms() ? items : 2 * ladi_num_sms();
    LADI_CUDA(ladi_launch(attention_short_persistent_kernel, dim3(grid), dim3(256), SMEM_SHORT_PERSIST, stream, tq, tk, tv, p, q_tiles, d->heads, items));
  } else if (variant == 1 && d->nkv <= BKV) {
    dim3 grid((d->nq + BQ - 1) / BQ, d->heads, d->batch);
    LADI_CUDA(ladi_launch(attention_single_kernel<true>, grid, dim3(256), SMEM_SHORT, stream, tq, tk, tv, p));
  } else if (variant == 1) {
    dim3 grid((d->nq + BQ - 1) / BQ, d->heads, d->batch);
    LADI_CUDA(ladi_launch(attention_single_kernel<false>, grid, dim3(256), SMEM_SINGLE, stream, tq, tk, tv, p));
  } else {
    dim3 grid((d->nq + 2 * BQ - 1) / (2 * BQ), d->heads, d->batch);
    if (variant == 4) LADI_CUDA(ladi_launch(attention_pair_kernel<false, true>, grid, dim3(640), SMEM_PAIR, stream, tq, tk, tv, p));
    else if (variant == 5) LADI_CUDA(ladi_launch(attention_pair_kernel<false, true, true>, grid, dim3(640), SMEM_PAIR, stream, tq, tk, tv, p));
    else if (variant == 6) LADI_CUDA(ladi_launch(attention_pair_kernel<false, true, true, true>, grid, dim3(640), SMEM_PAIR, stream, tq, tk, tv, p));
    else LADI_CUDA(ladi_launch(attention_pair_kernel<false, false>, grid, dim3(640), SMEM_PAIR, stream, tq, tk, tv, p));
  }
  return LADI_OK;
}

extern "C" int ladi_attention_d512_bf16(const ladi_attn_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LADI_CHECK(d != nullptr && d->q && d->k && d->v && d->out, "null attention operand");
  LADI_CHECK(d->batch > 0 && d->heads == 1 && d->nq > 0 && d->nkv > 0, "d512 attention: one head, non-empty extents");
  LADI_CHECK(d->q_pitch % 8 == 0 && d->k_pitch % 8 == 0 && d->v_pitch % 8 == 0 && d->out_pitch % 8 == 0, "pitches must be multiples of 8");
  const int D5 = d->head_dim == 0 ? 512 : d->head_dim;
  LADI_CHECK(D5 == 512 || D5 == 256, "wide attention: head_dim must be 512 (or 256), got %d", D5);
  LADI_CHECK(d->q_pitch >= D5 && d->k_pitch >= D5 && d->v_pitch >= D5 && d->out_pitch >= D5, "pitch < head_dim");
  LADI_CHECK((reinterpret_cast<uintptr_t>(d->out) & 15) == 0, "out must be 16-byte aligned");
  CUtensorMap tq, tk, tv;
  if (make_map(&tq, d->q, D5, d->nq, d->batch, d->q_pitch, d->q_batch_stride)) return LADI_ERR_CUDA;
  if (make_map(&tk, d->k, D5, d->nkv, d->batch, d->k_pitch, d->k_batch_stride)) return LADI_ERR_CUDA;
  if (make_map(&tv, d->v, D5, d->nkv, d->batch, d->v_pitch, d->v_batch_stride)) return LADI_ERR_CUDA;
  AttnParams p;
  p.nq = d->nq; p.nkv = d->nkv; p.n_kv_tiles = (d->nkv + BKV - 1) / BKV;
  p.out = reinterpret_cast<bf16*>(d->out); p.out_pitch = d->out_pitch; p.out_batch_stride = d->out_batch_stride;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.b_delay = 0; p.trace = nullptr;
  static bool attr_set = false;
  if (!attr_set) {
    LADI_CUDA(cudaFuncSetAttribute(attention_d512_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d512(512)));
    LADI_CUDA(cudaFuncSetAttribute(attention_d512_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d512(256)));
    attr_set = true;
  }
  dim3 grid((d->nq + BQ - 1) / BQ, D5 / D5_VHALF, d->batch);
  if (D5 == 512) LADI_CUDA(ladi_launch(attention_d512_kernel<512>, grid, dim3(384), smem_d512(512), stream, tq, tk, tv, p));
  else LADI_CUDA(ladi_launch(attention_d512_kernel<256>, grid, dim3(384), smem_d512(256), stream, tq, tk, tv, p));
  return LADI_OK;
}
