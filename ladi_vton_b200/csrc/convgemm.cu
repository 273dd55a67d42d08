// Implicit-GEMM convolution / GEMM for sm_100a: TMA (4-D shifted boxes, zero OOB fill = conv padding) -> 128B-swizzled smem
// -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> fused epilogue (bias / per-step bias / SiLU / GEGLU / residual / row
// scale) -> bf16 or fp32 NHWC.  One persistent CTA per SM; warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4-7 = epilogue; TMEM accumulator double-buffered so the epilogue of tile i overlaps the MMAs of i+1.
//
// Epilogue (bf16 outputs): accumulator slabs of 64 output columns go TMEM -> registers -> 128B-swizzled staging tile in
// shared memory -> ONE TMA store per slab (hardware clips rows/columns outside the tensor, so ragged tiles need no masks);
// the residual operand of the same slab is prefetched by TMA into a second staging tile.  Both are double-buffered.
// fp32 / tiny-N outputs (conv_out, VAE moments, the VAE S matrix) use a direct-store epilogue.
//
// Replaces, on the reference hot path: every nn.Conv2d / nn.Linear executed by UNet2DConditionModel.forward
// (/root/reference/src/vto_pipelines/tryon_pipe.py:732), AutoencoderKL.encode/decode (src/models/vae.py:99-119,183-212) and
// EMASC.forward (src/models/emasc.py:37-40), which the reference runs as cuDNN/cuBLAS library calls.
//
// A GEMM is the degenerate convolution: 1 tap, H=1, W=M.  A 3x3 stride-2 convolution reads the four parity planes of its
// input through tensor maps with doubled strides; a channel concat (UNet skip connections) or a fused 1x1 shortcut is just
// more K segments read through more tensor maps -- the concatenated tensor is never materialised.
#include "common.h"
#include "ptx.cuh"

namespace {

constexpr int BM = 128;   // output pixels (rows) per tile == UMMA M
constexpr int BK = 64;    // bf16 channels per K block == one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int MAX_SEG = 24;
constexpr int NUM_A_MAPS = 8;
constexpr int MAX_STAGES = 8;
constexpr int A_BYTES = BM * BK * 2;
constexpr int SLAB_BYTES = BM * 128;  // staging tile: 128 rows x 64 bf16
constexpr int RES_BUFS = 3;           // residual slabs in flight (TMA prefetch ring, runs ahead across tile boundaries)

struct Segment {       // one run of K blocks read from one tensor map with one spatial shift
  int16_t map, dx, dy, chunks;
  int32_t c_begin;
};

struct KParams {
  int n_img, H, W;            // output extent
  int bn, bh, bw;             // M tile = bn x bh x bw pixels (product 128)
  int tiles_x, tiles_y, tiles_b, tiles_m, tiles_n;
  int c_out;
  int nseg, total_kb;
  int stages;                 // smem pipeline depth (runtime: whatever fits beside the staging tiles)
  int splits, kb_per_split;   // split-K: work item = (tile, split); split s reduces K blocks [s*kb_per_split, ...) into fp32 partials
  long long split_stride;     // elements between the partial planes of consecutive splits
  int staged;                 // 1 = smem-staged TMA-store epilogue (bf16 out), 0 = direct stores
  Segment seg[MAX_SEG];
  const float* bias;          // [c_out] fp32 (per column) or [M] (per row) or null
  int bias_per_row;
  int bias_step_stride;       // bias += (*step_ptr) * stride   (per-DDIM-step time-embedding bias table)
  const int* step_ptr;
  const bf16* residual;       // [rows, residual_pitch] or null
  int residual_pitch;
  const float* row_scale;     // [rows] or null (EMASC (1-mask))
  int act;                    // 0 none, 1 SiLU, 2 GEGLU (interleaved value/gate columns -> c_out/2 outputs)
  void* out;
  int out_pitch;
  int out_fp32;
  // LayerNorm folded into the GEMMs on either side of it (BasicTransformerBlock.norm1/2/3):
  float2* rowstat_out;        // producer: per (row, 32-column chunk) {sum, sum of squares} of the values this GEMM stores
  int rowstat_chunks;         //           = c_out / 32
  const float2* ln_stats;     // consumer: the producer's table for this GEMM's A rows; out = rstd*(acc - mean*colsum[c]) + bias[c]
  int ln_chunks;              //           = K / 32 (the normalised width)
  const float* ln_colsum;     //           [c_out] fp32: sum_k of the (gamma-scaled, bf16-rounded) weight row
  float ln_eps;
  // nearest-2x upsample fused into a 3x3 convolution (diffusers Upsample2D): the 4 output parities are 4 groups of M tiles, each a
  // 2x2-tap convolution over the half-resolution input with its own merged weights and a strided output tensor map
  int up2x, tiles_pp;         // tiles_pp = M tiles per parity plane
};

struct AMaps {
  CUtensorMap m[NUM_A_MAPS];
};
struct EMaps {            // epilogue tensor maps: 64-column (SWIZZLE_128B) and 32-column (SWIZZLE_64B) boxes
  CUtensorMap out64[4], out32[4], res64, res32;   // out maps indexed by output parity (index 0 unless up2x)
};

// x * sigmoid(x) = h * (1 + tanh(h)), h = x / 2: one MUFU op (tanh.approx, relative error 2^-11, below bf16 output resolution)
// instead of ex2 + a guarded division
__device__ __forceinline__ float silu_f(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// GELU(erf): erf(z) = z * Q(z^2) for |z| <= 3 (degree-8 minimax fit in z^2, clamped to +-1 beyond): |erf error| <= 4.8e-5,
// |gelu error| <= 1e-4 absolute -- below bf16 output resolution -- in 13 FMA-pipe instructions and NO MUFU op.  The GEGLU epilogue
// evaluates it for every element of the widest GEMMs of the UNet, where erff (~30 instr) or an exp-based form made the
// epilogue, not the tensor core, the bottleneck.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fminf(fabsf(x) * 0.70710678118654752f, 3.0f);
  const float t = z * z;
  float q = 5.152028155e-08f;
  q = fmaf(q, t, -2.354745293e-06f);
  q = fmaf(q, t, 4.747267050e-05f);
  q = fmaf(q, t, -5.641741965e-04f);
  q = fmaf(q, t, 4.485662883e-03f);
  q = fmaf(q, t, -2.576915092e-02f);
  q = fmaf(q, t, 1.120152109e-01f);
  q = fmaf(q, t, -3.758995125e-01f);
  q = fmaf(q, t, 1.128372696e+00f);
  const float e = fminf(z * q, 1.0f);
  return 0.5f * x * (1.f + copysignf(e, x));
}

struct TileCoord {
  int nt, x0, y0, n0, par;
};
__device__ __forceinline__ TileCoord tile_coord(const KParams& p, int tile) {
  TileCoord t;
  int mt = tile / p.tiles_n;
  t.nt = tile - mt * p.tiles_n;
  t.par = 0;
  if (p.up2x) {
    t.par = mt / p.tiles_pp;
    mt -= t.par * p.tiles_pp;
  }
  const int tb = mt / (p.tiles_y * p.tiles_x);
  const int rem = mt - tb * (p.tiles_y * p.tiles_x);
  const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
  t.x0 = tx * p.bw; t.y0 = ty * p.bh; t.n0 = tb * p.bn;
  return t;
}

// PAIR = true: launched as clusters of two CTAs (the two SMs of a TPC) that execute tcgen05.mma.cta_group::2 with M = 256: the
// cluster owns two M tiles (one per CTA) of the same N tile; each CTA stages its own 128 A rows and HALF of the B rows, so the
// shared-memory traffic per MMA (TMA writes + operand reads) drops from A + B to A + B/2.  Only the leader issues MMAs; the
// epilogue is per CTA and identical to the single-CTA kernel.
template <int BN, bool PAIR>
__global__ void __launch_bounds__(384, 1)
convgemm_kernel(const __grid_constant__ AMaps amaps, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ EMaps emaps,
                const __grid_constant__ KParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int B_ROWS = PAIR ? BN / 2 : BN;   // B rows staged by this CTA
  constexpr int B_BYTES = B_ROWS * BK * 2;
  constexpr uint32_t ACC_STRIDE = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // TMEM columns per accumulator
  constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;
  const int STAGES = p.stages;
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * A_BYTES;
  uint8_t* sOut = sB + STAGES * B_BYTES;                         // 2 staging tiles (staged epilogue only)
  uint8_t* sRes = sOut + (p.staged ? 2 * SLAB_BYTES : 0);        // residual ring (staged + residual only)
  uint8_t* sEnd = sRes + ((p.staged && p.residual != nullptr) ? RES_BUFS * SLAB_BYTES : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEnd);
  const uint32_t full0 = ptx::smem_u32(bars), empty0 = full0 + 8 * MAX_STAGES, tfull0 = empty0 + 8 * MAX_STAGES,
                 tempty0 = tfull0 + 16, rfull0 = tempty0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4 + RES_BUFS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;      // 0 = leader (issues the MMAs)
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) ptx::prefetch_tmap(&amaps.m[p.seg[s].map]);
    ptx::prefetch_tmap(&tmB);
    if (p.staged) {
      ptx::prefetch_tmap(&emaps.out64[0]);
      if (p.residual != nullptr) ptx::prefetch_tmap(&emaps.res64);
    }
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full0 + 8 * s, 1);
      ptx::mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull0 + 8 * a, 1);
      ptx::mbar_init(tempty0 + 8 * a, PAIR ? 512 : 256);  // PAIR: the epilogue threads of BOTH CTAs release the leader's accumulator
    }
    for (int a = 0; a < RES_BUFS; ++a) ptx::mbar_init(rfull0 + 8 * a, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) ptx::tmem_alloc_pair(ptx::smem_u32(tmem_slot), TMEM_COLS);
    else ptx::tmem_alloc(ptx::smem_u32(tmem_slot), TMEM_COLS);
  }
  ptx::tc_fence_before();
  if constexpr (PAIR) ptx::cluster_sync();  // the peer's barriers are initialised before anything is signalled across the pair
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();  // everything above overlapped the previous kernel's tail; its results are visible from here on

  // work items: single CTA = (tile, K split); pair = (pair of M tiles, N tile), this CTA taking M tile 2*pm + rank (a phantom tile
  // past the end when tiles_m is odd: its loads are zero-filled and its stores clipped by the TMA unit / the row_ok mask)
  const int num_items = PAIR ? ((p.tiles_m + 1) >> 1) * p.tiles_n : p.tiles_m * p.tiles_n * p.splits;
  auto item_tile = [&](int item) -> int {
    if constexpr (PAIR) {
      const int pm = item / p.tiles_n;
      return (2 * pm + (int)rank) * p.tiles_n + (item - pm * p.tiles_n);
    } else {
      return item / p.splits;
    }
  };
  // the leader's barriers as shared::cluster addresses (what the peer signals)
  const uint32_t full_leader0 = PAIR ? ptx::mapa(full0, 0) : full0;
  const uint32_t tempty_leader0 = PAIR ? ptx::mapa(tempty0, 0) : tempty0;
  auto release_acc = [&](uint32_t as) {
    if constexpr (PAIR) ptx::mbar_arrive_cluster(tempty_leader0 + 8 * as);
    else ptx::mbar_arrive(tempty0 + 8 * as);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform loop, one elected lane issues)
    uint32_t stage = 0, phase = 0;
    for (int item = worker; item < num_items; item += nworkers) {
      const int tile = item_tile(item), split = PAIR ? 0 : item - tile * p.splits;
      const int kb0 = split * p.kb_per_split, kb1 = min(p.total_kb, kb0 + p.kb_per_split);
      const TileCoord tc = tile_coord(p, tile);
      int kb = 0;
      for (int s = 0; s < p.nseg && kb < kb1; ++s) {
        const Segment sg = p.seg[s];
        const CUtensorMap* am = &amaps.m[sg.map];
        if (kb + sg.chunks <= kb0) { kb += sg.chunks; continue; }
        for (int c = 0; c < sg.chunks && kb < kb1; ++c, ++kb) {
          if (kb < kb0) continue;
          ptx::mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t fb = full0 + 8 * stage;
          if (ptx::elect_one()) {
            if constexpr (PAIR) {
              // both CTAs' bytes are credited to the leader's barrier, which alone expects them
              if (rank == 0) ptx::mbar_expect_tx(fb, 2 * (A_BYTES + B_BYTES));
              const uint32_t fl = full_leader0 + 8 * stage;
              ptx::tma_load_4d_pair(am, ptx::smem_u32(sA + stage * A_BYTES), fl, sg.c_begin + c * BK, tc.x0 + sg.dx + (tc.par & 1),
                                    tc.y0 + sg.dy + (tc.par >> 1), tc.n0);
              ptx::tma_load_2d_pair(&tmB, ptx::smem_u32(sB + stage * B_BYTES), fl, kb * BK, tc.par * p.c_out + tc.nt * BN + (int)rank * B_ROWS);
            } else {
              ptx::mbar_expect_tx(fb, A_BYTES + B_BYTES);
              ptx::tma_load_4d(am, ptx::smem_u32(sA + stage * A_BYTES), fb, sg.c_begin + c * BK, tc.x0 + sg.dx + (tc.par & 1),
                               tc.y0 + sg.dy + (tc.par >> 1), tc.n0);
              ptx::tma_load_2d(&tmB, ptx::smem_u32(sB + stage * B_BYTES), fb, kb * BK, tc.par * p.c_out + tc.nt * BN);
            }
          }
          __syncwarp();
          if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues)
    constexpr uint32_t idesc = ptx::idesc_bf16(PAIR ? 2 * BM : BM, BN, 0, 0);
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int item = worker; item < num_items; item += nworkers, ++it) {
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const int split = PAIR ? 0 : item % p.splits;
      const int nkb = min(p.total_kb, (split + 1) * p.kb_per_split) - split * p.kb_per_split;
      ptx::mbar_wait(tempty0 + 8 * as, aphase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(full0 + 8 * stage, phase);
        ptx::tc_fence_after();
        const uint64_t adesc = ptx::smem_desc_sw128(ptx::smem_u32(sA + stage * A_BYTES));
        const uint64_t bdesc = ptx::smem_desc_sw128(ptx::smem_u32(sB + stage * B_BYTES));
        if (ptx::elect_one()) {
          if constexpr (PAIR) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              ptx::mma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            ptx::mma_commit_pair(empty0 + 8 * stage, 3);                    // both CTAs' slots are reusable
            if (kb == nkb - 1) ptx::mma_commit_pair(tfull0 + 8 * as, 3);    // both CTAs' accumulator halves are complete
          } else {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              ptx::mma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            ptx::mma_commit(empty0 + 8 * stage);  // smem slot reusable once these MMAs have read it
            if (kb == nkb - 1) ptx::mma_commit(tfull0 + 8 * as);  // accumulator complete
          }
        }
        __syncwarp();
        if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // -------------------------------------------------------------------- epilogue: 2 warpgroups x 128 threads; thread = one output row,
    // the two warpgroups split the accumulator columns of every slab between them
    const int ew = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = ew * 32 + lane;
    const float* bias = p.bias;
    if (bias != nullptr && p.step_ptr != nullptr) bias += (size_t)(*p.step_ptr) * p.bias_step_stride;
    const int rn = r / (p.bh * p.bw), rr = r - rn * (p.bh * p.bw);
    const int ry = rr / p.bw, rx = rr - ry * p.bw;

    if (p.staged) {
      // ================= staged epilogue: TMEM -> regs -> swizzled smem slab -> TMA store; residual slabs by TMA load ============
      const bool elected = (threadIdx.x == 128);  // first thread of epilogue warpgroup 0
      const bool has_res = p.residual != nullptr;
      const int geglu = p.act == 2;
      constexpr int BNo_full = BN;                       // accumulator columns per tile
      const int acc_per_slab = geglu ? 128 : 64;          // accumulator columns feeding one 64-column output slab
      const int c_eff = geglu ? p.c_out / 2 : p.c_out;    // output channels
      uint32_t rcount = 0;                                // residual slabs consumed so far (ring slot = rcount % RES_BUFS)
      uint32_t scount = 0;                                // slabs processed so far (staging buffer = scount & 1)
      const int sw128 = r & 7, sw64 = (r >> 1) & 3;
      const int BNo = geglu ? BN / 2 : BN;                // output columns per tile
      // residual producer cursor (elected thread): runs up to RES_BUFS slabs ahead of the consumers, across tile boundaries
      int pitem = worker, pslab = 0;
      uint32_t pcount = 0;
      auto issue_residual = [&]() {
        if (pitem >= num_items) return;
        const TileCoord pc = tile_coord(p, item_tile(pitem));
        const int pvalid = min(BN, p.c_out - pc.nt * BN);
        const int pnslabs = (pvalid + acc_per_slab - 1) / acc_per_slab;
        const int w = min(64, BNo - 64 * pslab);
        const uint32_t slot = pcount % RES_BUFS;
        const uint32_t rb = rfull0 + 8 * slot;
        ptx::mbar_expect_tx(rb, BM * w * 2);
        ptx::tma_load_4d(w == 64 ? &emaps.res64 : &emaps.res32, ptx::smem_u32(sRes + slot * SLAB_BYTES), rb, pc.nt * BNo + 64 * pslab, pc.x0,
                         pc.y0, pc.n0);
        ++pcount;
        if (++pslab == pnslabs) { pslab = 0; pitem += nworkers; }
      };
      if (has_res && elected)
        for (int i = 0; i < RES_BUFS; ++i) issue_residual();
      int it = 0;
      for (int item = worker; item < num_items; item += nworkers, ++it) {   // staged epilogue: never split-K, item == tile (or pair)
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        const TileCoord tc = tile_coord(p, item_tile(item));
        const int acc0 = tc.nt * BN;                                  // first accumulator column of this tile (global)
        const int acc_valid = min(BNo_full, p.c_out - acc0);
        const int nslabs = (acc_valid + acc_per_slab - 1) / acc_per_slab;
        const int out0 = geglu ? acc0 / 2 : acc0;                     // first output column of this tile
        const int n = tc.n0 + rn, y = tc.y0 + ry, x = tc.x0 + rx;
        const bool row_ok = (n < p.n_img) && (y < p.H) && (x < p.W);
        const size_t grow = ((size_t)n * p.H + y) * p.W + x;
        const float rscale = (p.row_scale != nullptr && row_ok) ? p.row_scale[grow] : 1.f;
        const float rbias = (p.bias_per_row && bias != nullptr && row_ok) ? bias[grow] : 0.f;
        // folded LayerNorm (consumer side): mean / rstd of this A row from the producer's per-chunk partial sums, summed in chunk
        // order (deterministic); issued before the accumulator wait so the loads overlap the main loop's tail
        float ln_a = 1.f, ln_b = 0.f;
        if (p.ln_stats != nullptr) {
          float s_ = 0.f, q_ = 0.f;
          if (row_ok) {
            const float4* st = reinterpret_cast<const float4*>(p.ln_stats + grow * p.ln_chunks);  // ln_chunks is even
            for (int c = 0; c < (p.ln_chunks >> 1); ++c) {
              const float4 u = __ldg(st + c);
              s_ += u.x; q_ += u.y; s_ += u.z; q_ += u.w;
            }
          }
          const float inv_c = 1.f / (32.f * (float)p.ln_chunks);
          const float mean = s_ * inv_c;
          const float var = fmaxf(q_ * inv_c - mean * mean, 0.f);
          ln_a = rsqrtf(var + p.ln_eps);
          ln_b = -ln_a * mean;
        }

        ptx::mbar_wait(tfull0 + 8 * as, aphase);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + as * ACC_STRIDE + ((uint32_t)(ew * 32) << 16);

        for (int s = 0; s < nslabs; ++s, ++scount) {
          const int wout = min(64, BNo - 64 * s);                    // 64, or 32 for the tail slab of BN=160/32 tiles
          const int ocol = out0 + 64 * s;                            // first output column of the slab (global)
          const uint32_t rbuf = rcount % RES_BUFS;
          if (has_res) ptx::mbar_wait(rfull0 + 8 * rbuf, (rcount / RES_BUFS) & 1);
          uint8_t* ostage = sOut + (scount & 1) * SLAB_BYTES;
          const uint8_t* rstage = sRes + rbuf * SLAB_BYTES;
          // ---- 64 output columns = 2 (or, GEGLU, 4) accumulator chunks of 32
          const int nchunks = geglu ? (wout * 2) / 32 : wout / 32;
          const bool last_slab = (s == nslabs - 1);
          if (last_slab && half >= nchunks) {                         // 32-column tail slab: warpgroup 1 has no chunk, release TMEM
            ptx::tc_fence_before();
            release_acc(as);
          }
#pragma unroll 1
          for (int c = half; c < nchunks; c += 2) {
            const int acol = (geglu ? 128 : 64) * s + 32 * c;       // accumulator column within the tile
            uint32_t v[32];
            ptx::tmem_ld32(t_row + acol, v);
            ptx::tmem_wait_ld();
            if (last_slab && c + 2 >= nchunks) {                     // this thread's last read of the accumulator: hand it back
              ptx::tc_fence_before();
              release_acc(as);
            }
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            const int gcol = acc0 + acol;                            // global accumulator column (bias index)
            if (p.ln_stats != nullptr && gcol + 32 <= p.c_out) {      // rstd * (x W'^T) - rstd * mean * colsum(W')   (c_out % 32 == 0)
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + gcol + j));
                f[j] = fmaf(ln_b, c4.x, ln_a * f[j]); f[j + 1] = fmaf(ln_b, c4.y, ln_a * f[j + 1]);
                f[j + 2] = fmaf(ln_b, c4.z, ln_a * f[j + 2]); f[j + 3] = fmaf(ln_b, c4.w, ln_a * f[j + 3]);
              }
            }
            if (bias != nullptr) {
              if (p.bias_per_row) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] += rbias;
              } else if (gcol + 32 <= p.c_out) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + gcol + j));
                  f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (gcol + j < p.c_out) f[j] += __ldg(bias + gcol + j);
              }
            }
            if (geglu) {
              // 32 accumulator columns -> 16 outputs = 2 pieces of 16 bytes at output columns 16c .. 16c+15 of the slab
              uint32_t o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                o[j] = ptx::pack_bf16(f[4 * j] * gelu_erf(f[4 * j + 1]) * rscale, f[4 * j + 2] * gelu_erf(f[4 * j + 3]) * rscale);
              const int q0 = 2 * c;                                  // 16-byte piece index within the 128-byte row
              *reinterpret_cast<uint4*>(ostage + r * 128 + (((q0) ^ sw128) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(ostage + r * 128 + (((q0 + 1) ^ sw128) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
              continue;
            }
            if (p.act == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
            } else if (p.act == 3) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
            } else if (p.act == 4) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            float rs0 = 0.f, rs1 = 0.f, rq0 = 0.f, rq1 = 0.f;        // row statistics of this 32-column chunk (rowstat_out)
#pragma unroll
            for (int q = 0; q < 4; ++q) {                            // 4 pieces of 8 columns (16 bytes of bf16)
              const int piece = 4 * c + q;                           // piece index within the slab row
              const int off = (wout == 64) ? r * 128 + ((piece ^ sw128) << 4) : r * 64 + ((piece ^ sw64) << 4);
              float* g = f + 8 * q;
              if (has_res) {
                const uint4 u = *reinterpret_cast<const uint4*>(rstage + off);
                g[0] += ptx::bf16_lo(u.x); g[1] += ptx::bf16_hi(u.x); g[2] += ptx::bf16_lo(u.y); g[3] += ptx::bf16_hi(u.y);
                g[4] += ptx::bf16_lo(u.z); g[5] += ptx::bf16_hi(u.z); g[6] += ptx::bf16_lo(u.w); g[7] += ptx::bf16_hi(u.w);
              }
              if (p.rowstat_out != nullptr) {
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                  rs0 += g[k]; rq0 = fmaf(g[k], g[k], rq0);
                  rs1 += g[k + 1]; rq1 = fmaf(g[k + 1], g[k + 1], rq1);
                }
              }
              *reinterpret_cast<uint4*>(ostage + off) =
                  make_uint4(ptx::pack_bf16(g[0] * rscale, g[1] * rscale), ptx::pack_bf16(g[2] * rscale, g[3] * rscale),
                             ptx::pack_bf16(g[4] * rscale, g[5] * rscale), ptx::pack_bf16(g[6] * rscale, g[7] * rscale));
            }
            if (p.rowstat_out != nullptr && row_ok && gcol + 32 <= p.c_out) p.rowstat_out[grow * p.rowstat_chunks + (gcol >> 5)] = make_float2(rs0 + rs1, rq0 + rq1);
          }
          if (has_res) ++rcount;
          ptx::fence_proxy_async_smem();                 // staging writes -> visible to the TMA engine
          if (elected) ptx::tma_store_wait_read0();      // the store that used the OTHER staging tile has drained it
          ptx::named_barrier_sync(1, 256);
          if (elected) {
            if (ocol < c_eff) {
              ptx::tma_store_4d(wout == 64 ? &emaps.out64[tc.par] : &emaps.out32[tc.par], ptx::smem_u32(ostage), ocol, tc.x0, tc.y0, tc.n0);
              ptx::tma_store_commit();
            }
            if (has_res) issue_residual();  // the ring slot every thread just finished reading is free again
          }
        }
      }
      if (elected) ptx::tma_store_wait_read0();
    } else {
      // ================= direct-store epilogue (fp32 outputs, tiny / unaligned N); the two warpgroups alternate 32-column chunks ==
      int it = 0;
      for (int item = worker; item < num_items; item += nworkers, ++it) {
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        const int tile = item_tile(item), split = PAIR ? 0 : item - tile * p.splits;
        const TileCoord tc = tile_coord(p, tile);
        const int nt = tc.nt;
        const int n = tc.n0 + rn, y = tc.y0 + ry, x = tc.x0 + rx;
        const bool row_ok = (n < p.n_img) && (y < p.H) && (x < p.W);
        const size_t grow = ((size_t)n * p.H + y) * p.W + x;
        const float rscale = (p.row_scale != nullptr && row_ok) ? p.row_scale[grow] : 1.f;
        const float rbias = (p.bias_per_row && bias != nullptr && row_ok) ? bias[grow] : 0.f;

        ptx::mbar_wait(tfull0 + 8 * as, aphase);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + as * ACC_STRIDE + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
        for (int c0 = 32 * half; c0 < BN; c0 += 64) {
          const int col0 = nt * BN + c0;
          if (col0 >= p.c_out) break;  // uniform across the warpgroup
          uint32_t v[32];
          ptx::tmem_ld32(t_row + c0, v);
          ptx::tmem_wait_ld();
          if (!row_ok) continue;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          const int ncols = min(32, p.c_out - col0);
          if (bias != nullptr) {
            if (p.bias_per_row) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] += rbias;
            } else if (ncols == 32) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) f[j] += __ldg(bias + col0 + j);
            }
          }
          if (p.act == 2) {
            bf16* orow = reinterpret_cast<bf16*>(p.out) + grow * p.out_pitch + (col0 >> 1);
            for (int j = 0; j < ncols / 2; ++j) orow[j] = __float2bfloat16(f[2 * j] * gelu_erf(f[2 * j + 1]) * rscale);
            continue;
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
          } else if (p.act == 3) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
          } else if (p.act == 4) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (p.residual != nullptr) {
            const bf16* rrow = p.residual + grow * p.residual_pitch + col0;
            for (int j = 0; j < ncols; ++j) f[j] += __bfloat162float(rrow[j]);
          }
          if (p.row_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= rscale;
          }
          if (p.out_fp32) {
            float* orow = reinterpret_cast<float*>(p.out) + (size_t)split * p.split_stride + grow * p.out_pitch + col0;
            if (ncols == 32 && (p.out_pitch & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(orow + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              for (int j = 0; j < ncols; ++j) orow[j] = f[j];
            }
          } else {
            bf16* orow = reinterpret_cast<bf16*>(p.out) + grow * p.out_pitch + col0;
            if (ncols == 32 && (p.out_pitch & 7) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                *reinterpret_cast<uint4*>(orow + j) = make_uint4(ptx::pack_bf16(f[j], f[j + 1]), ptx::pack_bf16(f[j + 2], f[j + 3]),
                                                                   ptx::pack_bf16(f[j + 4], f[j + 5]), ptx::pack_bf16(f[j + 6], f[j + 7]));
            } else {
              for (int j = 0; j < ncols; ++j) orow[j] = __float2bfloat16(f[j]);
            }
          }
        }
        ptx::tc_fence_before();
        release_acc(as);
      }
    }
  }

  __syncwarp();
  ptx::pdl_trigger();
  ptx::tc_fence_before();
  if constexpr (PAIR) ptx::cluster_sync();  // neither CTA frees tensor memory / exits while the other may still signal or read it
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if constexpr (PAIR) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Split-K second pass: out = act(sum_s partial[s] + bias) + residual, bf16.  Partials are summed in split order -> deterministic.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long split_stride, long long rows,
                                                            int c_out, const float* __restrict__ bias, int bias_step_stride,
                                                            const int* __restrict__ step_ptr, int act, const bf16* __restrict__ residual,
                                                            int residual_pitch, bf16* __restrict__ out, int out_pitch) {
  ptx::pdl_wait();
  if (bias != nullptr && step_ptr != nullptr) bias += (size_t)(*step_ptr) * bias_step_stride;
  const int cv = c_out >> 2;
  const long long total = rows * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cv;
    const int col = (int)(i - row * cv) * 4;
    float4 acc = *reinterpret_cast<const float4*>(ws + row * c_out + col);
    for (int s = 1; s < splits; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(ws + s * split_stride + row * c_out + col);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (bias != nullptr) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col));
      acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
    }
    if (act == 1) { acc.x = silu_f(acc.x); acc.y = silu_f(acc.y); acc.z = silu_f(acc.z); acc.w = silu_f(acc.w); }
    if (act == 3) { acc.x = gelu_erf(acc.x); acc.y = gelu_erf(acc.y); acc.z = gelu_erf(acc.z); acc.w = gelu_erf(acc.w); }
    if (act == 4) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    if (residual != nullptr) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(residual + row * residual_pitch + col));
      acc.x += ptx::bf16_lo(u.x); acc.y += ptx::bf16_hi(u.x); acc.z += ptx::bf16_lo(u.y); acc.w += ptx::bf16_hi(u.y);
    }
    *reinterpret_cast<uint2*>(out + row * out_pitch + col) = make_uint2(ptx::pack_bf16(acc.x, acc.y), ptx::pack_bf16(acc.z, acc.w));
  }
}

constexpr size_t SMEM_LIMIT = 232448;  // 227 KiB per CTA
constexpr size_t SMEM_TAIL = (2 * MAX_STAGES + 4 + RES_BUFS) * 8 + 16 + 1024;  // barriers + TMEM slot + alignment slack

template <int BN, bool PAIR>
int launch(const AMaps& am, const CUtensorMap& tmB, const EMaps& em, KParams& kp, cudaStream_t stream) {
  static bool attr_set = false;
  static int max_clusters = 0;
  if (!attr_set) {
    LADI_CUDA(cudaFuncSetAttribute(convgemm_kernel<BN, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    if (PAIR) {  // how many CTA pairs can be co-resident (one CTA per SM, both SMs of a TPC): the persistent grid size
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * ladi_num_sms()); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = SMEM_LIMIT;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, convgemm_kernel<BN, PAIR>, &cfg) != cudaSuccess || max_clusters <= 0) {
        (void)cudaGetLastError();
        max_clusters = ladi_num_sms() / 2;
      }
    }
    attr_set = true;
  }
  const size_t staging = kp.staged ? (size_t)(kp.residual != nullptr ? 2 + RES_BUFS : 2) * SLAB_BYTES : 0;
  const size_t per_stage = A_BYTES + (size_t)(PAIR ? BN / 2 : BN) * BK * 2;
  int stages = (int)((SMEM_LIMIT - SMEM_TAIL - staging) / per_stage);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  LADI_CHECK(stages >= 2, "not enough shared memory for a 2-stage pipeline");
  kp.stages = stages;
  const size_t smem = stages * per_stage + staging + SMEM_TAIL;
  if (PAIR) {
    const int items = ((kp.tiles_m + 1) / 2) * kp.tiles_n;
    const int clusters = items < max_clusters ? items : max_clusters;
    LADI_CUDA(ladi_launch_cluster(convgemm_kernel<BN, PAIR>, dim3(2 * clusters), dim3(384), 2u, smem, stream, am, tmB, em, kp));
    return LADI_OK;
  }
  const int tiles = kp.tiles_m * kp.tiles_n * kp.splits;
  const int grid = tiles < ladi_num_sms() ? tiles : ladi_num_sms();
  LADI_CUDA(ladi_launch(convgemm_kernel<BN, PAIR>, dim3(grid), dim3(384), smem, stream, am, tmB, em, kp));
  return LADI_OK;
}

int largest_pow2_divisor(int v, int cap) {
  int d = 1;
  while (d * 2 <= cap && v % (d * 2) == 0) d *= 2;
  return d;
}

int encode_nhwc_map(CUtensorMap* m, const void* base, int channels, int pitch, int w, int h, int n, const uint32_t* box, int swz) {
  const uint64_t dims[4] = {(uint64_t)channels, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  const uint64_t strides[3] = {(uint64_t)pitch * 2, (uint64_t)pitch * 2 * w, (uint64_t)pitch * 2 * w * h};
  return ladi_encode_tmap_bf16(m, base, 4, dims, strides, box, swz);
}

}  // namespace

extern "C" int ladi_conv2d_bf16(const ladi_conv_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LADI_CHECK(d != nullptr, "conv desc is null");
  LADI_CHECK(d->ksize == 1 || d->ksize == 3, "ksize must be 1 or 3 (got %d)", d->ksize);
  LADI_CHECK(d->stride == 1 || d->stride == 2, "stride must be 1 or 2 (got %d)", d->stride);
  LADI_CHECK(d->n_src >= 1 && d->n_src <= 2 && d->n_sc >= 0 && d->n_sc <= 2, "bad source counts");
  LADI_CHECK(d->n > 0 && d->h_out > 0 && d->w_out > 0 && d->c_out > 0, "bad output extent");
  LADI_CHECK(d->act >= 0 && d->act <= 4, "bad act");
  LADI_CHECK(d->act != 2 || (d->c_out % 2 == 0 && !d->out_fp32 && d->residual == nullptr), "GEGLU needs even c_out, bf16 out");
  LADI_CHECK(d->out != nullptr && d->weight != nullptr, "null out/weight");
  const int up = d->up2x ? 1 : 0;
  if (up) {
    LADI_CHECK(d->ksize == 3 && d->stride == 1 && d->n_sc == 0, "up2x: 3x3 stride-1 convolution without fused shortcut only");
    LADI_CHECK(d->h_out % 2 == 0 && d->w_out % 2 == 0, "up2x: output extent must be even");
    LADI_CHECK(d->residual == nullptr && d->row_scale == nullptr && !d->bias_per_row && d->act != 2 && !d->out_fp32,
               "up2x: bf16 output with per-channel bias / SiLU / GELU / ReLU only");
  }
  LADI_CHECK(d->rowstat_out == nullptr || (d->c_out % 32 == 0 && d->act != 2 && d->row_scale == nullptr && !d->out_fp32 && !up),
             "rowstat_out: c_out %% 32 == 0, bf16 out, no GEGLU / row_scale");
  LADI_CHECK(d->ln_stats == nullptr || (d->ln_colsum != nullptr && d->c_out % 32 == 0 && d->ksize == 1 && d->n_src == 1 && d->n_sc == 0 &&
                                        d->src_c[0] % 64 == 0 && !d->out_fp32 && !d->bias_per_row && !up),
             "ln_stats: GEMM over one source of K %% 64 == 0, c_out %% 32 == 0, bf16 out, ln_colsum given");

  KParams kp;
  memset(&kp, 0, sizeof(kp));
  AMaps am;
  EMaps em;
  // up2x: the tile grid covers ONE output parity plane (h_out/2 x w_out/2 pixels = the input extent); tiles_m counts all four
  const int HO = up ? d->h_out / 2 : d->h_out, WO = up ? d->w_out / 2 : d->w_out;
  kp.n_img = d->n; kp.H = HO; kp.W = WO; kp.c_out = d->c_out;
  kp.bw = largest_pow2_divisor(WO, 128);
  if (d->ksize == 1 && d->h_out == 1 && d->stride == 1) kp.bw = 128;  // plain GEMM: row tail handled by OOB fill
  kp.bh = largest_pow2_divisor(HO, 128 / kp.bw);
  kp.bn = 128 / (kp.bw * kp.bh);
  kp.tiles_x = (WO + kp.bw - 1) / kp.bw;
  kp.tiles_y = (HO + kp.bh - 1) / kp.bh;
  kp.tiles_b = (d->n + kp.bn - 1) / kp.bn;
  kp.tiles_m = kp.tiles_x * kp.tiles_y * kp.tiles_b;
  kp.up2x = up; kp.tiles_pp = kp.tiles_m;
  if (up) kp.tiles_m *= 4;

  // ---- A tensor maps: one per (source, parity plane)
  int nmaps = 0;
  const int s = d->stride;
  const int h_in = up ? HO : (d->h_in > 0 ? d->h_in : d->h_out * s), w_in = up ? WO : (d->w_in > 0 ? d->w_in : d->w_out * s);
  const uint32_t box[4] = {(uint32_t)BK, (uint32_t)kp.bw, (uint32_t)kp.bh, (uint32_t)kp.bn};
  int src_map0[2] = {0, 0}, sc_map0[2] = {0, 0};
  for (int i = 0; i < d->n_src; ++i) {
    LADI_CHECK(d->src[i] != nullptr && d->src_c[i] > 0 && d->src_pitch[i] % 8 == 0 && d->src_pitch[i] >= d->src_c[i],
               "source %d: need non-null, pitch multiple of 8 and >= C", i);
    src_map0[i] = nmaps;
    for (int py = 0; py < s; ++py)
      for (int px = 0; px < s; ++px) {
        LADI_CHECK(nmaps < NUM_A_MAPS, "too many tensor maps");
        const uint64_t pitch = (uint64_t)d->src_pitch[i];
        const uint64_t dims[4] = {(uint64_t)d->src_c[i], (uint64_t)((w_in - px + s - 1) / s), (uint64_t)((h_in - py + s - 1) / s),
                                  (uint64_t)d->n};
        const uint64_t strides[3] = {pitch * 2 * s, pitch * 2 * w_in * s, pitch * 2 * (uint64_t)w_in * h_in};
        const bf16* base = reinterpret_cast<const bf16*>(d->src[i]) + ((size_t)py * w_in + px) * pitch;
        if (ladi_encode_tmap_bf16(&am.m[nmaps], base, 4, dims, strides, box)) return LADI_ERR_CUDA;
        ++nmaps;
      }
  }
  for (int i = 0; i < d->n_sc; ++i) {
    LADI_CHECK(s == 1, "fused 1x1 shortcut requires stride 1");
    LADI_CHECK(d->sc[i] != nullptr && d->sc_c[i] > 0 && d->sc_pitch[i] % 8 == 0, "shortcut source %d invalid", i);
    LADI_CHECK(nmaps < NUM_A_MAPS, "too many tensor maps");
    sc_map0[i] = nmaps;
    if (encode_nhwc_map(&am.m[nmaps], d->sc[i], d->sc_c[i], d->sc_pitch[i], d->w_out, d->h_out, d->n, box, 128)) return LADI_ERR_CUDA;
    ++nmaps;
  }
  for (int i = nmaps; i < NUM_A_MAPS; ++i) am.m[i] = am.m[0];

  // ---- K segments: taps (row-major) x sources, then the shortcut sources.  Weight K order must match (see packer).
  int nseg = 0, total = 0;
  const int taps = up ? 4 : d->ksize * d->ksize;
  for (int t = 0; t < taps; ++t) {
    const int ky = up ? t / 2 : (d->ksize == 3 ? t / 3 : 0), kx = up ? t % 2 : (d->ksize == 3 ? t % 3 : 0);
    // input coordinate = out*stride + k - pad_lo;  parity plane p = (k - pad_lo) mod s, shift = floor((k - pad_lo) / s)
    // up2x: output pixel 2i+p reads input pixels i+p-1 and i+p (merged taps); the kernel adds the parity p, so the base shift is t-1
    const int oy = up ? ky - 1 : ky - (d->ksize == 3 ? d->pad_lo : 0), ox = up ? kx - 1 : kx - (d->ksize == 3 ? d->pad_lo : 0);
    const int py = ((oy % s) + s) % s, px = ((ox % s) + s) % s;
    const int dy = (oy - py) / s, dx = (ox - px) / s;
    for (int i = 0; i < d->n_src; ++i) {
      LADI_CHECK(nseg < MAX_SEG, "too many K segments");
      Segment& sg = kp.seg[nseg++];
      sg.map = (int16_t)(src_map0[i] + py * s + px);
      sg.dx = (int16_t)dx; sg.dy = (int16_t)dy;
      sg.chunks = (int16_t)((d->src_c[i] + BK - 1) / BK);
      sg.c_begin = 0;
      total += sg.chunks;
    }
  }
  for (int i = 0; i < d->n_sc; ++i) {
    LADI_CHECK(nseg < MAX_SEG, "too many K segments");
    Segment& sg = kp.seg[nseg++];
    sg.map = (int16_t)sc_map0[i]; sg.dx = 0; sg.dy = 0;
    sg.chunks = (int16_t)((d->sc_c[i] + BK - 1) / BK);
    sg.c_begin = 0;
    total += sg.chunks;
  }
  kp.nseg = nseg; kp.total_kb = total;
  LADI_CHECK(d->k_total == total * BK, "weight K (%d) != %d K-blocks x 64 implied by the sources", d->k_total, total);
  LADI_CHECK(d->weight_pitch % 8 == 0 && d->weight_pitch >= d->k_total, "weight pitch must be a multiple of 8 and >= K");

  kp.bias = d->bias; kp.bias_per_row = d->bias_per_row; kp.bias_step_stride = d->bias_step_stride; kp.step_ptr = d->step_ptr;
  kp.residual = reinterpret_cast<const bf16*>(d->residual); kp.residual_pitch = d->residual_pitch;
  kp.row_scale = d->row_scale; kp.act = d->act; kp.out = d->out; kp.out_pitch = d->out_pitch; kp.out_fp32 = d->out_fp32;
  kp.rowstat_out = reinterpret_cast<float2*>(d->rowstat_out); kp.rowstat_chunks = d->c_out / 32;
  kp.ln_stats = reinterpret_cast<const float2*>(d->ln_stats); kp.ln_chunks = d->src_c[0] / 32; kp.ln_colsum = d->ln_colsum; kp.ln_eps = d->ln_eps;

  // ---- epilogue flavour: staged TMA stores need bf16 output with 16-byte aligned rows
  const int c_eff = d->act == 2 ? d->c_out / 2 : d->c_out;
  kp.staged = (!d->out_fp32 && d->out_pitch % 8 == 0 && c_eff % 8 == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 &&
               (d->residual == nullptr || (d->residual_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0)) &&
               d->force_direct_epilogue == 0)
                  ? 1 : 0;

  // ---- N tile: minimise (waves x per-tile cost) with per-tile cost ~ BN + fixed A-operand cost
  const int sms = ladi_num_sms();
  int BN = 0;
  if (d->force_bn > 0) {
    BN = d->force_bn;
  } else {
    static const int cand[6] = {256, 192, 160, 128, 64, 32};
    long best = -1;
    for (int i = 0; i < 6; ++i) {
      if (d->act == 2 && kp.staged && cand[i] % 128 != 0) continue;  // GEGLU slabs consume 128 accumulator columns
      const long tiles = (long)kp.tiles_m * ((d->c_out + cand[i] - 1) / cand[i]);
      long cost = ((tiles + sms - 1) / sms) * (cand[i] + 64);
      if (total <= 8 && d->c_out >= 128) {
        // short reductions (K <= 512) are epilogue-bound: the per-tile fixed cost is small, narrower tiles balance better and the
        // 32-column tail slab of a 160-wide tile is its slowest part.  Measured for the 49152x320x320 GEMMs (profiles/r01_bn_sweep.jsonl):
        // 35.8 us (BN=128), 40.0 (160), 39.9 (192), 48.1 (256); 49152x320x960: 54 us (192) vs 59.5 (128) / 62.5 (160).
        if (cand[i] < 128) continue;
        cost = ((tiles + sms - 1) / sms) * (cand[i] + 16);
        if (cand[i] % 64 != 0) cost += cost / 4;
      }
      if (best < 0 || cost < best) { best = cost; BN = cand[i]; }
    }
  }
  if (d->act == 2 && kp.staged && BN % 128 != 0) kp.staged = 0;
  LADI_CHECK(kp.staged || (!up && d->rowstat_out == nullptr && d->ln_stats == nullptr),
             "up2x / rowstat_out / ln_stats need the staged epilogue (bf16 out, 16-byte aligned rows)");

  // ---- split-K: few output tiles but a long reduction (the 8x6 / 16x12 UNet levels): spread K over otherwise idle SMs,
  // fp32 partial planes in the caller's workspace, summed (in split order) by splitk_reduce_kernel
  kp.splits = 1; kp.kb_per_split = total; kp.split_stride = 0;
  const long long rows = (long long)d->n * d->h_out * d->w_out;
  bool split = false;
  if (d->splitk_ws != nullptr && d->force_bn == 0 && !d->out_fp32 && d->act != 2 && d->row_scale == nullptr && !d->bias_per_row && !up &&
      d->rowstat_out == nullptr && d->ln_stats == nullptr &&
      d->c_out % 4 == 0 && d->out_pitch % 4 == 0 && (d->residual == nullptr || d->residual_pitch % 4 == 0) && total >= 32) {
    const int bn_s = d->c_out >= 256 ? 256 : (d->c_out >= 128 ? 128 : 64);
    const long tiles_s = (long)kp.tiles_m * ((d->c_out + bn_s - 1) / bn_s);
    if (tiles_s * 2 <= sms) {
      int S = (int)(sms / tiles_s);
      if (S > total / 8) S = total / 8;
      if (S > 16) S = 16;
      if (S >= 2) {
        const int per = (total + S - 1) / S;
        S = (total + per - 1) / per;
        if (S >= 2 && (long long)S * rows * d->c_out * 4 <= d->splitk_ws_bytes) {
          split = true; BN = bn_s;
          kp.splits = S; kp.kb_per_split = per; kp.split_stride = rows * d->c_out;
          kp.staged = 0; kp.out = d->splitk_ws; kp.out_pitch = d->c_out; kp.out_fp32 = 1;
          kp.bias = nullptr; kp.step_ptr = nullptr; kp.act = 0; kp.residual = nullptr;
        }
      }
    }
  }
  kp.tiles_n = (d->c_out + BN - 1) / BN;

  // ---- CTA pairs (tcgen05 cta_group::2): two M tiles per cluster share one N tile; needs >= 2 M tiles, no split-K, and a B half
  // of whole 8-row swizzle groups.  pair_mode: 0 = library default (on; env LADI_CONV_2CTA=0 turns it off), 1 = force, 2 = never.
  bool pair = false;
  if (!split && kp.tiles_m >= 2 && BN >= 128 && BN % 16 == 0 && (!up || kp.tiles_pp % 2 == 0)) {  // up2x: a pair shares one parity's weights
    if (d->pair_mode == 1) pair = true;
    // default: pair whenever there are enough M tiles that the phantom tile of an odd count is noise (measured 1.03-1.26x on every
    // UNet shape, profiles/r01_pair_bench.jsonl)
    else if (d->pair_mode == 0 && ladi_conv_pair_default()) pair = kp.tiles_m >= 4 && (kp.tiles_m % 2 == 0 || kp.tiles_m >= 16);
  }
  LADI_CHECK(d->pair_mode != 1 || pair, "pair_mode=1 needs >= 2 M tiles, BN >= 128 and no split-K (BN=%d, tiles_m=%d)", BN, kp.tiles_m);

  CUtensorMap tmB;
  {
    const uint64_t dims[2] = {(uint64_t)d->k_total, (uint64_t)d->c_out * (up ? 4 : 1)};  // up2x: the four parities' merged weights, stacked
    const uint64_t strides[1] = {(uint64_t)d->weight_pitch * 2};
    const uint32_t bbox[2] = {(uint32_t)BK, (uint32_t)(pair ? BN / 2 : BN)};
    if (ladi_encode_tmap_bf16(&tmB, d->weight, 2, dims, strides, bbox)) return LADI_ERR_CUDA;
  }
  if (kp.staged) {
    const uint32_t box64[4] = {64u, (uint32_t)kp.bw, (uint32_t)kp.bh, (uint32_t)kp.bn};
    const uint32_t box32[4] = {32u, (uint32_t)kp.bw, (uint32_t)kp.bh, (uint32_t)kp.bn};
    if (up) {  // one strided map per output parity: pixel (2i+py, 2j+px) of the full-resolution tensor
      for (int par = 0; par < 4; ++par) {
        const uint64_t pitch = (uint64_t)d->out_pitch;
        const uint64_t dims[4] = {(uint64_t)c_eff, (uint64_t)WO, (uint64_t)HO, (uint64_t)d->n};
        const uint64_t strides[3] = {pitch * 2 * 2, pitch * 2 * d->w_out * 2, pitch * 2 * (uint64_t)d->w_out * d->h_out};
        const bf16* base = reinterpret_cast<const bf16*>(d->out) + ((size_t)(par >> 1) * d->w_out + (par & 1)) * pitch;
        if (ladi_encode_tmap_bf16(&em.out64[par], base, 4, dims, strides, box64, 128)) return LADI_ERR_CUDA;
        if (ladi_encode_tmap_bf16(&em.out32[par], base, 4, dims, strides, box32, 64)) return LADI_ERR_CUDA;
      }
    } else {
      if (encode_nhwc_map(&em.out64[0], d->out, c_eff, d->out_pitch, d->w_out, d->h_out, d->n, box64, 128)) return LADI_ERR_CUDA;
      if (encode_nhwc_map(&em.out32[0], d->out, c_eff, d->out_pitch, d->w_out, d->h_out, d->n, box32, 64)) return LADI_ERR_CUDA;
      for (int par = 1; par < 4; ++par) { em.out64[par] = em.out64[0]; em.out32[par] = em.out32[0]; }
    }
    if (d->residual != nullptr) {
      if (encode_nhwc_map(&em.res64, d->residual, c_eff, d->residual_pitch, d->w_out, d->h_out, d->n, box64, 128)) return LADI_ERR_CUDA;
      if (encode_nhwc_map(&em.res32, d->residual, c_eff, d->residual_pitch, d->w_out, d->h_out, d->n, box32, 64)) return LADI_ERR_CUDA;
    } else {
      em.res64 = em.out64[0]; em.res32 = em.out32[0];
    }
  } else {
    for (int par = 0; par < 4; ++par) em.out64[par] = em.out32[par] = tmB;
    em.res64 = em.res32 = tmB;
  }
  int rc = LADI_OK;
  switch (BN) {
    case 256: rc = pair ? launch<256, true>(am, tmB, em, kp, stream) : launch<256, false>(am, tmB, em, kp, stream); break;
    case 192: rc = pair ? launch<192, true>(am, tmB, em, kp, stream) : launch<192, false>(am, tmB, em, kp, stream); break;
    case 160: rc = pair ? launch<160, true>(am, tmB, em, kp, stream) : launch<160, false>(am, tmB, em, kp, stream); break;
    case 128: rc = pair ? launch<128, true>(am, tmB, em, kp, stream) : launch<128, false>(am, tmB, em, kp, stream); break;
    case 64: rc = launch<64, false>(am, tmB, em, kp, stream); break;
    case 32: rc = launch<32, false>(am, tmB, em, kp, stream); break;
    default: LADI_CHECK(false, "unsupported BN %d", BN);
  }
  if (rc != LADI_OK || !split) return rc;
  const long long vecs = rows * (d->c_out / 4);
  long long blocks = (vecs + 255) / 256;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  LADI_CUDA(ladi_launch(splitk_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, (const float*)d->splitk_ws, kp.splits,
                        kp.split_stride, rows, d->c_out, d->bias, d->bias_step_stride, d->step_ptr, d->act,
                        reinterpret_cast<const bf16*>(d->residual), d->residual_pitch, reinterpret_cast<bf16*>(d->out), d->out_pitch));
  return LADI_OK;
}
