// Persistent, warp-specialised bf16 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (fp32 accum).
//
// Replaces every F.linear on the hot path of the reference (moondream/torch/layers.py:34-35,
// vision.py:67, text.py:30,53,166, layers.py:130,139, region.py:43,57,71,93).
//
//   warp 0 : TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, mbarrier ring)
//   warp 1 : MMA issuer     (one elected lane issues tcgen05.mma 128 x BN x 16, accumulators in TMEM)
//   warp 2 : TMEM allocator
//   warps 4-7 : epilogue    (tcgen05.ld -> bias / GELU-tanh / residual -> bf16 -> global)
//
// TMEM holds two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
// Tails in M, N and K are handled by TMA out-of-bounds zero fill plus masked stores, so the
// awkward reference shapes (K=588->592, N=4304, M=B*729) need no padded copies of activations.
//
// Two kernels:
//   * gemm_bf16_kernel (row form, prefill / ViT sized M): bf16 out[M,N] with a fused epilogue.
//   * smallbatch_gemm_kernel (decode sized M <= 128 per launch): one CTA per (weight-row tile, K split) streams
//     its slice of the weight matrix once; fp32 partial sums ws[split][b][n] are finished by the consumer
//     (splitk_epilogue_kernel, the decode attention prologue, the residual + LayerNorm epilogue, argmax).
#include <vector>

#include "kernels.cuh"
#include "ptx.cuh"

namespace md {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;                        // two warps per TMEM lane quadrant
constexpr int kGemmThreads = 128 + kEpiWarps * 32;  // producer, MMA, TMEM-alloc, spare + epilogue

static int g_gemm_debug = 0;            // md_debug_gemm: timing experiments only
static int g_gemm_sm_cap = 0;           // md_debug_gemm_sm_cap: 0 = every SM (default)

struct GemmParams {
  int M, N, K;
  int m_blocks, n_blocks, k_blocks;   // m_blocks counts (BM * CG)-row tiles; k_blocks = ceil(K / BK)
  int mode;                            // EPI_BIAS / EPI_BIAS_GELU / EPI_BIAS_RESIDUAL / EPI_QKV_ROPE
  __nv_bfloat16* out;
  long long ldo;
  const __nv_bfloat16* bias;           // [N] or nullptr
  const __nv_bfloat16* res;            // residual [*, N]
  long long ldr;
  int res_mod;                         // residual row = row % res_mod when > 0 (pos_emb broadcast)
  int remap_gin, remap_gout, remap_goff;  // out row = (r / gin) * gout + r % gin + goff when gin > 0
  RopeEpilogue rope;                   // EPI_QKV_ROPE
  int early_trigger;                   // release the dependent grid at once (its prologue overlaps this kernel's tail wave)
};

// Fused epilogue of the prefill QKV projection: this thread owns token row `row` and the 32 output columns
// col0..col0+31 = one half of one head of q, k or v.  dims 0..31 of q and k rotate (pairs (j, j+16) are
// both in this thread), everything else passes through; q goes to q_out, k / v to the KV page of (seq, pos).
struct RopeRow { int pos; long long kv_row; };   // kv_row: element offset of (page, token) for head 0, k plane
__device__ __forceinline__ RopeRow rope_locate(const RopeEpilogue& e, int row) {
  int lo = 0, hi = e.n_seqs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (e.q_offsets[mid] <= row) lo = mid; else hi = mid;
  }
  RopeRow r;
  r.pos = e.start_pos[lo] + (row - e.q_offsets[lo]);
  const int page = e.block_tables[static_cast<long long>(lo) * e.max_blocks + (r.pos >> 6)];
  r.kv_row = ((static_cast<long long>(e.layer) * e.n_pages + page) * 2) * e.n_kv_heads * (64 * 64) + (r.pos & 63) * 64;
  return r;
}
__device__ __forceinline__ void rope_store_chunk(const GemmParams& p, const uint32_t (&acc)[32], int row, int col0,
                                                 const RopeRow& rr) {
  const RopeEpilogue& e = p.rope;
  // columns: q [0, D) | k [D, D + KVW) | v [D + KVW, D + 2 KVW), KVW = n_kv_heads * 64 (text.py:36-38)
  const int kvw = e.n_kv_heads * 64;
  const int part = col0 < e.D ? 0 : (col0 < e.D + kvw ? 1 : 2);
  const int within = col0 - (part == 0 ? 0 : e.D + (part - 1) * kvw);
  const int head = within >> 6, half = (within >> 5) & 1;
  float v[32];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint4 bq = *reinterpret_cast<const uint4*>(p.bias + col0 + g * 8);
    const uint32_t bw[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[g * 8 + 2 * j] = bf16_round(__uint_as_float(acc[g * 8 + 2 * j]) + bf16_lo(bw[j]));
      v[g * 8 + 2 * j + 1] = bf16_round(__uint_as_float(acc[g * 8 + 2 * j + 1]) + bf16_hi(bw[j]));
    }
  }
  uint32_t o[16];
  if (part < 2 && half == 0) {
    const float4* tab = reinterpret_cast<const float4*>(e.freqs + static_cast<long long>(rr.pos) * 32);   // 16 x (cos, sin)
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float4 cs = tab[j >> 1];                      // (cos_j, sin_j, cos_j+1, sin_j+1)
      const float r0 = __fsub_rn(__fmul_rn(v[j], cs.x), __fmul_rn(v[16 + j], cs.y));
      const float i0 = __fadd_rn(__fmul_rn(v[j], cs.y), __fmul_rn(v[16 + j], cs.x));
      const float r1 = __fsub_rn(__fmul_rn(v[j + 1], cs.z), __fmul_rn(v[17 + j], cs.w));
      const float i1 = __fadd_rn(__fmul_rn(v[j + 1], cs.w), __fmul_rn(v[17 + j], cs.z));
      o[j] = pack_bf16x2(r0, i0);                         // interleaved output (re', im')
      o[j + 1] = pack_bf16x2(r1, i1);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
  }
  __nv_bfloat16* dst;
  if (part == 0) dst = e.q_out + static_cast<long long>(row) * e.D + head * 64 + half * 32;
  else dst = e.kv_pool + rr.kv_row + (static_cast<long long>(part - 1) * e.n_kv_heads + head) * (64 * 64) + half * 32;
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(dst + g * 8) = make_uint4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
}

template <int BN, int STAGES, int CG>
struct GemmSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / CG) * BK * 2;   // a CTA pair splits the B tile between its CTAs
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 4) * 8 + 16 + 1024;  // +1024 align slack
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, cta_group::2) per 256 x BN
// tile; each CTA stages its own 128 rows of A and half of the B tile, the leader issues the MMAs,
// each CTA's TMEM holds the accumulator rows it will write.
template <int BN, int STAGES, int CG>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  using S = GemmSmem<BN, STAGES, CG>;
  constexpr uint32_t kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;  // power of two for BN in {32..256}
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN must be a power of two");
  static_assert(CG == 1 || CG == 2, "cta group");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = rank == 0;

  __shared__ unsigned long long tl_s[5];    // debug timeline stamps (see ptx.cuh), untouched unless installed
  const bool tl = tl_on();
  if (tl && threadIdx.x == 0) tl_s[0] = tl_now();
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], CG);            // pair: the leader's barrier collects both producers
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps * 32 * CG);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) { tmem_alloc_pair(tmem_ptr_smem, kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_ptr_smem, kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // every dependent kernel waits (griddepcontrol.wait) before it touches global memory, so releasing it here only moves
  // its launch latency and prologue under this kernel's last, partially filled wave of tiles
  if (p.early_trigger) pdl_launch_dependents();

  const int total_tiles = p.m_blocks * p.n_blocks;
  const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      pdl_wait();
      if (tl) tl_s[1] = tl_now();
      for (int tile = unit; tile < total_tiles; tile += n_units) {
        const int n_blk = tile % p.n_blocks;
        const int m_blk = tile / p.n_blocks;
        const int a_row = m_blk * (BM * CG) + rank * BM;
        const int b_row = n_blk * BN + rank * (BN / CG);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (CG == 2) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
            else mbar_arrive_leader(&full_bar[stage]);
            tma_load_2d_pair(sa, &tmA, &full_bar[stage], kb * BK, a_row);
            tma_load_2d_pair(sb, &tmB, &full_bar[stage], kb * BK, b_row);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, a_row);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, b_row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (lane == 0 && leader) {
      pdl_wait();
      constexpr uint32_t idesc = make_idesc_bf16_f32(BM * CG, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = unit; tile < total_tiles; tile += n_units, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (tl && it == 0 && kb == 0) tl_s[2] = tl_now();            // first operands landed
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
          const uint64_t da = make_desc_k_sw128(sa);
          const uint64_t db = make_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128-byte swizzle row: +2 in 16-B units
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            if (CG == 2) umma_bf16_pair(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, acc);
            else umma_bf16(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, acc);
          }
          // smem slot reusable (in both CTAs of a pair) once these MMAs have read it
          if (CG == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (CG == 2) umma_commit_pair(&tmem_full[as]); else umma_commit(&tmem_full[as]);   // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    pdl_wait();                             // global reads (residual) and writes must follow the predecessor
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int half = (warp - 4) >> 2;       // the two warps of a quadrant interleave 32-column chunks
    int it = 0;
    for (int tile = unit; tile < total_tiles; tile += n_units, ++it) {
      const int n_blk = tile % p.n_blocks;
      const int m_blk = tile / p.n_blocks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row = m_blk * (BM * CG) + rank * BM + q * 32 + lane;   // accumulator row
      const bool row_ok = row < p.M;
      long long out_row = row;
      if (p.remap_gin > 0)
        out_row = static_cast<long long>(row / p.remap_gin) * p.remap_gout + row % p.remap_gin +
                  p.remap_goff;
      const int res_row = (p.res_mod > 0) ? (row % p.res_mod) : row;
      const bool use_res = p.mode == EPI_BIAS_RESIDUAL && row_ok;
      const __nv_bfloat16* rbase = use_res ? p.res + static_cast<long long>(res_row) * p.ldr : nullptr;

      RopeRow rr{0, 0};
      if (p.mode == EPI_QKV_ROPE && row_ok) rr = rope_locate(p.rope, row);
      // the residual does not depend on the accumulator: fetch the first chunk before waiting for the MMAs
      uint4 rq[4], rq_next[4];
      auto load_res = [&](int cc, uint4 (&dst)[4]) {
        const int c0 = n_blk * BN + cc * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          dst[g] = (use_res && cc < BN / 32 && c0 + g * 8 < p.N)
                       ? *reinterpret_cast<const uint4*>(rbase + c0 + g * 8) : make_uint4(0, 0, 0, 0);
      };
      load_res(half, rq);

      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      if (tl && threadIdx.x == 128) tl_s[3] = tl_now();                  // (last) accumulator complete
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(as * BN);

#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        uint32_t acc[32];
        tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), acc);
        load_res(c + 2, rq_next);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (p.mode == EPI_QKV_ROPE) {
          if (row_ok && col0 < p.N) rope_store_chunk(p, acc, row, col0, rr);
          continue;
        }
        if (row_ok && col0 < p.N) {
          __nv_bfloat16* optr = p.out + out_row * p.ldo + col0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {          // 4 groups of 8 columns = 16-byte stores
            if (col0 + g * 8 >= p.N) break;       // N is a multiple of 8 (checked on the host)
            float v[8];
            uint4 bq = make_uint4(0, 0, 0, 0);
            if (p.bias) bq = *reinterpret_cast<const uint4*>(p.bias + col0 + g * 8);
            const uint32_t bw[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // the reference rounds the Linear output to bf16 before anything else touches it
              v[2 * j] = bf16_round(__uint_as_float(acc[g * 8 + 2 * j]) + bf16_lo(bw[j]));
              v[2 * j + 1] = bf16_round(__uint_as_float(acc[g * 8 + 2 * j + 1]) + bf16_hi(bw[j]));
            }
            if (p.mode == EPI_BIAS_GELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = gelu_tanh(v[j]);
            } else if (p.mode == EPI_BIAS_RESIDUAL) {
              const uint32_t rw[4] = {rq[g].x, rq[g].y, rq[g].z, rq[g].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                v[2 * j] += bf16_lo(rw[j]);
                v[2 * j + 1] += bf16_hi(rw[j]);
              }
            }
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]);
            o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]);
            o.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(optr + g * 8) = o;
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) rq[g] = rq_next[g];
      }
      tc_fence_before();
      if (CG == 2) mbar_arrive_leader(&tmem_empty[as]); else mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (tl && threadIdx.x == 0 && blockIdx.x < total_tiles * CG)
    tl_emit((1u << 28) | (static_cast<uint32_t>(p.mode) << 24) | (static_cast<uint32_t>(p.M) & 0xFFFFFFu),
            tl_s[0], tl_s[1], tl_s[2], tl_s[3], tl_now());
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------------
// Small-batch weight stream (decode-time Linear layers, batch <= 128 rows of activations).
//
// ws[split][b][n] = sum_{k in split} X[b][k] * W[n][k]  (fp32 partials, finished by the consumer kernel).
//
// The activations are the MMA's A operand (M = 128 lanes, only `batch` of them meaningful: the TMA box holds the
// real rows, the descriptor's remaining rows read whatever follows in shared memory and land in lanes that are
// never stored) and the weight tile is the B operand (N = tile_n rounded up to 16, up to 256 rows).
//
// Measured per k-block with tools/decode_timeline.py: a 64-wide k-block costs ~600 clocks in either orientation
// when the tile is narrow, because every K = 16 MMA re-reads its 128-row A operand from shared memory at
// ~32 B/clk.  With the weights as A (the first version of this stream) that made time proportional to k-blocks
// and independent of tile height (72-row tiles streamed at 4.3 TB/s chip-wide); as the B operand the weight
// bytes per k-block can grow to 256 rows for the same A-operand cost, which is what the split plans exploit.
// Halving the A-operand read with M = 64 MMAs paid (layer period 78 -> 71 us); feeding the activation tile through TMEM
// instead (tcgen05.st into a ring of A slots, tcgen05.mma with a TMEM A operand, M = 128) was built, is bit-identical,
// and is slower (stream 12 -> 17-18 us: it behaves like the M = 128 shared-memory form), so it was removed again
// (profiles/r02_decode_timeline_tsmode.json).
//
// One CTA per (weight-row tile, K split), all resident at once (the plan keeps tiles * splits <= #SMs where
// it can).  Under programmatic dependent launch the weight tiles of the first ring of stages are requested
// before the dependency wait (no earlier kernel writes weights); only the activation loads, and everything
// downstream of them, wait for the predecessor.
// ------------------------------------------------------------------------------------------------
constexpr int kSbMaxStages = 12;
constexpr int kSbThreads = 256;          // warp 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4..7 epilogue
constexpr int kSbTailPad = BM * BK * 2;  // the A descriptor spans 128 rows whatever the batch: keep it in bounds

struct SmallBatchParams {
  int batch, batch_total, n_out, k_blocks;   // this launch covers `batch` (<= 128) rows of a batch_total-row problem
  int tile_n, n_mma, k_splits;
  int kb_per_split, seg_splits, seg_kb, kb_per_split2;   // split -> k-block range, as in GemmParams
  int stages, a_bytes, stage_bytes;    // a_bytes: activation tile (box rows * 128 B), then the weight tile
  int tmem_cols;
  int trigger_early;
  float* ws;
};

// MROWS = 64 (the default for batches <= 64; md_debug_gemm bit 6 forces 128 for A/B runs) issues M = 64 MMAs: half the
// activation lanes per K = 16 step, accumulator rows 16q .. 16q+15 in lanes 0..15 of TMEM lane quadrant q (the 1-CTA
// M = 64 data-path layout).  MROWS = 128 serves batches of 65..128 rows.
template <int MROWS>
__global__ void __launch_bounds__(kSbThreads, 1)
smallbatch_gemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                       const SmallBatchParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes + kSbTailPad);
  uint64_t* empty_bar = full_bar + kSbMaxStages;
  uint64_t* tmem_full = empty_bar + kSbMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full + 1);
  __shared__ unsigned long long tl_s[5];    // debug timeline stamps (see ptx.cuh), untouched unless installed
  const bool tl = tl_on();
  if (tl && threadIdx.x == 0) tl_s[0] = tl_now();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (p.trigger_early) pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int split = blockIdx.x % p.k_splits;
  const int tile = blockIdx.x / p.k_splits;
  int kb0, kb1;
  if (split < p.seg_splits) {
    kb0 = split * p.kb_per_split;
    kb1 = min(p.seg_kb, kb0 + p.kb_per_split);
  } else {
    kb0 = p.seg_kb + (split - p.seg_splits) * p.kb_per_split2;
    kb1 = min(p.k_blocks, kb0 + p.kb_per_split2);
  }
  const int nk = kb1 - kb0;
  const int n0 = tile * p.tile_n;
  const uint32_t tx_bytes = static_cast<uint32_t>(p.a_bytes + p.tile_n * (BK * 2));

  if (warp == 0) {
    if (lane == 0) {
      const int pre = min(p.stages, nk);
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&full_bar[i], tx_bytes);
        tma_load_2d(smem + i * p.stage_bytes + p.a_bytes, &tmW, &full_bar[i], (kb0 + i) * BK, n0);
      }
      pdl_wait();
      if (tl) tl_s[1] = tl_now();
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nk; ++j) {
        uint8_t* sx = smem + stage * p.stage_bytes;
        if (j >= pre) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_2d(sx + p.a_bytes, &tmW, &full_bar[stage], (kb0 + j) * BK, n0);
        }
        tma_load_2d(sx, &tmX, &full_bar[stage], (kb0 + j) * BK, 0);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16_f32(MROWS, p.n_mma);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nk; ++j) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (tl && j == 0) tl_s[2] = tl_now();                 // first operands landed
        const uint32_t sx = smem_u32(smem + stage * p.stage_bytes);
        const uint64_t da = make_desc_k_sw128(sx);
        const uint64_t db = make_desc_k_sw128(sx + static_cast<uint32_t>(p.a_bytes));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                    (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full);
    }
  } else if (warp >= 4) {
    // lanes of TMEM = rows of the activation tile: warp q owns batch rows 32q .. 32q+31 (M = 128) or
    // 16q .. 16q+15 in its first 16 lanes (M = 64)
    constexpr int kRowsPerWarp = MROWS / 4;
    const int q = warp - 4;
    pdl_wait();                              // ws is still being read by the predecessor's consumers
    if (q * kRowsPerWarp < p.batch) {
      const int b = (lane < kRowsPerWarp) ? q * kRowsPerWarp + lane : p.batch;   // surplus lanes store nothing
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      if (tl && threadIdx.x == 128) tl_s[3] = tl_now();       // accumulator complete
      float* dst = p.ws + (static_cast<long long>(split) * p.batch_total + b) * p.n_out + n0;
      const bool vec = ((n0 | p.n_out) & 3) == 0;
      const int n_valid = min(p.tile_n, p.n_out - n0);
      for (int c = 0; c * 32 < n_valid; ++c) {
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), acc);
        tmem_ld_wait();
        if (b < p.batch) {
          if (vec && c * 32 + 32 <= n_valid) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<float4*>(dst + c * 32 + g * 4) =
                  make_float4(__uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]),
                              __uint_as_float(acc[4 * g + 2]), __uint_as_float(acc[4 * g + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < n_valid) dst[c * 32 + j] = __uint_as_float(acc[j]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tl && threadIdx.x == 0)
    tl_emit((1u << 28) | (static_cast<uint32_t>(EPI_PARTIAL) << 24) | (static_cast<uint32_t>(p.n_out) & 0xFFFFFFu),
            tl_s[0], tl_s[1], tl_s[2], tl_s[3], tl_now());
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// Finishes a small-batch GEMM: out[b][n] = epi( sum_s ws[s][b][n] + bias[n] ) in a fixed
// summation order (deterministic, unlike atomics).
__global__ void splitk_epilogue_kernel(const float* __restrict__ ws, int splits, int B, int N,
                                       int mode, const __nv_bfloat16* __restrict__ bias,
                                       const __nv_bfloat16* __restrict__ res, long long ldr,
                                       __nv_bfloat16* __restrict__ out, long long ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int b = blockIdx.y;
  if (n >= N) return;
  float a0 = 0.f, a1 = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float2 v = *reinterpret_cast<const float2*>(ws + (static_cast<long long>(s) * B + b) * N + n);
    a0 += v.x;
    a1 += v.y;
  }
  if (bias) {
    const uint32_t bw = *reinterpret_cast<const uint32_t*>(bias + n);
    a0 += bf16_lo(bw);
    a1 += bf16_hi(bw);
  }
  a0 = bf16_round(a0);
  a1 = bf16_round(a1);
  if (mode == EPI_BIAS_GELU) {
    a0 = gelu_tanh(a0);
    a1 = gelu_tanh(a1);
  } else if (mode == EPI_BIAS_RESIDUAL) {
    const uint32_t rw = *reinterpret_cast<const uint32_t*>(res + static_cast<long long>(b) * ldr + n);
    a0 += bf16_lo(rw);
    a1 += bf16_hi(rw);
  }
  *reinterpret_cast<uint32_t*>(out + static_cast<long long>(b) * ldo + n) = pack_bf16x2(a0, a1);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// bf16 row-major [rows, cols] with row stride ld (elements); box = [box_rows, 64] 128B-swizzled.
int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, long long rows, long long cols,
                      long long ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16)
    return set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed");
  return 0;
}

// raw bytes, row-major [rows][row_bytes] with row pitch pitch_bytes; box = [box_rows, box_bytes], no swizzle
// (staging tiles of packed int4 / int8 weights, gemm_quant.cu)
int make_tmap_u8_2d(CUtensorMap* tm, const void* base, long long rows, long long row_bytes, long long pitch_bytes,
                    int box_rows, int box_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || pitch_bytes % 16 || box_bytes % 16 || box_rows < 1 || box_rows > 256)
    return set_error("TMA byte operand must be 16-byte aligned with 16-byte multiple pitch and box");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(row_bytes), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch_bytes)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_bytes), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled (bytes) failed");
  return 0;
}

// ---- optional per-launch timing of the row-form GEMM (bench.py's roofline leg) ----
struct GemmProfile {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  double flops = 0.0;
  long long launches = 0;
};
static GemmProfile g_prof;

void gemm_profile_enable(int on) {
  g_prof.on = on != 0;
  if (on) { g_prof.used = 0; g_prof.flops = 0.0; g_prof.launches = 0; }
}
// Sums the recorded start/stop pairs (synchronises on the last event). Returns 0 on success.
int gemm_profile_read(double* total_ms, double* total_flops, long long* launches) {
  double ms = 0.0;
  for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
    if (cudaEventSynchronize(g_prof.pool[i + 1]) != cudaSuccess) return set_error("profile: event sync failed");
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof.pool[i], g_prof.pool[i + 1]) != cudaSuccess)
      return set_error("profile: elapsed time failed");
    ms += t;
  }
  *total_ms = ms; *total_flops = g_prof.flops; *launches = g_prof.launches;
  return 0;
}
static cudaEvent_t profile_event() {
  if (g_prof.used == g_prof.pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_prof.pool.push_back(e);
  }
  return g_prof.pool[g_prof.used++];
}

// bf16 3-D tensor [d2][d1][d0] (d0 contiguous) with byte strides s1, s2; box {b0, b1, b2}.
int make_tmap_bf16_3d(CUtensorMap* tm, const void* base, long long d0, long long d1, long long d2,
                      long long s1_bytes, long long s2_bytes, int b0, int b1, int b2, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (s1_bytes % 16) || (s2_bytes % 16))
    return set_error("TMA operand must be 16-byte aligned with 16-byte multiple strides");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(s1_bytes), static_cast<cuuint64_t>(s2_bytes)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(b0), static_cast<cuuint32_t>(b1), static_cast<cuuint32_t>(b2)};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                              : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled (3-D) failed");
  return 0;
}

static int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int STAGES, int CG>
static int launch_gemm(const CUtensorMap& tA, const CUtensorMap& tB, const GemmParams& p,
                       cudaStream_t stream) {
  using S = GemmSmem<BN, STAGES, CG>;
  static_assert(S::kTotal <= 227 * 1024, "shared memory budget");
  static DeviceOnce configured;
  if (configured.first()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, STAGES, CG>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  }
  const int total = p.m_blocks * p.n_blocks;
  int max_units = num_sms() / CG;
  if (g_gemm_sm_cap > 0 && g_gemm_sm_cap / CG < max_units)      // experiments: leave SMs to a concurrent stream
    max_units = g_gemm_sm_cap / CG > 0 ? g_gemm_sm_cap / CG : 1;
  const int units = total < max_units ? total : max_units;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(units * CG);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, STAGES, CG>, tA, tB, p);
  count_launch();
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  return 0;
}

int g_patch_embed_unfused = 0;   // md_debug_gemm bit 7: patchify kernel + row-form GEMM instead of the fused kernel (A/B, tests)
void gemm_debug_flags(int flags) { g_gemm_debug = flags; g_patch_embed_unfused = (flags & 128) ? 1 : 0; }
void gemm_debug_sm_cap(int sms) { g_gemm_sm_cap = sms < 0 ? 0 : sms; }
static int g_force_cg = 0;   // 0 = auto, 1 / 2 = force (tests and A/B timing)
void gemm_force_cta_group(int cg) { g_force_cg = cg; }
// Installs (buf != nullptr) or removes the debug timeline buffer in every kernel translation unit.
int timeline_install_all(unsigned long long* buf, unsigned int* count, unsigned int cap) {
  Timeline t{buf, count, buf ? cap : 0u};
  if (timeline_install(t) != cudaSuccess) return set_error("timeline: cudaMemcpyToSymbol failed");
  timeline_install_attention(t);
  timeline_install_elementwise(t);
  timeline_install_attention_tc(t);
  return 0;
}

// bn: B-tile rows; cg: CTAs per tile
static int dispatch_gemm(int bn, int cg, const CUtensorMap& tA, const CUtensorMap& tB, const GemmParams& p,
                         cudaStream_t stream) {
  if (cg == 2) {
    if (bn == 256) return launch_gemm<256, 6, 2>(tA, tB, p, stream);
    return set_error("pair GEMM needs BN = 256");
  }
  switch (bn) {
    case 256: return launch_gemm<256, 4, 1>(tA, tB, p, stream);
    case 128: return launch_gemm<128, 6, 1>(tA, tB, p, stream);
    case 64: return launch_gemm<64, 8, 1>(tA, tB, p, stream);
    case 32: return launch_gemm<32, 10, 1>(tA, tB, p, stream);
  }
  return set_error("unsupported BN");
}

static int pick_bn_rows(int N) {
  // 256-wide tiles give the tensor pipe the longest uninterrupted run; fall back for narrow N.
  if (N >= 192) return 256;
  if (N >= 96) return 128;
  if (N >= 48) return 64;
  return 32;
}

int gemm_rowform(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M,
                 int N, int K, int mode, const __nv_bfloat16* bias, const __nv_bfloat16* res,
                 long long ldr, int res_mod, __nv_bfloat16* out, long long ldo, int remap_gin,
                 int remap_gout, int remap_goff, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error("gemm: empty problem");
  if (N % 8 || K % 8) return set_error("gemm: N and K must be multiples of 8");
  if (mode == EPI_BIAS_RESIDUAL && !res) return set_error("gemm: residual mode without residual");
  if ((ldo % 8) || (res && (ldr % 8))) return set_error("gemm: ldo/ldr must be multiples of 8");
  const int bn = pick_bn_rows(N);
  int cg = (bn == 256 && M > BM) ? 2 : 1;          // CTA pairs for the large prefill / ViT GEMMs
  if (g_force_cg == 1) cg = 1;
  if (g_force_cg == 2 && bn == 256) cg = 2;
  CUtensorMap tA, tB;
  if (make_tmap_bf16_2d(&tA, A, M, K, lda, BM)) return 1;
  if (make_tmap_bf16_2d(&tB, W, N, K, ldw, bn / cg)) return 1;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.m_blocks = (M + BM * cg - 1) / (BM * cg);
  p.n_blocks = (N + bn - 1) / bn;
  p.k_blocks = (K + BK - 1) / BK;
  p.mode = mode;
  p.out = out; p.ldo = ldo; p.bias = bias; p.res = res; p.ldr = ldr; p.res_mod = res_mod;
  p.remap_gin = remap_gin; p.remap_gout = remap_gout; p.remap_goff = remap_goff;
  p.early_trigger = (g_pdl && !(g_gemm_debug & 32)) ? 1 : 0;      // md_debug_gemm bit 5: off (A/B)
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  const bool prof = g_prof.on && cudaStreamIsCapturing(stream, &cap) == cudaSuccess &&
                    cap == cudaStreamCaptureStatusNone;
  if (prof) cudaEventRecord(profile_event(), stream);
  const int rc = dispatch_gemm(bn, cg, tA, tB, p, stream);
  if (prof) {
    cudaEventRecord(profile_event(), stream);
    g_prof.flops += 2.0 * M * static_cast<double>(N) * K;
    g_prof.launches += 1;
  }
  return rc;
}


// Plan of a small-batch weight stream: tile height (rows of W per CTA, <= 128) and k-blocks per split such
// that (tiles x splits) fills whole waves of the SMs.  cost = waves x tile_rows x kb is the critical path in
// units of 128-byte weight rows; ties prefer fewer splits (less partial-sum traffic).
// kb_divisor > 0 restricts kb to divisors of it.
// Operand-read model of the stream (the default plan; md_debug_gemm bit 3 selects the previous one for A/B runs;
// DESIGN.md section 5.1): an SS-mode K = 16 MMA fetches its A
// (activation lanes) and B (weight rows) operands from shared memory at ~32 B/clk, one after the other, so a 64-wide
// k-block of a tile costs 4 * (m_rows + rows16) clocks and weights enter the tensor core at 32 * rows / (m_rows + rows)
// B/clk per SM.  That explains the measured streams (97-row tiles with M = 128 lanes: 14 of 32 B/clk = 4 TB/s
// chip-wide; 194 x 2 splits: 10.8 us mean) and says the stream turns HBM-bound (23 B/clk per SM) once
// rows >= 2.6 * m_rows: tiles up to 256 rows, paid for with K splits.  cost in clocks.
// Measured (tools/decode_timeline.py, 2B b32, profiles/r02_decode_timeline_wide.json): the [qkv ; fc1] stream
// 18.4 -> 12.3 us with 208-row tiles x 2 splits and M = 64 MMAs.
static StreamPlan plan_smallbatch_wide(int n_out, int K, int m_rows) {
  const int k_blocks = (K + BK - 1) / BK;
  const int sms = num_sms();
  const double hbm_b_per_clk = 3400.0, sm_b_per_clk = 34.0;       // ~6.5 TB/s at ~1.9 GHz; per-SM streaming ceiling
  StreamPlan best{BM, k_blocks, 1};
  double best_cost = -1.0;
  for (int splits = 1; splits <= 8 && splits <= k_blocks; ++splits) {
    const int kb = (k_blocks + splits - 1) / splits;
    if (kb * (splits - 1) >= k_blocks) continue;                  // a split would be empty
    if (kb < 4 && k_blocks >= 4) break;
    for (int rows = 256; rows >= 64; --rows) {
      const int tiles = (n_out + rows - 1) / rows;
      const long long ctas = 1LL * tiles * splits;
      const long long waves = (ctas + sms - 1) / sms;
      const int rows16 = (rows + 15) / 16 * 16;
      const double t_op = 4.0 * kb * (m_rows + rows16);
      const double t_sm = 128.0 * rows * kb / sm_b_per_clk;
      const double t_hbm = 128.0 * n_out * k_blocks / hbm_b_per_clk;
      double cost = waves * (t_op > t_sm ? t_op : t_sm);
      if (cost < t_hbm) cost = t_hbm;
      cost += 8.0 * splits * n_out * 32 / hbm_b_per_clk + 150.0 * splits;   // fp32 partials written + re-read (L2)
      if (best_cost < 0 || cost < best_cost - 1e-9) { best_cost = cost; best = StreamPlan{rows, kb, splits}; }
    }
  }
  return best;
}

StreamPlan plan_smallbatch(int n_out, int K, int kb_divisor, int m_rows) {
  if (!(g_gemm_debug & 8) && kb_divisor == 0) return plan_smallbatch_wide(n_out, K, m_rows);
  const int k_blocks = (K + BK - 1) / BK;
  const int sms = num_sms();
  StreamPlan best{BM, k_blocks, 1};
  long long best_cost = -1;
  for (int kb = k_blocks; kb >= 1; --kb) {
    if (kb_divisor > 0 ? (kb > kb_divisor || kb_divisor % kb) : false) continue;
    const int splits = (k_blocks + kb - 1) / kb;
    if (splits > 32) break;
    if (kb_divisor == 0 && kb < 4 && k_blocks >= 4) break;           // keep >= 256 of K per split
    for (int rows = BM; rows >= 64; --rows) {
      const int tiles = (n_out + rows - 1) / rows;
      const long long waves = (1LL * tiles * splits + sms - 1) / sms;
      // critical path in 128-byte weight rows + this SM's share of the fp32 partial-sum traffic
      // (written and re-read once per split, nominal batch 32)
      const long long cost = waves * rows * kb + (1LL * splits * n_out * 32 * 8) / (128LL * sms);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = StreamPlan{rows, kb, splits}; }
    }
  }
  return best;
}

// ws[splits][batch][n_out] (fp32) = W[n_out,K] * X[batch,K]^T  partial sums; split s owns k-blocks
// [s * kb_per_split, (s+1) * kb_per_split).
// With seg_kb > 0 the k-blocks split in two segments (see SmallBatchParams): [0, seg_kb) in pieces of
// kb_per_split, the rest in pieces of kb_per_split2.  Batches above 128 rows run as consecutive 128-row launches.
// Returns the number of splits (> 0), -1 on error.
static int gemm_smallbatch_impl(const __nv_bfloat16* W, long long ldw, const __nv_bfloat16* X, long long ldx,
                             int n_out, int batch, int K, int kb_per_split, int tile_n, float* ws,
                             cudaStream_t stream, int seg_kb = 0, int kb_per_split2 = 0) {
  if (n_out <= 0 || batch <= 0 || K <= 0) { set_error("small-batch GEMM: empty problem"); return -1; }
  if (K % 8) { set_error("small-batch GEMM: K must be a multiple of 8"); return -1; }
  if (tile_n < 1 || tile_n > 256) tile_n = BM;
  SmallBatchParams p{};
  p.batch_total = batch; p.n_out = n_out;
  p.k_blocks = (K + BK - 1) / BK;
  p.tile_n = tile_n;
  p.n_mma = (tile_n + 15) / 16 * 16;
  const int n_tiles = (n_out + tile_n - 1) / tile_n;
  if (kb_per_split < 1) kb_per_split = 1;
  if (kb_per_split > p.k_blocks) kb_per_split = p.k_blocks;
  p.kb_per_split = kb_per_split;
  if (seg_kb > 0 && seg_kb < p.k_blocks) {
    if (kb_per_split2 < 1) kb_per_split2 = 1;
    p.seg_kb = seg_kb;
    p.seg_splits = (seg_kb + kb_per_split - 1) / kb_per_split;
    p.kb_per_split2 = kb_per_split2;
    p.k_splits = p.seg_splits + (p.k_blocks - seg_kb + kb_per_split2 - 1) / kb_per_split2;
  } else {
    p.k_splits = (p.k_blocks + kb_per_split - 1) / kb_per_split;   // every split owns >= 1 k-block
    p.seg_splits = p.k_splits; p.seg_kb = p.k_blocks; p.kb_per_split2 = kb_per_split;
  }
  p.tmem_cols = p.n_mma <= 32 ? 32 : p.n_mma <= 64 ? 64 : p.n_mma <= 128 ? 128 : 256;
  p.trigger_early = g_pdl >= 2 ? 1 : 0;
  constexpr int kSbSmemMax = 227 * 1024 - 1024;      // leave room for the kernel's few static __shared__ words
  static DeviceOnce configured;
  if (configured.first()) {
    cudaError_t e = cudaFuncSetAttribute(smallbatch_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSbSmemMax);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return -1; }
  }
  // batches <= 64 use M = 64 MMAs: half the activation-operand shared-memory read per k-block, which is what bounds
  // the stream (decode layer period 78.1 -> 70.7 us, tools/decode_timeline.py, profiles/r02_decode_timeline_m64.json);
  // md_debug_gemm bit 6 forces the M = 128 instantiation for A/B runs.
  const bool m64 = !(g_gemm_debug & 64) && batch <= 64;
  static DeviceOnce configured64;
  if (m64 && configured64.first()) {
    cudaError_t e = cudaFuncSetAttribute(smallbatch_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSbSmemMax);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return -1; }
  }
  CUtensorMap tW;
  if (make_tmap_bf16_2d(&tW, W, n_out, K, ldw, tile_n)) return -1;
  for (int b0 = 0; b0 < batch; b0 += BM) {
    p.batch = batch - b0 < BM ? batch - b0 : BM;
    const int a_rows = (p.batch + 7) / 8 * 8;
    p.a_bytes = a_rows * BK * 2;                     // a multiple of 1024: the weight tile stays swizzle-aligned
    p.stage_bytes = p.a_bytes + p.n_mma * BK * 2;
    const int bar_bytes = (2 * kSbMaxStages + 1) * 8 + 16;
    p.stages = (kSbSmemMax - 1024 - kSbTailPad - bar_bytes) / p.stage_bytes;
    if (p.stages > kSbMaxStages) p.stages = kSbMaxStages;
    if (p.stages < 2) { set_error("small-batch GEMM: tile does not fit shared memory"); return -1; }
    p.ws = ws + static_cast<long long>(b0) * n_out;
    CUtensorMap tX;
    if (make_tmap_bf16_2d(&tX, X + static_cast<long long>(b0) * ldx, p.batch, K, ldx, a_rows)) return -1;
    const int smem_bytes = p.stages * p.stage_bytes + kSbTailPad + bar_bytes + 1024;
    count_launch();
    const dim3 grid(n_tiles * p.k_splits), block(kSbThreads);
    const size_t smem = static_cast<size_t>(smem_bytes);
    const cudaError_t e = m64 ? launch_k(smallbatch_gemm_kernel<64>, grid, block, smem, stream, tX, tW, p)
                              : launch_k(smallbatch_gemm_kernel<128>, grid, block, smem, stream, tX, tW, p);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return -1; }
  }
  return p.k_splits;
}

// upper bound over the plans a launch may pick (callers size their partial-sum workspace with it)
int gemm_smallbatch_splits(int n_out, int K) {
  const int saved = g_gemm_debug;
  int s = 1;
  for (int legacy : {0, 8})
    for (int m : {64, 128}) {
      g_gemm_debug = (saved & ~8) | legacy;
      const int w = plan_smallbatch(n_out, K, 0, m).splits;
      if (w > s) s = w;
    }
  g_gemm_debug = saved;
  return s;
}

int gemm_smallbatch(const __nv_bfloat16* W, long long ldw, const __nv_bfloat16* X, long long ldx,
                 int n_out, int batch, int K, int splits, float* ws, cudaStream_t stream) {
  (void)splits;                                  // the plan decides (callers size ws with gemm_smallbatch_splits)
  const StreamPlan pl = plan_smallbatch(n_out, K, 0, (!(g_gemm_debug & 64) && batch <= 64) ? 64 : 128);
  return gemm_smallbatch_impl(W, ldw, X, ldx, n_out, batch, K, pl.kb, pl.tile_rows, ws, stream);
}

int gemm_rowform_qkv_rope(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M,
                          int K, const __nv_bfloat16* bias, const RopeEpilogue& epi, cudaStream_t stream) {
  const int N = epi.D + 2 * epi.n_kv_heads * 64;
  if (M <= 0 || K <= 0) return set_error("gemm_qkv_rope: empty problem");
  if (K % 8 || epi.D % 64 || epi.D != epi.n_heads * 64 || epi.n_kv_heads <= 0 || epi.n_heads % epi.n_kv_heads || !bias)
    return set_error("gemm_qkv_rope: unsupported shape");
  const int bn = 256;                                // 4 heads per column tile; N = 3D is a multiple of 64
  int cg = (M > BM) ? 2 : 1;
  if (g_force_cg == 1) cg = 1;
  CUtensorMap tA, tB;
  if (make_tmap_bf16_2d(&tA, A, M, K, lda, BM)) return 1;
  if (make_tmap_bf16_2d(&tB, W, N, K, ldw, bn / cg)) return 1;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.m_blocks = (M + BM * cg - 1) / (BM * cg);
  p.n_blocks = (N + bn - 1) / bn;
  p.k_blocks = (K + BK - 1) / BK;
  p.mode = EPI_QKV_ROPE;
  p.bias = bias;
  p.rope = epi;
  p.early_trigger = (g_pdl && !(g_gemm_debug & 32)) ? 1 : 0;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  const bool prof = g_prof.on && cudaStreamIsCapturing(stream, &cap) == cudaSuccess &&
                    cap == cudaStreamCaptureStatusNone;
  if (prof) cudaEventRecord(profile_event(), stream);
  const int rc = dispatch_gemm(bn, cg, tA, tB, p, stream);
  if (prof) {
    cudaEventRecord(profile_event(), stream);
    g_prof.flops += 2.0 * M * static_cast<double>(N) * K;
    g_prof.launches += 1;
  }
  return rc;
}

// Plan for a K-concatenated weight stream [A | B] whose halves must not share a split (seg_K = width of A, a
// multiple of 64).  A k-block costs four M = 128 MMAs whatever the tile height (~0.3 us measured with
// tools/decode_timeline.py: the 128-lane activation operand is re-read from shared memory every K = 16 step),
// so the plan minimises the k-blocks of the longest CTA: full 128-row tiles and many short splits, as long as
// every CTA is resident at once.
StreamPlan2 plan_smallbatch_2seg(int n_out, int K, int seg_K) {
  const int k_blocks = (K + BK - 1) / BK;
  const int ka = seg_K / BK, kbt = k_blocks - ka;
  const int sms = num_sms();
  StreamPlan2 best{BM, 1, ka, 1, kbt};
  long long best_cost = -1;
  if (g_gemm_debug & 4) {                 // A/B timing: the previous plan (equal splits of seg_K, balanced tile rows)
    const StreamPlan old = plan_smallbatch(n_out, K, ka);
    return StreamPlan2{old.tile_rows, ka / old.kb, old.kb, (kbt + old.kb - 1) / old.kb, old.kb};
  }
  for (int rows = BM; rows >= 64; --rows) {
    const int tiles = (n_out + rows - 1) / rows;
    for (int sa = 1; sa <= ka && sa <= 8; ++sa) {
      const int kba = (ka + sa - 1) / sa;
      if (kba * (sa - 1) >= ka) continue;                       // a split would be empty
      for (int sb = 1; sb <= kbt && sb <= 24; ++sb) {
        const int kbb = (kbt + sb - 1) / sb;
        if (kbb * (sb - 1) >= kbt) continue;
        const long long waves = (1LL * tiles * (sa + sb) + sms - 1) / sms;
        const int eff_rows = rows > 96 ? rows : 96;              // MMA floor in units of streamed weight rows
        const long long cost = waves * eff_rows * (kba > kbb ? kba : kbb) +
                               (1LL * (sa + sb) * n_out * 32 * 8) / (128LL * sms);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = StreamPlan2{rows, sa, kba, sb, kbb}; }
      }
    }
  }
  return best;
}

// same as gemm_smallbatch for W = [A | B] along K with the split boundaries of plan_smallbatch_2seg: splits
// [0, splits_a) hold A's partial sums, the rest B's.  Returns the number of splits, -1 on error.
int gemm_smallbatch_2seg(const __nv_bfloat16* W, long long ldw, const __nv_bfloat16* X, long long ldx,
                      int n_out, int batch, int K, int seg_K, float* ws, cudaStream_t stream) {
  if (seg_K <= 0 || seg_K >= K || seg_K % BK) { set_error("gemm_smallbatch_2seg: segment boundary must be a multiple of 64 inside K"); return -1; }
  const StreamPlan2 pl = plan_smallbatch_2seg(n_out, K, seg_K);
  return gemm_smallbatch_impl(W, ldw, X, ldx, n_out, batch, K, pl.kb_a, pl.tile_rows, ws, stream, seg_K / BK, pl.kb_b);
}

int splitk_epilogue(const float* ws, int splits, int B, int N, int mode, const __nv_bfloat16* bias,
                    const __nv_bfloat16* res, long long ldr, __nv_bfloat16* out, long long ldo,
                    cudaStream_t stream) {
  if (N % 2) return set_error("splitk_epilogue: N must be even");
  dim3 grid((N / 2 + 127) / 128, B);
  cudaError_t e = launch_k(splitk_epilogue_kernel, grid, dim3(128), 0, stream, ws, splits, B, N, mode, bias, res, ldr, out, ldo);
  count_launch();
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  return 0;
}

}  // namespace md
