// tcgen05 / TMA flash attention for the decoder prefill (head_dim 64, paged K/V, prefix-LM mask).
//
// Replaces F.scaled_dot_product_attention at reference text.py:46-50 under the mask of
// moondream.py:138-146 for prefill-sized query blocks.
//
// Work item = 128 queries of one (sequence, head), key tiles of 128 (= two 64-token KV pages); persistent CTAs (two per
// SM) walk the items.
//   warp 0     : TMA loader      (Q per item; K and V page boxes into separately released 2-stage rings)
//   warp 1     : MMA issuer      (S = Q K^T -> TMEM; O += P V -> TMEM), one elected lane
//   warps 2-5  : softmax         (thread = query row = TMEM lane: no shuffles; SINGLE-PASS online softmax in the
//                                 exp2 domain: the row's 128 scores of a tile are read from TMEM once and stay in
//                                 registers; P written to 128B-swizzled smem as the A operand of PV; O rescaled in
//                                 TMEM lazily, only when the row maximum grew by more than 2^8; final normalise + store)
// TMEM: S 128 columns + O 64 columns (256 allocated) and ~113 KB of shared memory, so two CTAs fit on an
// SM and their softmax / MMA phases interleave on the shared tensor pipe.
// V is consumed as an MN-major B operand straight from the [tokens, 64] page image TMA writes: no
// transpose anywhere.
#include <math.h>

#include "kernels.cuh"
#include "ptx.cuh"

namespace md {

namespace fa {
constexpr int BM = 128;              // queries per CTA
constexpr int BN = 128;              // keys per tile
constexpr int HD = 64;
constexpr int kStages = 2;
constexpr int kQBytes = BM * HD * 2;            // 16 KB
constexpr int kKVBytes = BN * HD * 2;           // 16 KB each for K and V
constexpr int kPBytes = BM * BN * 2;            // 32 KB (two 64-key blocks of [128 x 64])
constexpr int kSmemTiles = kQBytes + kStages * 2 * kKVBytes + kPBytes;   // 112 KB
constexpr int kSmemTotal = kSmemTiles + 256;   // two CTAs per SM: 2 x (kSmemTotal + 1 KB reserved) <= 228 KB
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColS = 0, kColO = 128;
constexpr float kLazyLog2 = 8.0f;     // O is rescaled only when a row maximum has grown by more than 2^8 (see the kernel)
}  // namespace fa


__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// max over one 32-score chunk of a row; `full` (tile-uniform) skips the prefix-LM / length mask
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], bool full, int k_first, int kv_len,
                                           int qpos, int prefix_len) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  if (full) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {            // max(max(m, a), b) folds into one 3-input FMNMX
      m0 = fmaxf(fmaxf(m0, __uint_as_float(v[i])), __uint_as_float(v[i + 1]));
      m1 = fmaxf(fmaxf(m1, __uint_as_float(v[i + 2])), __uint_as_float(v[i + 3]));
      m2 = fmaxf(fmaxf(m2, __uint_as_float(v[i + 4])), __uint_as_float(v[i + 5]));
      m3 = fmaxf(fmaxf(m3, __uint_as_float(v[i + 6])), __uint_as_float(v[i + 7]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int kpos = k_first + i;
      const bool ok = kpos < kv_len && (kpos <= qpos || (kpos < prefix_len && qpos < prefix_len));
      if (ok) m0 = fmaxf(m0, __uint_as_float(v[i]));
    }
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// exp2(s * scale - base) for one 32-score chunk -> bf16 -> four swizzled 16-byte chunks of the row's
// 128-byte line (chunk slots chunk0 .. chunk0+3); returns the row-sum contribution.
// POLY: pairs 1, 2, 5 of every 8 go through exp2_fma2 (FMA pipe) instead of MUFU.EX2.  It bought nothing while the
// boundary tile and the TMEM round trips dominated (profiles/r02_attention_ab.json); with those gone the per-phase clocks
// (tools/attn_phases.py) show the exponentials + pack + store phase at 1490 of 3990 clocks per tile, paced by the
// 16-per-clock MUFU shared by four softmax warps per SM sub-partition.
// 2^x on the FMA pipe: x = n + f with n = round(x) taken from the low mantissa bits of x + 1.5 * 2^23 and f in
// [-0.5, 0.5]; 2^f by a minimax cubic (relative error <= 7.5e-5, a fiftieth of a bf16 ulp: the result is rounded to
// bf16 right after); 2^n by adding n to the exponent field.
__device__ __forceinline__ float2 exp2_fma2(float2 x) {
  const float kMagic = 12582912.f;                       // 1.5 * 2^23
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 t = fadd2(x, make_float2(kMagic, kMagic));
  const float2 n = fadd2(t, make_float2(-kMagic, -kMagic));
  const float2 f = ffma2(n, make_float2(-1.f, -1.f), x);
  float2 q = ffma2(f, make_float2(0.05517146f, 0.05517146f), make_float2(0.24261086f, 0.24261086f));
  q = ffma2(q, f, make_float2(0.69326099f, 0.69326099f));
  q = ffma2(q, f, make_float2(0.99992809f, 0.99992809f));
  return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23)));
}
template <int POLY = 0>       // 0: MUFU only; 3 / 4: that many of every 8 exponential pairs on the FMA pipe (4 measured slower)
__device__ __forceinline__ float chunk_probs(const uint32_t (&v)[32], bool full, int k_first, int kv_len, int qpos,
                                             int prefix_len, float scale_log2, float base, uint8_t* line,
                                             int chunk0, int r) {
  uint32_t pk[16];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (full) {
    const float2 sc = make_float2(scale_log2, scale_log2), nb = make_float2(-base, -base);
    float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float2 x0 = ffma2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc, nb);
      const float2 x1 = ffma2(make_float2(__uint_as_float(v[2 * i + 2]), __uint_as_float(v[2 * i + 3])), sc, nb);
      // i is a compile-time constant after unrolling: pairs i = 2, 10 and i + 1 = 1, 5, 9, 13 take the FMA-pipe path
      const bool poly0 = (POLY == 3 && (i & 7) == 2) || (POLY == 4 && (i & 3) == 2);
      const bool poly1 = POLY >= 3 && (i & 3) == 0;
      const float2 e0 = poly0 ? exp2_fma2(x0) : make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
      const float2 e1 = poly1 ? exp2_fma2(x1) : make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
      s01 = fadd2(s01, e0);
      s23 = fadd2(s23, e1);
      pk[i] = pack_bf16x2(e0.x, e0.y);
      pk[i + 1] = pack_bf16x2(e1.x, e1.y);
    }
    a0 = s01.x; a1 = s01.y; a2 = s23.x; a3 = s23.y;
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float e[2];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int kpos = k_first + 2 * i + h2;
        const bool ok = kpos < kv_len && (kpos <= qpos || (kpos < prefix_len && qpos < prefix_len));
        e[h2] = ok ? ex2_approx(fmaf(__uint_as_float(v[2 * i + h2]), scale_log2, -base)) : 0.f;
      }
      a0 += e[0];
      a1 += e[1];
      pk[i] = pack_bf16x2(e[0], e[1]);
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(line + (((chunk0 + g) ^ (r & 7)) << 4)) =
        make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
  return (a0 + a1) + (a2 + a3);
}

// Boundary tiles of the single-pass softmax: a key is visible to a query row iff its position is below the row's limit
// (kv_len for ViT; under the prefix-LM mask min(kv_len, qpos < prefix ? max(qpos + 1, prefix) : qpos + 1)), so masking
// a 32-score chunk is one compare + select per element against n_valid = limit - first key; the scores become -inf
// and the unmasked fast paths (packed FFMA2 / FADD2, ex2(-inf) = 0) run on every tile.  (The per-element predicate +
// scalar accumulation the two-pass kernel uses on boundary tiles made that one tile in six cost several full tiles.)
__device__ __forceinline__ void mask_chunk(uint32_t (&v)[32], int n_valid) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = (i < n_valid) ? v[i] : 0xff800000u;
}

struct FaTcParams {
  const int* q_offsets;     // [n_seqs + 1]
  const int* start_pos;     // [n_seqs]
  const int* block_tables;  // [n_seqs][max_blocks]
  int max_blocks, n_heads, n_kv_heads, layer, n_pages, prefix_len;
  int n_seqs, nq_tiles;     // work items = nq_tiles x n_heads x n_seqs (128-query tile, head, sequence)
  __nv_bfloat16* out;       // [T_total, n_heads * 64]
  float scale_log2;
};

// The softmax of these kernels was measured step by step this round (profiles/r02_attention_ab.json,
// r02_ncu_attention_summary.json, tools/attn_phases.py):
//   * two passes over S (row maximum, then probabilities) pull 160-200 KB per tile through the TMEM read path
//     (~64 B/clk per SM) and rewrite O whenever a maximum moved: tensor pipe 18-21 %, nothing else saturated;
//   * SINGLE = true (default): a thread reads its row's 128 scores ONCE and keeps them in registers (168 registers, two
//     CTAs per SM still fit), S is released to the MMA warp at once (the next Q K^T overlaps this tile's softmax), and
//     O is rescaled lazily as in FlashAttention-4: only when the row maximum grew by more than 2^8 — otherwise the
//     probabilities run up to 2^8, which the final O / l divides out exactly;
//   * boundary tiles are masked with one compare + select per score against the row's key limit and then take the same
//     packed fast path (the per-element predicate form cost 1049 instead of 228 instructions per thread on one tile in six);
//   * 3 of every 8 exponential pairs are evaluated on the FMA pipe (exp2_fma2): the exponential phase is paced by the
//     16-per-clock MUFU;
//   * two softmax warpgroups per CTA (64 scores per thread, row maximum agreed through TMEM columns or shared memory)
//     were built and measured too: the agreement barrier costs more than it saves (0.314 vs 0.300 ms) — removed.
// SINGLE = false is the round-1 two-pass form, kept for A/B runs (md_debug_attention_impl(2)).
// PROF adds per-phase clock sums of one softmax thread for tools/attn_phases.py (md_debug_attention_impl(3)).
template <bool SINGLE, bool PROF = false, int POLY = 3>
__global__ void __launch_bounds__(fa::kThreads, 2)
fa_tc_prefill_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     const FaTcParams p) {
  using namespace fa;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023) __trap();             // swizzled tiles need 1024-byte alignment
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;                       // [stage][16 KB]
  uint8_t* sV = sK + kStages * kKVBytes;            // [stage][16 KB]
  uint8_t* sP = sV + kStages * kKVBytes;            // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemTiles);
  // K and V stages are released separately: K right after its Q K^T (early in the PREVIOUS tile's softmax), V after
  // its P V, so a tile's loads are requested about two softmax durations ahead of their use.
  uint64_t* q_full = bars;                          // 1
  uint64_t* k_full = bars + 1;                      // [2]
  uint64_t* k_empty = bars + 3;                     // [2]
  uint64_t* s_full = bars + 5;                      // 1
  uint64_t* s_empty = bars + 6;                     // 1 (128 arrivals)
  uint64_t* p_full = bars + 7;                      // 1 (128 arrivals)
  uint64_t* p_empty = bars + 8;                     // 1
  uint64_t* v_full = bars + 9;                      // [2]
  uint64_t* v_empty = bars + 11;                    // [2]
  uint64_t* p_half = bars + 13;                     // 1: the first 64-key block of P has been read by P V
  uint64_t* q_empty = bars + 14;                    // 1: the item's last Q K^T has read Q
  uint64_t* o_free = bars + 15;                     // 1 (128 arrivals): the item's O has been read out of TMEM
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmKV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(p_empty, 1);
    mbar_init(p_half, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: barriers, TMEM and descriptors above are set up under the predecessor's tail
  pdl_launch_dependents();
  pdl_wait();

  // PERSISTENT over work items = (128-query tile, head, sequence): a CTA that lives for one item spends a third of its
  // life outside the key-tile loop (launch, barrier / TMEM set-up, first Q / K round trip, O read-out, exit: measured
  // with tools/attn_phases.py), so a CTA walks items blockIdx.x, blockIdx.x + gridDim.x, ...; the three roles iterate
  // the same list, barrier parities run on a global tile counter g and item counter ni, the loader runs ahead into the
  // next item as soon as Q (q_empty) and the K / V stages are free, and the first P V of an item waits until the
  // previous item's O has left TMEM (o_free).  With gridDim.x = number of items it degenerates to one item per CTA.
  const int nq_tiles = p.nq_tiles;
  const int total_items = nq_tiles * p.n_heads * p.n_seqs;
  struct Item { int q0, head, q_off, n_q, q_pos0, kv_len, n_tiles; const int* btab; bool ok; };
  auto get_item = [&](int it) {
    Item t;
    const int qt = it % nq_tiles, rest = it / nq_tiles;
    t.head = rest % p.n_heads;
    const int seq = rest / p.n_heads;
    t.q0 = qt * BM;
    t.q_off = p.q_offsets[seq];
    t.n_q = p.q_offsets[seq + 1] - t.q_off;
    t.ok = t.q0 < t.n_q;
    t.q_pos0 = p.start_pos[seq];
    t.kv_len = t.q_pos0 + t.n_q;
    t.btab = p.block_tables + static_cast<long long>(seq) * p.max_blocks;
    // keys this query tile can see (prefix rows see the whole prefix; later rows are causal)
    int reach = t.q_pos0 + min(t.q0 + BM, t.n_q);
    if (t.q_pos0 + t.q0 < p.prefix_len) reach = max(reach, p.prefix_len);
    reach = min(reach, t.kv_len);
    t.n_tiles = (reach + BN - 1) / BN;
    return t;
  };

  if (warp == 0) {
    // ------------------------------ TMA loader ------------------------------
    if (lane == 0) {
      const long long page_rows = 2LL * p.n_kv_heads * 64;              // rows of one page (k then v)
      uint32_t g = 0, ni = 0;
      for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
        const Item t = get_item(it);
        if (!t.ok) continue;
        const int kv_head = t.head / (p.n_heads / p.n_kv_heads);       // grouped-query attention (text.py:49)
        mbar_wait(q_empty, (ni & 1u) ^ 1u);                             // the previous item's Q K^Ts are done with Q
        mbar_arrive_expect_tx(q_full, kQBytes);
        tma_load_2d(sQ, &tmQ, q_full, t.head * HD, t.q_off + t.q0);
        for (int j = 0; j < t.n_tiles; ++j, ++g) {
          const int st = g & 1;
          const uint32_t u = g >> 1;
          long long row_k[2];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int blk = min(2 * j + h2, p.max_blocks - 1);
            row_k[h2] = (static_cast<long long>(p.layer) * p.n_pages + t.btab[blk]) * page_rows + static_cast<long long>(kv_head) * 64;
          }
          mbar_wait(&k_empty[st], (u & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[st], kKVBytes);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
            tma_load_2d(sK + st * kKVBytes + h2 * (kKVBytes / 2), &tmKV, &k_full[st], 0, static_cast<int32_t>(row_k[h2]));
          mbar_wait(&v_empty[st], (u & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[st], kKVBytes);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
            tma_load_2d(sV + st * kKVBytes + h2 * (kKVBytes / 2), &tmKV, &v_full[st], 0,
                        static_cast<int32_t>(row_k[h2] + static_cast<long long>(p.n_kv_heads) * 64));
        }
        ++ni;
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16_f32(BM, BN);          // S[128,128] = Q K^T
      constexpr uint32_t idesc_pv = make_idesc_bf16_f32_bmn(BM, HD);      // O[128,64] += P V (V MN-major)
      const uint32_t tS = tmem_base + kColS, tO = tmem_base + kColO;
      const uint64_t dQ = make_desc_k_sw128(smem_u32(sQ));
      uint32_t g = 0, ni = 0;
      // P V of the tile with global index gg; `first` = the item's first tile (overwrites O)
      auto issue_pv = [&](uint32_t gg, bool first) {
        const int st = gg & 1;
        mbar_wait(&v_full[st], (gg >> 1) & 1u);
        mbar_wait(p_full, gg & 1u);
        if (first) mbar_wait(o_free, (ni & 1u) ^ 1u);                   // the previous item's O has been read out
        tc_fence_after();
        const uint32_t sp = smem_u32(sP), sv = smem_u32(sV + st * kKVBytes);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          // A: P, K-major: 16 keys = 32 B inside the 128-byte row; the second 64-key block is 16 KB on
          // B: V, MN-major: 16 keys = 16 rows of 128 B = 2048 B
          const uint64_t da = make_desc_k_sw128(sp + (k >> 2) * (BM * 128)) + static_cast<uint64_t>(2 * (k & 3));
          const uint64_t db = make_desc_mn_sw128(sv + k * 2048, 16);
          umma_bf16(tO, da, db, idesc_pv, (!first || k > 0) ? 1u : 0u);
          if (k == 3) umma_commit(p_half);   // the first 64-key block of P may be overwritten already
        }
        umma_commit(p_empty);            // P buffer reusable, O updated
        umma_commit(&v_empty[st]);       // V stage reusable
      };
      for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
        const Item t = get_item(it);
        if (!t.ok) continue;
        mbar_wait(q_full, ni & 1u);
        for (int j = 0; j < t.n_tiles; ++j, ++g) {
          const int st = g & 1;
          mbar_wait(&k_full[st], (g >> 1) & 1u);
          mbar_wait(s_empty, (g & 1u) ^ 1u);                            // the previous tile's softmax has read S
          tc_fence_after();
          const uint64_t dK = make_desc_k_sw128(smem_u32(sK + st * kKVBytes));
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16(tS, dQ + static_cast<uint64_t>(2 * k), dK + static_cast<uint64_t>(2 * k), idesc_qk, k > 0 ? 1u : 0u);
          umma_commit(s_full);
          umma_commit(&k_empty[st]);       // K stage reusable as soon as this Q K^T has read it
          if (j + 1 == t.n_tiles) umma_commit(q_empty);                // Q reusable: the loader may fetch the next item's
          if (j > 0) issue_pv(g - 1, j == 1);
        }
        issue_pv(g - 1, t.n_tiles == 1);
        ++ni;
      }
    }
  } else {
    // ------------------------------ softmax (128 threads, one query row each) ------------------------------
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // row in the tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + kColS, tO = tmem_base + lane_addr + kColO;
    uint8_t* prow = sP + r * 128;
    // debug timeline (tools/attn_phases.py): clock sums of the phases of one softmax thread
    const bool tlp = PROF && tl_on() && threadIdx.x == 64;
    long long ph[6] = {0, 0, 0, 0, 0, 0}, tc0 = 0, tiles_done = 0;
    auto tick = [&](int i) { if (tlp) { const long long now = clock64(); ph[i] += now - tc0; tc0 = now; } };
    const long long t_begin = tlp ? clock64() : 0;
    uint32_t g = 0;
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      const Item t = get_item(it);
      if (!t.ok) continue;
      const int q0 = t.q0, head = t.head, q_off = t.q_off, n_q = t.n_q, q_pos0 = t.q_pos0, kv_len = t.kv_len, n_tiles = t.n_tiles;
      const int qpos = q_pos0 + q0 + r;
      const bool row_ok = q0 + r < n_q;
      float m_run = -INFINITY, l_run = 0.f;               // SINGLE: m_run = the maximum O and l are currently scaled by
      const int q_lo = q_pos0 + q0;                        // smallest query position of this item
      // first key position this row may NOT attend to (prefix-LM mask of moondream.py:138-146 + sequence length)
      const int row_lim = min(kv_len, qpos < p.prefix_len ? max(qpos + 1, p.prefix_len) : qpos + 1);
      for (int j = 0; j < n_tiles; ++j, ++g) {
      const int k0 = j * BN;
      // interior tiles need no masking: every key exists and every row of the CTA may attend to it
      const bool full = (k0 + BN <= kv_len) &&
                        (k0 + BN - 1 <= q_lo || (k0 + BN <= p.prefix_len && q_lo + BM <= p.prefix_len));
      if (tlp) tc0 = clock64();
      mbar_wait(s_full, g & 1u);
      tc_fence_after();
      if constexpr (SINGLE) {
        tick(0);                                          // waited for S
        // ---- the row's 128 scores are read from TMEM once and live in registers from here on ----
        uint32_t v0[32], v1[32], v2[32], v3[32];
        tmem_ld_32x32(tS, v0);
        tmem_ld_32x32(tS + 32, v1);
        tmem_ld_32x32(tS + 64, v2);
        tmem_ld_32x32(tS + 96, v3);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_empty);                             // S may be overwritten by the next Q K^T
        tick(1);                                          // TMEM load
        if (!full) {                                      // tile-uniform: boundary tiles only
          mask_chunk(v0, row_lim - k0);
          mask_chunk(v1, row_lim - (k0 + 32));
          mask_chunk(v2, row_lim - (k0 + 64));
          mask_chunk(v3, row_lim - (k0 + 96));
        }
        const float tile_max = fmaxf(fmaxf(chunk_max(v0, true, 0, 0, 0, 0), chunk_max(v1, true, 0, 0, 0, 0)),
                                     fmaxf(chunk_max(v2, true, 0, 0, 0, 0), chunk_max(v3, true, 0, 0, 0, 0)));
        float alpha = 1.f;
        bool grow = false;
        if (m_run == -INFINITY) {
          m_run = tile_max;                               // nothing accumulated yet
        } else if ((tile_max - m_run) * p.scale_log2 > kLazyLog2) {
          alpha = ex2_approx((m_run - tile_max) * p.scale_log2);
          m_run = tile_max;
          grow = true;
        }
        tick(2);                                          // mask + maximum
        // The previous P V reads P's first 64-key block, then its second: this tile's first block may be written as
        // soon as the first half of that P V is done (p_half); only an O rescale needs all of it.
        bool pv_done = j == 0;
        if (j > 0) {
          const bool any_grow = __any_sync(0xffffffffu, grow);      // warp-uniform; lanes that did not grow multiply by 1
          if (any_grow) {
            mbar_wait(p_empty, (g - 1u) & 1u);
            pv_done = true;
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < HD / 16; ++c) {           // 16 columns at a time: the scores stay in registers
              uint32_t o[16];
              tmem_ld_32x16(tO + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x16(tO + c * 16, o);
            }
            tmem_st_wait();
          } else {
            mbar_wait(p_half, (g - 1u) & 1u);
          }
        }
        l_run *= alpha;
        tick(3);                                          // waited for (half of) the previous P V (+ rare rescale)
        const float base = (m_run == -INFINITY) ? 0.f : m_run * p.scale_log2;
        l_run += chunk_probs<POLY>(v0, true, 0, 0, 0, 0, p.scale_log2, base, prow, 0, r);
        l_run += chunk_probs<POLY>(v1, true, 0, 0, 0, 0, p.scale_log2, base, prow, 4, r);
        if (!pv_done) mbar_wait(p_empty, (g - 1u) & 1u);
        l_run += chunk_probs<POLY>(v2, true, 0, 0, 0, 0, p.scale_log2, base, prow + BM * 128, 0, r);
        l_run += chunk_probs<POLY>(v3, true, 0, 0, 0, 0, p.scale_log2, base, prow + BM * 128, 4, r);
        tick(4);                                          // exponentials, pack, P stores
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(p_full);
        tick(5);                                          // fence + arrive
        continue;
      }

      // ---- two-pass form (A/B reference), pass 1: row maximum (two 32-column chunks in flight) ----
      float mx = m_run;
      {
        uint32_t va[32], vb[32];
        tmem_ld_32x32(tS, va);
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          tmem_ld_wait();
          tmem_ld_32x32(tS + (2 * cc + 1) * 32, vb);
          mx = fmaxf(mx, chunk_max(va, full, k0 + (2 * cc) * 32, kv_len, qpos, p.prefix_len));
          tmem_ld_wait();
          if (cc == 0) tmem_ld_32x32(tS + 64, va);
          mx = fmaxf(mx, chunk_max(vb, full, k0 + (2 * cc + 1) * 32, kv_len, qpos, p.prefix_len));
        }
      }
      const float base = (mx == -INFINITY) ? 0.f : mx * p.scale_log2;
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run * p.scale_log2 - base);
      // the previous P V must have completed before O is rescaled and P is overwritten
      if (j > 0) {
        mbar_wait(p_empty, (g - 1u) & 1u);
        tc_fence_after();
        if (__any_sync(0xffffffffu, mx != m_run)) {      // warp-uniform: rescale this warp's 32 rows of O
          uint32_t o0[32], o1[32];
          tmem_ld_32x32(tO, o0);
          tmem_ld_32x32(tO + 32, o1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            o0[i] = __float_as_uint(__uint_as_float(o0[i]) * alpha);
            o1[i] = __float_as_uint(__uint_as_float(o1[i]) * alpha);
          }
          tmem_st_32x32(tO, o0);
          tmem_st_32x32(tO + 32, o1);
          tmem_st_wait();
        }
      }
      l_run *= alpha;
      m_run = mx;

      // ---- pass 2: probabilities -> bf16 -> swizzled smem (A operand of P V) ----
      {
        uint32_t va[32], vb[32];
        tmem_ld_32x32(tS, va);
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          tmem_ld_wait();
          tmem_ld_32x32(tS + (2 * cc + 1) * 32, vb);
          l_run += chunk_probs(va, full, k0 + (2 * cc) * 32, kv_len, qpos, p.prefix_len, p.scale_log2, base,
                               prow + cc * (BM * 128), 0, r);
          tmem_ld_wait();
          if (cc == 0) tmem_ld_32x32(tS + 64, va);
          l_run += chunk_probs(vb, full, k0 + (2 * cc + 1) * 32, kv_len, qpos, p.prefix_len, p.scale_log2, base,
                               prow + cc * (BM * 128), 4, r);
        }
      }
      // S has been consumed; P is in shared memory: publish both
      tc_fence_before();
      mbar_arrive(s_empty);
      fence_proxy_async_smem();
      mbar_arrive(p_full);
    }
      tiles_done += n_tiles;
      // ---- epilogue: O / l -> bf16 -> global ----
      mbar_wait(p_empty, (g - 1u) & 1u);
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      __nv_bfloat16* orow = p.out + (static_cast<long long>(q_off) + q0 + r) * (static_cast<long long>(p.n_heads) * HD) + head * HD;
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(tO, o0);
      tmem_ld_32x32(tO + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free);                                  // the next item's first P V may overwrite O
      if (row_ok) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o0[8 * g8]) * inv, __uint_as_float(o0[8 * g8 + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o0[8 * g8 + 2]) * inv, __uint_as_float(o0[8 * g8 + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o0[8 * g8 + 4]) * inv, __uint_as_float(o0[8 * g8 + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o0[8 * g8 + 6]) * inv, __uint_as_float(o0[8 * g8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + g8 * 8) = w;
        }
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o1[8 * g8]) * inv, __uint_as_float(o1[8 * g8 + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o1[8 * g8 + 2]) * inv, __uint_as_float(o1[8 * g8 + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o1[8 * g8 + 4]) * inv, __uint_as_float(o1[8 * g8 + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o1[8 * g8 + 6]) * inv, __uint_as_float(o1[8 * g8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + 32 + g8 * 8) = w;
        }
      }
    }
    if (tlp) {     // two records: phase sums [0..4], then [5], total softmax-loop clocks, tiles
      tl_emit(5u << 28, ph[0], ph[1], ph[2], ph[3], ph[4]);
      tl_emit(6u << 28, ph[5], clock64() - t_begin, tiles_done, 0, 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, fa::kTmemCols);
  }
}

int prefill_attention_tc(const __nv_bfloat16* q, int n_heads, int n_kv_heads, int total_tokens, const int* q_offsets,
                         const int* start_pos, int n_seqs, int max_q, int prefix_len,
                         const __nv_bfloat16* kv_pool, int n_pages, int n_layers, const int* block_tables,
                         int max_blocks, int layer, __nv_bfloat16* out, cudaStream_t stream) {
  if (n_seqs <= 0 || max_q <= 0) return set_error("prefill_attention: empty batch");
  if (n_kv_heads <= 0) n_kv_heads = n_heads;
  if (n_heads % n_kv_heads) return set_error("prefill_attention: n_heads must be a multiple of n_kv_heads");
  CUtensorMap tQ, tKV;
  if (make_tmap_bf16_2d(&tQ, q, total_tokens, static_cast<long long>(n_heads) * 64,
                        static_cast<long long>(n_heads) * 64, fa::BM)) return 1;
  const long long pool_rows = static_cast<long long>(n_layers) * n_pages * 2 * n_kv_heads * 64;
  if (pool_rows >= (1LL << 31)) return set_error("prefill_attention: KV pool too large for 32-bit TMA coordinates");
  if (make_tmap_bf16_2d(&tKV, kv_pool, pool_rows, 64, 64, 64)) return 1;
  static DeviceOnce configured;
  if (configured.first()) {
    for (auto* fn : {fa_tc_prefill_kernel<true>, fa_tc_prefill_kernel<false>, fa_tc_prefill_kernel<true, true>}) {
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, fa::kSmemTotal);
      if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
      // two CTAs per SM need the full shared-memory carve-out
      e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    }
  }
  FaTcParams p{};
  p.q_offsets = q_offsets; p.start_pos = start_pos; p.block_tables = block_tables;
  p.max_blocks = max_blocks; p.n_heads = n_heads; p.n_kv_heads = n_kv_heads; p.layer = layer; p.n_pages = n_pages;
  p.prefix_len = prefix_len; p.out = out;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.n_seqs = n_seqs;
  p.nq_tiles = (max_q + fa::BM - 1) / fa::BM;
  const long long items = static_cast<long long>(p.nq_tiles) * n_heads * n_seqs;
  if (items >= (1LL << 31)) return set_error("prefill_attention: too many work items");
  // persistent: two CTAs per SM walk the items (md_debug_attention_impl(4): one CTA per item, the previous form)
  const long long resident = 2LL * num_sms();
  dim3 grid(static_cast<unsigned>(g_attention_impl == 4 || items < resident ? items : resident));
  count_launch();
  // default: the single-pass softmax; md_debug_attention_impl(2) = the two-pass form, (3) = single pass with phase clocks
  const dim3 block(fa::kThreads);
  const cudaError_t e =
      g_attention_impl == 2 ? launch_k(fa_tc_prefill_kernel<false>, grid, block, fa::kSmemTotal, stream, tQ, tKV, p)
      : g_attention_impl == 3 ? launch_k(fa_tc_prefill_kernel<true, true>, grid, block, fa::kSmemTotal, stream, tQ, tKV, p)
                              : launch_k(fa_tc_prefill_kernel<true>, grid, block, fa::kSmemTotal, stream, tQ, tKV, p);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  return 0;
}


// ------------------------------------------------------------------------------------------------
// ViT self-attention (reference layers.py:155-166): head_dim 72, 729 tokens per crop, no mask.
// Same warp roles and softmax as the prefill kernel.  The 72-wide head is staged as a 64-wide
// 128B-swizzled block plus a 16-wide 32B-swizzled block (dims 64..79; 72..79 are TMA out-of-bounds
// zero fill through a 3-D view [token][3*H heads][72]), so
//   S  = Q K^T : 4 k-steps on the 64-block + 1 k-step on the 16-block,
//   O += P V   : N = 64 columns from the 64-block and N = 16 columns from the 16-block (both MN-major).
// One K tile (released as soon as S is complete) and two V stages: 112 KB, 2 CTAs per SM.
// ------------------------------------------------------------------------------------------------
namespace fv {
constexpr int BM = 128, BN = 128, HD = 72;
constexpr int kBlk0 = 128 * 128;                 // [128 rows x 64] bf16, 16 KB
constexpr int kBlk1 = 128 * 32;                  // [128 rows x 16] bf16, 4 KB
constexpr int kTile = kBlk0 + kBlk1;             // 20 KB
constexpr int kPBytes = BM * BN * 2;             // 32 KB
constexpr int kSmemTiles = kTile /*Q*/ + kTile /*K*/ + 2 * kTile /*V*/ + kPBytes;   // 112 KB
constexpr int kSmemTotal = kSmemTiles + 256;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColS = 0, kColO = 128;       // O: 80 columns (64 + 16)
}  // namespace fv

struct FaVitParams {
  int seq, n_heads, n_crops;
  __nv_bfloat16* out;        // [n_crops * seq, n_heads * 72]
  float scale_log2;
};

// SINGLE / two-pass: as in fa_tc_prefill_kernel.
template <bool SINGLE, int POLY = 3>
__global__ void __launch_bounds__(fv::kThreads, 2)
fa_tc_vit_kernel(const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm16,
                 const FaVitParams p) {
  using namespace fv;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023) __trap();
  uint8_t* sQ = smem;                               // blk0 | blk1
  // ONE K tile and TWO V stages: the next Q K^T cannot start before this tile's softmax has read S anyway, and K is
  // free again as soon as its Q K^T has run (early in the previous tile's softmax), so a second K stage buys nothing;
  // V is needed at the END of its tile's softmax and, single-buffered, could only be requested at its start.
  uint8_t* sK = sQ + kTile;                         // [blk0 | blk1]
  uint8_t* sV = sK + kTile;                         // [2][blk0 | blk1]
  uint8_t* sP = sV + 2 * kTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemTiles);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;                      // [2]
  uint64_t* v_empty = bars + 5;                     // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* s_empty = bars + 8;                     // 128 arrivals
  uint64_t* p_full = bars + 9;                      // 128 arrivals
  uint64_t* p_empty = bars + 10;
  uint64_t* p_half = bars + 11;                     // the first 64-key block of P has been read by P V
  uint64_t* q_empty = bars + 12;                    // the item's last Q K^T has read Q
  uint64_t* o_free = bars + 13;                     // 128 arrivals: the item's O has been read out of TMEM
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_q = p.seq, kv_len = p.seq;
  const int n_tiles = (kv_len + BN - 1) / BN;
  const int H = p.n_heads;
  // persistent over work items (128-query tile, head, crop), as fa_tc_prefill_kernel
  const int nq_tiles = (p.seq + BM - 1) / BM;
  const int total_items = nq_tiles * H * p.n_crops;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm64);
    prefetch_tensormap(&tm16);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(p_empty, 1);
    mbar_init(p_half, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------ TMA loader ------------------------------
    if (lane == 0) {
      uint32_t g = 0, ni = 0;
      for (int it = blockIdx.x; it < total_items; it += gridDim.x, ++ni) {
        const int q0 = (it % nq_tiles) * BM, head = (it / nq_tiles) % H, row0 = (it / (nq_tiles * H)) * p.seq;
        mbar_wait(q_empty, (ni & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full, kTile);
        tma_load_3d(sQ, &tm64, q_full, 0, head, row0 + q0);
        tma_load_3d(sQ + kBlk0, &tm16, q_full, 64, head, row0 + q0);
        for (int j = 0; j < n_tiles; ++j, ++g) {
          const int st = g & 1;
          mbar_wait(k_empty, (g & 1u) ^ 1u);
          mbar_arrive_expect_tx(k_full, kTile);
          tma_load_3d(sK, &tm64, k_full, 0, H + head, row0 + j * BN);
          tma_load_3d(sK + kBlk0, &tm16, k_full, 64, H + head, row0 + j * BN);
          mbar_wait(&v_empty[st], ((g >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&v_full[st], kTile);
          tma_load_3d(sV + st * kTile, &tm64, &v_full[st], 0, 2 * H + head, row0 + j * BN);
          tma_load_3d(sV + st * kTile + kBlk0, &tm16, &v_full[st], 64, 2 * H + head, row0 + j * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16_f32(BM, BN);
      constexpr uint32_t idesc_pv64 = make_idesc_bf16_f32_bmn(BM, 64);
      constexpr uint32_t idesc_pv16 = make_idesc_bf16_f32_bmn(BM, 16);
      const uint32_t tS = tmem_base + kColS, tO = tmem_base + kColO;
      const uint64_t dQ0 = make_desc_k_sw128(smem_u32(sQ));
      const uint64_t dQ1 = make_desc_sw32(smem_u32(sQ + kBlk0));
      uint32_t g = 0, ni = 0;
      auto issue_pv = [&](uint32_t gg, bool first) {
        const int vs = gg & 1;
        mbar_wait(&v_full[vs], (gg >> 1) & 1u);
        mbar_wait(p_full, gg & 1u);
        if (first) mbar_wait(o_free, (ni & 1u) ^ 1u);              // the previous item's O has been read out
        tc_fence_after();
        const uint32_t sp = smem_u32(sP), sv0 = smem_u32(sV + vs * kTile), sv1 = smem_u32(sV + vs * kTile + kBlk0);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          const uint64_t da = make_desc_k_sw128(sp + (k >> 2) * (BM * 128)) + static_cast<uint64_t>(2 * (k & 3));
          const uint32_t acc = (!first || k > 0) ? 1u : 0u;
          umma_bf16(tO, da, make_desc_mn_sw128(sv0 + k * 2048, 16), idesc_pv64, acc);       // dims 0..63
          umma_bf16(tO + 64, da, make_desc_sw32(sv1 + k * 512), idesc_pv16, acc);           // dims 64..79
          if (k == 3) umma_commit(p_half);
        }
        umma_commit(p_empty);
        umma_commit(&v_empty[vs]);
      };
      for (int it = blockIdx.x; it < total_items; it += gridDim.x, ++ni) {
        mbar_wait(q_full, ni & 1u);
        for (int j = 0; j < n_tiles; ++j, ++g) {
          mbar_wait(k_full, g & 1u);
          mbar_wait(s_empty, (g & 1u) ^ 1u);
          tc_fence_after();
          const uint64_t dK0 = make_desc_k_sw128(smem_u32(sK));
          const uint64_t dK1 = make_desc_sw32(smem_u32(sK + kBlk0));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tS, dQ0 + static_cast<uint64_t>(2 * k), dK0 + static_cast<uint64_t>(2 * k), idesc_qk, k > 0 ? 1u : 0u);
          umma_bf16(tS, dQ1, dK1, idesc_qk, 1u);                     // dims 64..79 (72..79 are zero)
          umma_commit(s_full);
          umma_commit(k_empty);
          if (j + 1 == n_tiles) umma_commit(q_empty);
          if (j > 0) issue_pv(g - 1, j == 1);
        }
        issue_pv(g - 1, n_tiles == 1);
      }
    }
  } else {
    // ------------------------------ softmax (128 threads, one query row each) ------------------------------
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + kColS, tO = tmem_base + lane_addr + kColO;
    uint8_t* prow = sP + r * 128;
    const int qpos = 1 << 30;                                  // no causal structure: every key is allowed
    uint32_t g = 0;
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      const int q0 = (it % nq_tiles) * BM, head = (it / nq_tiles) % H, row0 = (it / (nq_tiles * H)) * p.seq;
      const bool row_ok = q0 + r < n_q;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++g) {
      const int k0 = j * BN;
      const bool full = k0 + BN <= kv_len;
      mbar_wait(s_full, g & 1u);
      tc_fence_after();
      if constexpr (SINGLE) {
        uint32_t v0[32], v1[32], v2[32], v3[32];
        tmem_ld_32x32(tS, v0);
        tmem_ld_32x32(tS + 32, v1);
        tmem_ld_32x32(tS + 64, v2);
        tmem_ld_32x32(tS + 96, v3);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_empty);
        if (!full) {
          mask_chunk(v0, kv_len - k0);
          mask_chunk(v1, kv_len - (k0 + 32));
          mask_chunk(v2, kv_len - (k0 + 64));
          mask_chunk(v3, kv_len - (k0 + 96));
        }
        const float tile_max = fmaxf(fmaxf(chunk_max(v0, true, 0, 0, 0, 0), chunk_max(v1, true, 0, 0, 0, 0)),
                                     fmaxf(chunk_max(v2, true, 0, 0, 0, 0), chunk_max(v3, true, 0, 0, 0, 0)));
        float alpha = 1.f;
        bool grow = false;
        if (m_run == -INFINITY) {
          m_run = tile_max;
        } else if ((tile_max - m_run) * p.scale_log2 > fa::kLazyLog2) {
          alpha = ex2_approx((m_run - tile_max) * p.scale_log2);
          m_run = tile_max;
          grow = true;
        }
        bool pv_done = j == 0;
        if (j > 0) {
          if (__any_sync(0xffffffffu, grow)) {
            mbar_wait(p_empty, (g - 1u) & 1u);
            pv_done = true;
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 5; ++c) {                      // 80 O columns, 16 at a time
              uint32_t o[16];
              tmem_ld_32x16(tO + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x16(tO + c * 16, o);
            }
            tmem_st_wait();
          } else {
            mbar_wait(p_half, (g - 1u) & 1u);
          }
        }
        l_run *= alpha;
        const float base = (m_run == -INFINITY) ? 0.f : m_run * p.scale_log2;
        l_run += chunk_probs<POLY>(v0, true, 0, 0, 0, 0, p.scale_log2, base, prow, 0, r);
        l_run += chunk_probs<POLY>(v1, true, 0, 0, 0, 0, p.scale_log2, base, prow, 4, r);
        if (!pv_done) mbar_wait(p_empty, (g - 1u) & 1u);
        l_run += chunk_probs<POLY>(v2, true, 0, 0, 0, 0, p.scale_log2, base, prow + BM * 128, 0, r);
        l_run += chunk_probs<POLY>(v3, true, 0, 0, 0, 0, p.scale_log2, base, prow + BM * 128, 4, r);
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(p_full);
        continue;
      }
      // ---- two-pass form (A/B reference) ----
      float mx = m_run;
      {
        uint32_t va[32], vb[32];
        tmem_ld_32x32(tS, va);
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          tmem_ld_wait();
          tmem_ld_32x32(tS + (2 * cc + 1) * 32, vb);
          mx = fmaxf(mx, chunk_max(va, full, k0 + (2 * cc) * 32, kv_len, qpos, 0));
          tmem_ld_wait();
          if (cc == 0) tmem_ld_32x32(tS + 64, va);
          mx = fmaxf(mx, chunk_max(vb, full, k0 + (2 * cc + 1) * 32, kv_len, qpos, 0));
        }
      }
      const float base = (mx == -INFINITY) ? 0.f : mx * p.scale_log2;
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run * p.scale_log2 - base);
      if (j > 0) {
        mbar_wait(p_empty, (g - 1u) & 1u);
        tc_fence_after();
        if (__any_sync(0xffffffffu, mx != m_run)) {
          uint32_t o0[32], o1[32], o2[16];
          tmem_ld_32x32(tO, o0);
          tmem_ld_32x32(tO + 32, o1);
          tmem_ld_32x16(tO + 64, o2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            o0[i] = __float_as_uint(__uint_as_float(o0[i]) * alpha);
            o1[i] = __float_as_uint(__uint_as_float(o1[i]) * alpha);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) o2[i] = __float_as_uint(__uint_as_float(o2[i]) * alpha);
          tmem_st_32x32(tO, o0);
          tmem_st_32x32(tO + 32, o1);
          tmem_st_32x16(tO + 64, o2);
          tmem_st_wait();
        }
      }
      l_run *= alpha;
      m_run = mx;
      {
        uint32_t va[32], vb[32];
        tmem_ld_32x32(tS, va);
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          tmem_ld_wait();
          tmem_ld_32x32(tS + (2 * cc + 1) * 32, vb);
          l_run += chunk_probs(va, full, k0 + (2 * cc) * 32, kv_len, qpos, 0, p.scale_log2, base,
                               prow + cc * (BM * 128), 0, r);
          tmem_ld_wait();
          if (cc == 0) tmem_ld_32x32(tS + 64, va);
          l_run += chunk_probs(vb, full, k0 + (2 * cc + 1) * 32, kv_len, qpos, 0, p.scale_log2, base,
                               prow + cc * (BM * 128), 4, r);
        }
      }
      tc_fence_before();
      mbar_arrive(s_empty);
      fence_proxy_async_smem();
      mbar_arrive(p_full);
    }
      // ---- epilogue: 72 of the 80 O columns ----
      mbar_wait(p_empty, (g - 1u) & 1u);
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      __nv_bfloat16* orow = p.out + (static_cast<long long>(row0) + q0 + r) * (static_cast<long long>(H) * HD) + head * HD;
      auto st8 = [&](const uint32_t* o, int col) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
        *reinterpret_cast<uint4*>(orow + col) = w;
      };
      uint32_t o0[32], o1[32], o2[16];
      tmem_ld_32x32(tO, o0);
      tmem_ld_32x32(tO + 32, o1);
      tmem_ld_32x16(tO + 64, o2);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free);                                       // the next item's first P V may overwrite O
      if (row_ok) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) { st8(o0 + 8 * g8, 8 * g8); st8(o1 + 8 * g8, 32 + 8 * g8); }
        st8(o2, 64);                                             // dims 64..71; 72..79 are padding
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, fv::kTmemCols);
  }
}

void timeline_install_attention_tc(const Timeline& t) { timeline_install(t); }

int vit_attention_tc(const __nv_bfloat16* qkv, int n_crops, int seq, int n_heads, __nv_bfloat16* out,
                     cudaStream_t stream) {
  if (n_crops <= 0) return set_error("vit_attention: empty batch");
  const long long D = static_cast<long long>(n_heads) * 72;
  const long long T = static_cast<long long>(n_crops) * seq;
  CUtensorMap t64, t16;
  // view [token][3*H heads][72]: reading dims 64..79 of a head zero-fills 72..79
  if (make_tmap_bf16_3d(&t64, qkv, 72, 3LL * n_heads, T, 144, 3 * D * 2, 64, 1, 128, 128)) return 1;
  if (make_tmap_bf16_3d(&t16, qkv, 72, 3LL * n_heads, T, 144, 3 * D * 2, 16, 1, 128, 32)) return 1;
  static DeviceOnce configured;
  if (configured.first()) {
    for (auto* fn : {fa_tc_vit_kernel<true>, fa_tc_vit_kernel<false>}) {
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, fv::kSmemTotal);
      if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
      e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    }
  }
  FaVitParams p{};
  p.seq = seq; p.n_heads = n_heads; p.n_crops = n_crops; p.out = out;
  p.scale_log2 = (1.0f / sqrtf(72.0f)) * 1.4426950408889634f;
  const long long items = static_cast<long long>((seq + fv::BM - 1) / fv::BM) * n_heads * n_crops;
  if (items >= (1LL << 31)) return set_error("vit_attention: too many work items");
  const long long resident = 2LL * num_sms();                   // persistent: two CTAs per SM walk the items
  dim3 grid(static_cast<unsigned>(g_attention_impl == 4 || items < resident ? items : resident));
  count_launch();
  const cudaError_t e = g_attention_impl == 2
      ? launch_k(fa_tc_vit_kernel<false>, grid, dim3(fv::kThreads), fv::kSmemTotal, stream, t64, t16, p)
      : launch_k(fa_tc_vit_kernel<true>, grid, dim3(fv::kThreads), fv::kSmemTotal, stream, t64, t16, p);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  return 0;
}

}  // namespace md
