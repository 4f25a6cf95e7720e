// Patch embedding as ONE im2col-fused tcgen05 GEMM (reference vision.py:25-41 prepare_crops' normalisation, :44-61
// create_patches, :67-68 `patch_emb(x) + pos_emb`):
//
//     x[r][n] = bf16( bf16( sum_k lut[pixel(r, k)] * W[n][k] + bias[n] ) + pos_emb[r % tokens][n] )
//
// r = (crop, patch row, patch column), k = (channel, py, px).  The [tokens, 588] patch matrix never exists in HBM: the
// A operand of the MMAs is gathered from the uint8 NHWC crops (256-entry LUT = the reference's bf16 rounding chain per
// byte, so the pixel path stays bit-exact) straight into 128B-swizzled shared-memory tiles by the CTA's own warps.
//
//   warp 0     : TMA producer for the weight tiles W[n chunk of 128][k-block of 64]  (3-stage ring, K tail zero-filled)
//   warp 1     : MMA issuer  (M = 128 patches x N = 128 x K = 16, two TMEM accumulator stages)
//   warp 2     : TMEM allocator
//   warps 4-11 : gather the 128 x 640 A tile of a row tile (10 k-blocks; columns 588..639 zero), then drain the
//                accumulators of its n chunks: + bias, round, + pos_emb, round, 16-byte stores
// Persistent over row tiles (one CTA per SM).  The accumulation order over K is the row-form GEMM's (the same K = 16
// MMAs in the same order), so the result equals patchify + gemm_rowform bit for bit (tests/test_kernels_gpu.py).
#include "kernels.cuh"
#include "ptx.cuh"

namespace md {

namespace pe {
constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 3;
constexpr int kThreads = 384;
constexpr int kWorkers = 256;                    // warps 4..11
constexpr int kMaxKBlocks = 10;                  // K <= 640 (patch 14 x 14 x 3 = 588)
constexpr int kATile = BM * BK * 2;              // 16 KB per k-block
constexpr int kWTile = BN * BK * 2;              // 16 KB
constexpr uint32_t kTmemCols = 2 * BN;
}  // namespace pe

struct PatchEmbedParams {
  const uint8_t* crops;                          // [n_crops, crop, crop, 3]
  int n_crops, crop, patch, grid, tokens;        // tokens = grid * grid
  int T, N, K, k_blocks, n_chunks, row_tiles;    // T = n_crops * tokens, N = enc_dim, K = 3 * patch^2
  const __nv_bfloat16* lut;                      // [256]
  const __nv_bfloat16* bias;                     // [N]
  const __nv_bfloat16* pos_emb;                  // [tokens, N]
  __nv_bfloat16* out;                            // [T, N]
};

__global__ void __launch_bounds__(pe::kThreads, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmW, const PatchEmbedParams p) {
  using namespace pe;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                                            // [k_blocks][128 x 64] bf16, 128B-swizzled
  uint8_t* sW = sA + kMaxKBlocks * kATile;                       // [kStages][128 x 64]
  int* s_koff = reinterpret_cast<int*>(sW + kStages * kWTile);   // [k_blocks * 64] byte offset of feature k in a patch, -1 = pad
  __nv_bfloat16* s_lut = reinterpret_cast<__nv_bfloat16*>(s_koff + kMaxKBlocks * BK);
  uint64_t* wfull = reinterpret_cast<uint64_t*>(s_lut + 256);
  uint64_t* wempty = wfull + kStages;
  uint64_t* a_ready = wempty + kStages;          // 1 (kWorkers arrivals per row tile)
  uint64_t* a_free = a_ready + 1;                // 1 (all MMAs of a row tile have read A)
  uint64_t* acc_full = a_free + 1;               // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2] (kWorkers arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) prefetch_tensormap(&tmW);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    mbar_init(a_ready, kWorkers);
    mbar_init(a_free, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kWorkers); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, kTmemCols);
    tmem_relinquish();
  }
  // feature k = (c, py, px) -> byte offset inside the patch's pixel window; constants of the launch
  for (int k = threadIdx.x; k < p.k_blocks * BK; k += kThreads) {
    int off = -1;
    if (k < p.K) {
      const int pp = p.patch * p.patch;
      const int c = k / pp, rem = k - c * pp, py = rem / p.patch, px = rem - py * p.patch;
      off = (py * p.crop + px) * 3 + c;
    }
    s_koff[k] = off;
  }
  for (int i = threadIdx.x; i < 256; i += kThreads) s_lut[i] = p.lut[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();

  const int n_my_tiles = (p.row_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ------------------------------ weight tiles (constants: no dependency wait) ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < n_my_tiles; ++t)
        for (int nc = 0; nc < p.n_chunks; ++nc)
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&wempty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&wfull[stage], kWTile);
            tma_load_2d(sW + stage * kWTile, &tmW, &wfull[stage], kb * BK, nc * BN);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;                                   // accumulator use counter
      for (int t = 0; t < n_my_tiles; ++t) {
        mbar_wait(a_ready, static_cast<uint32_t>(t & 1));
        tc_fence_after();
        for (int nc = 0; nc < p.n_chunks; ++nc, ++it) {
          const int as = it & 1;
          mbar_wait(&acc_empty[as], (static_cast<uint32_t>(it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&wfull[stage], phase);
            tc_fence_after();
            const uint64_t da = make_desc_k_sw128(smem_u32(sA + kb * kATile));
            const uint64_t db = make_desc_k_sw128(smem_u32(sW + stage * kWTile));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&wempty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&acc_full[as]);
        }
        umma_commit(a_free);                        // every MMA of this row tile has read A
      }
    }
  } else if (warp >= 4) {
    const int wt = threadIdx.x - 128;
    const int q = warp & 3;                         // TMEM lane quadrant
    const int half = (warp - 4) >> 2;               // the two warps of a quadrant take 64 columns each
    pdl_wait();                                     // crops may come from a preceding kernel (device preprocessing)
    int it = 0;
    for (int t = 0; t < n_my_tiles; ++t) {
      const int tile = static_cast<int>(blockIdx.x) + t * static_cast<int>(gridDim.x);
      const int r0 = tile * BM;
      // ---- gather: (row, 16-byte chunk of 8 features) items; a warp writes four full 128-byte lines per store ----
      if (t > 0) mbar_wait(a_free, static_cast<uint32_t>((t - 1) & 1));
      const int chunks = p.k_blocks * 8;
      for (int i = wt; i < BM * chunks; i += kWorkers) {
        const int row = i / chunks, cc = i - row * chunks;
        const int r = r0 + row;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (r < p.T) {
          const int n = r / p.tokens, tok = r - n * p.tokens, gy = tok / p.grid, gx = tok - gy * p.grid;
          const uint8_t* base = p.crops + ((static_cast<long long>(n) * p.crop + gy * p.patch) * p.crop + gx * p.patch) * 3;
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int o0 = s_koff[cc * 8 + 2 * j], o1 = s_koff[cc * 8 + 2 * j + 1];
            const uint32_t lo = o0 >= 0 ? static_cast<uint32_t>(__bfloat16_as_ushort(s_lut[base[o0]])) : 0u;
            const uint32_t hi = o1 >= 0 ? static_cast<uint32_t>(__bfloat16_as_ushort(s_lut[base[o1]])) : 0u;
            w[j] = lo | (hi << 16);
          }
          o = make_uint4(w[0], w[1], w[2], w[3]);
        }
        *reinterpret_cast<uint4*>(sA + (cc >> 3) * kATile + row * 128 + (((cc & 7) ^ (row & 7)) << 4)) = o;
      }
      fence_proxy_async_smem();
      mbar_arrive(a_ready);
      // ---- epilogue of this row tile's n chunks ----
      const int row = r0 + q * 32 + lane;
      const bool row_ok = row < p.T;
      const __nv_bfloat16* prow = p.pos_emb + static_cast<long long>(row_ok ? row % p.tokens : 0) * p.N;
      __nv_bfloat16* orow = p.out + static_cast<long long>(row) * p.N;
      for (int nc = 0; nc < p.n_chunks; ++nc, ++it) {
        const int as = it & 1;
        mbar_wait(&acc_full[as], static_cast<uint32_t>(it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN + half * 64);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), acc);
          tmem_ld_wait();
          const int col0 = nc * BN + half * 64 + c * 32;
          if (row_ok && col0 < p.N) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (col0 + g * 8 >= p.N) break;                       // N is a multiple of 8
              const uint4 bq = *reinterpret_cast<const uint4*>(p.bias + col0 + g * 8);
              const uint4 pq = *reinterpret_cast<const uint4*>(prow + col0 + g * 8);
              const uint32_t bw[4] = {bq.x, bq.y, bq.z, bq.w}, pw[4] = {pq.x, pq.y, pq.z, pq.w};
              uint32_t ow[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // the Linear output is a bf16 tensor before pos_emb is added (vision.py:67-68)
                const float v0 = bf16_round(__uint_as_float(acc[g * 8 + 2 * j]) + bf16_lo(bw[j])) + bf16_lo(pw[j]);
                const float v1 = bf16_round(__uint_as_float(acc[g * 8 + 2 * j + 1]) + bf16_hi(bw[j])) + bf16_hi(pw[j]);
                ow[j] = pack_bf16x2(v0, v1);
              }
              *reinterpret_cast<uint4*>(orow + col0 + g * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[as]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, pe::kTmemCols);
  }
}

// x [n_crops * grid^2, N] = patch_emb(patches(crops)) + pos_emb   (W: [N, k_pad] with zero columns >= 3 * patch^2)
int patch_embed_fused(const uint8_t* crops, int n_crops, int crop, int patch, const __nv_bfloat16* lut,
                      const __nv_bfloat16* W, int k_pad, const __nv_bfloat16* bias, const __nv_bfloat16* pos_emb, int N,
                      __nv_bfloat16* out, cudaStream_t stream) {
  using namespace pe;
  if (n_crops <= 0) return set_error("patch_embed: empty batch");
  const int K = 3 * patch * patch;
  if (crop % patch || k_pad < K || k_pad % 8 || N % 8) return set_error("patch_embed: bad geometry");
  const int k_blocks = (k_pad + BK - 1) / BK;
  if (k_blocks > kMaxKBlocks) return set_error("patch_embed: patch too large for the shared-memory A tile");
  PatchEmbedParams p{};
  p.crops = crops; p.n_crops = n_crops; p.crop = crop; p.patch = patch; p.grid = crop / patch;
  p.tokens = p.grid * p.grid;
  p.T = n_crops * p.tokens; p.N = N; p.K = K; p.k_blocks = k_blocks;
  p.n_chunks = (N + BN - 1) / BN;
  p.row_tiles = (p.T + BM - 1) / BM;
  p.lut = lut; p.bias = bias; p.pos_emb = pos_emb; p.out = out;
  CUtensorMap tW;
  if (make_tmap_bf16_2d(&tW, W, N, k_pad, k_pad, BN)) return 1;
  constexpr int kSmem = kMaxKBlocks * kATile + kStages * kWTile + kMaxKBlocks * BK * 4 + 512 + 16 * 8 + 16 + 1024;
  static_assert(kSmem <= 227 * 1024, "shared memory budget");
  static DeviceOnce configured;
  if (configured.first()) {
    cudaError_t e = cudaFuncSetAttribute(patch_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  }
  const int ctas = p.row_tiles < num_sms() ? p.row_tiles : num_sms();
  count_launch();
  cudaError_t e = launch_k(patch_embed_kernel, dim3(ctas), dim3(kThreads), static_cast<size_t>(kSmem), stream, tW, p);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  return 0;
}

}  // namespace md
