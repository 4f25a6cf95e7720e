// Weight-only quantised weight stream for the decode-time Linear layers: int4 with
// group-128 scale / zero point — the reference's QuantizedLinear checkpoint format (moondream/torch/layers.py:38-110),
// re-laid out at load time — and int8 (BASELINE.json config 5).
//
// The reference dequantises once (`dequantize_tensor`, layers.py:38-44):
//     W = bf16( bf16( nibble - zero_point ) * scale )                     (group = 128 consecutive input features)
// and runs its bf16 Linear on W.  These kernels keep the packed bytes in HBM (a quarter / half of the bf16 traffic and
// footprint), rebuild exactly that bf16 W tile by tile in shared memory and feed the same tcgen05 MMAs as the bf16
// stream (gemm_tcgen05.cu: smallbatch_gemm_kernel), with the same split-K plan and summation order — so the fp32
// partial sums equal the bf16 stream's on the dequantised matrix bit for bit (tests/test_quant_gpu.py).
//
// Stream layout (moondream_b200/quant.py converts the reference layout):
//     int4: wq[n][k / 2]  low nibble = input feature 2j, high nibble = 2j + 1 (unsigned 0..15)
//     int8: wq[n][k]      signed bytes
//     scale[n][k / 128], zero[n][k / 128] fp32  (int8: zero = 0 and one scale per row, repeated per group)
//
//   warp 0      : TMA producer (packed weight tile -> staging ring before the dependency wait; activations after)
//   warp 1      : MMA issuer   (M = 64 / 128 activation lanes x tile rows x 16)
//   warp 2      : TMEM allocator
//   warps 4-11  : dequantisers (staging -> bf16 -> 128B-swizzled B tile, generic-proxy stores + fence.proxy.async);
//                 warps 4-7 then drain the accumulator to the fp32 partial-sum workspace
//
// The MMA operand reads bound this stream exactly like the bf16 one, so it is not faster than
// bf16 at the HBM roofline; what it buys is the footprint and three quarters of the decode weight traffic.
#include "kernels.cuh"
#include "ptx.cuh"

namespace md {

namespace wq {
constexpr int BK = 64;
constexpr int kThreads = 384;
constexpr int kWorkers = 256;             // warps 4..11
constexpr int kMaxStages = 8;
constexpr int kTailPad = 128 * BK * 2;    // the A descriptor spans up to 128 rows whatever the batch
}  // namespace wq

struct QuantStreamParams {
  int batch, batch_total, n_out, k_blocks;
  int tile_n, n_mma, k_splits;
  int kb_per_split, seg_splits, seg_kb, kb_per_split2;
  int stages, a_bytes, b_bytes, q_bytes, stage_bytes;
  int groups, max_groups;                  // scale / zero row pitch (K / 128); groups one split can touch
  int tmem_cols;
  float* ws;
  const float* scale;
  const float* zero;
};

// 8 packed weights -> 8 fp32 values (exact)
template <int BITS>
__device__ __forceinline__ void unpack8(const uint8_t* src, float (&v)[8]) {
  if (BITS == 4) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(src);
#pragma unroll
    for (int j = 0; j < 8; ++j)       // 2^23 + nibble is exact in fp32; so is the difference
      v[j] = __uint_as_float(0x4B000000u | ((w >> (4 * j)) & 15u)) - 8388608.f;
  } else {
    const uint2 w = *reinterpret_cast<const uint2*>(src);
    const uint32_t ws2[2] = {w.x ^ 0x80808080u, w.y ^ 0x80808080u};     // signed byte + 128
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] = __uint_as_float(0x4B000000u | ((ws2[j >> 2] >> (8 * (j & 3))) & 255u)) - 8388736.f;   // 2^23 + 128
  }
}

// the reference's dequantisation of one value (layers.py:42-43: `W_r.sub_(zero).mul_(scale)` on a bf16 tensor with
// fp32 parameters: each step is evaluated in fp32 and rounded to bf16)
__device__ __forceinline__ float dequant1(float v, float zero, float scale) {
  return __fmul_rn(bf16_round(__fsub_rn(v, zero)), scale);      // the caller rounds the product to bf16 when packing
}

template <int MROWS, int BITS>
__global__ void __launch_bounds__(wq::kThreads, 1)
smallbatch_gemm_quant_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmQ,
                             const QuantStreamParams p) {
  using namespace wq;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  // [stage: A | B | Q] x stages | tail pad | scale table | zero table | barriers
  float* s_scale = reinterpret_cast<float*>(smem + p.stages * p.stage_bytes + kTailPad);
  float* s_zero = s_scale + p.tile_n * p.max_groups;
  uint64_t* qfull = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(s_zero + p.tile_n * p.max_groups) + 7) & ~static_cast<uintptr_t>(7));
  uint64_t* afull = qfull + kMaxStages;
  uint64_t* bready = afull + kMaxStages;
  uint64_t* empty_bar = bready + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmQ);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&qfull[i], 1);
      mbar_init(&afull[i], 1);
      mbar_init(&bready[i], kWorkers);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }

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
  const int g0 = kb0 >> 1;                       // first quantisation group of this split (128 = two k-blocks)

  if (warp >= 4) {
    // scale / zero-point table of this (tile, split): constants, so no dependency wait
    const int wt = threadIdx.x - 128;
    for (int i = wt; i < p.tile_n * p.max_groups; i += kWorkers) {
      const int row = i / p.max_groups, g = i - row * p.max_groups;
      const int n = min(n0 + row, p.n_out - 1), gg = min(g0 + g, p.groups - 1);
      s_scale[i] = p.scale[static_cast<long long>(n) * p.groups + gg];
      s_zero[i] = p.zero[static_cast<long long>(n) * p.groups + gg];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      const int pre = min(p.stages, nk);
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&qfull[i], static_cast<uint32_t>(p.q_bytes));
        tma_load_2d(smem + i * p.stage_bytes + p.a_bytes + p.b_bytes, &tmQ, &qfull[i], (kb0 + i) * (BK * BITS / 8), n0);
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nk; ++j) {
        uint8_t* sx = smem + stage * p.stage_bytes;
        if (j >= pre) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&qfull[stage], static_cast<uint32_t>(p.q_bytes));
          tma_load_2d(sx + p.a_bytes + p.b_bytes, &tmQ, &qfull[stage], (kb0 + j) * (BK * BITS / 8), n0);
        }
        mbar_arrive_expect_tx(&afull[stage], static_cast<uint32_t>(p.a_bytes));
        tma_load_2d(sx, &tmX, &afull[stage], (kb0 + j) * BK, 0);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16_f32(MROWS, p.n_mma);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nk; ++j) {
        mbar_wait(&afull[stage], phase);
        mbar_wait(&bready[stage], phase);
        tc_fence_after();
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
    // ------------------------------ dequantisers ------------------------------
    const int wt = threadIdx.x - 128;
    const int c = wt & 7;                            // 8 weights = one 16-byte chunk of the 128-byte B row
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < nk; ++j) {
      const int gl = ((kb0 + j) >> 1) - g0;
      uint8_t* sb = smem + stage * p.stage_bytes + p.a_bytes;
      const uint8_t* sq = sb + p.b_bytes;
      mbar_wait(&qfull[stage], phase);
      // 8 lanes per row: a warp reads 4 x 32 (or 4 x 64) contiguous staging bytes and writes 4 full 128-byte lines
      for (int row = wt >> 3; row < p.tile_n; row += kWorkers / 8) {
        const float sc = s_scale[row * p.max_groups + gl], z = s_zero[row * p.max_groups + gl];
        float v[8];
        unpack8<BITS>(sq + row * (BK * BITS / 8) + c * BITS, v);
        uint4 o;
        o.x = pack_bf16x2(dequant1(v[0], z, sc), dequant1(v[1], z, sc));
        o.y = pack_bf16x2(dequant1(v[2], z, sc), dequant1(v[3], z, sc));
        o.z = pack_bf16x2(dequant1(v[4], z, sc), dequant1(v[5], z, sc));
        o.w = pack_bf16x2(dequant1(v[6], z, sc), dequant1(v[7], z, sc));
        *reinterpret_cast<uint4*>(sb + row * 128 + ((c ^ (row & 7)) << 4)) = o;
      }
      fence_proxy_async_smem();                      // generic-proxy stores -> visible to the tensor core's reads
      mbar_arrive(&bready[stage]);
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    // ------------------------------ accumulator -> fp32 partial sums (warps 4..7) ------------------------------
    if (warp < 8) {
      constexpr int kRowsPerWarp = MROWS / 4;
      const int q = warp - 4;
      pdl_wait();                              // ws is still being read by the predecessor's consumers
      if (q * kRowsPerWarp < p.batch) {
        const int b = (lane < kRowsPerWarp) ? q * kRowsPerWarp + lane : p.batch;   // surplus lanes store nothing
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        float* dst = p.ws + (static_cast<long long>(split) * p.batch_total + b) * p.n_out + n0;
        const bool vec = ((n0 | p.n_out) & 3) == 0;
        const int n_valid = min(p.tile_n, p.n_out - n0);
        for (int cc = 0; cc * 32 < n_valid; ++cc) {
          uint32_t acc[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(cc * 32), acc);
          tmem_ld_wait();
          if (b < p.batch) {
            if (vec && cc * 32 + 32 <= n_valid) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                *reinterpret_cast<float4*>(dst + cc * 32 + g * 4) =
                    make_float4(__uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]),
                                __uint_as_float(acc[4 * g + 2]), __uint_as_float(acc[4 * g + 3]));
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj)
                if (cc * 32 + jj < n_valid) dst[cc * 32 + jj] = __uint_as_float(acc[jj]);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// W[n][k] = bf16(bf16(q - zero) * scale) for a whole matrix (prefill runs its row-form GEMMs on this scratch copy)
template <int BITS>
__global__ void dequant_weights_kernel(const uint8_t* __restrict__ q, const float* __restrict__ scale,
                                       const float* __restrict__ zero, int N, int K, __nv_bfloat16* __restrict__ out,
                                       long long ldo) {
  pdl_launch_dependents();
  pdl_wait();                                    // the scratch may still be read by the previous block's GEMMs
  const long long chunks = static_cast<long long>(N) * (K >> 3);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < chunks;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / (K >> 3)), c = static_cast<int>(i - static_cast<long long>(n) * (K >> 3));
    const float sc = scale[static_cast<long long>(n) * (K >> 7) + (c >> 4)];
    const float z = zero[static_cast<long long>(n) * (K >> 7) + (c >> 4)];
    float v[8];
    unpack8<BITS>(q + (static_cast<long long>(n) * K * BITS >> 3) + c * BITS, v);
    uint4 o;
    o.x = pack_bf16x2(dequant1(v[0], z, sc), dequant1(v[1], z, sc));
    o.y = pack_bf16x2(dequant1(v[2], z, sc), dequant1(v[3], z, sc));
    o.z = pack_bf16x2(dequant1(v[4], z, sc), dequant1(v[5], z, sc));
    o.w = pack_bf16x2(dequant1(v[6], z, sc), dequant1(v[7], z, sc));
    *reinterpret_cast<uint4*>(out + static_cast<long long>(n) * ldo + c * 8) = o;
  }
}

int dequant_weights(int bits, const uint8_t* q, const float* scale, const float* zero, int N, int K,
                    __nv_bfloat16* out, long long ldo, cudaStream_t stream) {
  if (bits != 4 && bits != 8) return set_error("dequant_weights: bits must be 4 or 8");
  if (N <= 0 || K <= 0 || K % 128) return set_error("dequant_weights: K must be a positive multiple of the group size 128");
  if (ldo % 8 || (reinterpret_cast<uintptr_t>(out) & 15)) return set_error("dequant_weights: output must be 16-byte aligned with ldo % 8 == 0");
  const long long chunks = static_cast<long long>(N) * (K >> 3);
  const int block = 256;
  long long grid = (chunks + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (grid > cap) grid = cap;
  count_launch();
  cudaError_t e = bits == 4
      ? launch_k(dequant_weights_kernel<4>, dim3(static_cast<unsigned>(grid)), dim3(block), 0, stream, q, scale, zero, N, K, out, ldo)
      : launch_k(dequant_weights_kernel<8>, dim3(static_cast<unsigned>(grid)), dim3(block), 0, stream, q, scale, zero, N, K, out, ldo);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  return 0;
}

// Same contract as gemm_smallbatch / gemm_smallbatch_2seg (gemm_tcgen05.cu) with the weights given as packed bytes:
// ws[split][batch][n_out] fp32 partial sums under the same plan (seg_K > 0: W = [A | B] along K, no split straddles
// the boundary).  Returns the number of splits, -1 on error.
int gemm_smallbatch_quant(int bits, const uint8_t* Wq, const float* scale, const float* zero, const __nv_bfloat16* X,
                          long long ldx, int n_out, int batch, int K, int seg_K, float* ws, cudaStream_t stream) {
  using namespace wq;
  if (bits != 4 && bits != 8) { set_error("quantised stream: bits must be 4 or 8"); return -1; }
  if (n_out <= 0 || batch <= 0 || K <= 0 || K % 128) { set_error("quantised stream: K must be a positive multiple of 128"); return -1; }
  if (seg_K < 0 || seg_K >= K || seg_K % 128) { set_error("quantised stream: the segment boundary must be a multiple of 128 inside K"); return -1; }
  const bool m64 = batch <= 64;
  QuantStreamParams p{};
  p.batch_total = batch; p.n_out = n_out;
  p.k_blocks = K / BK;
  p.groups = K / 128;
  int kb_max;
  if (seg_K > 0) {
    const StreamPlan2 pl = plan_smallbatch_2seg(n_out, K, seg_K);
    p.tile_n = pl.tile_rows;
    p.kb_per_split = pl.kb_a; p.seg_kb = seg_K / BK; p.seg_splits = (p.seg_kb + pl.kb_a - 1) / pl.kb_a;
    p.kb_per_split2 = pl.kb_b;
    p.k_splits = p.seg_splits + (p.k_blocks - p.seg_kb + pl.kb_b - 1) / pl.kb_b;
    kb_max = pl.kb_a > pl.kb_b ? pl.kb_a : pl.kb_b;
  } else {
    const StreamPlan pl = plan_smallbatch(n_out, K, 0, m64 ? 64 : 128);
    p.tile_n = pl.tile_rows;
    p.kb_per_split = pl.kb < 1 ? 1 : (pl.kb > p.k_blocks ? p.k_blocks : pl.kb);
    p.k_splits = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    p.seg_splits = p.k_splits; p.seg_kb = p.k_blocks; p.kb_per_split2 = p.kb_per_split;
    kb_max = p.kb_per_split;
  }
  if (p.tile_n < 1 || p.tile_n > 256) p.tile_n = 128;
  p.n_mma = (p.tile_n + 15) / 16 * 16;
  p.max_groups = kb_max / 2 + 2;                       // a split starting on an odd k-block touches one more group
  p.tmem_cols = p.n_mma <= 32 ? 32 : p.n_mma <= 64 ? 64 : p.n_mma <= 128 ? 128 : 256;
  p.scale = scale; p.zero = zero;
  const int n_tiles = (n_out + p.tile_n - 1) / p.tile_n;
  constexpr int kSmemMax = 227 * 1024 - 1024;
  static DeviceOnce configured;
  if (configured.first()) {
    cudaError_t e = cudaSuccess;
    for (auto* fn : {smallbatch_gemm_quant_kernel<64, 4>, smallbatch_gemm_quant_kernel<128, 4>,
                     smallbatch_gemm_quant_kernel<64, 8>, smallbatch_gemm_quant_kernel<128, 8>})
      if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return -1; }
  }
  CUtensorMap tQ;
  const long long row_bytes = static_cast<long long>(K) * bits / 8;
  if (make_tmap_u8_2d(&tQ, Wq, n_out, row_bytes, row_bytes, p.tile_n, BK * bits / 8)) return -1;
  for (int b0 = 0; b0 < batch; b0 += 128) {
    p.batch = batch - b0 < 128 ? batch - b0 : 128;
    const int a_rows = (p.batch + 7) / 8 * 8;
    p.a_bytes = a_rows * BK * 2;                       // a multiple of 1024: the B tile stays swizzle-aligned
    p.b_bytes = p.n_mma * BK * 2;                      // a multiple of 2048
    p.q_bytes = p.tile_n * BK * bits / 8;
    p.stage_bytes = (p.a_bytes + p.b_bytes + p.q_bytes + 1023) / 1024 * 1024;
    const int fixed = 1024 + kTailPad + 2 * p.tile_n * p.max_groups * 4 + 8 + (4 * kMaxStages + 1) * 8 + 16;
    p.stages = (kSmemMax - fixed) / p.stage_bytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    if (p.stages < 2) { set_error("quantised stream: tile does not fit shared memory"); return -1; }
    p.ws = ws + static_cast<long long>(b0) * n_out;
    CUtensorMap tX;
    if (make_tmap_bf16_2d(&tX, X + static_cast<long long>(b0) * ldx, p.batch, K, ldx, a_rows)) return -1;
    const size_t smem = static_cast<size_t>(p.stages) * p.stage_bytes + fixed;
    const dim3 grid(n_tiles * p.k_splits), block(kThreads);
    count_launch();
    cudaError_t e;
    if (bits == 4) e = m64 ? launch_k(smallbatch_gemm_quant_kernel<64, 4>, grid, block, smem, stream, tX, tQ, p)
                           : launch_k(smallbatch_gemm_quant_kernel<128, 4>, grid, block, smem, stream, tX, tQ, p);
    else e = m64 ? launch_k(smallbatch_gemm_quant_kernel<64, 8>, grid, block, smem, stream, tX, tQ, p)
                 : launch_k(smallbatch_gemm_quant_kernel<128, 8>, grid, block, smem, stream, tX, tQ, p);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return -1; }
  }
  return p.k_splits;
}

}  // namespace md
