// Bandwidth-bound kernels around the GEMMs and attention: LayerNorm, patch extraction (+ pixel
// normalisation), crop-feature stitching + adaptive average pooling, token embedding, partial RoPE +
// paged KV-cache write, greedy argmax over the vocabulary, Fourier features of the region head.
#include <math.h>

#include "decode_epilogue.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace md {

#define MD_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
  do {                                                                                      \
    count_launch();                                                                         \
    cudaError_t e__ = launch_k(kernel, grid, block, smem, stream, __VA_ARGS__);             \
    if (e__ != cudaSuccess) return set_error(cudaGetErrorString(e__));                      \
  } while (0)

#define MD_CHECK_LAUNCH()                                        \
  do {                                                           \
    count_launch();                                              \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return set_error(cudaGetErrorString(e__)); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// LayerNorm with bias, eps 1e-5, fp32 statistics (reference layers.py:118-119 -> F.layer_norm).
// One warp per row; the row lives in registers between the mean, variance and normalise passes.
// The output form (x * rstd + (-mean * rstd)) * w + b mirrors ATen's CPU kernel, the oracle's.
// ------------------------------------------------------------------------------------------------
constexpr int kLnMaxChunks = 16;   // per lane, 8 elements each -> dim <= 4096

// CH = 16-byte chunks per lane this instantiation keeps in registers (dim <= CH * 256): sizing it to the row
// width instead of the 4096 maximum keeps the register count low enough for full occupancy.
template <int CH>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                 const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ b,
                 __nv_bfloat16* __restrict__ y, long long ldy, int rows, int dim, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int chunks = dim >> 3;
  const __nv_bfloat16* xr = x + static_cast<long long>(row) * ldx;
  uint4 v[CH];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + i * 32;
    if (c < chunks) {
      v[i] = *reinterpret_cast<const uint4*>(xr + c * 8);
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) sum += bf16_lo(u[j]) + bf16_hi(u[j]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / dim;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + i * 32;
    if (c < chunks) {
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = bf16_lo(u[j]) - mean, c2 = bf16_hi(u[j]) - mean;
        sq += a * a + c2 * c2;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = 1.0f / sqrtf(sq / dim + eps);
  const float shift = -rstd * mean;
  __nv_bfloat16* yr = y + static_cast<long long>(row) * ldy;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + i * 32;
    if (c < chunks) {
      const uint4 wq = *reinterpret_cast<const uint4*>(w + c * 8);
      const uint4 bq = *reinterpret_cast<const uint4*>(b + c * 8);
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      const uint32_t wu[4] = {wq.x, wq.y, wq.z, wq.w};
      const uint32_t bu[4] = {bq.x, bq.y, bq.z, bq.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = (bf16_lo(u[j]) * rstd + shift) * bf16_lo(wu[j]) + bf16_lo(bu[j]);
        const float hi = (bf16_hi(u[j]) * rstd + shift) * bf16_hi(wu[j]) + bf16_hi(bu[j]);
        o[j] = pack_bf16x2(lo, hi);
      }
      *reinterpret_cast<uint4*>(yr + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

int layernorm(const __nv_bfloat16* x, long long ldx, const __nv_bfloat16* w, const __nv_bfloat16* b,
              __nv_bfloat16* y, long long ldy, int rows, int dim, float eps, cudaStream_t stream) {
  if (rows <= 0) return set_error("layernorm: empty input");
  if (dim % 8 || dim > kLnMaxChunks * 32 * 8) return set_error("layernorm: dim must be a multiple of 8 and <= 4096");
  const int need = (dim / 8 + 31) / 32;             // chunks per lane
  const dim3 grid((rows + 7) / 8), block(256);
  if (need <= 2) MD_LAUNCH(layernorm_kernel<2>, grid, block, 0, stream, x, ldx, w, b, y, ldy, rows, dim, eps);
  else if (need <= 4) MD_LAUNCH(layernorm_kernel<4>, grid, block, 0, stream, x, ldx, w, b, y, ldy, rows, dim, eps);
  else if (need <= 5) MD_LAUNCH(layernorm_kernel<5>, grid, block, 0, stream, x, ldx, w, b, y, ldy, rows, dim, eps);
  else if (need <= 8) MD_LAUNCH(layernorm_kernel<8>, grid, block, 0, stream, x, ldx, w, b, y, ldy, rows, dim, eps);
  else MD_LAUNCH(layernorm_kernel<16>, grid, block, 0, stream, x, ldx, w, b, y, ldy, rows, dim, eps);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// patchify: uint8 NHWC crops -> bf16 patch matrix [n_crops * grid^2, k_pad], feature order (c, py, px)
// (reference vision.py:25-61: prepare_crops' bf16 normalisation + create_patches).  The 256-entry
// `lut` holds the reference's rounding chain bf16(bf16(bf16(v)/255) - 0.5) * 2 computed on the host
// with the very ops the reference uses, so the pixel path is bit-exact.  One CTA = one row of patches.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
patchify_kernel(const uint8_t* __restrict__ crops, int crop, int patch, int k_pad,
                const __nv_bfloat16* __restrict__ lut, __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint8_t px_smem[];                 // [patch][crop*3] bytes
  const int grid = crop / patch;
  const int prow = blockIdx.x;                         // patch row within the crop
  const int ci = blockIdx.y;
  const int row_bytes = crop * 3;
  const uint8_t* src = crops + (static_cast<long long>(ci) * crop + prow * patch) * row_bytes;
  for (int i = threadIdx.x; i < patch * row_bytes; i += blockDim.x) px_smem[i] = src[i];
  __syncthreads();
  const int feat = 3 * patch * patch;
  __nv_bfloat16* dst = out + (static_cast<long long>(ci) * grid * grid + prow * grid) * k_pad;
  for (int i = threadIdx.x; i < grid * k_pad; i += blockDim.x) {
    const int pc = i / k_pad, f = i % k_pad;
    __nv_bfloat16 v = __float2bfloat16(0.f);
    if (f < feat) {
      const int c = f / (patch * patch), rem = f % (patch * patch);
      const int py = rem / patch, pxx = rem % patch;
      v = lut[px_smem[py * row_bytes + (pc * patch + pxx) * 3 + c]];
    }
    dst[i] = v;
  }
}

int patchify(const uint8_t* crops, int n_crops, int crop, int patch, int k_pad,
             const __nv_bfloat16* lut, __nv_bfloat16* out, cudaStream_t stream) {
  if (n_crops <= 0) return set_error("patchify: empty batch");
  if (crop % patch || k_pad < 3 * patch * patch) return set_error("patchify: bad geometry");
  const int smem = patch * crop * 3;
  dim3 grid(crop / patch, n_crops);
  patchify_kernel<<<grid, 256, smem, stream>>>(crops, crop, patch, k_pad, lut, out);
  MD_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// stitch + adaptive average pool + concat (reference image_crops.py:170-231 with patch_size=1,
// vision.py:83-88).  For image i with tiling (th, tw): the stitched map is [(g-2m)th+2m, (g-2m)tw+2m]
// cells, each cell taken from the local crop that owns it; output cell (oy, ox) of the g x g grid
// averages the stitched cells in [floor(oy*H/g), ceil((oy+1)*H/g)) x [...], fp32 sum / count.
// out[i, oy*g+ox, 0:D] = global crop features, out[..., D:2D] = pooled local features.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stitch_pool_concat_kernel(const __nv_bfloat16* __restrict__ feats, const int* __restrict__ crop_offsets,
                          const int* __restrict__ tilings, int g, int margin, int dim,
                          __nv_bfloat16* __restrict__ out) {
  const int cell = blockIdx.x, img = blockIdx.y;
  const int oy = cell / g, ox = cell % g;
  const int crop0 = crop_offsets[img];
  const int th = tilings[2 * img], tw = tilings[2 * img + 1];
  const int inner = g - 2 * margin;
  const int H = inner * th + 2 * margin, W = inner * tw + 2 * margin;
  const int y0 = (oy * H) / g, y1 = ((oy + 1) * H + g - 1) / g;
  const int x0 = (ox * W) / g, x1 = ((ox + 1) * W + g - 1) / g;
  const float inv = 1.0f / static_cast<float>((y1 - y0) * (x1 - x0));
  const long long crop_elems = static_cast<long long>(g) * g * dim;
  __nv_bfloat16* o = out + (static_cast<long long>(img) * g * g + cell) * (2 * dim);
  const __nv_bfloat16* gsrc = feats + crop0 * crop_elems + static_cast<long long>(cell) * dim;
  for (int d = threadIdx.x * 2; d < dim; d += blockDim.x * 2) {
    *reinterpret_cast<uint32_t*>(o + d) = *reinterpret_cast<const uint32_t*>(gsrc + d);
    float a0 = 0.f, a1 = 0.f;
    for (int y = y0; y < y1; ++y) {
      int ty = y < margin ? 0 : (y - margin) / inner;
      if (ty > th - 1) ty = th - 1;
      const int py = y - ty * inner;
      for (int x = x0; x < x1; ++x) {
        int tx = x < margin ? 0 : (x - margin) / inner;
        if (tx > tw - 1) tx = tw - 1;
        const int pxx = x - tx * inner;
        const __nv_bfloat16* s = feats + (crop0 + 1 + ty * tw + tx) * crop_elems +
                                 (static_cast<long long>(py) * g + pxx) * dim + d;
        const uint32_t u = *reinterpret_cast<const uint32_t*>(s);
        a0 += bf16_lo(u);
        a1 += bf16_hi(u);
      }
    }
    // ATen divides the fp32 sum by the element count
    *reinterpret_cast<uint32_t*>(o + dim + d) = pack_bf16x2(a0 / ((y1 - y0) * (x1 - x0)), a1 / ((y1 - y0) * (x1 - x0)));
  }
  (void)inv;
}

// The same pooling + concat for an ALREADY stitched map (the reference's `_vis_proj(g, r)` seam, vision.py:77-89):
// global [g*g, dim], stitched [H, W, dim] -> out [g*g, 2*dim].
__global__ void __launch_bounds__(256)
pool_concat_kernel(const __nv_bfloat16* __restrict__ global_feats, const __nv_bfloat16* __restrict__ stitched,
                   int H, int W, int g, int dim, __nv_bfloat16* __restrict__ out) {
  const int cell = blockIdx.x;
  const int oy = cell / g, ox = cell % g;
  const int y0 = (oy * H) / g, y1 = ((oy + 1) * H + g - 1) / g;
  const int x0 = (ox * W) / g, x1 = ((ox + 1) * W + g - 1) / g;
  const float cnt = static_cast<float>((y1 - y0) * (x1 - x0));
  __nv_bfloat16* o = out + static_cast<long long>(cell) * (2 * dim);
  const __nv_bfloat16* gsrc = global_feats + static_cast<long long>(cell) * dim;
  for (int d = threadIdx.x * 2; d < dim; d += blockDim.x * 2) {
    *reinterpret_cast<uint32_t*>(o + d) = *reinterpret_cast<const uint32_t*>(gsrc + d);
    float a0 = 0.f, a1 = 0.f;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(stitched + (static_cast<long long>(y) * W + x) * dim + d);
        a0 += bf16_lo(u);
        a1 += bf16_hi(u);
      }
    *reinterpret_cast<uint32_t*>(o + dim + d) = pack_bf16x2(a0 / cnt, a1 / cnt);
  }
}

int pool_concat(const __nv_bfloat16* global_feats, const __nv_bfloat16* stitched, int H, int W, int grid, int dim,
                __nv_bfloat16* out, cudaStream_t stream) {
  if (H <= 0 || W <= 0) return set_error("pool_concat: empty map");
  if (dim % 2) return set_error("pool_concat: dim must be even");
  pool_concat_kernel<<<grid * grid, 256, 0, stream>>>(global_feats, stitched, H, W, grid, dim, out);
  MD_CHECK_LAUNCH();
  return 0;
}

int stitch_pool_concat(const __nv_bfloat16* feats, const int* crop_offsets, const int* tilings,
                       int n_images, int grid, int margin, int dim, __nv_bfloat16* out,
                       cudaStream_t stream) {
  if (n_images <= 0) return set_error("stitch_pool_concat: empty batch");
  if (dim % 2) return set_error("stitch_pool_concat: dim must be even");
  dim3 g(grid * grid, n_images);
  stitch_pool_concat_kernel<<<g, 256, 0, stream>>>(feats, crop_offsets, tilings, grid, margin, dim, out);
  MD_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// token embedding gather (reference text.py:12-13).  ids may be strided (ids[i * id_stride]) so the
// decode loop can read column `step` of the [batch, max_tokens] token matrix without a copy.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
embed_kernel(const int* __restrict__ ids, long long id_stride, const __nv_bfloat16* __restrict__ wte,
             int dim, int vocab, __nv_bfloat16* __restrict__ out, long long ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  int id = ids[row * id_stride];
  if (id < 0 || id >= vocab) id = 0;
  const uint4* src = reinterpret_cast<const uint4*>(wte + static_cast<long long>(id) * dim);
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<long long>(row) * ldo);
  for (int i = threadIdx.x; i < dim / 8; i += blockDim.x) dst[i] = src[i];
}

int embed_tokens(const int* ids, long long id_stride, int n, const __nv_bfloat16* wte, int dim,
                 int vocab, __nv_bfloat16* out, long long ldo, cudaStream_t stream) {
  if (n <= 0) return set_error("embed_tokens: empty input");
  MD_LAUNCH(embed_kernel, dim3(n), dim3(128), 0, stream, ids, id_stride, wte, dim, vocab, out, ldo);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// partial RoPE + KV-cache write (reference rope.py:20-48 via text.py:42-43, moondream.py:74-78).
// First 32 of 64 dims rotate: re = x[0:16], im = x[16:32] (split-half input), output interleaved
// (re0', im0', re1', im1', ...); fp32 products and sums rounded separately (no FMA contraction) like
// the eager reference; result cast to bf16.  q goes to q_out [tokens, H*64]; k and v go to the paged
// pool at (block_table[seq][pos / 64], pos % 64).  One warp per (token, head).
//   token -> (seq, pos): prefill: seq from q_offsets (binary search), pos = start_pos[seq] + i;
//                        decode (q_offsets == nullptr): seq = token, pos = start_pos[seq].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rope_kv_write_kernel(const __nv_bfloat16* __restrict__ qkv, int n_tokens, int n_heads,
                     const int* __restrict__ q_offsets, const int* __restrict__ start_pos, int n_seqs,
                     const float* __restrict__ freqs, __nv_bfloat16* __restrict__ q_out,
                     __nv_bfloat16* __restrict__ kv_pool, int n_pages,
                     const int* __restrict__ block_tables, int max_blocks, int layer) {
  pdl_launch_dependents();
  pdl_wait();
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= n_tokens * n_heads) return;
  const int tok = gw / n_heads, head = gw % n_heads;
  int seq, pos;
  if (q_offsets) {
    int lo = 0, hi = n_seqs;                     // find seq with q_offsets[seq] <= tok < q_offsets[seq+1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (q_offsets[mid] <= tok) lo = mid; else hi = mid;
    }
    seq = lo;
    pos = start_pos[seq] + (tok - q_offsets[seq]);
  } else {
    seq = tok;
    pos = start_pos[seq];
  }
  const int D = n_heads * 64;
  const __nv_bfloat16* qr = qkv + static_cast<long long>(tok) * 3 * D + head * 64;
  const __nv_bfloat16* kr = qr + D;
  const __nv_bfloat16* vr = kr + D;
  float q0, q1, k0, k1;
  if (lane < 16) {
    const float c = freqs[(pos * 16 + lane) * 2], s = freqs[(pos * 16 + lane) * 2 + 1];
    const float qre = __bfloat162float(qr[lane]), qim = __bfloat162float(qr[16 + lane]);
    const float kre = __bfloat162float(kr[lane]), kim = __bfloat162float(kr[16 + lane]);
    q0 = __fsub_rn(__fmul_rn(qre, c), __fmul_rn(qim, s));
    q1 = __fadd_rn(__fmul_rn(qre, s), __fmul_rn(qim, c));
    k0 = __fsub_rn(__fmul_rn(kre, c), __fmul_rn(kim, s));
    k1 = __fadd_rn(__fmul_rn(kre, s), __fmul_rn(kim, c));
  } else {
    q0 = __bfloat162float(qr[2 * lane]); q1 = __bfloat162float(qr[2 * lane + 1]);
    k0 = __bfloat162float(kr[2 * lane]); k1 = __bfloat162float(kr[2 * lane + 1]);
  }
  const uint32_t vv = *reinterpret_cast<const uint32_t*>(vr + 2 * lane);
  *reinterpret_cast<uint32_t*>(q_out + static_cast<long long>(tok) * D + head * 64 + 2 * lane) = pack_bf16x2(q0, q1);
  const int page = block_tables[static_cast<long long>(seq) * max_blocks + (pos >> 6)];
  __nv_bfloat16* kdst = kv_pool +
      (((static_cast<long long>(layer) * n_pages + page) * 2) * n_heads + head) * (64 * 64) +
      (pos & 63) * 64 + 2 * lane;
  *reinterpret_cast<uint32_t*>(kdst) = pack_bf16x2(k0, k1);
  *reinterpret_cast<uint32_t*>(kdst + static_cast<long long>(n_heads) * (64 * 64)) = vv;
}

int rope_kv_write(const __nv_bfloat16* qkv, int n_tokens, int n_heads, const int* q_offsets,
                  const int* start_pos, int n_seqs, const float* freqs, __nv_bfloat16* q_out,
                  __nv_bfloat16* kv_pool, int n_pages, const int* block_tables, int max_blocks,
                  int layer, cudaStream_t stream) {
  if (n_tokens <= 0) return set_error("rope_kv_write: empty input");
  const int warps = n_tokens * n_heads;
  MD_LAUNCH(rope_kv_write_kernel, dim3((warps + 7) / 8), dim3(256), 0, stream, qkv, n_tokens, n_heads, q_offsets,
            start_pos, n_seqs, freqs, q_out, kv_pool, n_pages, block_tables, max_blocks, layer);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// greedy argmax over logits (reference text.py:163-167 + moondream.py:313-314,517-524).
// logits arrive as fp32 partial sums [splits][B][V] from the small-batch GEMM; the reference's
// logits are bf16, so values are rounded to bf16 before comparing and ties go to the lowest index
// (torch.argmax).  `mask_id` / `mask_id2` >= 0 are forced to -inf (answer_id from the 2nd generated token on,
// moondream.py:517; eos_id and size_id inside _generate_reasoning, :395-396).
// Optionally writes the top1 - top2 margin and the bf16-rounded logits.
// ------------------------------------------------------------------------------------------------
struct ArgTop {
  float best, second;
  int idx;
};
__device__ __forceinline__ void argtop_merge(ArgTop& a, float ob, float os, int oi) {
  // value descending, index ascending (torch.argmax returns the first maximum)
  if (ob > a.best || (ob == a.best && oi < a.idx)) { a.second = fmaxf(a.best, os); a.best = ob; a.idx = oi; }
  else a.second = fmaxf(a.second, ob);
}
__device__ __forceinline__ ArgTop argtop_block_reduce(ArgTop t, float* sb, float* ss, int* si) {
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, t.best, o);
    const float os = __shfl_xor_sync(0xffffffffu, t.second, o);
    const int oi = __shfl_xor_sync(0xffffffffu, t.idx, o);
    argtop_merge(t, ob, os, oi);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sb[warp] = t.best; ss[warp] = t.second; si[warp] = t.idx; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    t.best = lane < nw ? sb[lane] : -INFINITY;
    t.second = lane < nw ? ss[lane] : -INFINITY;
    t.idx = lane < nw ? si[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, t.best, o);
      const float os = __shfl_xor_sync(0xffffffffu, t.second, o);
      const int oi = __shfl_xor_sync(0xffffffffu, t.idx, o);
      argtop_merge(t, ob, os, oi);
    }
  }
  return t;   // valid in warp 0
}

// stage 1: block (part, b) scans vocabulary slice `part` of row b.  parts == 1 writes the result directly.
__global__ void __launch_bounds__(512)
argmax_partial_kernel(const float* __restrict__ ws, int splits, int B, int V, int parts,
                      const __nv_bfloat16* __restrict__ bias, int bias_period, int mask_id, int mask_id2,
                      float* __restrict__ part_best, float* __restrict__ part_second, int* __restrict__ part_idx,
                      int* __restrict__ out_ids, long long out_stride, const int* __restrict__ out_index,
                      float* __restrict__ out_margin, __nv_bfloat16* __restrict__ out_logits) {
  pdl_launch_dependents();
  pdl_wait();
  const int part = blockIdx.x, b = blockIdx.y;
  const int per = (V + parts - 1) / parts;
  const int v0 = part * per, v1 = min(V, v0 + per);
  const __nv_bfloat16* brow = bias ? bias + static_cast<long long>(b % bias_period) * V : nullptr;
  ArgTop t{-INFINITY, -INFINITY, 0x7fffffff};
  for (int v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += ws[(static_cast<long long>(s) * B + b) * V + v];
    if (brow) a += __bfloat162float(brow[v]);
    a = bf16_round(a);
    if (v == mask_id || v == mask_id2) a = -INFINITY;
    if (out_logits) out_logits[static_cast<long long>(b) * V + v] = __float2bfloat16_rn(a);
    if (a > t.best) { t.second = t.best; t.best = a; t.idx = v; }
    else if (a > t.second) t.second = a;
  }
  __shared__ float sb[32], ss[32];
  __shared__ int si[32];
  t = argtop_block_reduce(t, sb, ss, si);
  if (threadIdx.x == 0) {
    if (parts == 1) {
      const long long at = static_cast<long long>(b) * out_stride + (out_index ? *out_index : 0);
      out_ids[at] = t.idx;
      if (out_margin) out_margin[at] = t.best - t.second;
    } else {
      part_best[b * parts + part] = t.best;
      part_second[b * parts + part] = t.second;
      part_idx[b * parts + part] = t.idx;
    }
  }
}

// stage 2: one warp per row merges the slices (slices are visited in index order, ties keep the lowest)
__global__ void __launch_bounds__(128)
argmax_final_kernel(const float* __restrict__ part_best, const float* __restrict__ part_second,
                    const int* __restrict__ part_idx, int B, int parts, int* __restrict__ out_ids,
                    long long out_stride, const int* __restrict__ out_index, float* __restrict__ out_margin) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  ArgTop t{-INFINITY, -INFINITY, 0x7fffffff};
  for (int p = lane; p < parts; p += 32)
    argtop_merge(t, part_best[b * parts + p], part_second[b * parts + p], part_idx[b * parts + p]);
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, t.best, o);
    const float os = __shfl_xor_sync(0xffffffffu, t.second, o);
    const int oi = __shfl_xor_sync(0xffffffffu, t.idx, o);
    argtop_merge(t, ob, os, oi);
  }
  if (lane == 0) {
    const long long at = static_cast<long long>(b) * out_stride + (out_index ? *out_index : 0);
    out_ids[at] = t.idx;
    if (out_margin) out_margin[at] = t.best - t.second;
  }
}

constexpr int kArgmaxMaxParts = 16;
long long argmax_scratch_floats(int B) { return 3LL * B * kArgmaxMaxParts; }

int argmax_logits(const float* ws, int splits, int B, int V, const __nv_bfloat16* bias, int bias_period,
                  int mask_id, int mask_id2, int* out_ids, long long out_stride, const int* out_index,
                  float* out_margin, __nv_bfloat16* out_logits, float* scratch, cudaStream_t stream) {
  if (B <= 0 || V <= 0) return set_error("argmax_logits: empty input");
  if (bias_period < 1) bias_period = 1;
  int parts = V >= 8192 ? kArgmaxMaxParts : 1;
  if (!scratch) parts = 1;
  float* pb = scratch;
  float* ps = scratch ? scratch + 1LL * B * kArgmaxMaxParts : nullptr;
  int* pi = scratch ? reinterpret_cast<int*>(scratch + 2LL * B * kArgmaxMaxParts) : nullptr;
  MD_LAUNCH(argmax_partial_kernel, dim3(parts, B), dim3(512), 0, stream, ws, splits, B, V, parts, bias,
            bias_period, mask_id, mask_id2, pb, ps, pi, out_ids, out_stride, out_index, out_margin, out_logits);
  if (parts > 1) {
    MD_LAUNCH(argmax_final_kernel, dim3((B + 3) / 4), dim3(128), 0, stream, pb, ps, pi, B, parts, out_ids,
              out_stride, out_index, out_margin);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// decode-loop bookkeeping, row gather, region bin -> value maps
// ------------------------------------------------------------------------------------------------
__global__ void decode_advance_kernel(int* cur_tok, int* pos, int* step, const int* preds,
                                      const int* forced, long long stride, int batch, int eos_id,
                                      int* finished) {
  pdl_launch_dependents();
  pdl_wait();
  const int s = *step;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    const int t = (forced ? forced : preds)[b * stride + s + 1];
    cur_tok[b] = t;
    pos[b] += 1;
    if (finished && t == eos_id) finished[b] = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) *step = s + 1;
}

int decode_advance(int* cur_tok, int* pos, int* step, const int* preds, const int* forced,
                   long long stride, int batch, int eos_id, int* finished, cudaStream_t stream) {
  MD_LAUNCH(decode_advance_kernel, dim3(1), dim3(256), 0, stream, cur_tok, pos, step, preds, forced, stride, batch,
            eos_id, finished);
  return 0;
}

__global__ void __launch_bounds__(128)
gather_rows_kernel(const __nv_bfloat16* __restrict__ src, long long ld_src, const int* __restrict__ idx,
                   int dim, __nv_bfloat16* __restrict__ out, long long ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const uint4* s = reinterpret_cast<const uint4*>(src + static_cast<long long>(idx[row]) * ld_src);
  uint4* d = reinterpret_cast<uint4*>(out + static_cast<long long>(row) * ldo);
  for (int i = threadIdx.x; i < dim / 8; i += blockDim.x) d[i] = s[i];
}

int gather_rows(const __nv_bfloat16* src, long long ld_src, const int* idx, int n, int dim,
                __nv_bfloat16* out, long long ldo, cudaStream_t stream) {
  if (n <= 0) return set_error("gather_rows: empty input");
  if (dim % 8) return set_error("gather_rows: dim must be a multiple of 8");
  MD_LAUNCH(gather_rows_kernel, dim3(n), dim3(128), 0, stream, src, ld_src, idx, dim, out, ldo);
  return 0;
}

__global__ void bins_to_values_kernel(int which, const int* bins, int n, int n_bins, float* out) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float b = static_cast<float>(bins[i]);
  if (which == 0) out[i] = __fdiv_rn(b, static_cast<float>(n_bins));     // argmax / logits.size(-1), moondream.py:673
  else out[i] = exp2f(__fsub_rn(__fmul_rn(__fdiv_rn(b, 1023.0f), 10.0f), 10.0f));
}

int bins_to_values(int which, const int* bins, int n, int n_bins, float* out, cudaStream_t stream) {
  if (n <= 0) return set_error("bins_to_values: empty input");
  if (n_bins <= 0) return set_error("bins_to_values: n_bins must be positive");
  MD_LAUNCH(bins_to_values_kernel, dim3((n + 127) / 128), dim3(128), 0, stream, which, bins, n, n_bins, out);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Fourier features of the region head (reference region.py:12-29): f = bf16(2*pi*x) @ w (bf16 matmul,
// fp32 accumulate, bf16 result), out = [cos f, sin f] in bf16.  x: [B, n_in] fp32 values that are
// first cast to bf16 exactly like the reference casts coordinates (moondream.py:674-676).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fourier_features_kernel(const float* __restrict__ x, int n_in, const __nv_bfloat16* __restrict__ w,
                        int half, __nv_bfloat16* __restrict__ out, long long ldo) {
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    float f = 0.f;
    for (int i = 0; i < n_in; ++i) {
      // 2 * math.pi * x: python float times a bf16 tensor -> bf16 result
      const float xi = bf16_round(6.283185307179586f * bf16_round(x[b * n_in + i]));
      f = fmaf(xi, __bfloat162float(w[i * half + j]), f);
    }
    f = bf16_round(f);
    out[b * ldo + j] = __float2bfloat16_rn(cosf(f));
    out[b * ldo + half + j] = __float2bfloat16_rn(sinf(f));
  }
}

int fourier_features(const float* x, int B, int n_in, const __nv_bfloat16* w, int half,
                     __nv_bfloat16* out, long long ldo, cudaStream_t stream) {
  if (B <= 0) return set_error("fourier_features: empty input");
  fourier_features_kernel<<<B, 256, 0, stream>>>(x, n_in, w, half, out, ldo);
  MD_CHECK_LAUNCH();
  return 0;
}

}  // namespace md

namespace md {

// ------------------------------------------------------------------------------------------------
// Fused decode-step epilogues (one token per sequence).
//
// (The [qkv ; fc1] stream is finished inside decode_attention_kernel<true>, attention.cu.)
// decode_residual_ln_epilogue: finishes the K-concatenated [proj | fc2] GEMM.
//   splits [0, proj_splits) belong to proj(att), the rest to fc2(hid); each group is summed in fixed
//   order, biased and rounded to bf16 separately, then x = bf16(bf16(x + attn) + mlp) exactly as
//   text.py:158 evaluates `x + l_attn + l_mlp`; finally LayerNorm (next block's ln, or post_ln).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
decode_residual_ln_epilogue_kernel(const float* __restrict__ ws, int splits, int proj_splits, int B, int D,
                                   const __nv_bfloat16* __restrict__ bias_proj,
                                   const __nv_bfloat16* __restrict__ bias_fc2, __nv_bfloat16* __restrict__ x,
                                   const __nv_bfloat16* __restrict__ ln_w, const __nv_bfloat16* __restrict__ ln_b,
                                   __nv_bfloat16* __restrict__ ln_out, float eps) {
  const bool tl = tl_on() && threadIdx.x == 0;
  unsigned long long tl0 = 0, tl1 = 0;
  if (tl) tl0 = tl_now();
  pdl_launch_dependents();
  // the kernel is one latency chain (32 CTAs, a few loads per thread): parameters do not depend on the
  // predecessor, so their loads are issued before the dependency wait
  ResLnParams P;
  residual_ln_load_params(P, D, bias_proj, bias_fc2, ln_w, ln_b);
  pdl_wait();
  if (tl) tl1 = tl_now();
  __shared__ float red[2][8];
  residual_ln_row<false>(P, ws, splits, proj_splits, B, D, x, ln_out, eps, blockIdx.x, red);
  if (tl) tl_emit(3u << 28, tl0, tl1, tl1, 0ull, tl_now());
}

void timeline_install_elementwise(const Timeline& t) { timeline_install(t); }

// ------------------------------------------------------------------------------------------------
// Elementwise pieces of the LoRA path (the GEMM epilogues take ONE residual; a LoRA block needs two sums):
//   gelu_rows:  y = bf16(gelu_tanh(x))                                  layers.py:137 after `x0 + x1`
//   add3_rows:  y = bf16(bf16(x + a) + b)                               text.py:158 `x + l_attn + l_mlp`
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gelu_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 v = reinterpret_cast<const uint4*>(x)[i];
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(gelu_tanh(bf16_lo(w[j])), gelu_tanh(bf16_hi(w[j])));
  reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(256)
add3_rows_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ a,
                 const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 xv = reinterpret_cast<const uint4*>(x)[i], av = reinterpret_cast<const uint4*>(a)[i],
              bv = reinterpret_cast<const uint4*>(b)[i];
  const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, aw[4] = {av.x, av.y, av.z, av.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = pack_bf16x2(bf16_round(bf16_lo(xw[j]) + bf16_lo(aw[j])) + bf16_lo(bw[j]),
                       bf16_round(bf16_hi(xw[j]) + bf16_hi(aw[j])) + bf16_hi(bw[j]));
  reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

int gelu_rows(const __nv_bfloat16* x, __nv_bfloat16* y, long long n, cudaStream_t stream) {
  if (n <= 0 || n % 8) return set_error("gelu_rows: element count must be a positive multiple of 8");
  gelu_rows_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, stream>>>(x, y, n / 8);
  MD_CHECK_LAUNCH();
  return 0;
}

int add3_rows(const __nv_bfloat16* x, const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* y, long long n,
              cudaStream_t stream) {
  if (n <= 0 || n % 8) return set_error("add3_rows: element count must be a positive multiple of 8");
  add3_rows_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, stream>>>(x, a, b, y, n / 8);
  MD_CHECK_LAUNCH();
  return 0;
}

int decode_residual_ln_epilogue(const float* ws, int splits, int proj_splits, int B, int D,
                                const __nv_bfloat16* bias_proj, const __nv_bfloat16* bias_fc2,
                                __nv_bfloat16* x, const __nv_bfloat16* ln_w, const __nv_bfloat16* ln_b,
                                __nv_bfloat16* ln_out, cudaStream_t stream) {
  if (D > 4096 || D % 8) return set_error("decode epilogue: dim must be a multiple of 8 and <= 4096");
  MD_LAUNCH(decode_residual_ln_epilogue_kernel, dim3(B), dim3(256), 0, stream, ws, splits, proj_splits, B, D,
            bias_proj, bias_fc2, x, ln_w, ln_b, ln_out, 1e-5f);
  return 0;
}

}  // namespace md
