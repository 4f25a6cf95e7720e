// On-device token selection for the decode loop, so that sampling (the reference's DEFAULT: temperature 0.5,
// top_p 0.3, moondream.py:51-52) stays inside the CUDA graph like greedy decoding does:
//   sample_top_p      softmax(logits / T) -> _apply_top_p -> one draw   (moondream.py:270-278, 312-318, 524-530)
//   embed_select      text_encoder, except rows whose token is `sel_id` take a caller-provided row
//                     (the coord_id -> encode_coordinate substitution of _generate_reasoning, moondream.py:381-391)
//   store_column_f32  dst[b, *index] = src[b]   (per-step record of the decoded coordinates)
// All HBM/latency-trivial CUDA-core work: one CTA per sequence over a 51200-entry row.
#include <math.h>

#include "kernels.cuh"
#include "ptx.cuh"

namespace md {

#define MD_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
  do {                                                                                      \
    count_launch();                                                                         \
    cudaError_t e__ = launch_k(kernel, grid, block, smem, stream, __VA_ARGS__);             \
    if (e__ != cudaSuccess) return set_error(cudaGetErrorString(e__));                      \
  } while (0)

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al.): counter-based, so a draw is a pure function of (seed, step, row) and a graph
// replay needs no RNG state besides the device-side step counter.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

constexpr int kSampleThreads = 1024;
constexpr int kProbBins = 16384;       // bf16 bit patterns of probabilities in [0, 1]: 0x0000 .. 0x3F80

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = red[0];
  for (int i = 1; i < kSampleThreads / 32; ++i) t = fmaxf(t, red[i]);
  __syncthreads();
  return t;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < kSampleThreads / 32; ++i) t += red[i];
  __syncthreads();
  return t;
}
// exclusive prefix over the block in thread order (thread 0 first); also returns the total
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* red, T* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const T n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) red[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    T w = red[lane];
    T winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const T n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    red[lane] = winc - w;                 // exclusive warp offsets
    if (lane == 31) red[32] = winc;       // grand total
  }
  __syncthreads();
  const T out = red[warp] + inc - v;
  *total = red[32];
  __syncthreads();
  return out;
}

// The reference's drop rule for the element at sorted position i (moondream.py:270-275, all in bf16):
//   c_i = bf16(sum_{k <= i} p_k) (torch's bf16 cumsum: sequential fp32 accumulation, each output rounded),
//   dropped  <=>  bf16(c_i - p_i) > bf16(top_p).
// `mass_incl` is the fp32 prefix sum including this element.
__device__ __forceinline__ bool top_p_keeps(float mass_incl, float p, float top_p_b) {
  return !(bf16_round(bf16_round(mass_incl) - p) > top_p_b);
}

// One CTA per sequence.  logits: bf16 [B, V] (already masked with -inf where the caller excludes ids).
// scratch: bf16 [B, V] (the softmax probabilities; becomes the reference's `next_probs` when keep_probs != 0).
__global__ void __launch_bounds__(kSampleThreads, 1)
sample_top_p_kernel(const __nv_bfloat16* __restrict__ logits, int V, float temperature, float top_p,
                    const unsigned long long* __restrict__ seed, const int* __restrict__ step,
                    const float* __restrict__ uniforms, __nv_bfloat16* __restrict__ scratch, int keep_probs,
                    int* __restrict__ out_ids, long long out_stride, int out_offset) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint32_t hist[];                    // [kProbBins]
  __shared__ float red_f[40];
  __shared__ int red_i[40];
  __shared__ unsigned int thr_bits_s;
  __shared__ int thr_keep_s, pick_s, last_kept_s;
  const int b = blockIdx.x, tid = threadIdx.x;
  const __nv_bfloat16* lrow = logits + static_cast<long long>(b) * V;
  __nv_bfloat16* prow = scratch + static_cast<long long>(b) * V;
  const int per = (V + kSampleThreads - 1) / kSampleThreads;
  const int i0 = min(V, tid * per), i1 = min(V, i0 + per);        // contiguous segment: index order = thread order

  for (int i = tid; i < kProbBins; i += kSampleThreads) hist[i] = 0;
  if (tid == 0) { thr_bits_s = 0; thr_keep_s = 0; pick_s = 0x7fffffff; last_kept_s = -1; }
  // ---- softmax(logits / T), computed like torch's CPU kernels: bf16 quotient, fp32 exp / sum, bf16 result ----
  float mx = -INFINITY;
  for (int i = i0; i < i1; ++i) mx = fmaxf(mx, bf16_round(__bfloat162float(lrow[i]) / temperature));
  mx = block_reduce_max(mx, red_f);
  float sum = 0.f;
  for (int i = i0; i < i1; ++i) sum += expf(bf16_round(__bfloat162float(lrow[i]) / temperature) - mx);
  sum = block_reduce_sum(sum, red_f);
  for (int i = i0; i < i1; ++i) {
    const float p = bf16_round(expf(bf16_round(__bfloat162float(lrow[i]) / temperature) - mx) / sum);
    const __nv_bfloat16 pb = __float2bfloat16_rn(p);
    prow[i] = pb;
    const int bits = min(static_cast<int>(__bfloat16_as_ushort(pb)), kProbBins - 1);
    if (bits) atomicAdd(&hist[bits], 1u);          // probability 0 never enters the nucleus (and would serialise the atomics)
  }
  __syncthreads();
  // ---- threshold search over the histogram of probability VALUES, highest first (no sort) ----
  // sorted order of the reference = probability descending, index ascending (torch.sort is stable); the kept set
  // is a prefix of that order: whole bins above a threshold value, the first `thr_keep` elements of the threshold bin.
  constexpr int kBinsPer = kProbBins / kSampleThreads;   // 16
  const int bin_hi = kProbBins - 1 - tid * kBinsPer;     // thread 0 owns the largest values
  float local = 0.f;
#pragma unroll
  for (int k = 0; k < kBinsPer; ++k) {
    const int bits = bin_hi - k;
    local += static_cast<float>(hist[bits]) * __uint_as_float(static_cast<uint32_t>(bits) << 16);
  }
  float total_mass;
  float mass = block_exclusive_scan<float>(local, red_f, &total_mass);
  const float top_p_b = bf16_round(top_p);
  int my_thr_bits = 0, my_thr_keep = 0;
  for (int k = 0; k < kBinsPer; ++k) {
    const int bits = bin_hi - k;
    const int cnt = static_cast<int>(hist[bits]);
    if (cnt == 0) continue;
    const float val = __uint_as_float(static_cast<uint32_t>(bits) << 16);
    // elements j = 0 .. cnt-1 of this bin have inclusive mass `mass + (j + 1) * val`; the keep predicate is monotone
    int keep;
    if (!top_p_keeps(mass + val, val, top_p_b)) keep = 0;
    else if (top_p_keeps(mass + cnt * val, val, top_p_b)) keep = cnt;
    else {
      int lo = 0, hi = cnt - 1;                          // lo kept, hi dropped
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (top_p_keeps(mass + (mid + 1) * val, val, top_p_b)) lo = mid; else hi = mid;
      }
      keep = lo + 1;
    }
    if (keep < cnt && my_thr_bits == 0) { my_thr_bits = bits; my_thr_keep = keep; }
    mass += cnt * val;
  }
  if (my_thr_bits) atomicMax(&thr_bits_s, static_cast<unsigned int>(my_thr_bits));
  __syncthreads();
  const int thr_bits = static_cast<int>(thr_bits_s);    // 0: nothing is dropped
  if (my_thr_bits && my_thr_bits == thr_bits) thr_keep_s = my_thr_keep;
  // ---- kept set (ties inside the threshold bin go to the lowest indices) and its mass ----
  int eq = 0;
  for (int i = i0; i < i1; ++i) eq += static_cast<int>(__bfloat16_as_ushort(prow[i])) == thr_bits;
  int eq_total;
  int rank = block_exclusive_scan<int>(eq, red_i, &eq_total);
  const int thr_keep = thr_keep_s;
  float kept = 0.f;
  for (int i = i0; i < i1; ++i) {
    const int bits = __bfloat16_as_ushort(prow[i]);
    bool k = bits > thr_bits;
    if (bits == thr_bits && thr_bits != 0) k = rank++ < thr_keep;
    const float p = __bfloat162float(prow[i]);
    if (k) kept += p;
    else prow[i] = __float2bfloat16_rn(0.f);
  }
  const float S = bf16_round(block_reduce_sum(kept, red_f));      // probs_sort.sum() is a bf16 tensor
  // ---- renormalise (probs_sort.div_, bf16) and draw by inverse CDF in index order ----
  float q_local = 0.f;
  for (int i = i0; i < i1; ++i) {
    const float q = bf16_round(__bfloat162float(prow[i]) / S);
    if (keep_probs) prow[i] = __float2bfloat16_rn(q);
    q_local += q;
  }
  float q_total;
  float q_before = block_exclusive_scan<float>(q_local, red_f, &q_total);
  float u;
  if (uniforms) u = uniforms[b];
  else {
    const unsigned long long sd = seed ? *seed : 0ull;
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(b), static_cast<uint32_t>((step ? *step : 0) + out_offset), 0u, 0u),
                                  make_uint2(static_cast<uint32_t>(sd), static_cast<uint32_t>(sd >> 32)));
    u = static_cast<float>(r.x >> 8) * (1.0f / 16777216.0f);        // [0, 1)
  }
  const float target = u * q_total;
  float run = q_before;
  for (int i = i0; i < i1; ++i) {
    const float q = keep_probs ? __bfloat162float(prow[i]) : bf16_round(__bfloat162float(prow[i]) / S);
    if (q > 0.f) {
      atomicMax(&last_kept_s, i);
      if (run <= target && target < run + q) atomicMin(&pick_s, i);
    }
    run += q;
  }
  __syncthreads();
  if (tid == 0) {
    int tok = pick_s;
    if (tok == 0x7fffffff) tok = max(last_kept_s, 0);   // u * total rounded past the last kept entry
    out_ids[static_cast<long long>(b) * out_stride + (step ? *step : 0) + out_offset] = tok;
  }
}

int sample_top_p(const __nv_bfloat16* logits, int B, int V, float temperature, float top_p,
                 const unsigned long long* seed, const int* step, const float* uniforms, __nv_bfloat16* scratch,
                 int keep_probs, int* out_ids, long long out_stride, int out_offset, cudaStream_t stream) {
  if (B <= 0 || V <= 0) return set_error("sample_top_p: empty input");
  if (!(temperature > 0.f)) return set_error("sample_top_p: temperature must be > 0 (0 is the greedy argmax path)");
  if (!scratch) return set_error("sample_top_p: a [B, V] bf16 scratch buffer is required");
  constexpr int smem = kProbBins * 4;
  static DeviceOnce configured;
  if (configured.first()) {
    cudaError_t e = cudaFuncSetAttribute(sample_top_p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  }
  MD_LAUNCH(sample_top_p_kernel, dim3(B), dim3(kSampleThreads), smem, stream, logits, V, temperature, top_p, seed, step,
            uniforms, scratch, keep_probs, out_ids, out_stride, out_offset);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// text_encoder with substitution: out[r] = (ids[r] == sel_id) ? alt[r] : wte[ids[r]]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
embed_select_kernel(const int* __restrict__ ids, long long id_stride, const __nv_bfloat16* __restrict__ wte, int dim,
                    int vocab, int sel_id, const __nv_bfloat16* __restrict__ alt, long long ld_alt,
                    __nv_bfloat16* __restrict__ out, long long ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  int id = ids[row * id_stride];
  const uint4* src;
  if (id == sel_id && alt) src = reinterpret_cast<const uint4*>(alt + static_cast<long long>(row) * ld_alt);
  else {
    if (id < 0 || id >= vocab) id = 0;
    src = reinterpret_cast<const uint4*>(wte + static_cast<long long>(id) * dim);
  }
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<long long>(row) * ldo);
  for (int i = threadIdx.x; i < dim / 8; i += blockDim.x) dst[i] = src[i];
}

int embed_select(const int* ids, long long id_stride, int n, const __nv_bfloat16* wte, int dim, int vocab, int sel_id,
                 const __nv_bfloat16* alt, long long ld_alt, __nv_bfloat16* out, long long ldo, cudaStream_t stream) {
  if (n <= 0) return set_error("embed_select: empty input");
  MD_LAUNCH(embed_select_kernel, dim3(n), dim3(128), 0, stream, ids, id_stride, wte, dim, vocab, sel_id, alt, ld_alt,
            out, ldo);
  return 0;
}

__global__ void store_column_f32_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, long long stride,
                                        const int* __restrict__ index, int offset) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i * stride + (index ? *index : 0) + offset] = src[i];
}

int store_column_f32(const float* src, int n, float* dst, long long stride, const int* index, int offset,
                     cudaStream_t stream) {
  if (n <= 0) return set_error("store_column_f32: empty input");
  MD_LAUNCH(store_column_f32_kernel, dim3((n + 127) / 128), dim3(128), 0, stream, src, n, dst, stride, index, offset);
  return 0;
}

}  // namespace md
