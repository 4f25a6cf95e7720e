// The row computation of the decode step's residual + LayerNorm epilogue, shared by the stand-alone kernel
// (elementwise.cu) and the tail of the [proj | fc2] weight stream (gemm_tcgen05.cu), where the last CTAs to publish
// their partial sums finish the rows themselves.  256 threads per row; contains block-wide barriers.
//   splits [0, proj_splits) belong to proj(att), the rest to fc2(hid); each group is summed in fixed order, biased and
//   rounded to bf16 separately, then x = bf16(bf16(x + attn) + mlp) exactly as text.py:158 evaluates
//   `x + l_attn + l_mlp`; finally LayerNorm (next block's ln, or post_ln).
// PARAMS_EARLY: parameters were loaded by the caller before its dependency wait (stand-alone kernel).
#pragma once
#include "kernels.cuh"
#include "ptx.cuh"

namespace md {

struct ResLnParams {
  uint4 bp[2], bf[2], wq[2], bq[2];
};

__device__ __forceinline__ void residual_ln_load_params(ResLnParams& P, int D, const __nv_bfloat16* bias_proj,
                                                        const __nv_bfloat16* bias_fc2, const __nv_bfloat16* ln_w,
                                                        const __nv_bfloat16* ln_b) {
  const int chunks = D >> 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c >= chunks) continue;
    P.bp[i] = *reinterpret_cast<const uint4*>(bias_proj + c * 8);
    P.bf[i] = *reinterpret_cast<const uint4*>(bias_fc2 + c * 8);
    P.wq[i] = *reinterpret_cast<const uint4*>(ln_w + c * 8);
    P.bq[i] = *reinterpret_cast<const uint4*>(ln_b + c * 8);
  }
}

// L2 = true: read the partial sums with ld.global.cg (written by other CTAs of the SAME grid)
template <bool L2>
__device__ __forceinline__ void residual_ln_row(const ResLnParams& P, const float* __restrict__ ws, int splits,
                                                int proj_splits, int B, int D, __nv_bfloat16* __restrict__ x,
                                                __nv_bfloat16* __restrict__ ln_out, float eps, int b, float (*red)[8]) {
  constexpr int kMaxChunks = 2;                        // 8-element chunks per thread: D <= 4096
  constexpr int kBatch = 10;                           // splits fetched per round of independent loads
  const int tid = threadIdx.x;
  const int chunks = D >> 3;
  float v[kMaxChunks][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int c = tid + i * 256;
    if (c >= chunks) continue;
    const int d0 = c * 8;
    float a[8], m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = m[k] = 0.f;
    // all loads of a thread are independent 16-byte reads: issue them together (kBatch splits per round) so
    // the L2 round trips overlap instead of forming a chain; the summation order stays split 0, 1, 2, ...
    const long long sstride = static_cast<long long>(B) * D;
    const float* base_ptr = ws + static_cast<long long>(b) * D + d0;
    const uint4 xq = *reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * D + d0);
    for (int s0 = 0; s0 < splits; s0 += kBatch) {
      float4 lo[kBatch], hi[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        if (s0 + u < splits) {
          const float4* src = reinterpret_cast<const float4*>(base_ptr + (s0 + u) * sstride);
          lo[u] = L2 ? __ldcg(src) : src[0];
          hi[u] = L2 ? __ldcg(src + 1) : src[1];
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        if (s0 + u < splits) {
          float* dst = (s0 + u) < proj_splits ? a : m;
          dst[0] += lo[u].x; dst[1] += lo[u].y; dst[2] += lo[u].z; dst[3] += lo[u].w;
          dst[4] += hi[u].x; dst[5] += hi[u].y; dst[6] += hi[u].z; dst[7] += hi[u].w;
        }
      }
    }
    const uint32_t bpw[4] = {P.bp[i].x, P.bp[i].y, P.bp[i].z, P.bp[i].w}, bfw[4] = {P.bf[i].x, P.bf[i].y, P.bf[i].z, P.bf[i].w},
                   xw[4] = {xq.x, xq.y, xq.z, xq.w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a0 = bf16_round(a[2 * k] + bf16_lo(bpw[k])), a1 = bf16_round(a[2 * k + 1] + bf16_hi(bpw[k]));
      const float m0 = bf16_round(m[2 * k] + bf16_lo(bfw[k])), m1 = bf16_round(m[2 * k + 1] + bf16_hi(bfw[k]));
      const float r0 = bf16_round(bf16_round(bf16_lo(xw[k]) + a0) + m0);
      const float r1 = bf16_round(bf16_round(bf16_hi(xw[k]) + a1) + m1);
      v[i][2 * k] = r0; v[i][2 * k + 1] = r1;
      sum += r0 + r1;
      ow[k] = pack_bf16x2(r0, r1);
    }
    *reinterpret_cast<uint4*>(x + static_cast<long long>(b) * D + d0) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
  auto block_sum = [&](float val, int which) {         // one array per reduction: a single barrier each
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if ((tid & 31) == 0) red[which][tid >> 5] = val;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[which][i];
    return t;
  };
  const float mean = block_sum(sum, 0) / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    if (tid + i * 256 >= chunks) continue;
#pragma unroll
    for (int k = 0; k < 8; ++k) sq += (v[i][k] - mean) * (v[i][k] - mean);
  }
  const float var = block_sum(sq, 1) / D;
  const float rstd = 1.0f / sqrtf(var + eps);
  const float shift = -rstd * mean;
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int c = tid + i * 256;
    if (c >= chunks) continue;
    const int d0 = c * 8;
    const uint32_t ww[4] = {P.wq[i].x, P.wq[i].y, P.wq[i].z, P.wq[i].w}, bw[4] = {P.bq[i].x, P.bq[i].y, P.bq[i].z, P.bq[i].w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float y0 = (v[i][2 * k] * rstd + shift) * bf16_lo(ww[k]) + bf16_lo(bw[k]);
      const float y1 = (v[i][2 * k + 1] * rstd + shift) * bf16_hi(ww[k]) + bf16_hi(bw[k]);
      ow[k] = pack_bf16x2(y0, y1);
    }
    *reinterpret_cast<uint4*>(ln_out + static_cast<long long>(b) * D + d0) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

}  // namespace md
