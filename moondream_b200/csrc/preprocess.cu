// Image preprocessing on the device: the reference's overlap_crop_image
// (moondream/torch/image_crops.py:58-167) resizes with PIL's Lanczos on one host thread per image.  Pillow's 8-bit
// resampler (libImaging/Resample.c) is integer arithmetic — per output coordinate a window of 22-bit fixed-point
// coefficients, accumulate from 2^21, shift right by 22, clamp to 0..255, horizontal pass then vertical pass with a
// uint8 image in between — so applying host-computed coefficient tables (moondream_b200/resample.py, the same
// expressions Pillow evaluates) in these kernels reproduces PIL bit for bit.  HBM-trivial byte work on CUDA cores:
// one thread per output pixel, coalesced along x.
#include "kernels.cuh"
#include "ptx.cuh"

namespace md {

constexpr int kResampleShift = 22;      // Resample.c PRECISION_BITS = 32 - 8 - 2

// axis 1 (horizontal): dst[y][xo][c] from src[y][xmin .. xmin + n)[c];  axis 0 (vertical): dst[yo][x][c] from
// src[ymin .. ymin + n)[x][c].  C = 3 interleaved channels.
template <int AXIS>
__global__ void __launch_bounds__(256)
resample_u8_kernel(const uint8_t* __restrict__ src, int in_h, int in_w, const int* __restrict__ bounds,
                   const int* __restrict__ coeffs, int ksize, int out_h, int out_w, uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= out_w) return;
  const int o = AXIS == 1 ? x : y;
  const int lo = bounds[2 * o], n = bounds[2 * o + 1];
  const int* k = coeffs + static_cast<long long>(o) * ksize;
  int s0 = 1 << (kResampleShift - 1), s1 = s0, s2 = s0;
  if (AXIS == 1) {
    const uint8_t* row = src + (static_cast<long long>(y) * in_w + lo) * 3;
    for (int t = 0; t < n; ++t) {
      const int w = k[t];
      s0 += row[3 * t] * w;
      s1 += row[3 * t + 1] * w;
      s2 += row[3 * t + 2] * w;
    }
  } else {
    const uint8_t* col = src + (static_cast<long long>(lo) * in_w + x) * 3;
    for (int t = 0; t < n; ++t) {
      const int w = k[t];
      const uint8_t* p = col + static_cast<long long>(t) * in_w * 3;
      s0 += p[0] * w;
      s1 += p[1] * w;
      s2 += p[2] * w;
    }
  }
  // clip8(): arithmetic shift, then clamp (the lookup table of Resample.c)
  uint8_t* d = dst + (static_cast<long long>(y) * out_w + x) * 3;
  d[0] = static_cast<uint8_t>(min(max(s0 >> kResampleShift, 0), 255));
  d[1] = static_cast<uint8_t>(min(max(s1 >> kResampleShift, 0), 255));
  d[2] = static_cast<uint8_t>(min(max(s2 >> kResampleShift, 0), 255));
}

int resample_u8(const uint8_t* src, int in_h, int in_w, int axis, const int* bounds, const int* coeffs, int ksize,
                int out_size, uint8_t* dst, cudaStream_t stream) {
  if (in_h <= 0 || in_w <= 0 || out_size <= 0 || ksize <= 0) return set_error("resample_u8: empty input");
  if (axis != 0 && axis != 1) return set_error("resample_u8: axis must be 0 (vertical) or 1 (horizontal)");
  const int out_h = axis == 0 ? out_size : in_h, out_w = axis == 1 ? out_size : in_w;
  const dim3 grid((out_w + 255) / 256, out_h);
  count_launch();
  if (axis == 1) resample_u8_kernel<1><<<grid, 256, 0, stream>>>(src, in_h, in_w, bounds, coeffs, ksize, out_h, out_w, dst);
  else resample_u8_kernel<0><<<grid, 256, 0, stream>>>(src, in_h, in_w, bounds, coeffs, ksize, out_h, out_w, dst);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  return 0;
}

// crops[r * cols + c] = canvas[r * stride : r * stride + crop, c * stride : c * stride + crop]
// (image_crops.py:152-165; regions past the canvas stay zero like the reference's pre-zeroed crops)
__global__ void __launch_bounds__(256)
extract_windows_kernel(const uint8_t* __restrict__ canvas, int h, int w, int cols, int stride, int crop,
                       uint8_t* __restrict__ crops) {
  const int win = blockIdx.z, y = blockIdx.y;
  const int r = win / cols, c = win % cols;
  const int sy = r * stride + y;
  uint8_t* d = crops + (static_cast<long long>(win) * crop + y) * crop * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < crop * 3; i += gridDim.x * blockDim.x) {
    const int sx = c * stride + i / 3;
    d[i] = (sy < h && sx < w) ? canvas[(static_cast<long long>(sy) * w + sx) * 3 + i % 3] : 0;
  }
}

int extract_windows_u8(const uint8_t* canvas, int h, int w, int rows, int cols, int stride, int crop, uint8_t* crops,
                       cudaStream_t stream) {
  if (rows <= 0 || cols <= 0 || crop <= 0) return set_error("extract_windows_u8: empty tiling");
  const dim3 grid((crop * 3 + 255) / 256, crop, rows * cols);
  count_launch();
  extract_windows_kernel<<<grid, 256, 0, stream>>>(canvas, h, w, cols, stride, crop, crops);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  return 0;
}

}  // namespace md
