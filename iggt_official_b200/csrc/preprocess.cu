// GPU image pre-processing: the reference resizes every decoded view on the host with Pillow's
// 8-bit bicubic resampler and converts it with torchvision's ToTensor (iggt/utils/load_fn.py:82-83).  These two
// kernels reproduce Pillow's ImagingResample for 8-bit channels BIT-EXACTLY on the device: a horizontal pass and a
// vertical pass over fixed-point coefficient tables (22 fractional bits, built on the host in double precision the way
// Pillow's precompute_coeffs / normalize_coeffs_8bpc do), a u8 intermediate with Pillow's rounding (+2^21, >>22, clamp),
// and the vertical pass writes x / 255 as planar fp32 (ToTensor) straight into the [S, 3, H, W] batch, including the
// crop window and pad offset of load_fn.py:86-98.  HBM-bound byte work: one thread per output pixel, taps are
// contiguous bytes of a row (horizontal) or the same column of successive rows (vertical, coalesced across threads).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/iggt_b200.h"

namespace iggt {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;                       // arithmetic shift, like Pillow's clip8 lookup
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// dst[y, xx, c] = clip8(2^21 + sum_x src[y, xmin(xx) + x, c] * kk[xx, x]),  x < xmax(xx)
__global__ void __launch_bounds__(256)
resample_h_u8_kernel(const uint8_t* __restrict__ src, int64_t src_row_stride, int w_out, const int32_t* __restrict__ kk,
                     const int32_t* __restrict__ bounds, int ksize, uint8_t* __restrict__ dst) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= w_out) return;
  const int y = blockIdx.y;
  const int xmin = bounds[xx * 2], xmax = bounds[xx * 2 + 1];
  const int32_t* k = kk + static_cast<int64_t>(xx) * ksize;
  const uint8_t* s = src + y * src_row_stride + static_cast<int64_t>(xmin) * 3;
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  for (int x = 0; x < xmax; ++x) {
    const int w = k[x];
    a0 += s[x * 3 + 0] * w;
    a1 += s[x * 3 + 1] * w;
    a2 += s[x * 3 + 2] * w;
  }
  uint8_t* d = dst + (static_cast<int64_t>(y) * w_out + xx) * 3;
  d[0] = static_cast<uint8_t>(clip8(a0));
  d[1] = static_cast<uint8_t>(clip8(a1));
  d[2] = static_cast<uint8_t>(clip8(a2));
}

// out[c, r, x] = clip8(2^21 + sum_y tmp[ymin(oy0 + r) - y_shift + y, x, c] * kk[oy0 + r, y]) / 255
__global__ void __launch_bounds__(256)
resample_v_u8_f32_kernel(const uint8_t* __restrict__ tmp, int w_out, const int32_t* __restrict__ kk,
                         const int32_t* __restrict__ bounds, int ksize, int y_shift, int oy0, float* __restrict__ dst,
                         int64_t plane_stride, int64_t row_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w_out) return;
  const int r = blockIdx.y;
  const int oy = oy0 + r;
  const int ymin = bounds[oy * 2] - y_shift, ymax = bounds[oy * 2 + 1];
  const int32_t* k = kk + static_cast<int64_t>(oy) * ksize;
  const uint8_t* s = tmp + (static_cast<int64_t>(ymin) * w_out + x) * 3;
  const int64_t step = static_cast<int64_t>(w_out) * 3;
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  for (int y = 0; y < ymax; ++y, s += step) {
    const int w = k[y];
    a0 += s[0] * w;
    a1 += s[1] * w;
    a2 += s[2] * w;
  }
  float* d = dst + r * row_stride + x;
  d[0] = __fdiv_rn(static_cast<float>(clip8(a0)), 255.0f);
  d[plane_stride] = __fdiv_rn(static_cast<float>(clip8(a1)), 255.0f);
  d[2 * plane_stride] = __fdiv_rn(static_cast<float>(clip8(a2)), 255.0f);
}

}  // namespace iggt

using namespace iggt;

extern "C" int iggt_resample_h_u8(const uint8_t* src, int64_t src_row_stride, int rows, int w_out, const int32_t* kk,
                                  const int32_t* bounds, int ksize, uint8_t* dst, iggt_stream_t stream) {
  if (!src || !kk || !bounds || !dst || rows <= 0 || w_out <= 0 || ksize <= 0 || rows > 65535) return -1;
  dim3 grid((w_out + 255) / 256, rows);
  resample_h_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_row_stride, w_out, kk, bounds, ksize, dst);
  return (int)cudaGetLastError();
}

extern "C" int iggt_resample_v_u8_f32(const uint8_t* tmp, int w_out, const int32_t* kk, const int32_t* bounds,
                                      int ksize, int y_shift, int oy0, int out_rows, float* dst, int64_t plane_stride,
                                      int64_t row_stride, iggt_stream_t stream) {
  if (!tmp || !kk || !bounds || !dst || out_rows <= 0 || w_out <= 0 || ksize <= 0 || out_rows > 65535) return -1;
  dim3 grid((w_out + 255) / 256, out_rows);
  resample_v_u8_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tmp, w_out, kk, bounds, ksize, y_shift, oy0, dst,
                                                                 plane_stride, row_stride);
  return (int)cudaGetLastError();
}
