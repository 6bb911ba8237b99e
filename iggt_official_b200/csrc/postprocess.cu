// GPU post-processing of the predictions: pose encoding -> extrinsics / intrinsics and depth
// un-projection to world points, which the reference does on the host after a full D2H copy
// (demo.py:340-355; iggt/utils/pose_enc.py:65-130, iggt/utils/rotation.py:14-44, iggt/utils/geometry.py:151-300).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/iggt_b200.h"

namespace iggt {

// One thread per view: pose_enc [n, 9] = (T xyz, quaternion xyzw (scalar last), fov_h, fov_w)
//   extrinsics [n, 3, 4] = [R | T] (camera from world), intrinsics [n, 3, 3] with the principal point at (W/2, H/2).
__global__ void pose_to_cameras_kernel(const float* __restrict__ pose, float* __restrict__ extr,
                                       float* __restrict__ intr, int n, float H, float W) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const float* p = pose + v * 9;
  const float i = p[3], j = p[4], k = p[5], r = p[6];
  const float two_s = 2.0f / (i * i + j * j + k * k + r * r);
  float R[9];
  R[0] = 1.f - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r); R[2] = two_s * (i * k + j * r);
  R[3] = two_s * (i * j + k * r); R[4] = 1.f - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
  R[6] = two_s * (i * k - j * r); R[7] = two_s * (j * k + i * r); R[8] = 1.f - two_s * (i * i + j * j);
  float* e = extr + v * 12;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    e[a * 4 + 0] = R[a * 3 + 0]; e[a * 4 + 1] = R[a * 3 + 1]; e[a * 4 + 2] = R[a * 3 + 2]; e[a * 4 + 3] = p[a];
  }
  if (intr) {
    float* q = intr + v * 9;
    const float fy = (H * 0.5f) / tanf(p[7] * 0.5f);
    const float fx = (W * 0.5f) / tanf(p[8] * 0.5f);
    q[0] = fx; q[1] = 0.f; q[2] = W * 0.5f;
    q[3] = 0.f; q[4] = fy; q[5] = H * 0.5f;
    q[6] = 0.f; q[7] = 0.f; q[8] = 1.f;
  }
}

// depth [n, H, W] -> world points [n, H, W, 3]:  X_cam = ((u - cu) d / fu, (v - cv) d / fv, d),
// X_world = R^T (X_cam - t)  (closed-form inverse of the camera-from-world extrinsic), mask = eps < d < z_far.
__global__ void __launch_bounds__(256)
unproject_depth_kernel(const float* __restrict__ depth, const float* __restrict__ extr, const float* __restrict__ intr,
                       float* __restrict__ world, uint8_t* __restrict__ mask, int H, int W, float eps, float z_far) {
  __shared__ float cam[21];
  const int v = blockIdx.y;
  if (threadIdx.x < 12) cam[threadIdx.x] = extr[v * 12 + threadIdx.x];
  else if (threadIdx.x < 21) cam[threadIdx.x] = intr[v * 9 + threadIdx.x - 12];
  __syncthreads();
  const float fu = cam[12], fv = cam[16], cu = cam[14], cv = cam[17];
  const int hw = H * W;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += gridDim.x * blockDim.x) {
    const float d = depth[static_cast<int64_t>(v) * hw + pix];
    const float u = static_cast<float>(pix % W), vv = static_cast<float>(pix / W);
    const float x = (u - cu) * d / fu - cam[3];
    const float y = (vv - cv) * d / fv - cam[7];
    const float z = d - cam[11];
    float* o = world + (static_cast<int64_t>(v) * hw + pix) * 3;
    o[0] = cam[0] * x + cam[4] * y + cam[8] * z;     // R^T row 0
    o[1] = cam[1] * x + cam[5] * y + cam[9] * z;
    o[2] = cam[2] * x + cam[6] * y + cam[10] * z;
    if (mask) mask[static_cast<int64_t>(v) * hw + pix] = (d > eps && (z_far <= 0.f || d < z_far)) ? 1 : 0;
  }
}

}  // namespace iggt

using namespace iggt;

extern "C" int iggt_pose_to_cameras(const float* pose_enc, float* extrinsics, float* intrinsics, int n, int H, int W,
                                    iggt_stream_t stream) {
  if (n <= 0 || !pose_enc || !extrinsics) return -1;
  pose_to_cameras_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pose_enc, extrinsics, intrinsics, n,
                                                                          (float)H, (float)W);
  return (int)cudaGetLastError();
}

extern "C" int iggt_unproject_depth(const float* depth, const float* extrinsics, const float* intrinsics, float* world,
                                    uint8_t* mask, int n, int H, int W, float eps, float z_far, iggt_stream_t stream) {
  if (n <= 0 || H <= 0 || W <= 0 || !depth || !extrinsics || !intrinsics || !world) return -1;
  int gx = (H * W + 255) / 256;
  if (gx > 592) gx = 592;
  unproject_depth_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(depth, extrinsics, intrinsics, world, mask, H, W,
                                                                       eps, z_far);
  return (int)cudaGetLastError();
}
