// Host-side launch helpers shared by the GEMM / conv translation units.
#pragma once
#include "gemm.cuh"
#include "tmap.cuh"

namespace iggt {

inline int device_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Pick the N tile: fewest "wave-quantised" tile-slots; ties go to the wider tile (less smem traffic
// per MAC: a 128x256 tile reads 96 B/clk of operands, 128x128 reads 128 B/clk).
inline int choose_bn(int m_tiles, int N) {
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  const int sms = device_sm_count();
  auto cost = [&](int bn) {
    long tiles = (long)m_tiles * ((N + bn - 1) / bn);
    long waves = (tiles + sms - 1) / sms;
    return waves * bn;  // time ~ waves x tile width
  };
  long c256 = cost(256), c128 = cost(128);
  return (c128 * 100 < c256 * 92) ? 128 : 256;
}

template <int BN, int EPI, bool BF16, bool CONV, int G = (EPI == EPI_QKV ? 1 : 2)>
inline int launch_gemm_kernel(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC,
                              const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_tcgen05_kernel<BN, EPI, BF16, CONV, G>;
  static bool configured = false;
  constexpr int smem = GemmSmem<BN>::TOTAL;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  if (p.stream_k) grid = device_sm_count();
  if (grid <= 0) return 0;
  return (int)launch_pdl(kern, dim3(grid), dim3(128 + 128 * G), smem, stream, tA, tB, tC, p);
}

}  // namespace iggt
