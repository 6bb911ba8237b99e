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

// cta_group::2 pairs for the 256-wide tiles; `m_sub` = 128-row sub-tiles of the problem.  IGGT_PAIR is a bit mask over
// the callers (1 residual, 2 qkv, 4 plain store, 8 convolution; 0 = one CTA per tile everywhere).  Measured on the
// C2 step (profiles/r01_ncu_notes.md): pairs win for the residual / qkv / conv epilogues (-1.4 ms per step) and lose
// ~3 % on the GELU store epilogue, whose two coupled epilogues gate the pair's next tile - hence the default 11.
enum PairUser : int { PAIR_RESID = 1, PAIR_QKV = 2, PAIR_STORE = 4, PAIR_CONV = 8 };
inline bool use_pair(int user, int bn, int m_sub) {
  static const int v = [] { const char* e = getenv("IGGT_PAIR"); return e ? atoi(e) : 11; }();
  return (v & user) != 0 && bn == 256 && m_sub >= 2;
}

// PAIR: p.num_m_tiles counts 256-row tile pairs and tB's box holds BN/2 weight rows (see gemm.cuh).
template <int BN, int EPI, bool BF16, bool CONV, bool PAIR = false, int G = (EPI == EPI_QKV ? 1 : 2)>
inline int launch_gemm_kernel(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC,
                              const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_tcgen05_kernel<BN, EPI, BF16, CONV, G, PAIR>;
  static bool configured = false;
  constexpr int smem = GemmSmem<BN, PAIR>::TOTAL;
  static_assert(smem <= 232448, "shared memory budget");
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int workers_max = PAIR ? device_sm_count() / 2 : device_sm_count();
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int workers = tiles < workers_max ? tiles : workers_max;
  if (p.stream_k) workers = workers_max;
  if (workers <= 0) return 0;
  return (int)launch_pdl_cluster(kern, dim3(workers * (PAIR ? 2 : 1)), dim3(128 + 128 * G), smem, stream,
                                 PAIR ? 2 : 1, tA, tB, tC, p);
}

}  // namespace iggt
