// Flash-attention forward for head_dim 64 on sm_100a: QK^T and PV on tcgen05 tensor cores with TMEM
// accumulators, K/V tiles staged by TMA, online softmax in registers (one query row per thread).
//
// Replaces F.scaled_dot_product_attention at iggt/layers/attention.py:61-66 (non-causal, no mask,
// scale 1/sqrt(64)) for the three uses in the trunk: DINOv2 blocks and frame blocks (Lq = Lk = tokens of
// one view) and global blocks (Lk = all views' tokens; with view sharding Lq = local views' tokens and
// K/V come from the all-gathered buffer).
//
// Layout: q/k/v/o are token-major [rows, ld] 16-bit matrices, head h occupies columns [64h, 64h+64).
// Sequence s owns q/o rows [s*Lq, (s+1)*Lq) and k/v rows [s*Lk, (s+1)*Lk).
//
// CTA = 128 query rows x 1 head. Warp 0: TMA producer (Q once, K/V ring). Warp 1: MMA issuer.
// Warp 2: TMEM alloc. Warps 4-7: softmax/correction (thread = query row). S is double-buffered in
// TMEM so QK^T of tile j+1 overlaps the softmax of tile j; each P_j V_j product lands in its own TMEM
// buffer and is folded into the register accumulator with the usual exp(m_old - m_new) rescale.
#include "ptx.cuh"
#include "tmap.cuh"
#include "../../include/iggt_b200.h"

namespace iggt {

constexpr int ATT_BQ = 128;
constexpr int ATT_BK = 128;
constexpr int ATT_D = 64;
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_THREADS = 256;
constexpr int ATT_TILE_BYTES = ATT_BK * ATT_D * 2;  // 16 KB (Q, K and V tiles)
constexpr int ATT_P_BYTES = ATT_BQ * ATT_BK * 2;    // 32 KB
constexpr int ATT_SMEM = ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES) + 2 * ATT_P_BYTES + 1024 + 256;

struct AttnParams {
  int Lq, Lk, H;
  int64_t ldo;
  void* o;
  float scale_log2;  // softmax scale * log2(e)
};

template <bool BF16>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;
  uint8_t* sP = sV + ATT_KV_STAGES * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_P_BYTES);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // 3
  uint64_t* k_empty = k_full + ATT_KV_STAGES;
  uint64_t* v_full = k_empty + ATT_KV_STAGES;
  uint64_t* v_empty = v_full + ATT_KV_STAGES;
  uint64_t* s_full = v_empty + ATT_KV_STAGES;   // 2
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 2;
  uint64_t* o_full = p_empty + 2;
  uint64_t* o_empty = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int n_kv = (p.Lk + ATT_BK - 1) / ATT_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);  mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 128); mbar_init(&p_empty[i], 1);
      mbar_init(&o_full[i], 1);  mbar_init(&o_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;         // 2 x 128 columns
  const uint32_t tO = tmem_base + 256;   // 2 x 64 columns

  if (warp == 0) {
    if (lane == 0) {
      const int col = head * ATT_D;
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, col, seq * p.Lq + qt * ATT_BQ);
      int st = 0; uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        const int row = seq * p.Lk + j * ATT_BK;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], ATT_TILE_BYTES);
        tma_load_2d(sK + st * ATT_TILE_BYTES, &tmK, &k_full[st], col, row);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], ATT_TILE_BYTES);
        tma_load_2d(sV + st * ATT_TILE_BYTES, &tmV, &v_full[st], col, row);
        if (++st == ATT_KV_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(ATT_BQ, ATT_BK, BF16, false, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(ATT_BQ, ATT_D, BF16, false, true);  // V is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_pv = [&](int j, int vst, uint32_t vph) {
        const int b = j & 1;
        const uint32_t bph = (j >> 1) & 1;
        mbar_wait(&p_full[b], bph);
        mbar_wait(&v_full[vst], vph);
        mbar_wait(&o_empty[b], bph ^ 1);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(sP + b * ATT_P_BYTES);
        const uint32_t v_addr = smem_u32(sV + vst * ATT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_BK / 16; ++kk) {
          const uint64_t da = make_desc_sw128(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 1024);
          const uint64_t db = make_desc_sw128(v_addr + kk * 2048, 1024);
          umma_f16(tO + b * ATT_D, da, db, idesc_pv, kk != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[b]);
        umma_commit(&v_empty[vst]);
        umma_commit(&p_empty[b]);
      };
      mbar_wait(q_full, 0);
      int kst = 0; uint32_t kph = 0;
      int vst = 0; uint32_t vph = 0;
      for (int j = 0; j < n_kv; ++j) {
        const int b = j & 1;
        const uint32_t bph = (j >> 1) & 1;
        mbar_wait(&k_full[kst], kph);
        mbar_wait(&s_empty[b], bph ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + kst * ATT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_D / 16; ++kk) {
          const uint64_t da = make_desc_sw128(q_addr + kk * 32, 1024);
          const uint64_t db = make_desc_sw128(k_addr + kk * 32, 1024);
          umma_f16(tS + b * ATT_BK, da, db, idesc_qk, kk != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[b]);
        umma_commit(&k_empty[kst]);
        if (++kst == ATT_KV_STAGES) { kst = 0; kph ^= 1; }
        if (j > 0) {
          issue_pv(j - 1, vst, vph);
          if (++vst == ATT_KV_STAGES) { vst = 0; vph ^= 1; }
        }
      }
      issue_pv(n_kv - 1, vst, vph);
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(ew * 32) << 16;
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    float o[ATT_D];
#pragma unroll
    for (int i = 0; i < ATT_D; ++i) o[i] = 0.f;
    const float c = p.scale_log2;

    auto fold_pv = [&](int j) {
      const int b = j & 1;
      const uint32_t bph = (j >> 1) & 1;
      mbar_wait(&o_full[b], bph);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tO + lane_off + b * ATT_D, r0);
      tmem_ld_32x32(tO + lane_off + b * ATT_D + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_empty[b]);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o[i] = o[i] * alpha_prev + __uint_as_float(r0[i]);
        o[32 + i] = o[32 + i] * alpha_prev + __uint_as_float(r1[i]);
      }
    };

    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      const uint32_t bph = (j >> 1) & 1;
      mbar_wait(&s_full[b], bph);
      tc_fence_after();
      float s[ATT_BK];
      {
        uint32_t r[32];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          tmem_ld_32x32(tS + lane_off + b * ATT_BK + ch * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) s[ch * 32 + i] = __uint_as_float(r[i]);
        }
      }
      tc_fence_before();
      mbar_arrive(&s_empty[b]);
      const int valid = p.Lk - j * ATT_BK;  // keys of this tile that exist
      if (valid < ATT_BK) {
#pragma unroll
        for (int i = 0; i < ATT_BK; ++i) if (i >= valid) s[i] = -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < ATT_BK; ++i) mx = fmaxf(mx, s[i]);
      const float m_new = fmaxf(m, mx);
      const float alpha = exp2f((m - m_new) * c);
      const float mc = m_new * c;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < ATT_BK; ++i) { s[i] = exp2f(fmaf(s[i], c, -mc)); sum += s[i]; }
      l = l * alpha + sum;
      m = m_new;
      // P_j -> smem (K-major, 128B swizzle, two 64-key atoms)
      mbar_wait(&p_empty[b], bph ^ 1);
      uint8_t* pb = sP + b * ATT_P_BYTES + row * 128;
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        uint4 u;
        u.x = pack16x2<BF16>(s[ch * 8 + 0], s[ch * 8 + 1]);
        u.y = pack16x2<BF16>(s[ch * 8 + 2], s[ch * 8 + 3]);
        u.z = pack16x2<BF16>(s[ch * 8 + 4], s[ch * 8 + 5]);
        u.w = pack16x2<BF16>(s[ch * 8 + 6], s[ch * 8 + 7]);
        *reinterpret_cast<uint4*>(pb + (ch >> 3) * 16384 + (((ch & 7) ^ (row & 7)) << 4)) = u;
      }
      fence_proxy_async_smem();
      mbar_arrive(&p_full[b]);
      if (j > 0) fold_pv(j - 1);
      alpha_prev = alpha;
    }
    fold_pv(n_kv - 1);
    const int qrow = qt * ATT_BQ + row;
    if (qrow < p.Lq) {
      const float inv = 1.0f / l;
      uint16_t* dst = reinterpret_cast<uint16_t*>(p.o) + (static_cast<int64_t>(seq) * p.Lq + qrow) * p.ldo +
                      head * ATT_D;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 u;
        u.x = pack16x2<BF16>(o[ch * 8 + 0] * inv, o[ch * 8 + 1] * inv);
        u.y = pack16x2<BF16>(o[ch * 8 + 2] * inv, o[ch * 8 + 3] * inv);
        u.z = pack16x2<BF16>(o[ch * 8 + 4] * inv, o[ch * 8 + 5] * inv);
        u.w = pack16x2<BF16>(o[ch * 8 + 6] * inv, o[ch * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + ch * 8) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <bool BF16>
int launch_attention(const CUtensorMap& tQ, const CUtensorMap& tK, const CUtensorMap& tV,
                     const AttnParams& p, int num_seq, cudaStream_t stream) {
  auto kern = attention_fwd_kernel<BF16>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dim3 grid((p.Lq + ATT_BQ - 1) / ATT_BQ, p.H, num_seq);
  kern<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tQ, tK, tV, p);
  return (int)cudaGetLastError();
}

}  // namespace iggt

using namespace iggt;

extern "C" int iggt_attention_fwd_v1(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                     int64_t ldv, void* o, int64_t ldo, int num_seq, int Lq, int Lk,
                                     int H, int head_dim, float scale, int dtype, iggt_stream_t stream) {
  if (head_dim != 64) return -1;
  if (num_seq <= 0 || Lq <= 0 || Lk <= 0 || H <= 0) return -1;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return -2;
  if (dtype != 0 && dtype != 1) return -3;
  const TmDtype dt = dtype ? TM_BF16 : TM_F16;
  CUtensorMap tQ, tK, tV;
  if (make_tmap_2d(&tQ, dt, q, (uint64_t)num_seq * Lq, (uint64_t)H * 64, ldq, 64, ATT_BQ)) return -4;
  if (make_tmap_2d(&tK, dt, k, (uint64_t)num_seq * Lk, (uint64_t)H * 64, ldk, 64, ATT_BK)) return -4;
  if (make_tmap_2d(&tV, dt, v, (uint64_t)num_seq * Lk, (uint64_t)H * 64, ldv, 64, ATT_BK)) return -4;
  AttnParams p;
  p.Lq = Lq; p.Lk = Lk; p.H = H; p.ldo = ldo; p.o = o;
  p.scale_log2 = scale * 1.4426950408889634f;
  return dtype ? launch_attention<true>(tQ, tK, tV, p, num_seq, (cudaStream_t)stream)
               : launch_attention<false>(tQ, tK, tV, p, num_seq, (cudaStream_t)stream);
}
