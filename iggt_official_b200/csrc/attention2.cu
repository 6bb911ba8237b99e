// Flash-attention forward v2 for head_dim 64 on sm_100a (persistent, two query tiles in flight).
//
// Each CTA loops over work items (sequence, head, pair of 128-row query tiles).  Three warpgroups:
//   WG0: warp 0 = TMA producer (Q pair once per item, K/V ring), warp 1 = tcgen05.mma issuer, warp 2 = TMEM
//        allocator; registers released with setmaxnreg.dec
//   WG1 / WG2: online softmax + output accumulation for query tile A / B (thread = query row),
//        registers raised with setmaxnreg.inc
// The two query tiles ping-pong on the tensor pipe (S_A(j+1) = Q_A K_{j+1}^T is issued while WG2 is still in
// the softmax of tile j and vice versa), so the MUFU-bound softmax of one tile hides the MMAs of the other.
// P goes through 128B-swizzled shared memory as the A operand of P V; every P_j V_j lands in its own TMEM
// buffer (double-buffered) and is folded into the register accumulator with the exp2(m_old - m_new) rescale.
//
// TMEM map (512 columns): S_A [0,128) S_B [128,256) PV_A0 [256,320) PV_A1 [320,384) PV_B0 [384,448) PV_B1 [448,512)
//
// Replaces F.scaled_dot_product_attention at iggt/layers/attention.py:61-66 (see include/iggt_b200.h).
#include <stdlib.h>
#include "ptx.cuh"
#include "tmap.cuh"
#include "../../include/iggt_b200.h"

namespace iggt {

constexpr int A2_BQ = 128;          // rows per query tile
constexpr int A2_BK = 128;          // keys per kv tile
constexpr int A2_D = 64;
constexpr int A2_STAGES = 3;
constexpr int A2_THREADS = 384;
constexpr int A2_TILE = A2_BK * A2_D * 2;     // 16 KB
constexpr int A2_P = A2_BQ * A2_BK * 2;       // 32 KB
constexpr int A2_SMEM = A2_TILE * (2 + 2 * A2_STAGES) + 2 * A2_P + 1024 + 512;

struct Attn2Params {
  int Lq, Lk, H, num_seq;
  int q_pairs;          // ceil(ceil(Lq/128)/2)
  int total_items;      // num_seq * H * q_pairs
  int64_t ldo;
  void* o;
  float scale_log2;
};

__device__ __forceinline__ void setmaxnreg_dec_56() { asm volatile("setmaxnreg.dec.sync.aligned.u32 72;"); }
__device__ __forceinline__ void setmaxnreg_inc_224() { asm volatile("setmaxnreg.inc.sync.aligned.u32 216;"); }

template <bool BF16>
__global__ void __launch_bounds__(A2_THREADS, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 2 tiles
  uint8_t* sK = sQ + 2 * A2_TILE;
  uint8_t* sV = sK + A2_STAGES * A2_TILE;
  uint8_t* sP = sV + A2_STAGES * A2_TILE;               // 2 x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * A2_P);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* q_empty = q_full + 2;          // [2]
  uint64_t* k_full = q_empty + 2;          // [3]
  uint64_t* k_empty = k_full + A2_STAGES;
  uint64_t* v_full = k_empty + A2_STAGES;
  uint64_t* v_empty = v_full + A2_STAGES;
  uint64_t* s_full = v_empty + A2_STAGES;  // [2]
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 2;
  uint64_t* o_full = p_empty + 2;          // [2 tiles][2 bufs]
  uint64_t* o_empty = o_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = (p.Lk + A2_BK - 1) / A2_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);  mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);  mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 128); mbar_init(&p_empty[i], 1);
    }
    for (int i = 0; i < A2_STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec_56();
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      int st = 0; uint32_t ph = 0;
      uint32_t item_cnt = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++item_cnt) {
        const int qp = item % p.q_pairs;
        const int head = (item / p.q_pairs) % p.H;
        const int seq = item / (p.q_pairs * p.H);
        const int col = head * A2_D;
        const uint32_t qpar = item_cnt & 1;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&q_empty[t], qpar ^ 1);
          mbar_expect_tx(&q_full[t], A2_TILE);
          tma_load_2d(sQ + t * A2_TILE, &tmQ, &q_full[t], col, seq * p.Lq + (qp * 2 + t) * A2_BQ);
        }
        for (int j = 0; j < n_kv; ++j) {
          const int row = seq * p.Lk + j * A2_BK;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_expect_tx(&k_full[st], A2_TILE);
          tma_load_2d(sK + st * A2_TILE, &tmK, &k_full[st], col, row);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_expect_tx(&v_full[st], A2_TILE);
          tma_load_2d(sV + st * A2_TILE, &tmV, &v_full[st], col, row);
          if (++st == A2_STAGES) { st = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc_qk = make_idesc_f16(A2_BQ, A2_BK, BF16, false, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(A2_BQ, A2_D, BF16, false, true);   // V is MN-major
      int kst = 0; uint32_t kph = 0;
      int vst = 0; uint32_t vph = 0;
      uint32_t s_cnt[2] = {0, 0};     // QK issues per tile  (s_empty parity)
      uint32_t pv_cnt[2] = {0, 0};    // PV issues per tile  (p_full / o buffers)
      uint32_t item_cnt = 0;
      auto issue_qk = [&](int t, uint32_t k_addr) {
        mbar_wait(&s_empty[t], (s_cnt[t] & 1) ^ 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(sQ + t * A2_TILE);
#pragma unroll
        for (int kk = 0; kk < A2_D / 16; ++kk)
          umma_f16(tmem_base + t * A2_BK, make_desc_sw128(q_addr + kk * 32, 1024),
                   make_desc_sw128(k_addr + kk * 32, 1024), idesc_qk, kk != 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
        ++s_cnt[t];
      };
      auto issue_pv = [&](int t, uint32_t v_addr) {
        const uint32_t c = pv_cnt[t];
        const int b = c & 1;
        mbar_wait(&p_full[t], c & 1);
        mbar_wait(&o_empty[t * 2 + b], ((c >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(sP + t * A2_P);
        const uint32_t d_tmem = tmem_base + 256 + t * 128 + b * 64;
#pragma unroll
        for (int kk = 0; kk < A2_BK / 16; ++kk)
          umma_f16(d_tmem, make_desc_sw128(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 1024),
                   make_desc_sw128(v_addr + kk * 2048, 1024), idesc_pv, kk != 0 ? 1u : 0u);
        umma_commit(&o_full[t * 2 + b]);
        umma_commit(&p_empty[t]);
        ++pv_cnt[t];
      };
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++item_cnt) {
        const uint32_t qpar = item_cnt & 1;
        mbar_wait(&q_full[0], qpar);
        mbar_wait(&q_full[1], qpar);
        // S_A(0), S_B(0)
        mbar_wait(&k_full[kst], kph);
        {
          const uint32_t k_addr = smem_u32(sK + kst * A2_TILE);
          issue_qk(0, k_addr);
          issue_qk(1, k_addr);
          umma_commit(&k_empty[kst]);
          if (++kst == A2_STAGES) { kst = 0; kph ^= 1; }
        }
        for (int j = 0; j < n_kv; ++j) {
          // S(j+1) for both tiles is issued as soon as the softmax warpgroups have pulled S(j) into registers
          // (s_empty), i.e. long before they finish tile j: the softmax never waits for the tensor pipe.
          if (j + 1 < n_kv) {
            mbar_wait(&k_full[kst], kph);
            const uint32_t k_addr = smem_u32(sK + kst * A2_TILE);
            issue_qk(0, k_addr);
            issue_qk(1, k_addr);
            umma_commit(&k_empty[kst]);
            if (++kst == A2_STAGES) { kst = 0; kph ^= 1; }
          } else {
            umma_commit(&q_empty[0]);
            umma_commit(&q_empty[1]);
          }
          mbar_wait(&v_full[vst], vph);
          const uint32_t v_addr = smem_u32(sV + vst * A2_TILE);
          issue_pv(0, v_addr);
          issue_pv(1, v_addr);
          umma_commit(&v_empty[vst]);
          if (++vst == A2_STAGES) { vst = 0; vph ^= 1; }
        }
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax warpgroups
    setmaxnreg_inc_224();
    const int t = (warp - 4) >> 2;                 // query tile 0 (A) / 1 (B)
    const int ew = (warp - 4) & 3;
    const int row = ew * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t tS = tmem_base + t * A2_BK + lane_off;
    const uint32_t tO = tmem_base + 256 + t * 128 + lane_off;
    uint8_t* const pb = sP + t * A2_P + row * 128;
    const float c = p.scale_log2;
    uint32_t kv_cnt = 0;                           // kv tiles processed by this warpgroup (all items)
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int qp = item % p.q_pairs;
      const int head = (item / p.q_pairs) % p.H;
      const int seq = item / (p.q_pairs * p.H);
      float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
      float o[A2_D];
#pragma unroll
      for (int i = 0; i < A2_D; ++i) o[i] = 0.f;

      auto fold_pv = [&](uint32_t cnt) {          // cnt = global index of the PV being folded
        const int b = cnt & 1;
        mbar_wait(&o_full[t * 2 + b], (cnt >> 1) & 1);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(tO + b * 64, r0);
        tmem_ld_32x32(tO + b * 64 + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&o_empty[t * 2 + b]);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          o[i] = fmaf(o[i], alpha_prev, __uint_as_float(r0[i]));
          o[32 + i] = fmaf(o[32 + i], alpha_prev, __uint_as_float(r1[i]));
        }
      };

      for (int j = 0; j < n_kv; ++j, ++kv_cnt) {
        mbar_wait(&s_full[t], kv_cnt & 1);
        tc_fence_after();
        float s[A2_BK];
        {
          uint32_t r0[32], r1[32], r2[32], r3[32];
          tmem_ld_32x32(tS, r0);
          tmem_ld_32x32(tS + 32, r1);
          tmem_ld_32x32(tS + 64, r2);
          tmem_ld_32x32(tS + 96, r3);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s[i] = __uint_as_float(r0[i]); s[32 + i] = __uint_as_float(r1[i]);
            s[64 + i] = __uint_as_float(r2[i]); s[96 + i] = __uint_as_float(r3[i]);
          }
        }
        tc_fence_before();
        mbar_arrive(&s_empty[t]);
        const int valid = p.Lk - j * A2_BK;
        if (valid < A2_BK) {
#pragma unroll
          for (int i = 0; i < A2_BK; ++i) if (i >= valid) s[i] = -INFINITY;
        }
        float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
#pragma unroll
        for (int i = 4; i < A2_BK; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1]));
          mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
        }
        const float m_new = fmaxf(m, fmaxf(mx0, mx1));
        const float alpha = ex2_approx((m - m_new) * c);
        const float mc = m_new * c;
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
        for (int i = 0; i < A2_BK; i += 4) {
          s[i] = ex2_approx(fmaf(s[i], c, -mc));
          s[i + 1] = ex2_approx(fmaf(s[i + 1], c, -mc));
          s[i + 2] = ex2_approx(fmaf(s[i + 2], c, -mc));
          s[i + 3] = ex2_approx(fmaf(s[i + 3], c, -mc));
          sum0 += s[i];
          sum1 += s[i + 1];
          sum2 += s[i + 2];
          sum3 += s[i + 3];
        }
        l = fmaf(l, alpha, (sum0 + sum1) + (sum2 + sum3));
        m = m_new;
        mbar_wait(&p_empty[t], (kv_cnt & 1) ^ 1);
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
        mbar_arrive(&p_full[t]);
        if (j > 0) fold_pv(kv_cnt - 1);
        alpha_prev = alpha;
      }
      fold_pv(kv_cnt - 1);
      const int qrow = (qp * 2 + t) * A2_BQ + row;
      if (qrow < p.Lq) {
        const float inv = 1.0f / l;
        uint16_t* dst = reinterpret_cast<uint16_t*>(p.o) + (static_cast<int64_t>(seq) * p.Lq + qrow) * p.ldo + head * A2_D;
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <bool BF16>
int launch_attention2(const CUtensorMap& tQ, const CUtensorMap& tK, const CUtensorMap& tV, const Attn2Params& p,
                      cudaStream_t stream) {
  auto kern = attention2_kernel<BF16>;
  static bool configured = false;
  static int sms = 148;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM);
    if (e != cudaSuccess) return (int)e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  const int grid = p.total_items < sms ? p.total_items : sms;
  kern<<<grid, A2_THREADS, A2_SMEM, stream>>>(tQ, tK, tV, p);
  return (int)cudaGetLastError();
}

}  // namespace iggt

using namespace iggt;

extern "C" int iggt_attention_fwd_v1(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                     int64_t ldv, void* o, int64_t ldo, int num_seq, int Lq, int Lk, int H,
                                     int head_dim, float scale, int dtype, iggt_stream_t stream);
extern "C" int iggt_attention_fwd_v3(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                     int64_t ldv, void* o, int64_t ldo, int num_seq, int Lq, int Lk, int H,
                                     int head_dim, float scale, int dtype, iggt_stream_t stream);

extern "C" int iggt_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                  int64_t ldv, void* o, int64_t ldo, int num_seq, int Lq, int Lk, int H,
                                  int head_dim, float scale, int dtype, iggt_stream_t stream) {
  static const int ver = [] { const char* e = getenv("IGGT_ATTN"); return e ? atoi(e) : 3; }();
  if (ver == 1) return iggt_attention_fwd_v1(q, ldq, k, ldk, v, ldv, o, ldo, num_seq, Lq, Lk, H, head_dim, scale, dtype, stream);
  if (ver == 3) return iggt_attention_fwd_v3(q, ldq, k, ldk, v, ldv, o, ldo, num_seq, Lq, Lk, H, head_dim, scale, dtype, stream);
  if (head_dim != 64) return -1;
  if (num_seq <= 0 || Lq <= 0 || Lk <= 0 || H <= 0) return -1;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return -2;
  if (dtype != 0 && dtype != 1) return -3;
  const TmDtype dt = dtype ? TM_BF16 : TM_F16;
  CUtensorMap tQ, tK, tV;
  if (make_tmap_2d(&tQ, dt, q, (uint64_t)num_seq * Lq, (uint64_t)H * 64, ldq, 64, A2_BQ)) return -4;
  if (make_tmap_2d(&tK, dt, k, (uint64_t)num_seq * Lk, (uint64_t)H * 64, ldk, 64, A2_BK)) return -4;
  if (make_tmap_2d(&tV, dt, v, (uint64_t)num_seq * Lk, (uint64_t)H * 64, ldv, 64, A2_BK)) return -4;
  Attn2Params p;
  p.Lq = Lq; p.Lk = Lk; p.H = H; p.num_seq = num_seq;
  const int q_tiles = (Lq + A2_BQ - 1) / A2_BQ;
  p.q_pairs = (q_tiles + 1) / 2;
  p.total_items = num_seq * H * p.q_pairs;
  p.ldo = ldo; p.o = o;
  p.scale_log2 = scale * 1.4426950408889634f;
  return dtype ? launch_attention2<true>(tQ, tK, tV, p, (cudaStream_t)stream)
               : launch_attention2<false>(tQ, tK, tV, p, (cudaStream_t)stream);
}
