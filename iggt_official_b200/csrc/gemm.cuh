// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] (16-bit, K-major) x W[N,K]^T (16-bit, K-major), fp32 accumulate in TMEM )
//
// One CTA per SM loops over 128 x BN output tiles (or, PAIR, two CTAs of a cluster over 256 x BN tiles with one
// cta_group::2 MMA). Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warp 2 = TMEM
// allocator, warps 4.. = one or two 4-warp epilogue groups (TMEM -> registers -> swizzled smem staging -> TMA store /
// TMA reduce-add). Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1; work is
// handed out by WorkIter (whole tiles round-robin, or stream-K ranges of k-blocks for the residual epilogue).
//
// Reference call sites this replaces (all via torch.nn on the reference side):
//   qkv   iggt/layers/attention.py:52-58   (+ q/k LayerNorm(64) and 2-D RoPE, rope.py:154-188)
//   proj  iggt/layers/attention.py:74-75 + layer_scale.py:27 + block.py:105 (residual)
//   fc1   iggt/layers/mlp.py:35-36 (exact-erf GELU)      fc2  mlp.py:38 + block.py:106
//   1x1 / 3x3 convolutions of the dense heads, iggt/heads/dpt_head.py:234-316
#pragma once
#include "ptx.cuh"
#include "launch.cuh"

namespace iggt {

enum GemmEpi : int {
  EPI_STORE16 = 0,   // out16 = act(acc + bias) [+ addend]
  EPI_RESID32 = 1,   // out32 += gamma * (acc + bias)           (TMA reduce-add into the fp32 residual)
  EPI_QKV = 2,       // out16 = [rope(ln(q)) | rope(ln(k)) | v]  (per 64-wide head)
  EPI_STORE32 = 3,   // out32 = acc + bias
  EPI_QKV_GATHER = 4,  // EPI_QKV + the K | V chunks also stored into every rank's gathered buffer (view sharding); a separate
                       // instantiation so that the single-GPU qkv kernel carries none of it (it cost 22 % when it did)
};
__host__ __device__ constexpr bool epi_is_qkv(int e) { return e == EPI_QKV || e == EPI_QKV_GATHER; }

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_k_blocks;
  const float* bias;   // [N] or nullptr
  const float* gamma;  // [N] (EPI_RESID32) or nullptr (=1)
  int act;             // 0 none, 1 exact GELU, 2 ReLU, 3 LeakyReLU(0.01), 5 exact GELU (one-MUFU erfc form)
  // EPI_QKV
  int qk_norm;         // apply LayerNorm(64) (eps 1e-5, affine) + RoPE to the q and k column ranges
  int C;               // embedding width: q = cols [0,C), k = [C,2C), v = [2C,3C)
  const float* qn_w; const float* qn_b; const float* kn_w; const float* kn_b;  // [64]
  const float* rope_cos; const float* rope_sin;  // [npos][16]
  const int* pos_yx;   // [T][2] (y,x) RoPE positions of one view's tokens
  int T;               // tokens per view (row r of A is token r % T)
  // optional elementwise addend (EPI_STORE16): out += addend[(row % add_rows) * add_ld + col] (16-bit)
  const void* addend; int add_rows; int add_ld;
  // convolution mode (A is an NHWC tensor; K loop runs taps x C/64)
  int conv_taps;       // 1 (1x1 through the 4-D path) or 9 (3x3, pad 1)
  int conv_C;          // input channels
  int H, W, NB;        // spatial size / images
  int tiles_x, tiles_y;
  // residuals for the conv epilogue: out = act_post(act(acc+bias) + resid16[pixel,col] + resid2_16[pixel,col])
  // (NHWC, same H,W, ld = N)
  const void* resid;
  const void* resid2;
  int act_post;
  // EPI_RESID32: round (acc + bias) to 16 bit before the LayerScale multiply (autocast Linear output)
  int round_out16;
  // EPI_QKV, view sharding: the K | V column chunks (col >= gather_col0) of every tile are ALSO stored through these
  // tensor maps (device array; one per rank of the box, each describing THIS rank's row window of that rank's gathered
  // K|V buffer, reached over NVLink peer mappings): the all-gather of the global attention's keys and values is fused
  // into the producing GEMM, tile by tile (parallel.py, FusedKVGather)
  // The maps are 3-D {2048 columns, rows of one scene, scenes}: the rank's rows (scene, view, token) land at
  // (scene, rank, view, token) in the gathered buffer.  The TMA store of a 128-row tile is clipped to the scene of the
  // tile's first row; rows of a tile that belong to a later scene (B - 1 tiles per GEMM) are written by their epilogue
  // threads with plain 16-byte stores through the raw peer pointers that follow the maps in the device array
  // ([n maps][n pointers][ld, scene_ld as int64]).
  const CUtensorMap* gather_maps;
  int n_gather;
  int gather_col0;
  int gather_rows;     // rows of this rank per scene (S_loc * T)
  // stream-K (EPI_RESID32 only): the (tile, k-block) space is cut into gridDim.x equal contiguous ranges; every
  // CTA reduce-adds the partial product of each tile segment it owns (fp32 atomics in L2 make the pieces add up)
  int stream_k;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;
constexpr int CONV_TW = 16;
constexpr int CONV_TH = 8;

// PAIR: the CTA is one half of a cta_group::2 pair working on a 256 x BN tile; it stages its own 128 rows of A and
// HALF of the B tile (BN/2 weight rows) per k-block, so the per-SM smem fill drops from 48 KB to 32 KB per k-block
// at BN = 256 (the 1-CTA kernel is bound by exactly that traffic: 97 B/clk of TMA writes + 96 B/clk of MMA reads).
template <int BN, bool PAIR = false>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KB
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES = 16384;                 // 128 rows x 128 B staging tile
  static constexpr int STAGES = PAIR ? (BN == 256 ? 6 : 8) : ((BN == 256) ? 4 : ((BN == 128) ? 6 : 8));
  static constexpr int VEC_BYTES = BN * 4 + (BN * 4 > 1024 ? BN * 4 : 1024);   // bias[BN] + (gamma[BN] | q/k-norm vectors [4][64])
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 2 * STG_BYTES + 256 /*barriers*/ + VEC_BYTES;
};

template <bool BF16>
__device__ __forceinline__ float cvt16_to_f32(uint16_t h) {
  if constexpr (BF16) return __uint_as_float(static_cast<uint32_t>(h) << 16);
  else return __half2float(__ushort_as_half(h));
}
template <bool BF16>
__device__ __forceinline__ float round16(float x) {
  if constexpr (BF16) return __bfloat162float(__float2bfloat16_rn(x));
  else return __half2float(__float2half_rn(x));
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return gelu_fast(x);
  if (act == 2) return relu_nan(x);
  if (act == 3) return x > 0.0f ? x : 0.01f * x;
  return x;
}

// Work iterator shared by the three warp roles.  Default: whole tiles, static round-robin over the CTAs.
// stream_k: CTA b owns k-blocks [b*per, (b+1)*per) of the linearised (tile, k-block) space.
struct WorkIter {
  int pos, end, step, kb_per_tile, num_tiles;
  bool sk;
  // `worker` of `workers`: the CTA (or CTA pair) index and count
  __device__ WorkIter(const GemmParams& p, int worker, int workers) {
    kb_per_tile = p.num_k_blocks;
    num_tiles = p.num_m_tiles * p.num_n_tiles;
    sk = p.stream_k != 0;
    if (sk) {
      const long total = static_cast<long>(num_tiles) * kb_per_tile;
      const long per = (total + workers - 1) / workers;
      const long b = static_cast<long>(worker) * per;
      pos = static_cast<int>(b < total ? b : total);
      end = static_cast<int>(b + per < total ? b + per : total);
      step = 0;
    } else {
      pos = worker; end = num_tiles; step = workers;
    }
  }
  // next segment: tile index and k-block range [kb0, kb1)
  __device__ bool next(int& tile, int& kb0, int& kb1) {
    if (pos >= end) return false;
    if (sk) {
      tile = pos / kb_per_tile;
      kb0 = pos - tile * kb_per_tile;
      const int room = end - pos;
      kb1 = kb0 + room < kb_per_tile ? kb0 + room : kb_per_tile;
      pos += kb1 - kb0;
    } else {
      tile = pos; kb0 = 0; kb1 = kb_per_tile; pos += step;
    }
    return true;
  }
};

// CONV: A operand comes from a 4-D NHWC tensor map (box {64, TW, TH, 1}); otherwise 2-D [M,K].
// G = epilogue warpgroups (1 or 2). With G = 2 the column chunks of a tile alternate between two 4-warp
// groups (each TMEM lane quarter is then read by two warps), doubling epilogue issue slots and halving
// the registers available per thread (384 threads) -- used for every epilogue except the qkv one.
// PAIR = cta_group::2: launched as clusters of two CTAs; p.num_m_tiles then counts 256-row tile pairs and CTA `rank`
// of the pair owns the 128-row sub-tile 2 * mt + rank (a sub-tile past the end of the problem is all TMA zero fill
// on the way in and clipped on the way out).
template <int BN, int EPI, bool BF16, bool CONV, int G, bool PAIR>
__global__ void __launch_bounds__(128 + 128 * G, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using SM = GemmSmem<BN, PAIR>;
  constexpr int STAGES = SM::STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzled tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * SM::A_BYTES;
  uint8_t* staging = smem + STAGES * SM::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * SM::STG_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* epi_vec = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;          // 0 = the CTA that issues the pair's MMAs
  const int worker = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int workers = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * G * (PAIR ? 2 : 1));   // PAIR: the epilogue warps of both CTAs release rank 0's
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_pair<2 * BN>(tmem_slot);
    else tmem_alloc<2 * BN>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();      // PDL: everything above overlapped the previous kernel's tail
  griddep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      WorkIter work(p, worker, workers);
      int tile, kb0, kb1;
      while (work.next(tile, kb0, kb1)) {
        const int mt = (tile / p.num_n_tiles) * (PAIR ? 2 : 1) + static_cast<int>(rank);
        const int nt = tile % p.num_n_tiles;
        int img = 0, y0 = 0, x0 = 0;
        if constexpr (CONV) {
          const int per_img = p.tiles_x * p.tiles_y;
          img = mt / per_img;
          const int r = mt % per_img;
          y0 = (r / p.tiles_x) * CONV_TH;
          x0 = (r % p.tiles_x) * CONV_TW;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          int dy = 0, dx = 0, c0 = 0;
          if constexpr (CONV) {
            const int cblocks = p.conv_C / GEMM_BK;
            const int tap = kb / cblocks;
            c0 = (kb % cblocks) * GEMM_BK;
            if (p.conv_taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
          }
          if constexpr (PAIR) {
            // both CTAs' bytes are credited to rank 0's barrier; rank 0 posts the expectation for the pair
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * SM::STAGE_BYTES);
            const uint32_t fb = mapa_u32(&full_bar[stage], 0);
            if constexpr (CONV) tma_load_4d_pair(smem_a + stage * SM::A_BYTES, &tmA, fb, c0, x0 + dx, y0 + dy, img);
            else tma_load_2d_pair(smem_a + stage * SM::A_BYTES, &tmA, fb, kb * GEMM_BK, mt * GEMM_BM);
            tma_load_2d_pair(smem_b + stage * SM::B_BYTES, &tmB, fb, kb * GEMM_BK,
                             nt * BN + static_cast<int>(rank) * (BN / 2));
          } else {
            mbar_expect_tx(&full_bar[stage], SM::STAGE_BYTES);
            if constexpr (CONV) tma_load_4d(smem_a + stage * SM::A_BYTES, &tmA, &full_bar[stage], c0, x0 + dx, y0 + dy, img);
            else tma_load_2d(smem_a + stage * SM::A_BYTES, &tmA, &full_bar[stage], kb * GEMM_BK, mt * GEMM_BM);
            tma_load_2d(smem_b + stage * SM::B_BYTES, &tmB, &full_bar[stage], kb * GEMM_BK, nt * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(PAIR ? 2 * GEMM_BM : GEMM_BM, BN, BF16, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      WorkIter work(p, worker, workers);
      int tile, kb0, kb1;
      while (work.next(tile, kb0, kb1)) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * SM::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * SM::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t da = make_desc_sw128(a_addr + k * 32, 1024);
            const uint64_t db = make_desc_sw128(b_addr + k * 32, 1024);
            if constexpr (PAIR) umma_f16_pair(d_tmem, da, db, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            else umma_f16(d_tmem, da, db, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          // frees the smem slot (of both CTAs) when these MMAs retire
          if constexpr (PAIR) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (PAIR) umma_commit_pair(&tfull_bar[acc]);
        else umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (128 threads, 1 row each)
    const int grp = (warp - 4) >> 2;       // epilogue group
    const int ew = (warp - 4) & 3;         // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    const int row = ew * 32 + lane;        // row inside the tile
    const bool leader = (ew == 0 && lane == 0);
    constexpr int NBUF = 2 / G;            // staging buffers per group
    uint8_t* const stg_grp = staging + grp * NBUF * SM::STG_BYTES;
    const uint32_t bar_id = 1 + grp;
    const int gtid = ew * 32 + lane;       // thread index inside the group
    float* const vb = epi_vec;                  // this tile's bias   [BN]  (staged before the accumulator is ready)
    float* const vg = vb + BN;                  // this tile's gamma  [BN]  (EPI_RESID32)
    float* const vn = vb + BN;                  // q_norm w,b | k_norm w,b  [4][64]  (EPI_QKV, BN >= 128)
    if constexpr (epi_is_qkv(EPI)) {
      if (p.qk_norm) {
        for (int i = gtid; i < 64; i += 128) {
          vn[i] = p.qn_w[i]; vn[64 + i] = p.qn_b[i]; vn[128 + i] = p.kn_w[i]; vn[192 + i] = p.kn_b[i];
        }
      }
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t store_count = 0;
    WorkIter work(p, worker, workers);
    int tile, kb0, kb1;
    // releasing an accumulator stage: one arrival per epilogue warp on the MMA issuer's (rank 0's) barrier
    auto release_acc = [&](int a) {
      if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(&tempty_bar[a], 0));
      else mbar_arrive(&tempty_bar[a]);
    };
    while (work.next(tile, kb0, kb1)) {
      const int mt = (tile / p.num_n_tiles) * (PAIR ? 2 : 1) + static_cast<int>(rank);
      const int nt = tile % p.num_n_tiles;
      const int n0 = nt * BN;
      int img = 0, y0 = 0, x0 = 0;
      long grow;                           // global row (pixel) index of this thread, -1 if out of range
      if constexpr (CONV) {
        const int per_img = p.tiles_x * p.tiles_y;
        img = mt / per_img;
        const int r = mt % per_img;
        y0 = (r / p.tiles_x) * CONV_TH;
        x0 = (r % p.tiles_x) * CONV_TW;
        const int yy = y0 + row / CONV_TW, xx = x0 + row % CONV_TW;
        grow = (yy < p.H && xx < p.W && img < p.NB) ? ((long)img * p.H + yy) * p.W + xx : -1;
      } else {
        grow = (long)mt * GEMM_BM + row;
        if (grow >= p.M) grow = -1;
      }
      // stage the tile's per-column vectors in smem while the MMAs of this tile are still running
      // (with the smem carve-out this kernel uses there is no L1 left to cache them)
      named_bar_sync(3, 128 * G);          // every epilogue thread is done with the previous tile's vectors
      for (int i = grp * 128 + gtid; i < BN; i += 128 * G) {
        const int col = n0 + i;
        vb[i] = (p.bias && col < p.N && kb0 == 0) ? __ldg(p.bias + col) : 0.f;   // bias rides with the first K segment
        if constexpr (EPI == EPI_RESID32) vg[i] = (p.gamma && col < p.N) ? __ldg(p.gamma + col) : 1.f;
      }
      named_bar_sync(3, 128 * G);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;

      if constexpr (EPI == EPI_STORE16 || epi_is_qkv(EPI)) {
        const int nvalid = min(BN / 64, (p.N - n0 + 63) / 64);
        if (grp >= nvalid) {               // nothing to read for this group: release immediately
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(acc);
        }
#pragma unroll 1
        for (int c64 = grp; c64 < nvalid; c64 += G) {
          const int col0 = n0 + c64 * 64;
          uint8_t* stg = stg_grp + (store_count % NBUF) * SM::STG_BYTES;
          if (leader) tma_store_wait_read<NBUF - 1>();
          named_bar_sync(bar_id, 128);
          float v[64];
          {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(t_row + c64 * 64, r0);
            tmem_ld_32x32(t_row + c64 * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(r0[i]); v[32 + i] = __uint_as_float(r1[i]); }
          }
          if (c64 + G >= nvalid) {
            // last TMEM read of this tile by this warp: release the accumulator stage
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
          }
          {
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
              const float4 b = *reinterpret_cast<const float4*>(vb + c64 * 64 + i);
              const float2 lo = fadd2(make_float2(v[i], v[i + 1]), make_float2(b.x, b.y));
              const float2 hi = fadd2(make_float2(v[i + 2], v[i + 3]), make_float2(b.z, b.w));
              v[i] = lo.x; v[i + 1] = lo.y; v[i + 2] = hi.x; v[i + 3] = hi.y;
            }
          }
          if constexpr (epi_is_qkv(EPI)) {
            if (p.qk_norm && col0 < 2 * p.C) {
              const bool is_k = col0 >= p.C;
              const float* nw = vn + (is_k ? 128 : 0);
              const float* nb = nw + 64;
              // the reference rounds the Linear output to 16 bit before the fp32 LayerNorm (autocast)
              float s = 0.f;
#pragma unroll
              for (int i = 0; i < 64; ++i) { v[i] = round16<BF16>(v[i]); s += v[i]; }
              const float mean = s * (1.0f / 64.0f);
              float q = 0.f;
#pragma unroll
              for (int i = 0; i < 64; ++i) { const float d = v[i] - mean; q += d * d; }
              const float rstd = rsqrtf(q * (1.0f / 64.0f) + 1e-5f);
#pragma unroll
              for (int i = 0; i < 64; ++i) v[i] = (v[i] - mean) * rstd * nw[i] + nb[i];
              // 2-D RoPE: dims [0,32) rotate with the y position, [32,64) with x; halves of 16
              const int t = (grow >= 0) ? static_cast<int>(grow % p.T) : 0;
              const int py = __ldg(p.pos_yx + 2 * t), px = __ldg(p.pos_yx + 2 * t + 1);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int ps = h == 0 ? py : px;
                const float* cs = p.rope_cos + ps * 16;
                const float* sn = p.rope_sin + ps * 16;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float c = __ldg(cs + i), s_ = __ldg(sn + i);
                  const float a = v[h * 32 + i], b = v[h * 32 + 16 + i];
                  v[h * 32 + i] = a * c - b * s_;
                  v[h * 32 + 16 + i] = b * c + a * s_;
                }
              }
            }
          } else {
            if (p.act == 1) {
              // autocast: GELU is evaluated on the 16-bit Linear output (iggt/layers/mlp.py:35-36)
#pragma unroll
              for (int i = 0; i < 64; i += 2) {
                const float2 g = gelu_fast2(make_float2(round16<BF16>(v[i]), round16<BF16>(v[i + 1])));
                v[i] = g.x; v[i + 1] = g.y;
              }
            } else if (p.act == 5) {
              // the same GELU through the one-MUFU erfc form (ptx.cuh gelu_erfc2)
#pragma unroll
              for (int i = 0; i < 64; i += 2) {
                const float2 g = gelu_erfc2(make_float2(round16<BF16>(v[i]), round16<BF16>(v[i + 1])));
                v[i] = g.x; v[i + 1] = g.y;
              }
            } else if (p.act) {
#pragma unroll
              for (int i = 0; i < 64; ++i) v[i] = apply_act(v[i], p.act);
            }
            if (p.addend && grow >= 0) {
              const uint16_t* ad = reinterpret_cast<const uint16_t*>(p.addend) +
                                   (grow % p.add_rows) * (long)p.add_ld + col0;
#pragma unroll
              for (int i = 0; i < 64; i += 8) {
                if (col0 + i < p.N) {
                  const uint4 u = __ldg(reinterpret_cast<const uint4*>(ad + i));
                  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    v[i + 2 * j] += cvt16_to_f32<BF16>(static_cast<uint16_t>(w[j] & 0xFFFF));
                    v[i + 2 * j + 1] += cvt16_to_f32<BF16>(static_cast<uint16_t>(w[j] >> 16));
                  }
                }
              }
            }
            if constexpr (CONV) {
#pragma unroll
              for (int rr = 0; rr < 2; ++rr) {
                const void* rp = rr == 0 ? p.resid : p.resid2;
                if (rp && grow >= 0) {
                  const uint16_t* ad = reinterpret_cast<const uint16_t*>(rp) + grow * (long)p.N + col0;
#pragma unroll
                  for (int i = 0; i < 64; i += 8) {
                    if (col0 + i < p.N) {
                      const uint4 u = __ldg(reinterpret_cast<const uint4*>(ad + i));
                      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                      for (int j = 0; j < 4; ++j) {
                        v[i + 2 * j] += cvt16_to_f32<BF16>(static_cast<uint16_t>(w[j] & 0xFFFF));
                        v[i + 2 * j + 1] += cvt16_to_f32<BF16>(static_cast<uint16_t>(w[j] >> 16));
                      }
                    }
                  }
                }
              }
              if (p.act_post) {
#pragma unroll
                for (int i = 0; i < 64; ++i) v[i] = apply_act(v[i], p.act_post);
              }
            }
          }
          // registers -> 128B-swizzled staging tile (row = 128 B = 64 x 16 bit)
          bool later_scene = false;                                       // EPI_QKV gather: this row is not in the tile's first scene
          if constexpr (EPI == EPI_QKV_GATHER) {
            if (p.n_gather > 0 && col0 >= p.gather_col0 && grow >= 0)
              later_scene = grow / p.gather_rows != (static_cast<long>(mt) * GEMM_BM) / p.gather_rows;
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 u;
            u.x = pack16x2<BF16>(v[c * 8 + 0], v[c * 8 + 1]);
            u.y = pack16x2<BF16>(v[c * 8 + 2], v[c * 8 + 3]);
            u.z = pack16x2<BF16>(v[c * 8 + 4], v[c * 8 + 5]);
            u.w = pack16x2<BF16>(v[c * 8 + 6], v[c * 8 + 7]);
            *reinterpret_cast<uint4*>(stg + row * 128 + ((c ^ (row & 7)) << 4)) = u;
            if constexpr (EPI == EPI_QKV_GATHER) {
              if (later_scene) {
                void* const* ptrs = reinterpret_cast<void* const*>(p.gather_maps + p.n_gather);
                const long* lds = reinterpret_cast<const long*>(ptrs + p.n_gather);
                const long scene = grow / p.gather_rows, rr = grow - scene * p.gather_rows;
                const long off = scene * __ldg(lds + 1) + rr * __ldg(lds) + (col0 - p.gather_col0) + c * 8;
                for (int r = 0; r < p.n_gather; ++r)
                  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(ptrs[r]) + off) = u;
              }
            }
          }
          if constexpr (EPI == EPI_QKV_GATHER) {
            if (later_scene) __threadfence_system();                      // peer stores visible before the cross-rank barrier
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (leader) {
            if constexpr (CONV) tma_store_4d(&tmC, stg, col0, x0, y0, img);
            else tma_store_2d(&tmC, stg, col0, mt * GEMM_BM);
            if constexpr (EPI == EPI_QKV_GATHER) {
              if (p.n_gather > 0 && col0 >= p.gather_col0) {             // K | V chunk: to every rank's gathered buffer too
                const int row0 = mt * GEMM_BM, scene = row0 / p.gather_rows, r0 = row0 - scene * p.gather_rows;
                for (int r = 0; r < p.n_gather; ++r)
                  tma_store_3d(&p.gather_maps[r], stg, col0 - p.gather_col0, r0, scene);      // clipped to this scene's rows
              }
            }
            tma_store_commit();
          }
          ++store_count;
        }
      } else {
        // fp32 outputs: 32 columns (128 B) per staging tile
        const int nvalid = min(BN / 32, (p.N - n0 + 31) / 32);
        if (grp >= nvalid) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(acc);
        }
#pragma unroll 1
        for (int c32 = grp; c32 < nvalid; c32 += G) {
          const int col0 = n0 + c32 * 32;
          uint8_t* stg = stg_grp + (store_count % NBUF) * SM::STG_BYTES;
          if (leader) tma_store_wait_read<NBUF - 1>();
          named_bar_sync(bar_id, 128);
          uint32_t r0[32];
          tmem_ld_32x32(t_row + c32 * 32, r0);
          tmem_ld_wait();
          if (c32 + G >= nvalid) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
          }
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(vb + c32 * 32 + i);
            float2 lo = fadd2(make_float2(v[i], v[i + 1]), make_float2(b.x, b.y));
            float2 hi = fadd2(make_float2(v[i + 2], v[i + 3]), make_float2(b.z, b.w));
            if constexpr (EPI == EPI_RESID32) {
              if (p.round_out16 && !p.stream_k) {
                lo.x = round16<BF16>(lo.x); lo.y = round16<BF16>(lo.y);
                hi.x = round16<BF16>(hi.x); hi.y = round16<BF16>(hi.y);
              }
              const float4 g = *reinterpret_cast<const float4*>(vg + c32 * 32 + i);
              lo = fmul2(lo, make_float2(g.x, g.y));
              hi = fmul2(hi, make_float2(g.z, g.w));
            }
            v[i] = lo.x; v[i + 1] = lo.y; v[i + 2] = hi.x; v[i + 3] = hi.y;
          }
          if constexpr (EPI == EPI_STORE32) {
            if (p.act) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = apply_act(v[i], p.act);
            }
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 u = make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
            *reinterpret_cast<float4*>(stg + row * 128 + ((c ^ (row & 7)) << 4)) = u;
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (leader) {
            if constexpr (EPI == EPI_RESID32) tma_reduce_add_2d(&tmC, stg, col0, mt * GEMM_BM);
            else tma_store_2d(&tmC, stg, col0, mt * GEMM_BM);
            tma_store_commit();
          }
          ++store_count;
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (leader) {
      tma_store_wait_all<0>();
      if constexpr (EPI == EPI_QKV_GATHER) {
        if (p.n_gather > 0) __threadfence_system();      // peer stores visible before the cross-rank barrier
      }
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // neither CTA may exit (or free TMEM) while its peer can still touch it
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair<2 * BN>(tmem_base);
    else tmem_dealloc<2 * BN>(tmem_base);
  }
}

}  // namespace iggt
