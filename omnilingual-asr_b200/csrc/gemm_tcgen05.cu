// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (accumulators in TMEM, double buffered) -> tcgen05.ld epilogue with fused bias / GELU / residual /
// LayerNorm+GELU / arg-max.  One CTA per SM, 384 threads:
//   warp 0      TMA producer (one elected lane)
//   warp 1      MMA issuer   (one elected lane)
//   warp 2      TMEM allocator
//   warp 3      idle
//   warps 4-11  epilogue: warp w owns TMEM lanes [32*(w%4), +32) and column half (w-4)/4 of the tile
// Used for: FE conv layers 1-6 (implicit GEMM over overlapping rows), feature projection, QKV, out-proj,
// FFN1, FFN2 and the CTC head.
#include "gemm.cuh"
#include "host_util.h"
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

namespace oasr {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // one 128-byte swizzle atom of bf16
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;
constexpr int EPI_WARPS = 8;

struct GemmKernelParams {
  int rows_per_batch, batches, groups, N;
  int m_tiles_per_batch, n_tiles, total_tiles, k_blocks;
  int kb_per_tap, P;       // k-block -> (tap, column block); tap -> (parity, position offset)
  int slab_stages, slab_stage_bytes, slab_chunk_bytes;   // SLAB kernels: ring depth / stage (= one tap of W) / chunk size
  int k16_last;            // K = 16 steps that hold data in the LAST k-block of a tap (1 .. 4): the columns beyond
                           // a_inner are zero padding (pos-conv at 80 channels per group: 64 + 16 of 128), and
                           // multiplying zeros cost 3 of every 8 MMAs there
  int umma_n;              // N of one tcgen05.mma (<= 256, multiple of 16)
  int b_box_rows;          // rows of W fetched per TMA box
  uint32_t stage_tx_bytes; // bytes landing on a full barrier per stage
  int ldo;
  long long out_batch_rows;
  void* out;
  const float* bias;
  const float* resid;
  const float* ln_gamma;
  const float* ln_beta;
  unsigned long long* argmax;
  const int* n_valid;
  int frames_per_seq;
  int route_n, route_per;            // EPI_BF16: rows routed to their owners' (peer) buffers, see GemmArgs
  __nv_bfloat16* route_base[8];
};

// CG = CTAs cooperating on one tile (tcgen05 cta_group): 1, or 2 = a CTA pair computing a 256 x BN tile with each
// CTA holding 128 rows of A, half of the B rows and 128 rows of the accumulator.
constexpr bool epi_has_resid(int epi) { return epi == EPI_F32_RESID || epi == EPI_F32_GELU_RESID; }
// LayerNorm epilogue with BN = 256: the 512 channels of a row are split over the two CTAs of a cluster (CTA r computes
// columns [256 r, +256) of the same 128 rows with its own cta_group::1 pipeline), so that each CTA's accumulator is
// 256 TMEM columns and can be double-buffered: the next tile's MMAs run under the three-pass LayerNorm epilogue, which
// with a 512-column accumulator had the tensor pipe idle 43 % of the time.  The row statistics are completed through
// distributed shared memory (each warp writes its partial into both CTAs, one cluster-scope mbarrier per pass).
constexpr bool epi_nsplit(int bn, int epi) { return epi == EPI_LN_GELU_BF16; }
// Positional conv (the one user of EPI_F32_GELU_RESID): tap j of an output tile reads input rows [m0 + j, m0 + j + 128) -
// the same rows as tap j - 1 shifted by one.  Fetching a 128-row A tile per tap moved 32 KB of A beside 20 KB of W per
// tap and tile through L2 -> shared memory: 41 GB per launch at ~18 TB/s, the kernel's bound (tensor pipe 26 %).
// SLAB: the 256 input rows [m0, m0 + 256) a tile can touch (128 + up to 128 taps) are fetched ONCE per tile into a
// 128B-swizzled slab, and tap j's A operand is the slab read from row j on: descriptor start address + j * 128 bytes.
// The start is then not aligned to the swizzle pattern's 1024-byte repeat; measured on B200, tap by tap: the tensor
// core swizzles on ABSOLUTE shared-memory address bits (as TMA does when it writes the slab), so the descriptor's
// matrix-base-offset field stays 0 - setting it to j mod 8 gives wrong products for every j not a multiple of 8.
// The ring carries W alone.
constexpr bool epi_slab(int epi) { return epi == EPI_F32_GELU_RESID; }
constexpr int SLAB_ROWS = 256;

template <int BN, int CG, int EPI>
struct GemmCfg {
  static constexpr bool SLAB = epi_slab(EPI);
  static constexpr int SLAB_CHUNK_BYTES = SLAB_ROWS * BK * 2;          // one 64-column chunk of the slab: 32 KB
  static constexpr int SLAB_BYTES = SLAB ? 2 * SLAB_CHUNK_BYTES : 0;   // at most two chunks (a_inner <= 128)
  static constexpr int A_BYTES = SLAB ? 0 : BM * BK * 2;
  static constexpr int B_BYTES = (BN / CG) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static_assert(BN <= 256, "a tile is one tcgen05.mma wide; the accumulator is double-buffered in TMEM");
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS_RAW = BN * ACC_STAGES;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128 : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int BAR_BYTES = 5120;  // barriers (<256 B), LN scratch red1 (2 KB at +256), epilogue barriers (+2304), red2 (2 KB at +2560)
  // epilogue staging: per-warp [32][20] word transpose buffers, or (residual epilogues) per warp two 64B-swizzled
  // [32 rows][RES_CW fp32] tiles that TMA fills with the residual and stores back as the output (16 columns: 32 KB
  // in all, which leaves five ring stages; 32-column tiles left four and cost FFN2, K = 5120, 3 %)
  static constexpr int RES_CW = 16;
  static constexpr int RES_TILE_BYTES = 32 * RES_CW * 4;
  static constexpr int EPI_STAGE_BYTES = epi_has_resid(EPI) ? EPI_WARPS * 2 * RES_TILE_BYTES : EPI_WARPS * 32 * 20 * 4;
  static constexpr int EPI_PARAM_BYTES = 8192;  // per-warp bias slices [8][128] f32, or bias|gamma|beta [3][512] (LN)
  static constexpr int FIXED_BYTES = BAR_BYTES + EPI_STAGE_BYTES + EPI_PARAM_BYTES + 1024 /*align slack*/;
  static constexpr int STAGES_RAW = (227 * 1024 - FIXED_BYTES - SLAB_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = SLAB_BYTES + STAGES * STAGE_BYTES + FIXED_BYTES;
};

struct TileCoord {
  int n_tile, m0, b, g;
};
template <int CG>
__device__ __forceinline__ TileCoord decode_tile(const GemmKernelParams& p, int tile, int cta_rank) {
  TileCoord t;
  t.n_tile = tile % p.n_tiles;
  int mt = tile / p.n_tiles;
  t.m0 = (mt % p.m_tiles_per_batch) * (BM * CG) + cta_rank * BM;
  mt /= p.m_tiles_per_batch;
  t.b = mt % p.batches;
  t.g = mt / p.batches;
  return t;
}

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO, const GemmKernelParams p) {
  using C = GemmCfg<BN, CG, EPI>;
  constexpr bool NSPLIT = epi_nsplit(BN, EPI);
  static_assert(!NSPLIT || CG == 1, "the N-split LayerNorm kernel runs independent cta_group::1 pipelines");
  const int cluster_rank = (CG == 2 || NSPLIT) ? (int)cluster_ctarank() : 0;
  const int cta_rank = CG == 2 ? cluster_rank : 0;   // rank within a cta_group::2 pair (row half of the pair tile)
  const bool leader = cta_rank == 0;
  const int first_tile = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr bool SLAB = C::SLAB;
  uint8_t* slab = smem_al;                       // [chunk][SLAB_ROWS][64] bf16, 128B swizzle (SLAB kernels only)
  uint8_t* smem = smem_al + C::SLAB_BYTES;       // the TMA ring
  // ring geometry: compile-time, except for the slab kernel whose stages hold only the W rows really fetched (80 of 128
  // at the 1B width: eleven 10 KB stages instead of seven of 16 KB - the ring has to cover the TMA latency with half-taps
  // of ~100 MMA cycles each)
  const int NSTAGES = SLAB ? p.slab_stages : C::STAGES;
  const int STAGE_B = SLAB ? p.slab_stage_bytes : C::STAGE_BYTES;
  uint8_t* bar_base = smem + C::STAGES * C::STAGE_BYTES;   // the ring's region is sized at compile time either way
  uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty = full + NSTAGES;
  uint64_t* tfull = empty + NSTAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* slab_full = tempty + 2;              // the tile's slab has landed / the tile's MMAs have read it
  uint64_t* slab_empty = slab_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slab_empty + 1);
  float* red1 = reinterpret_cast<float*>(bar_base + 256);  // [2][128] LN partial sums
  float* red2 = reinterpret_cast<float*>(bar_base + 2560); // [2][128]; N-split: red1 / red2 are [cta 2][half 2][128]
  uint64_t* ln_bar = reinterpret_cast<uint64_t*>(bar_base + 2304);   // N-split: [row quarter 4][pass 2]
  constexpr int STAGE_LD = 20;                             // words per staged row: 16 payload + 4 pad (80 B)
  uint32_t* stage = reinterpret_cast<uint32_t*>(bar_base + C::BAR_BYTES) + ((threadIdx.x >> 5) & 7) * (32 * STAGE_LD);
  float* epi_params = reinterpret_cast<float*>(bar_base + C::BAR_BYTES + C::EPI_STAGE_BYTES);
  uint64_t* res_full = reinterpret_cast<uint64_t*>(bar_base + 2304);   // [EPI_WARPS][2], residual epilogues only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], EPI_WARPS * CG);  // the leader's barrier collects the epilogue warps of both CTAs
    }
    mbar_init(slab_full, 1);
    mbar_init(slab_empty, 1);
    if constexpr (epi_has_resid(EPI))
      for (int a = 0; a < EPI_WARPS * 2; ++a) mbar_init(&res_full[a], 1);
    if constexpr (NSPLIT)
      for (int a = 0; a < 8; ++a) mbar_init(&ln_bar[a], 4);   // the two column-half warps of either CTA
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) {
      tmem_alloc_cg2(tmem_slot, C::TMEM_COLS);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2 || NSPLIT) cluster_sync_all(); else __syncthreads();   // peer barriers are initialised past this point
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      // This single thread must issue one stage per ~512 MMA cycles: the loop carries its coordinates incrementally
      // (no division, nothing recomputed per k-block) - a 113-instruction body with two integer divisions measured
      // ~675 cycles per iteration and starved the tensor pipe (profiles/r1_notes.md).
      uint32_t slab_it = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
        const TileCoord t = decode_tile<CG>(p, tile, cta_rank);
        const int w_row = t.n_tile * BN + (CG == 2 ? cta_rank * p.b_box_rows : 0);
        int par = 0, pos = t.m0, c0 = 0, wk = 0;   // parity / position / column of the current tap, K offset in W
        if constexpr (SLAB) {
          // the tile's input rows, once: kb_per_tap chunks of [SLAB_ROWS][64]; rows / columns beyond the tensor read as 0
          if (slab_it > 0) mbar_wait(slab_empty, (slab_it - 1) & 1);
          mbar_arrive_expect_tx(slab_full, (uint32_t)p.kb_per_tap * C::SLAB_CHUNK_BYTES);
          for (int ch = 0; ch < p.kb_per_tap; ++ch)
            tma_load_5d(slab + ch * C::SLAB_CHUNK_BYTES, &tmA, slab_full, ch * BK, 0, t.m0, t.b, t.g);
          ++slab_it;
          // W: one ring stage per TAP (its kb_per_tap 64-column chunks behind one barrier): the issuing thread pays one
          // barrier round trip and one commit per tap instead of two
          const int n_taps = p.k_blocks / p.kb_per_tap;
          for (int tap = 0; tap < n_taps; ++tap) {
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sB = smem + s * STAGE_B;
            mbar_arrive_expect_tx(&full[s], p.stage_tx_bytes * (uint32_t)p.kb_per_tap);
            for (int ch = 0; ch < p.kb_per_tap; ++ch)
              tma_load_3d(sB + ch * p.slab_chunk_bytes, &tmB, &full[s], wk + ch * BK, w_row, t.g);
            wk += p.kb_per_tap * BK;
            if (++s == NSTAGES) {
              s = 0;
              ph ^= 1;
            }
          }
        } else for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sA = smem + s * STAGE_B;
          uint8_t* sB = sA + C::A_BYTES;
          if constexpr (CG == 2) {
            // both CTAs' bytes are credited to the leader's barrier, which the leader arms for the pair
            const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
            if (leader) mbar_arrive_expect_tx(&full[s], p.stage_tx_bytes);
            tma_load_5d_cg2(sA, &tmA, bar, c0, par, pos, t.b, t.g);
            tma_load_3d_cg2(sB, &tmB, bar, wk, w_row, t.g);
          } else {
            mbar_arrive_expect_tx(&full[s], p.stage_tx_bytes);
            tma_load_5d(sA, &tmA, &full[s], c0, par, pos, t.b, t.g);
            tma_load_3d(sB, &tmB, &full[s], wk, w_row, t.g);
          }
          wk += BK;
          c0 += BK;
          if (c0 == p.kb_per_tap * BK) {   // next tap
            c0 = 0;
            if (++par == p.P) {
              par = 0;
              ++pos;
            }
          }
          if (++s == NSTAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the warp-uniform control flow (descriptor arithmetic stays in the uniform datapath);
    // one elected lane of the leader CTA issues the tcgen05 instructions.
    const bool issuer = elect_one();
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(BM * CG, p.umma_n, 0, 0);
      constexpr uint32_t desc_hi = uint32_t(1024 >> 4) | (1u << 14) | (uint32_t(SWZ_128B) << 29);  // SBO, v1, layout
      const uint32_t smem_lo = ((smem_u32(smem) & 0x3FFFF) >> 4) | (1u << 16);                      // LBO field = 1
      const uint32_t slab_lo = ((smem_u32(slab) & 0x3FFFF) >> 4) | (1u << 16);
      uint32_t slab_it = 0;
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        int kt = 0;    // k-block within the tap
        if constexpr (SLAB) {
          mbar_wait(slab_full, slab_it & 1);
          tc_fence_after();
          const int n_taps = p.k_blocks / p.kb_per_tap;
          for (int tap = 0; tap < n_taps; ++tap) {   // tap = the row of the slab its A operand starts at
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t w_lo = smem_lo + s * (STAGE_B >> 4);
            for (int ch = 0; ch < p.kb_per_tap; ++ch) {
              const uint32_t a_lo = slab_lo + ch * (C::SLAB_CHUNK_BYTES >> 4) + tap * (128 >> 4);
              const uint32_t b_lo = w_lo + ch * (p.slab_chunk_bytes >> 4);
              const int k16 = ch + 1 == p.kb_per_tap ? p.k16_last : BK / UMMA_K;
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t adesc = (uint64_t(desc_hi) << 32) | (a_lo + k * (UMMA_K * 2 >> 4));
                const uint64_t bdesc = (uint64_t(desc_hi) << 32) | (b_lo + k * (UMMA_K * 2 >> 4));
                if (issuer && k < k16) umma_ss(d_tmem, adesc, bdesc, idesc, (tap | ch | k) != 0 ? 1u : 0u);
              }
            }
            if (issuer) umma_commit(&empty[s]);   // frees the W stage when these MMAs retire
            if (++s == NSTAGES) {
              s = 0;
              ph ^= 1;
            }
          }
        }
        if constexpr (!SLAB) for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_lo = smem_lo + s * (STAGE_B >> 4);
          const uint32_t b_lo = a_lo + (C::A_BYTES >> 4);
          const int k16 = kt + 1 == p.kb_per_tap ? p.k16_last : BK / UMMA_K;   // warp-uniform
          if (++kt == p.kb_per_tap) kt = 0;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = (uint64_t(desc_hi) << 32) | (a_lo + k * (UMMA_K * 2 >> 4));
            const uint64_t bdesc = (uint64_t(desc_hi) << 32) | (b_lo + k * (UMMA_K * 2 >> 4));
            const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
            if (issuer && k < k16) {
              if constexpr (CG == 2) umma_ss_cg2(d_tmem, adesc, bdesc, idesc, acc);
              else umma_ss(d_tmem, adesc, bdesc, idesc, acc);
            }
          }
          // frees the smem stage (in both CTAs of a pair) when these MMAs retire
          if (issuer) {
            if constexpr (CG == 2) umma_commit_cg2(&empty[s], 3); else umma_commit(&empty[s]);
          }
          if (++s == NSTAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        if (issuer) {   // accumulator complete
          if constexpr (CG == 2) umma_commit_cg2(&tfull[as], 3); else umma_commit(&tfull[as]);
          if constexpr (SLAB) umma_commit(slab_empty);   // ... and the slab has been read: the next tile's may land
        }
        if constexpr (SLAB) ++slab_it;
        __syncwarp();
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    // Row-per-thread: lane l of warp w reads TMEM lane 32*(w%4) + l, so every lane of a warp works on the SAME
    // columns and the per-column parameters (bias, gamma, beta) are warp-uniform: they are staged in shared memory
    // and read back as broadcast 16-byte loads.  Arithmetic runs on packed fp32 pairs (FADD2 / FFMA2); the TMEM
    // read of chunk k+1 is in flight while chunk k is processed.
    const int q = warp & 3;          // TMEM lane quarter this warp may touch
    const int half = (warp - 4) >> 2;
    constexpr int HALF_N = BN / 2;
    constexpr bool IS_LN = EPI == EPI_LN_GELU_BF16;
    // column parameters of this warp's half tile: [HALF_N] bias (IS_LN: + gamma + beta of the whole 512-wide row)
    float* wbias = IS_LN ? epi_params + half * HALF_N : epi_params + (warp - 4) * 128;
    const ulonglong2* wbias2 = reinterpret_cast<const ulonglong2*>(wbias);
    float ln_bias_sum = 0.f;   // IS_LN: sum of the bias over this warp's columns
    if constexpr (IS_LN) {
      const int col0 = NSPLIT ? cluster_rank * BN : 0;   // N-split: this CTA's 256 of the 512 channels, every tile
      for (int i = (int)threadIdx.x - 128; i < BN; i += EPI_WARPS * 32) {
        epi_params[i] = __ldg(p.bias + col0 + i);
        epi_params[BN + i] = __ldg(p.ln_gamma + col0 + i);
        epi_params[2 * BN + i] = __ldg(p.ln_beta + col0 + i);
      }
      named_bar_sync(5, EPI_WARPS * 32);
      for (int i = lane; i < HALF_N; i += 32) ln_bias_sum += wbias[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ln_bias_sum += __shfl_xor_sync(0xffffffffu, ln_bias_sum, o);
    }
    int as = 0;
    uint32_t aph = 0;
    const uint32_t tempty_leader0 = CG == 2 ? mapa_u32(smem_u32(&tempty[0]), 0) : 0u;
    // N-split LayerNorm: the row statistics need the partial sums of four warps, two of them in the peer CTA.  A warp
    // writes its partial into the same slot of both CTAs, arrives (release.cluster) on the pass's barrier in both, and
    // waits (acquire.cluster) on its own CTA's: count 4.  Slots are reused a tile later, which the next pass's barrier
    // orders (a CTA cannot write pass k of tile i+1 before its peer has arrived at pass k' of tile i, after its reads).
    uint32_t ln_ph = 0;
    const uint32_t peer_red1 = NSPLIT ? mapa_u32(smem_u32(red1), cluster_rank ^ 1) : 0u;
    const uint32_t peer_red2 = NSPLIT ? mapa_u32(smem_u32(red2), cluster_rank ^ 1) : 0u;
    const uint32_t peer_ln_bar = NSPLIT ? mapa_u32(smem_u32(&ln_bar[q * 2]), cluster_rank ^ 1) : 0u;
    constexpr int LN_N = 2 * BN;   // channels of a row (LayerNorm epilogue: split over the two CTAs of the cluster)
    auto ln_exchange = [&](float* red, uint32_t peer_red, int pass, float part, int row) -> float {
      const int slot = (cluster_rank * 2 + half) * 128 + row;
      red[slot] = part;
      st_shared_cluster_f32(peer_red + slot * 4, part);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(peer_ln_bar + pass * 8);
        mbar_arrive(&ln_bar[q * 2 + pass]);
      }
      mbar_wait_cluster(&ln_bar[q * 2 + pass], ln_ph);
      return (red[row] + red[128 + row]) + (red[256 + row] + red[384 + row]);   // same order in both CTAs
    };
    // Residual epilogues (out-proj, FFN2, pos-conv) never touch global memory from the epilogue threads: per warp,
    // TMA brings the residual of a [32 rows x 16 columns] chunk into a 64B-swizzled shared tile one chunk ahead
    // (across tile boundaries too), the threads add accumulator + bias (+GELU) in place - row per thread, the TMEM
    // layout - and TMA stores the tile as the output.  Completion is tracked by mbarriers / bulk groups, not by the
    // warp's load scoreboards (a register prefetch of the next chunk was serialised by them: profiles/r1_notes.md).
    // out may alias resid: a chunk is loaded, then stored, by the same warp, and no other tile touches it.
    constexpr bool HAS_RESID = epi_has_resid(EPI);
    uint8_t* res_buf = reinterpret_cast<uint8_t*>(bar_base + C::BAR_BYTES) + (warp - 4) * (2 * C::RES_TILE_BYTES);
    uint64_t* my_res_full = res_full + (warp - 4) * 2;
    int res_g = 0;                 // chunks consumed so far: buffer res_g & 1, barrier parity (res_g >> 1) & 1
    int pf_tile = first_tile, pf_ci = 0;   // next chunk to request
    auto chunk_valid = [&](int tile_i, int ci) {
      return (tile_i % p.n_tiles) * BN + half * (BN / 2) + ci * C::RES_CW < p.N;
    };
    auto pf_skip_invalid = [&]() {
      while (pf_tile < p.total_tiles && !chunk_valid(pf_tile, pf_ci)) {
        if (++pf_ci == (BN / 2) / C::RES_CW) {
          pf_ci = 0;
          pf_tile += tile_step;
        }
      }
    };
    auto pf_issue = [&](int buf) {   // lane 0: request the next valid chunk into buffer `buf`
      pf_skip_invalid();
      if (pf_tile < p.total_tiles) {
        const TileCoord t = decode_tile<CG>(p, pf_tile, cta_rank);
        mbar_arrive_expect_tx(&my_res_full[buf], C::RES_TILE_BYTES);
        tma_load_4d(res_buf + buf * C::RES_TILE_BYTES, &tmR, &my_res_full[buf], t.n_tile * BN + half * (BN / 2) + pf_ci * C::RES_CW,
                    t.g, t.m0 + q * 32, t.b);
        if (++pf_ci == (BN / 2) / C::RES_CW) {
          pf_ci = 0;
          pf_tile += tile_step;
        }
      }
    };
    if constexpr (HAS_RESID) {
      if (lane == 0) pf_issue(0);
    }
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile<CG>(p, tile, cta_rank);
      const int row_in_tile = q * 32 + lane;
      const int r = tc.m0 + row_in_tile;
      const bool row_ok = r < p.rows_per_batch;
      const long long out_row = (long long)tc.b * p.out_batch_rows + r;
      const int n_base = tc.n_tile * BN + half * HALF_N;   // column within the group
      const int gcol = tc.g * p.N;                         // first output column of the group
      const uint32_t t_base = tmem_base + (uint32_t(q * 32) << 16) + as * BN + half * HALF_N;

      if constexpr (!IS_LN) {
        // stage this tile's bias slice (zero beyond N); the previous tile's readers are past their __syncwarp
        if (lane * 4 < HALF_N) {
          const int n = n_base + lane * 4;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias != nullptr && n < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol + n));
          reinterpret_cast<float4*>(wbias)[lane] = b4;
        }
        __syncwarp();
      }

      mbar_wait(&tfull[as], aph);
      tc_fence_after();

      // hands the accumulator stage back to the MMA warp; called right after the last TMEM read of the tile
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster_relaxed(tempty_leader0 + as * 8);
          else mbar_arrive(&tempty[as]);
        }
      };
      // Software-pipelined walk over the HALF_N columns in chunks of 32: body(v, c) sees chunk c in registers.
      auto for_each_chunk = [&](auto&& body, bool release_after_last_read) {
        uint32_t va[32], vb[32];
        tmem_ld32(t_base, va);
#pragma unroll
        for (int c = 0; c < HALF_N; c += 64) {
          tmem_ld_wait_on(va);
          tmem_ld32(t_base + c + 32, vb);
          body(va, c);
          tmem_ld_wait_on(vb);
          if (c + 64 < HALF_N) tmem_ld32(t_base + c + 64, va);
          else if (release_after_last_read) release_acc();
          body(vb, c + 32);
        }
      };

      // Rows of this warp: [m0 + 32q, +32).  Global stores (and residual loads) go through a warp-private
      // smem transpose so that one instruction covers 8 rows x 64 contiguous bytes instead of 32 rows x 16 B.
      const int rows_valid = min(32, max(0, p.rows_per_batch - (tc.m0 + q * 32)));
      const long long out_row0 = (long long)tc.b * p.out_batch_rows + tc.m0 + q * 32;
      const int t_r8 = lane & 7, t_piece = lane >> 3;   // transposed role: row-in-group, 16-byte piece

      // store a [32 rows x 32 cols] bf16 chunk held row-per-thread as 16 packed words
      auto store_bf16_chunk = [&](const uint32_t (&o)[16], int col0 /* within group */) {
        uint4* srow = reinterpret_cast<uint4*>(stage + lane * STAGE_LD);
#pragma unroll
        for (int j = 0; j < 4; ++j) srow[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        __syncwarp();
        const int n = col0 + t_piece * 8;
        if (p.route_n > 0) {   // warp-uniform: every row goes to the rank that owns it (peer memory)
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int row = it * 8 + t_r8;
            const uint4 val = *reinterpret_cast<const uint4*>(stage + row * STAGE_LD + t_piece * 4);
            if (row < rows_valid && n < p.N) {
              const int m = (int)out_row0 + row;
              const int o = min(m / p.route_per, p.route_n - 1);
              *reinterpret_cast<uint4*>(p.route_base[o] + (long long)(m - o * p.route_per) * p.ldo + gcol + n) = val;
            }
          }
        } else {
          __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row0 * p.ldo + gcol + n;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int row = it * 8 + t_r8;
            const uint4 val = *reinterpret_cast<const uint4*>(stage + row * STAGE_LD + t_piece * 4);
            if (row < rows_valid && n < p.N) *reinterpret_cast<uint4*>(obase + (long long)row * p.ldo) = val;
          }
        }
        __syncwarp();
      };

      if constexpr (IS_LN) {
        // LayerNorm over the 512 channels of the row: 3 passes over TMEM, partner warp holds the other half.
        f32x2 acc0 = 0ull, acc1 = 0ull;
        for_each_chunk([&](const uint32_t (&v)[32], int) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            acc0 = f2_add(acc0, f2_pack_u(v[j], v[j + 1]));
            acc1 = f2_add(acc1, f2_pack_u(v[j + 2], v[j + 3]));
          }
        }, false);
        float sa, sb;
        f2_unpack(f2_add(acc0, acc1), sa, sb);
        const float mean = ln_exchange(red1, peer_red1, 0, sa + sb + ln_bias_sum, row_in_tile) * (1.0f / LN_N);
        const f32x2 nmean2 = f2_splat(-mean);
        acc0 = 0ull;
        acc1 = 0ull;
        for_each_chunk([&](const uint32_t (&v)[32], int c) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const ulonglong2 b = wbias2[(c + j) >> 2];
            const f32x2 d0 = f2_add(f2_add(f2_pack_u(v[j], v[j + 1]), b.x), nmean2);
            const f32x2 d1 = f2_add(f2_add(f2_pack_u(v[j + 2], v[j + 3]), b.y), nmean2);
            acc0 = f2_fma(d0, d0, acc0);
            acc1 = f2_fma(d1, d1, acc1);
          }
        }, false);
        f2_unpack(f2_add(acc0, acc1), sa, sb);
        const float var = ln_exchange(red2, peer_red2, 1, sa + sb, row_in_tile) * (1.0f / LN_N);
        ln_ph ^= 1;
        const f32x2 rstd2 = f2_splat(rsqrtf(var + 1e-5f));
        const ulonglong2* wgamma2 = reinterpret_cast<const ulonglong2*>(wbias + BN);
        const ulonglong2* wbeta2 = reinterpret_cast<const ulonglong2*>(wbias + 2 * BN);
        for_each_chunk([&](const uint32_t (&v)[32], int c) {
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const ulonglong2 b = wbias2[(c + j) >> 2], g = wgamma2[(c + j) >> 2], be = wbeta2[(c + j) >> 2];
            const f32x2 d0 = f2_add(f2_add(f2_pack_u(v[j], v[j + 1]), b.x), nmean2);
            const f32x2 d1 = f2_add(f2_add(f2_pack_u(v[j + 2], v[j + 3]), b.y), nmean2);
            float y0, y1, y2, y3;
            gelu_erf_x2(f2_fma(d0, f2_mul(g.x, rstd2), be.x), y0, y1);
            gelu_erf_x2(f2_fma(d1, f2_mul(g.y, rstd2), be.y), y2, y3);
            o[j >> 1] = pack_bf16x2(y0, y1);
            o[(j >> 1) + 1] = pack_bf16x2(y2, y3);
          }
          store_bf16_chunk(o, n_base + c);
        }, true);
      } else if constexpr (EPI == EPI_ARGMAX) {
        float best = -INFINITY;
        int best_i = 0x7fffffff;
        for_each_chunk([&](const uint32_t (&v)[32], int c) {
          const int n0 = n_base + c;
          if (n0 < p.N) {   // warp-uniform
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + j;
              const float x = __uint_as_float(v[j]) + wbias[c + j];
              if (n < p.N && x > best) {  // strict: lowest index wins among equals (torch.argmax)
                best = x;
                best_i = n;
              }
            }
          }
        }, true);
        if (row_ok && best_i != 0x7fffffff) atomicMax(p.argmax + out_row, argmax_pack(best, best_i));
      } else if constexpr (EPI == EPI_BF16 || EPI == EPI_BF16_GELU) {
        for_each_chunk([&](const uint32_t (&v)[32], int c) {
          if (n_base + c < p.N) {   // warp-uniform; columns >= N of a partial chunk are computed but not stored
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const ulonglong2 b = wbias2[(c + j) >> 2];
              const f32x2 x01 = f2_add(f2_pack_u(v[j], v[j + 1]), b.x);
              const f32x2 x23 = f2_add(f2_pack_u(v[j + 2], v[j + 3]), b.y);
              float y0, y1, y2, y3;
              if constexpr (EPI == EPI_BF16_GELU) {
                gelu_erf_x2(x01, y0, y1);
                gelu_erf_x2(x23, y2, y3);
              } else {
                f2_unpack(x01, y0, y1);
                f2_unpack(x23, y2, y3);
              }
              o[j >> 1] = pack_bf16x2(y0, y1);
              o[(j >> 1) + 1] = pack_bf16x2(y2, y3);
            }
            store_bf16_chunk(o, n_base + c);
          }
        }, true);
      } else {
        if constexpr (HAS_RESID) {
          constexpr int CW = C::RES_CW;
          constexpr int NCH = HALF_N / CW;
#pragma unroll 1
          for (int ci = 0; ci < NCH; ++ci) {
            const int c = ci * CW;
            uint32_t v[CW];
            tmem_ld16(t_base + c, v);
            if (n_base + c < p.N) {   // warp-uniform
              const int buf = res_g & 1;
              if (lane == 0) {
                // the other buffer was handed to a TMA store one chunk ago: wait until that store has read it, then
                // request the next chunk into it
                bulk_wait_group_read<0>();
                pf_issue(buf ^ 1);
              }
              tmem_ld_wait();
              if (ci + 1 == NCH) release_acc();
              mbar_wait(&my_res_full[buf], (res_g >> 1) & 1);
              uint8_t* row = res_buf + buf * C::RES_TILE_BYTES + lane * (CW * 4);
#pragma unroll
              for (int k = 0; k < CW / 4; ++k) {
                float4* cell = reinterpret_cast<float4*>(row + ((k ^ ((lane >> 1) & 3)) << 4));   // 64B swizzle
                const float4 r4 = *cell;
                const float4 b4 = reinterpret_cast<const float4*>(wbias)[(c >> 2) + k];
                float4 a;
                a.x = __uint_as_float(v[4 * k]) + b4.x;
                a.y = __uint_as_float(v[4 * k + 1]) + b4.y;
                a.z = __uint_as_float(v[4 * k + 2]) + b4.z;
                a.w = __uint_as_float(v[4 * k + 3]) + b4.w;
                if constexpr (EPI == EPI_F32_GELU_RESID) {
                  gelu_erf_x2(f2_pack(a.x, a.y), a.x, a.y);
                  gelu_erf_x2(f2_pack(a.z, a.w), a.z, a.w);
                }
                a.x += r4.x; a.y += r4.y; a.z += r4.z; a.w += r4.w;
                *cell = a;
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&tmO, res_buf + buf * C::RES_TILE_BYTES, n_base + c, tc.g, tc.m0 + q * 32, tc.b);
                bulk_commit_group();
              }
              ++res_g;
            } else {
              tmem_ld_wait();
              if (ci + 1 == NCH) release_acc();
            }
          }
        } else {
        // fp32 output without residual (EPI_F32), 16 columns per transpose round
#pragma unroll 1
        for (int c = 0; c < HALF_N; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_base + c, v);
          tmem_ld_wait();
          if (c + 32 >= HALF_N) release_acc();
          if (n_base + c < p.N) {   // warp-uniform
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              float4* srow = reinterpret_cast<float4*>(stage + lane * STAGE_LD);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                srow[j] = make_float4(__uint_as_float(v[hh * 16 + 4 * j]), __uint_as_float(v[hh * 16 + 4 * j + 1]),
                                      __uint_as_float(v[hh * 16 + 4 * j + 2]), __uint_as_float(v[hh * 16 + 4 * j + 3]));
              __syncwarp();
              const int n = n_base + c + hh * 16 + t_piece * 4;
              const bool col_ok = n < p.N;
              const float4 b4 = reinterpret_cast<const float4*>(wbias)[(c + hh * 16 + t_piece * 4) >> 2];
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int row = it * 8 + t_r8;
                float4 a = *reinterpret_cast<const float4*>(stage + row * STAGE_LD + t_piece * 4);
                if (row < rows_valid && col_ok) {
                  const long long grow = out_row0 + row;
                  a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
                  if (p.n_valid != nullptr) {
                    const int seq = int(grow / p.frames_per_seq);
                    const int t = int(grow - (long long)seq * p.frames_per_seq);
                    if (t >= __ldg(p.n_valid + seq)) a = make_float4(0.f, 0.f, 0.f, 0.f);
                  }
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + grow * p.ldo + gcol + n) = a;
                }
              }
              __syncwarp();
            }
          }
        }
        }
      }

      __syncwarp();   // every lane is done with this tile's bias slice
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
    if constexpr (HAS_RESID) {
      if (lane == 0) bulk_wait_group_read<0>();   // shared memory must outlive the last TMA store's read
    }
  }

  tc_fence_before();
  if constexpr (CG == 2 || NSPLIT) cluster_sync_all(); else __syncthreads();   // nobody touches a peer's smem/TMEM past this
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct TmapKey {
  const void* base;
  int rank;
  bool f32;
  CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
  uint64_t dims[5];
  uint64_t strides[4];
  uint32_t box[5];
  bool operator<(const TmapKey& o) const {
    if (base != o.base) return base < o.base;
    if (rank != o.rank) return rank < o.rank;
    if (f32 != o.f32) return f32 < o.f32;
    if (swizzle != o.swizzle) return swizzle < o.swizzle;
    for (int i = 0; i < 5; ++i) if (dims[i] != o.dims[i]) return dims[i] < o.dims[i];
    for (int i = 0; i < 4; ++i) if (strides[i] != o.strides[i]) return strides[i] < o.strides[i];
    for (int i = 0; i < 5; ++i) if (box[i] != o.box[i]) return box[i] < o.box[i];
    return false;
  }
};
std::map<TmapKey, CUtensorMap> g_tmap_cache;
std::mutex g_tmap_mu;

int cached_tmap(CUtensorMap* out, const TmapKey& key) {
  std::lock_guard<std::mutex> g(g_tmap_mu);
  auto it = g_tmap_cache.find(key);
  if (it != g_tmap_cache.end()) {
    *out = it->second;
    return OASR_OK;
  }
  OASR_TRY(make_tmap(out, key.base, key.f32, key.rank, key.dims, key.strides, key.box, key.swizzle));
  if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
  g_tmap_cache[key] = *out;
  return OASR_OK;
}

inline uint64_t nz(long long v, uint64_t fallback) { return v > 0 ? (uint64_t)v : fallback; }

template <int BN, int EPI, int CG>
int launch_cg(const GemmArgs& a, cudaStream_t stream) {
  using C = GemmCfg<BN, CG, EPI>;
  const int k_pad = a.k_pad > 0 ? a.k_pad : ((a.a_inner + BK - 1) / BK) * BK;
  GemmKernelParams p;
  p.rows_per_batch = a.rows_per_batch;
  p.batches = a.batches;
  p.groups = a.groups;
  p.N = a.N;
  p.m_tiles_per_batch = (a.rows_per_batch + BM * CG - 1) / (BM * CG);
  p.n_tiles = (a.N + BN - 1) / BN;
  p.total_tiles = p.m_tiles_per_batch * a.batches * a.groups * p.n_tiles;
  p.kb_per_tap = k_pad / BK;
  p.k_blocks = p.kb_per_tap * a.taps;
  {
    const int tail = a.a_inner - (p.kb_per_tap - 1) * BK;   // valid columns of a tap's last k-block (<= BK)
    p.k16_last = tail >= BK ? BK / UMMA_K : (tail + UMMA_K - 1) / UMMA_K;
    if (p.k16_last < 1) p.k16_last = 1;
  }
  p.P = a.P;
  const int n_cap = a.N < BN ? ((a.N + 15) / 16) * 16 : BN;   // columns one tile really computes
  p.umma_n = n_cap;
  p.b_box_rows = p.umma_n / CG;   // W rows one CTA fetches per MMA-N chunk
  p.stage_tx_bytes = (uint32_t)CG * (C::A_BYTES + (uint32_t)p.b_box_rows * BK * 2);
  // slab kernel: a ring stage holds one TAP of W = kb_per_tap chunks of the rows actually fetched (80 of 128 at the 1B
  // width), each chunk aligned to the 1024-byte swizzle repeat
  p.slab_chunk_bytes = ((p.b_box_rows * BK * 2 + 1023) / 1024) * 1024;
  p.slab_stage_bytes = p.slab_chunk_bytes * p.kb_per_tap;
  p.slab_stages = C::SLAB ? std::min(12, (C::STAGES * C::STAGE_BYTES) / p.slab_stage_bytes) : 0;
  p.ldo = a.ldo;
  p.out_batch_rows = a.out_batch_rows > 0 ? a.out_batch_rows : a.rows_per_batch;
  p.out = a.out;
  p.bias = a.bias;
  p.resid = a.resid;
  p.ln_gamma = a.ln_gamma;
  p.ln_beta = a.ln_beta;
  p.argmax = a.argmax;
  p.n_valid = a.n_valid;
  p.frames_per_seq = a.frames_per_seq;
  p.route_n = a.route_n;
  p.route_per = a.route_per;
  for (int i = 0; i < 8; ++i) p.route_base[i] = reinterpret_cast<__nv_bfloat16*>(a.route_base[i]);
  if (p.total_tiles == 0) return OASR_OK;

  // A: {a_inner, P, U, batches, groups}; a unit-extent dimension still needs a legal (16-byte multiple) stride
  TmapKey ka{};
  ka.base = a.A;
  ka.rank = 5;
  const uint64_t pos_stride_b = (uint64_t)a.a_pos_stride * 2;
  ka.dims[0] = (uint64_t)a.a_inner; ka.dims[1] = (uint64_t)a.P; ka.dims[2] = (uint64_t)a.a_positions;
  ka.dims[3] = (uint64_t)a.batches; ka.dims[4] = (uint64_t)a.groups;
  ka.strides[0] = nz(a.a_p_stride * 2, pos_stride_b);
  ka.strides[1] = pos_stride_b;
  ka.strides[2] = nz(a.a_batch_stride * 2, pos_stride_b * (uint64_t)a.a_positions);
  ka.strides[3] = nz(a.a_group_stride * 2, ka.strides[2] * (uint64_t)a.batches);
  ka.box[0] = BK; ka.box[1] = 1; ka.box[2] = C::SLAB ? SLAB_ROWS : BM; ka.box[3] = 1; ka.box[4] = 1;
  // W: {taps*k_pad, N, groups}
  TmapKey kw{};
  kw.base = a.W;
  kw.rank = 3;
  const uint64_t wk = (uint64_t)a.taps * k_pad;
  kw.dims[0] = wk; kw.dims[1] = (uint64_t)a.N; kw.dims[2] = (uint64_t)a.groups;
  kw.strides[0] = wk * 2;
  kw.strides[1] = wk * 2 * (uint64_t)a.N;
  kw.box[0] = BK; kw.box[1] = (uint32_t)p.b_box_rows; kw.box[2] = 1;
  CUtensorMap tmA, tmB;
  OASR_TRY(cached_tmap(&tmA, ka));
  OASR_TRY(cached_tmap(&tmB, kw));
  CUtensorMap tmR = tmA, tmO = tmA;   // placeholders unless the epilogue moves fp32 tiles through TMA
  if constexpr (epi_has_resid(EPI)) {
    // fp32 {N, groups, rows_per_batch, batches}: a [32 rows x 16 columns] box is clipped at the group's last column
    // and at the batch's last row (no write beyond either), 64B swizzle
    const void* bases[2] = {a.resid, a.out};
    CUtensorMap* maps[2] = {&tmR, &tmO};
    for (int i = 0; i < 2; ++i) {
      TmapKey kr{};
      kr.base = bases[i];
      kr.rank = 4;
      kr.f32 = true;
      kr.dims[0] = (uint64_t)a.N; kr.dims[1] = (uint64_t)a.groups; kr.dims[2] = (uint64_t)a.rows_per_batch;
      kr.dims[3] = (uint64_t)a.batches;
      kr.strides[0] = (uint64_t)a.N * 4;
      kr.strides[1] = (uint64_t)a.ldo * 4;
      kr.strides[2] = (uint64_t)p.out_batch_rows * a.ldo * 4;
      kr.box[0] = C::RES_CW; kr.box[1] = 1; kr.box[2] = 32; kr.box[3] = 1;
      kr.swizzle = CU_TENSOR_MAP_SWIZZLE_64B;
      OASR_TRY(cached_tmap(maps[i], kr));
    }
  }

  static unsigned long long attr_mask = 0;
  if (first_use_on_this_device(&attr_mask))
    OASR_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<BN, EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES));
  constexpr bool NSPLIT = epi_nsplit(BN, EPI);
  const int units = device_sm_count() / CG;   // CTAs or CTA pairs that can be resident
  int grid_units = p.total_tiles < units ? p.total_tiles : units;
  if (NSPLIT) grid_units &= ~1;   // clusters of two CTAs: tile 2m + r goes to the CTA of rank r (total_tiles is even)
  if constexpr (CG == 2 || NSPLIT) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid_units * CG);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OASR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_kernel<BN, EPI, CG>, tmA, tmB, tmR, tmO, p));
  } else {
    gemm_kernel<BN, EPI, CG><<<grid_units, GEMM_THREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, tmR, tmO, p);
  }
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

// CTA pairs whenever a tile spans full 256-column MMAs (every encoder / FE / CTC GEMM), single CTAs for narrower outputs
// (pos-conv groups, the small test shapes)
// A single 30 s window (M = 1499) gives the encoder GEMMs too few 256 x 256 pair tiles for 74 CTA pairs: out-proj and FFN2
// 30 tiles (41 % of the SMs busy), QKV 90 (two waves, the second a quarter full).  When the pair tiling fills less
// than ~two waves and 128 x 128 single-CTA tiles fill their waves better, the small tiles run instead (81 % for all four
// encoder shapes at M = 1499); the epilogues are the same code at BN = 128.
template <int EPI>
constexpr bool has_small_tile() { return EPI == EPI_BF16 || EPI == EPI_BF16_GELU || EPI == EPI_F32_RESID; }

inline bool prefer_small_tiles(const GemmArgs& a) {
  if (a.N < 256 || a.taps != 1 || a.groups != 1) return false;
  const long long sms = device_sm_count();
  const long long mp = (a.rows_per_batch + 255) / 256, np = (a.N + 255) / 256;
  const long long ms = (a.rows_per_batch + 127) / 128, ns = (a.N + 127) / 128;
  const long long tp = mp * np * a.batches, ts = ms * ns * a.batches;
  const long long pairs = sms / 2;
  if (tp > 2 * pairs) return false;
  const double eff_p = (double)tp / (double)(((tp + pairs - 1) / pairs) * pairs);
  const double eff_s = (double)ts / (double)(((ts + sms - 1) / sms) * sms);
  return eff_s > 1.15 * eff_p;
}

template <int BN, int EPI>
int launch(const GemmArgs& a, cudaStream_t stream) {
  if constexpr (BN == 256 && has_small_tile<EPI>()) {
    if (prefer_small_tiles(a)) return launch_cg<128, EPI, 1>(a, stream);
  }
  if (BN >= 256 && a.N >= BN) return launch_cg<BN, EPI, 2>(a, stream);
  return launch_cg<BN, EPI, 1>(a, stream);
}

}  // namespace

int gemm_bf16_tcgen05(const GemmArgs& a, cudaStream_t stream) {
  OASR_REQUIRE(a.A && a.W && a.rows_per_batch >= 0 && a.batches >= 1 && a.groups >= 1 && a.N > 0 && a.a_inner > 0 &&
                   a.taps >= 1 && a.P >= 1 && a.a_positions > 0,
               "gemm: bad arguments");
  OASR_REQUIRE(a.a_inner % 8 == 0, "gemm: inner extent must be a multiple of 8 (16-byte TMA rows)");
  OASR_REQUIRE(a.k_pad % 64 == 0, "gemm: k_pad must be a multiple of 64");
  OASR_REQUIRE(a.a_pos_stride % 8 == 0 && a.a_p_stride % 8 == 0 && a.a_batch_stride % 8 == 0 &&
                   a.a_group_stride % 8 == 0 && a.a_pos_stride > 0,
               "gemm: A strides must be multiples of 8 elements");
  OASR_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  OASR_REQUIRE(a.N % 8 == 0 || (a.epilogue == EPI_ARGMAX && a.N % 4 == 0), "gemm: N must be a multiple of 8 (4 for arg-max)");
  OASR_REQUIRE((reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "gemm: bias must be 16-byte aligned");
  if (a.route_n > 0) {
    OASR_REQUIRE(a.epilogue == EPI_BF16 && a.batches == 1 && a.groups == 1 && a.taps == 1 && a.route_n <= 8 &&
                     a.route_per > 0 && a.out_batch_rows == 0,
                 "gemm: routed rows need a plain bf16 GEMM, at most 8 owners and a positive share");
    for (int i = 0; i < a.route_n; ++i)
      OASR_REQUIRE(a.route_base[i] != nullptr && (reinterpret_cast<uintptr_t>(a.route_base[i]) & 15) == 0,
                   "gemm: routed destinations must be 16-byte aligned");
  }
  if (a.epilogue != EPI_ARGMAX) {
    OASR_REQUIRE((a.out != nullptr || a.route_n > 0) && a.ldo >= a.N * a.groups, "gemm: output missing");
    OASR_REQUIRE(a.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0,
                 "gemm: output rows must be 16-byte aligned");
  }
  switch (a.epilogue) {
    case EPI_BF16: return launch<256, EPI_BF16>(a, stream);
    case EPI_BF16_GELU: return launch<256, EPI_BF16_GELU>(a, stream);
    case EPI_F32: return launch<256, EPI_F32>(a, stream);
    case EPI_F32_RESID:
      OASR_REQUIRE(a.resid != nullptr, "gemm: residual missing");
      return launch<256, EPI_F32_RESID>(a, stream);
    case EPI_F32_GELU_RESID:
      OASR_REQUIRE(a.resid != nullptr && a.N <= 128, "gemm: gelu+residual epilogue needs a residual and N <= 128");
      // the slab kernel: a tile's taps read rows [m0, m0 + 128 + taps - 1) of ONE parity, at most two 64-column chunks
      OASR_REQUIRE(a.P == 1 && a.taps + BM - 1 <= SLAB_ROWS && a.a_inner <= 2 * BK && (a.k_pad == 0 || a.k_pad <= 2 * BK),
                   "gemm: gelu+residual (pos-conv) epilogue needs stride 1, at most 129 taps and at most 128 input channels per group");
      return launch<128, EPI_F32_GELU_RESID>(a, stream);
    case EPI_ARGMAX:
      OASR_REQUIRE(a.argmax != nullptr && a.bias != nullptr && a.groups == 1, "gemm: argmax buffer / bias missing");
      return launch<256, EPI_ARGMAX>(a, stream);
    case EPI_LN_GELU_BF16:
      OASR_REQUIRE(a.N == 512 && a.groups == 1 && a.bias && a.ln_gamma && a.ln_beta,
                   "gemm: LN epilogue needs N == 512 and bias/gamma/beta");
      return launch_cg<256, EPI_LN_GELU_BF16, 1>(a, stream);
    default: return fail(OASR_ERR_INVALID, "gemm: unknown epilogue");
  }
}

}  // namespace oasr
