// Bidirectional self-attention on tcgen05, seventh version (a15), head_dim <= 80 (300M, 1B): a PERSISTENT kernel, one
// CTA per SM walking a static list of (query group, head, window) work items; NT query tiles of 128 rows per CTA, one
// MMA-issuing warp per tile, the K/V ring, the barrier phases and the tile stagger running on across items.
//
// Shape: NT = 3 query tiles, BKV = 48 keys per block (template parameters).  Round 2 also built and measured S
// double-buffered with 2 x 64-key and 3 x 32-key blocks, and four tiles with 32-key blocks (all of TMEM): same outputs,
// 11 - 24 % slower (fewer warps, or more hand-offs per key), so they are not in the tree (profiles/r2_notes.md).
// TMEM columns: S[tile] (BKV each) | P[tile] (BKV/2 used, rounded to 16) | O[tile] (head_dim each).
// Warps: 4 per tile softmax + epilogue (warp w owns TMEM lanes [32(w%4), +32)), then the TMA producer, then one MMA
// issuer per tile (the first also allocates TMEM).
#include "host_util.h"
#include "kernels.cuh"
#include "attention_common.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <map>
#include <type_traits>
#include <mutex>

namespace oasr {
namespace {

using namespace att;

constexpr int BQ = 128;
__host__ __device__ constexpr int att_threads(int nt) { return (nt * 4 + 1 + nt) * 32; }
constexpr int MAX_KV_STAGES = 8;
// TMEM columns a CTA allocates: a power of two covering S + P + O of its NT tiles (one tile: 256, so that two such CTAs
// share an SM; three tiles: all 512)
__host__ __device__ constexpr int att_tmem_cols(int nt, int bkv, int hd) {
  const int need = nt * bkv + nt * ((bkv / 2 + 15) & ~15) + nt * hd;
  return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}
constexpr int KV_PREFETCH = 4;             // K/V blocks of the next work item requested before its Q
constexpr int START_OFFSET_CYCLES = 450;   // tile X issues its first S this many cycles after tile X-1

struct Attn7Params {
  int n_qt, n_items;   // query triples per (head, window); work items = n_qt * H * B
  int kv_stages;
  int T, H, d;
  float scale_log2e;
  const int* n_frames;
  __nv_bfloat16* out;
  long long* trace;   // debug: SM-clock timestamps of CTA (0,0,0), [role][event] (OASR_ATT_TRACE=file)
  int start_offset;   // tile X issues its first S this many cycles after tile X-1
  int relay;          // exponential phases in relay (see the softmax warps): 0 off, 1 / 2: hand on after the first / second 16-column piece
};
constexpr int TRACE_EVENTS = 256;   // per role: 0 MMA warp, 1 + X: first softmax warp of tile X
// Tracing is a compile-time option (-DOASR_ATT_TRACING): even a never-taken stamp costs the softmax warps a branch,
// and six of them per key block were ~10 % of the kernel.
#ifdef OASR_ATT_TRACING
#define ATT_TRACE(role, ev)                                                                                 \
  do {                                                                                                      \
    if (p.trace != nullptr && blockIdx.x == 0 && (ev) < TRACE_EVENTS) \
      p.trace[(role) * TRACE_EVENTS + (ev)] = clock64();                                                    \
  } while (0)
#else
#define ATT_TRACE(role, ev) \
  do {                      \
  } while (0)
#endif

template <int HD, int NT, int BKV>
__global__ void __launch_bounds__(att_threads(NT), NT == 1 ? 2 : 1)
attention_v7_kernel(const __grid_constant__ CUtensorMap tmq64, const __grid_constant__ CUtensorMap tmq32,
                    const __grid_constant__ CUtensorMap tmq16, const __grid_constant__ CUtensorMap tmk64,
                    const __grid_constant__ CUtensorMap tmk32, const __grid_constant__ CUtensorMap tmk16,
                    const __grid_constant__ CUtensorMap tmo, const Attn7Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  static_assert(BKV == 32 || BKV == 48 || BKV == 64, "key block");
  constexpr int PSLOT = round16(BKV / 2);
  constexpr int TM_S = 0, TM_P = NT * BKV, TM_O = TM_P + NT * PSLOT;
  constexpr int TMEM_COLS = att_tmem_cols(NT, BKV, HD);
  static_assert(TM_O + NT * HD <= TMEM_COLS && TMEM_COLS <= 512, "TMEM budget");
  constexpr int NQK = qk_nchunks(HD);
  constexpr int VW = v_w(HD);
  constexpr int NV = HD / VW;
  constexpr int q_tile_bytes = BQ * HD * 2;
  constexpr int kv_tile_bytes = BKV * HD * 2;
  const int KS = p.kv_stages;
  // Barriers first, at fixed offsets from the aligned base (the ring length is a run-time value: anything placed
  // behind it has an address the compiler re-derives from kernel parameters at every use once registers are short,
  // which put ~100 cycles of dependent latency in front of every barrier operation of the softmax warps).
  // Per tile X, bars[TB X + k]: k = 0 s_full (S_X(j) is in TMEM), 1 s_free (S_X(j) has been read into registers),
  // 2 p_full (P_X(j) is in TMEM), 3 o_done (P.V_X(j) has retired: O updated, P buffer free), 4 first_s (tile X has
  // issued its first S of the work item), 5 q_full (Q_X has landed), 6 q_empty (tile X's last S of the item is done).
  // The parity of a barrier for block g of a tile (counted across work items) is g & 1.
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  constexpr int TB = 8;
  constexpr int S_FULL = 0, S_FREE = 1, P_FULL = 2, O_DONE = 3, FIRST_S = 4, Q_FULL = 5, Q_EMPTY = 6;
  uint64_t* kv_full = bars + TB * NT;             // MAX_KV_STAGES
  uint64_t* kv_empty = kv_full + MAX_KV_STAGES;   // MAX_KV_STAGES
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + MAX_KV_STAGES);
  static_assert((TB * NT + 2 * MAX_KV_STAGES + 1) * 8 <= 1024, "barrier block");
  uint8_t* sQ = smem + 1024;              // [NT tiles]
  uint8_t* sKV = sQ + NT * q_tile_bytes;  // [stage][K | V]
  uint8_t* sO = sKV + KS * 2 * kv_tile_bytes;   // [softmax warp][32 rows][HD] bf16: staging of the output TMA stores

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == NT * 4 && lane == 0) {
    tma_prefetch_desc(&tmq64);
    tma_prefetch_desc(&tmk64);
    tma_prefetch_desc(&tmq16);
    tma_prefetch_desc(&tmk16);
    tma_prefetch_desc(&tmo);
    for (int i = 0; i < NT; ++i) {
      mbar_init(&bars[TB * i + Q_FULL], 1);
      mbar_init(&bars[TB * i + Q_EMPTY], 1);
    }
    for (int i = 0; i < MAX_KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], NT);  // one commit per tile's MMA issuer
    }
    for (int i = 0; i < NT; ++i) {
      mbar_init(&bars[TB * i + S_FULL], 1);
      mbar_init(&bars[TB * i + S_FREE], 4);
      mbar_init(&bars[TB * i + P_FULL], 4);
      mbar_init(&bars[TB * i + O_DONE], 1);
      mbar_init(&bars[TB * i + FIRST_S], 1);
    }
    fence_barrier_init();
  }
  if (warp == NT * 4 + 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Work items: (query triple qt, head h, window b), qt fastest so that the CTAs running side by side share K/V in L2.
  // Every role walks the same static sequence blockIdx.x, blockIdx.x + gridDim.x, ... and derives the same block
  // counts, so the mbarrier phases (which run on across items) stay in step without any further hand-shake.  Items of
  // a fully padded window (no keys) touch no barrier: the softmax warps write their zeros and move on.
  auto item = [&](int w, int& q0, int& h, int& b, int& n_keys) {
    const int qt = w % p.n_qt;
    const int hb = w / p.n_qt;
    h = hb % p.H;
    b = hb / p.H;
    q0 = qt * (NT * BQ);
    n_keys = min(p.n_frames ? p.n_frames[b] : p.T, p.T);
  };

  if (warp == NT * 4) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      auto qmap = [&](int w) { return w == 64 ? &tmq64 : (w == 32 ? &tmq32 : &tmq16); };
      auto kmap = [&](int w) { return w == 64 ? &tmk64 : (w == 32 ? &tmk32 : &tmk16); };
      int s = 0;
      uint32_t ph = 0, it = 0;
      for (int w = blockIdx.x; w < p.n_items; w += gridDim.x) {
        int q0, h, b, n_keys;
        item(w, q0, h, b, n_keys);
        const int nblk = (n_keys + BKV - 1) / BKV;
        if (nblk == 0) continue;
        const int qcol = h * HD, kcol = p.d + h * HD, vcol = 2 * p.d + h * HD;
        // The first K/V blocks go out before Q: the ring has room for them long before the tiles finish the previous
        // item, and a tile whose Q arrives can start at once.  Q is per tile (own buffer, own barriers), so that the
        // tiles keep the distance they started with instead of meeting at every item boundary.
        auto load_kv = [&](int j) {
          mbar_wait(&kv_empty[s], ph ^ 1);
          uint8_t* sK = sKV + s * 2 * kv_tile_bytes;
          mbar_arrive_expect_tx(&kv_full[s], 2 * kv_tile_bytes);
#pragma unroll
          for (int c = 0; c < NQK; ++c)    // K: same chunking as Q
            tma_load_3d(sK + 2 * BKV * qk_col(HD, c), kmap(qk_w(HD, c)), &kv_full[s], kcol + qk_col(HD, c), j * BKV, b);
#pragma unroll
          for (int c = 0; c < NV; ++c)     // V: NV uniform chunks of VW columns
            tma_load_3d(sK + kv_tile_bytes + c * (2 * BKV * VW), kmap(VW), &kv_full[s], vcol + c * VW, j * BKV, b);
          if (++s == KS) {
            s = 0;
            ph ^= 1;
          }
        };
        const int pre = min(nblk, KV_PREFETCH);
        ATT_TRACE(4, it * 8);
        for (int j = 0; j < pre; ++j) load_kv(j);
        ATT_TRACE(4, it * 8 + 1);
#pragma unroll
        for (int X = 0; X < NT; ++X) {
          if (it > 0) mbar_wait(&bars[TB * X + Q_EMPTY], (it - 1) & 1);   // tile X's S MMAs of the previous item have read Q_X
          mbar_arrive_expect_tx(&bars[TB * X + Q_FULL], q_tile_bytes);
#pragma unroll
          for (int c = 0; c < NQK; ++c)
            tma_load_3d(sQ + X * q_tile_bytes + 2 * BQ * qk_col(HD, c), qmap(qk_w(HD, c)), &bars[TB * X + Q_FULL],
                        qcol + qk_col(HD, c), q0 + X * BQ, b);
          ATT_TRACE(4, it * 8 + 2 + X);
        }
        ++it;
        // the next item usually belongs to another (head, window): its Q and first K/V blocks would come from HBM
        // when the tiles are already waiting for them.  Ask for them now, a whole item ahead, into L2 only.
        if (w + (int)gridDim.x < p.n_items) {
          int q0n, hn, bn, nkn;
          item(w + gridDim.x, q0n, hn, bn, nkn);
          const int nblkn = (nkn + BKV - 1) / BKV;
          if (nblkn > 0) {
#pragma unroll
            for (int X = 0; X < NT; ++X)
#pragma unroll
              for (int c = 0; c < NQK; ++c)
                tma_prefetch_l2_3d(qmap(qk_w(HD, c)), hn * HD + qk_col(HD, c), q0n + X * BQ, bn);
            for (int j = 0; j < min(nblkn, KV_PREFETCH); ++j) {
#pragma unroll
              for (int c = 0; c < NQK; ++c)
                tma_prefetch_l2_3d(kmap(qk_w(HD, c)), p.d + hn * HD + qk_col(HD, c), j * BKV, bn);
#pragma unroll
              for (int c = 0; c < NV; ++c) tma_prefetch_l2_3d(kmap(VW), 2 * p.d + hn * HD + c * VW, j * BKV, bn);
            }
          }
        }
        for (int j = pre; j < nblk; ++j) load_kv(j);
      }
    }
  } else if (warp > NT * 4) {
    // ---------------------------------------------------------------- MMA issuers: one warp per query tile
    // (see attention_v6.cu for why each tile has its own issuing warp.)  Across work items: the first S of an item
    // waits for Q, for its K tile and for the softmax warps to have read the last S of the previous item; the first
    // P.V of an item overwrites O, which the softmax warps of the same tile have read (their epilogue) before they
    // hand over the first P of the new item, so p_full orders that too.  After its last S of an item an issuer
    // commits to q_empty: the producer may then overwrite Q.
    const int X = warp - (NT * 4 + 1);
    const bool issuer = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0, 1);  // B = V is MN-major
    const uint32_t sq_lo = (smem_u32(sQ) & 0x3FFFF) >> 4;        // descriptor start-address fields (16-byte units)
    const uint32_t skv_lo = (smem_u32(sKV) & 0x3FFFF) >> 4;
    auto issue_s = [&](int st) {   // S_X = Q_X K^T for the K tile in stage st: HD/16 MMAs
      const uint32_t q_lo = sq_lo + X * (q_tile_bytes >> 4) + (1u << 16);            // LBO field = 1 (unused)
      const uint32_t k_lo = skv_lo + st * (2 * kv_tile_bytes >> 4) + (1u << 16);
      const uint32_t d_tmem = tmem_base + TM_S + X * BKV;
      bool first = true;
#pragma unroll
      for (int c = 0; c < NQK; ++c) {
        const int w = qk_w(HD, c);
        const uint32_t hi = desc_hi(16 * w, swz_of(w));   // K-major: rows of 2w bytes, 8-row groups of 16w bytes
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (kk < w / 16) {
            const uint32_t qoff = (2 * BQ * qk_col(HD, c) + kk * 32) >> 4;
            const uint32_t koff = (2 * BKV * qk_col(HD, c) + kk * 32) >> 4;
            if (issuer) umma_ss(d_tmem, desc64(hi, q_lo + qoff), desc64(hi, k_lo + koff), idesc_s, first ? 0u : 1u);
            first = false;
          }
        }
      }
      if (issuer) umma_commit(&bars[TB * X + S_FULL]);
    };
    // O_X += P_X V for the V tile in stage st.  V is MN-major: kv rows of 2*VW bytes, 8-row groups SBO = 16*VW
    // apart, the NV column chunks LBO = 2*BKV*VW apart.
    auto issue_pv = [&](int st, int j) {
      constexpr uint32_t hi = desc_hi(16 * VW, swz_of(VW));
      const uint32_t v_lo = skv_lo + ((st * 2 * kv_tile_bytes + kv_tile_bytes) >> 4) + (uint32_t((2 * BKV * VW) >> 4) << 16);
      const uint32_t d_tmem = tmem_base + TM_O + X * HD;
      const uint32_t p_tmem = tmem_base + TM_P + X * PSLOT;
#pragma unroll
      for (int kk = 0; kk < BKV / 16; ++kk)
        if (issuer)
          umma_ts(d_tmem, p_tmem + kk * 8, desc64(hi, v_lo + ((kk * 32 * VW) >> 4)), idesc_o, (j | kk) != 0 ? 1u : 0u);
      if (issuer) umma_commit(&bars[TB * X + O_DONE]);
    };
    // Two cursors walk the K/V ring, one block per step and across work items like the producer's: the stage whose
    // K the next S reads (s_*) and the stage whose V the next P.V reads (pv_*); the S cursor runs one block ahead.
    int s_st = 0, pv_st = 0;
    uint32_t s_ph = 0;
    uint32_t it = 0, gb = 0;   // items / key blocks of this tile finished so far (barrier phases run on across items)
    // S of block b of the current item (global block g = gb + b).  Waits for its K tile and for the softmax warps to
    // have read the previous block's S (the first block of an item: the last one of the previous item).
    auto next_s = [&](int b, int nblk) {
      const uint32_t g = gb + b;
      mbar_wait(&kv_full[s_st], s_ph);
      if (g >= 1) mbar_wait(&bars[TB * X + S_FREE], (g - 1) & 1);
      tc_fence_after();
      issue_s(s_st);
      if (b + 1 == nblk && issuer) umma_commit(&bars[TB * X + Q_EMPTY]);   // this tile's last S of the item has been issued
      if (++s_st == KS) {
        s_st = 0;
        s_ph ^= 1;
      }
    };
    for (int w = blockIdx.x; w < p.n_items; w += gridDim.x) {
      int q0, h, b, n_keys;
      item(w, q0, h, b, n_keys);
      const int nblk = (n_keys + BKV - 1) / BKV;
      if (nblk == 0) continue;
      mbar_wait(&bars[TB * X + Q_FULL], it & 1);
      if (lane == 0 && X == 0) ATT_TRACE(4, 128 + it * 4);
      mbar_wait(&kv_full[s_st], s_ph);
      if (lane == 0 && X == 0) ATT_TRACE(4, 128 + it * 4 + 1);
      if (X > 0) {
        // Every item starts the tiles a fraction of a block period apart, in the order A, B, C: the cycles a softmax
        // warp spends outside its exponentials per block (P hand-off, TMEM loads) then fall into the exponential
        // phases of the other tiles.  The distance drifts over an item; this brings it back.
        mbar_wait(&bars[TB * (X - 1) + FIRST_S], it & 1);
        const long long t_start = clock64();
        while (clock64() - t_start < (long long)p.start_offset) {
        }
      }
      next_s(0, nblk);
      if (issuer) mbar_arrive(&bars[TB * X + FIRST_S]);
      if (lane == 0 && X == 0 && gb > 0) ATT_TRACE(0, (gb - 1) * 2);
      for (int j = 0; j < nblk; ++j) {
        // S(j+1) has to go out BEFORE this thread blocks on P(j), or the softmax warps would find no scores after
        // their hand-off
        if (j + 1 < nblk) {
          next_s(j + 1, nblk);
          if (lane == 0 && X == 0) ATT_TRACE(0, (gb + j) * 2);
        }
        mbar_wait(&bars[TB * X + P_FULL], (gb + j) & 1);
        tc_fence_after();
        issue_pv(pv_st, j);
        if (lane == 0 && X == 0) ATT_TRACE(0, (gb + j) * 2 + 1);
        if (issuer) umma_commit(&kv_empty[pv_st]);   // K/V of block j: this tile's MMAs reading them have been issued
        __syncwarp();
        if (++pv_st == KS) pv_st = 0;
      }
      gb += nblk;
      ++it;
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue (warps 0 .. 4 NT - 1)
    const int X = warp >> 2;                     // query tile of this warpgroup
    const int r = (warp & 3) * 32 + lane;        // row within the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t t_s = t_lane + TM_S + X * BKV;
    const uint32_t t_p = t_lane + TM_P + X * PSLOT;
    constexpr int VB = BKV - 32;                           // columns beyond the first 32-column chunk: 0, 16 or 32
    const float c = p.scale_log2e;
    RowState<HD> rs;
    rs.t_o = t_lane + TM_O + X * HD;
    rs.o_done = &bars[TB * X + O_DONE];
    auto signal_s_free = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[TB * X + S_FREE]);
    };
    uint32_t va[32], vb[VB > 0 ? VB : 1];
    // waits for S of block g (counted across items) and requests it from TMEM into va / vb
    auto request_s = [&](uint32_t g) {
      mbar_wait(&bars[TB * X + S_FULL], g & 1);
      tc_fence_after();
      tmem_ld32(t_s, va);
      if constexpr (VB == 16) tmem_ld16(t_s + 32, reinterpret_cast<uint32_t(&)[16]>(vb));
      if constexpr (VB == 32) tmem_ld32(t_s + 32, reinterpret_cast<uint32_t(&)[32]>(vb));
    };
    uint32_t gb = 0;
    for (int w = blockIdx.x; w < p.n_items; w += gridDim.x) {
      // only the key count stays live over the key blocks; the item's coordinates are derived again at the epilogue
      // (this path runs at the register cap of a 512-thread CTA)
      int n_keys;
      {
        int q0, h, b;
        item(w, q0, h, b, n_keys);
      }
      const int nblk = (n_keys + BKV - 1) / BKV;
      if (nblk == 0) {  // fully padded window: attention output is defined as zero
        int q0, h, b;
        item(w, q0, h, b, n_keys);
        const int qrow = q0 + X * BQ + r;
        if (qrow < p.T) {
          __nv_bfloat16* orow = p.out + ((long long)b * p.T + qrow) * p.d + h * HD;
          for (int c8 = 0; c8 < HD / 8; ++c8) reinterpret_cast<uint4*>(orow)[c8] = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      rs.m_ref = 0.f;
      rs.sum = 0.f;
      rs.gb = gb;
      if (warp == 0 && lane == 0) ATT_TRACE(4, 192 + (gb / 32) * 8 + 2);
      request_s(gb);
      if (warp == 0 && lane == 0) ATT_TRACE(4, 192 + (gb / 32) * 8 + 3);
      for (int j = 0; j < nblk; ++j) {
        const int ncols = min(BKV, n_keys - j * BKV);  // valid keys in this block
        rs.j = j;
        uint32_t pk[BKV / 2];
        const bool tr = (warp & 3) == 0 && lane == 0;
        if (tr) ATT_TRACE(1 + X, (gb + j) * 6);
        // va / vb: columns [0,32) / [32,BKV) of S(j), requested at the end of the previous iteration: S goes back to
        // the MMA warp before the first exponential
        tmem_ld_wait_on(va);
        if constexpr (VB == 16) tmem_ld_wait_on16(reinterpret_cast<uint32_t(&)[16]>(vb));
        if constexpr (VB == 32) tmem_ld_wait_on(reinterpret_cast<uint32_t(&)[32]>(vb));
        signal_s_free();
        if (tr) ATT_TRACE(1 + X, (gb + j) * 6 + 1);
        // Exponential phases in RELAY.  Left alone, the NT softmax warps that share an SM sub-partition fall into
        // lock-step: they queue at the MUFU pipe together (24 cycles per column for three warps, the pipe's limit) and
        // then leave it idle together for the ~600 cycles each spends on its hand-off (TMEM store, barrier round trips)
        // - the timeline showed period = exponentials at the shared rate + hand-off, 2030 cycles per 48-key block
        // against a MUFU floor of 1152 (profiles/r2_notes.md).  The relay keeps them apart: tile X may start the
        // exponentials of its block g only when tile X-1 (tile NT-1's block g-1 for tile 0) is `relay` 16-column
        // pieces into its own, so that one tile's hand-off falls into the others' exponentials.  Hardware named
        // barriers 1 + X (arrive by the 128 threads of tile X, sync by the 128 of its successor): a few cycles each.
        // No tile can run two blocks ahead of its successor (its next start waits, around the ring, for the
        // successor's start), so a barrier generation never receives arrivals of two blocks.
        const uint32_t g_blk = gb + j;
        // The whole block is in registers, so nothing stands between the scores and their exponentials: no maximum, no
        // vote.  (A per-chunk reference check put ~80 cycles of dependent latency - 4 FMNMX levels, FFMA, vote, branch -
        // in front of each 16-column piece.)  The check comes AFTERWARDS, on the block's sum: a P above 2^REF_MARGIN
        // (+inf included) makes the sum exceed it, and then - a few times per row at most, usually never - the
        // reference moves by the exact power of two the block's maximum asks for and the block is computed again.
        auto block = [&](auto masked_tag) {
          constexpr bool MASKED = decltype(masked_tag)::value;
          auto block_max = [&]() {
            float cm = fmaxf(chunk_max<0, 16, MASKED>(va, ncols), chunk_max<16, 16, MASKED>(va + 16, ncols));
            if constexpr (VB > 0) cm = fmaxf(cm, chunk_max<32, VB, MASKED>(vb, ncols));
            return cm;
          };
          if (j == 0) rs.m_ref = ceilf(block_max() * c);   // column 0 is always a valid key
          if (p.relay > 0 && (X > 0 || g_blk > 0)) named_bar_sync(1 + (X + NT - 1) % NT, 256);
#pragma unroll 1
          for (int pass = 0;; ++pass) {
            rs.sm[0] = make_float2(0.f, 0.f);
            rs.sm[1] = make_float2(0.f, 0.f);
            // The relay hand-over sits between 16-column pieces, behind a run-time condition.  Where it lands in the
            // instruction stream is ptxas's choice (SASS of this build: after ~7 / ~25 of the block's 48 MUFUs for
            // relay 1 / 2), and the kernel time follows that placement: see profiles/r2_notes.md for the variants
            // that tried to control it (compile-time position, data-dependent barrier number, phases cut by loops).
            const float neg_ref = -rs.m_ref;
            exp_chunk<HD, 0, 16, MASKED, BKV / 2>(va, ncols, c, neg_ref, rs, pk);
            if (p.relay == 1 && pass == 0) named_bar_arrive(1 + X, 256);
            exp_chunk<HD, 16, 16, MASKED, BKV / 2>(va + 16, ncols, c, neg_ref, rs, pk);
            if (p.relay >= 2 && pass == 0) named_bar_arrive(1 + X, 256);
            if constexpr (VB > 0) exp_chunk<HD, 32, VB, MASKED, BKV / 2>(vb, ncols, c, neg_ref, rs, pk);
            const float2 t = fadd2(rs.sm[0], rs.sm[1]);
            const float bs = t.x + t.y;
            if (pass > 0 || !__any_sync(0xffffffffu, !(bs <= 0x1p80f))) {
              rs.sum += bs;
              break;
            }
            move_reference<HD, 0>(fmaf(block_max(), c, -rs.m_ref), rs, pk);
          }
        };
        if (ncols == BKV) block(std::false_type{}); else block(std::true_type{});
        if (tr) ATT_TRACE(1 + X, (gb + j) * 6 + 4);
        // S(j+1) was issued when s_free(j) arrived, i.e. long ago: its chunks are requested now so that the TMEM read
        // latency hides under the P hand-off below
        if (j + 1 < nblk) {
          request_s(gb + j + 1);
          if (tr) ATT_TRACE(1 + X, (gb + j) * 6 + 3);
        }
        // The P buffer is free once P.V_X(j-1) has retired.  S_X(j+1) was issued after P.V_X(j-1) by the same thread
        // and tcgen05.commit covers every earlier MMA, so the s_full(j+1) wait above already implies it; only the
        // last block has to ask o_done (an mbarrier round trip costs ~100 cycles on this critical path).  Block 0 of
        // a later item: the epilogue below has waited for the previous item's last P.V.
        if (j > 0 && j + 1 >= nblk) {
          mbar_wait(&bars[TB * X + O_DONE], (gb + j - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int q4 = 0; q4 < (BKV / 2) / 16; ++q4) {
          uint32_t w16[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w16[i] = pk[q4 * 16 + i];
          tmem_st16(t_p + q4 * 16, w16);
        }
        if constexpr ((BKV / 2) % 16 == 8) {
          uint32_t w8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w8[i] = pk[(BKV / 2) - 8 + i];
          tmem_st8(t_p + (BKV / 2) - 8, w8);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[TB * X + P_FULL]);
        if (tr) ATT_TRACE(1 + X, (gb + j) * 6 + 5);
      }
      // epilogue: O / rowsum -> bf16
      mbar_wait(&bars[TB * X + O_DONE], (gb + nblk - 1) & 1);
      tc_fence_after();
      if (warp == 0 && lane == 0) ATT_TRACE(4, 192 + (gb / 32) * 8);
      const float inv = 1.0f / rs.sum;
      // A lane owns a row, and rows are d*2 bytes apart in `out`: direct stores would put 16 bytes into each of 32
      // lines per instruction (measured: ~5000 cycles for the epilogue of three tiles).  The rows go to shared memory
      // instead and one TMA store per warp writes its [32 x HD] box; rows >= T are clipped by the tensor map.
      uint8_t* stage = sO + warp * (32 * HD * 2);
      if (lane == 0) bulk_wait_group_read<0>();   // the previous item's store has read the staging buffer
      __syncwarp();
#pragma unroll 1
      for (int cc = 0; cc < HD; cc += 16) {
        uint32_t v[16];
        tmem_ld16(rs.t_o + cc, v);
        tmem_ld_wait();
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2)
          o[i >> 1] = pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
        uint4* dst = reinterpret_cast<uint4*>(stage + lane * (HD * 2) + cc * 2);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {   // lane 0 holds the first row of the warp's box
        int q0, h, b, nk;
        item(w, q0, h, b, nk);
        const int qrow = q0 + X * BQ + r;
        if (qrow < p.T) {
          tma_store_3d(&tmo, stage, h * HD, qrow, b);
          bulk_commit_group();
        }
      }
      tc_fence_before();
      if (warp == 0 && lane == 0) ATT_TRACE(4, 192 + (gb / 32) * 8 + 1);
      gb += nblk;
    }
  }

  if (warp < NT * 4 && lane == 0) bulk_wait_group_read<0>();   // shared memory must outlive the last store's read
  __syncthreads();
  if (warp == NT * 4 + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

struct AttKey {
  const void* base;
  const void* out;
  int B, T, d3, bkv;
  bool operator<(const AttKey& o) const {
    if (base != o.base) return base < o.base;
    if (out != o.out) return out < o.out;
    if (B != o.B) return B < o.B;
    if (T != o.T) return T < o.T;
    if (d3 != o.d3) return d3 < o.d3;
    return bkv < o.bkv;
  }
};
struct AttMaps {
  CUtensorMap tm[7];
};
std::map<AttKey, AttMaps> g_att7_tmaps;
std::mutex g_att7_mu;

}  // namespace

int attention_bf16_v7(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
                      cudaStream_t stream) {
  OASR_REQUIRE(qkv && out && B > 0 && T > 0 && H > 0, "attention: bad arguments");
  OASR_REQUIRE(hd % 16 == 0 && hd >= 16 && hd <= 80, "attention v7: head_dim must be a multiple of 16 in [16, 80]");
  OASR_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "attention: buffers must be 16-byte aligned");
  const int d = H * hd;
  constexpr int bkv = 48;   // three tiles x 48-key blocks: the measured best of the shapes TMEM allows (file header)
  // A single window does not give the three-tile shape enough work items for one CTA per SM (1B, one 30 s window: 4 query
  // triples x 16 heads = 64 items for 148 SMs, 42 us per launch).  Then every CTA takes ONE tile - 192 threads, 256 TMEM
  // columns, a short K/V ring - so that two CTAs share an SM and 192 items run in one wave.
  const int items3 = ((T + 3 * BQ - 1) / (3 * BQ)) * H * B;
  const int NT = items3 < device_sm_count() ? 1 : 3;
  AttMaps m;
  {
    std::lock_guard<std::mutex> g(g_att7_mu);
    AttKey key{qkv, out, B, T, 3 * d, bkv};
    auto it = g_att7_tmaps.find(key);
    if (it == g_att7_tmaps.end()) {
      uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)B};
      uint64_t strides[2] = {(uint64_t)3 * d * 2, (uint64_t)T * 3 * d * 2};
      const uint32_t widths[3] = {64, 32, 16};
      const CUtensorMapSwizzle swz[3] = {CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_32B};
      for (int i = 0; i < 3; ++i) {
        uint32_t qbox[3] = {widths[i], (uint32_t)BQ, 1};
        uint32_t kbox[3] = {widths[i], (uint32_t)bkv, 1};
        OASR_TRY(make_tmap_bf16(&m.tm[i], qkv, 3, dims, strides, qbox, swz[i]));
        OASR_TRY(make_tmap_bf16(&m.tm[3 + i], qkv, 3, dims, strides, kbox, swz[i]));
      }
      {
        uint64_t odims[3] = {(uint64_t)d, (uint64_t)T, (uint64_t)B};
        uint64_t ostrides[2] = {(uint64_t)d * 2, (uint64_t)T * d * 2};
        uint32_t obox[3] = {(uint32_t)hd, 32, 1};
        OASR_TRY(make_tmap_bf16(&m.tm[6], out, 3, odims, ostrides, obox, CU_TENSOR_MAP_SWIZZLE_NONE));
      }
      if (g_att7_tmaps.size() > 1024) g_att7_tmaps.clear();
      g_att7_tmaps[key] = m;
    } else {
      m = it->second;
    }
  }
  Attn7Params p;
  const int q_tile_bytes = BQ * hd * 2, kv_tile_bytes = bkv * hd * 2;
  int kv_stages = (227 * 1024 - 2048 - 2 * NT * q_tile_bytes) / (2 * kv_tile_bytes);   // Q tiles + output staging of the same size
  kv_stages = kv_stages > MAX_KV_STAGES ? MAX_KV_STAGES : kv_stages;
  if (NT == 1 && kv_stages > 4) kv_stages = 4;   // two CTAs per SM: 2 x (Q + staging + 4 K/V stages) fits 227 KB up to head_dim 80
  OASR_REQUIRE(kv_stages >= 3, "attention: tile does not fit shared memory");
  p.kv_stages = kv_stages;
  const int smem_bytes = 2 * NT * q_tile_bytes + 2 * kv_tile_bytes * kv_stages + 1024 + 1024;
  p.T = T;
  p.H = H;
  p.d = d;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.n_frames = n_frames;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.trace = nullptr;
  p.start_offset = START_OFFSET_CYCLES;
  static const int relay = [] {
    const char* e = std::getenv("OASR_ATT_RELAY");   // 0: free-running tiles; 1 / 2: hand on after the first / second 16-column piece
    return e != nullptr ? std::atoi(e) : 2;
  }();
  p.relay = NT == 1 ? 0 : (relay < 0 ? 0 : (relay > 2 ? 2 : relay));   // one tile: nobody to take turns with
  const char* trace_path = std::getenv("OASR_ATT_TRACE");
  if (trace_path != nullptr) {
    OASR_CUDA_CHECK(cudaMalloc(&p.trace, 5 * TRACE_EVENTS * sizeof(long long)));
    OASR_CUDA_CHECK(cudaMemset(p.trace, 0, 5 * TRACE_EVENTS * sizeof(long long)));
  }
  p.n_qt = (T + NT * BQ - 1) / (NT * BQ);
  OASR_REQUIRE((long long)p.n_qt * H * B < (1ll << 31), "attention: too many work items");
  p.n_items = p.n_qt * H * B;
  const int slots = device_sm_count() * (NT == 1 ? 2 : 1);   // persistent: one CTA per SM (two of the one-tile shape)
  dim3 grid(p.n_items < slots ? p.n_items : slots);
  cudaError_t attr_err = cudaSuccess;
#define OASR_ATT_LAUNCH(HDV, NTV)                                                                                      \
  {                                                                                                                    \
    static unsigned long long attr_mask = 0;                                                                           \
    if (first_use_on_this_device(&attr_mask))                                                                          \
      attr_err = cudaFuncSetAttribute(attention_v7_kernel<HDV, NTV, bkv>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      227 * 1024);                                                                     \
    if (attr_err == cudaSuccess)                                                                                       \
      attention_v7_kernel<HDV, NTV, bkv><<<grid, att_threads(NTV), smem_bytes, stream>>>(                              \
          m.tm[0], m.tm[1], m.tm[2], m.tm[3], m.tm[4], m.tm[5], m.tm[6], p);                                           \
  }
#define OASR_ATT_CASE(HDV)                  \
  case HDV:                                 \
    if (NT == 1) OASR_ATT_LAUNCH(HDV, 1)    \
    else OASR_ATT_LAUNCH(HDV, 3)            \
    break;
  switch (hd) {
    OASR_ATT_CASE(16)
    OASR_ATT_CASE(32)
    OASR_ATT_CASE(48)
    OASR_ATT_CASE(64)
    OASR_ATT_CASE(80)
    default: return fail(OASR_ERR_UNSUPPORTED, "attention v7: head_dim must be a multiple of 16 in [16, 80]");
  }
#undef OASR_ATT_CASE
#undef OASR_ATT_LAUNCH
  OASR_CUDA_CHECK(attr_err);
  OASR_CUDA_CHECK(cudaGetLastError());
  if (p.trace != nullptr) {
    static long long host[5 * TRACE_EVENTS];
    OASR_CUDA_CHECK(cudaStreamSynchronize(stream));
    OASR_CUDA_CHECK(cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int r = 0; r < 5; ++r) {
        for (int e = 0; e < TRACE_EVENTS; ++e) fprintf(f, "%lld ", host[r * TRACE_EVENTS + e]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return OASR_OK;
}

}  // namespace oasr
