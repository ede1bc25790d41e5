// Bidirectional self-attention on tcgen05, fourth version (a14).  Same numerics contract and tile scheme as v3
// (one (sequence, head, PAIR of 128-query tiles) per CTA, one pass over the keys, integer log2-domain softmax
// reference so that a reference move rescales by an exact power of two), but the two MMAs of a key block are
// decoupled so that the softmax warps - the MUFU.EX2 pipe is the floor of this kernel - never wait for the tensor
// pipe (v3 timeline, profiles/r1_notes.md: softmax 1700 cycles, then 440 hand-off + 975 for P.V(j) and S(j+1) with
// the softmax warps idle, because P was aliased onto S and S(j+1) could not start before P.V(j) had consumed P(j)):
//
//   * P has its own TMEM columns.  The key block is shortened so that 2 S + 2 P + 2 O fit the 512 columns:
//     BKV = 128 (head_dim <= 64), 96 (<= 96), 80 (<= 128).
//   * The softmax warps signal `s_free` as soon as the LAST chunk of S(j) is in registers (after 1/3 - 1/2 of the
//     block's exponentials); the MMA thread then issues S(j+1) = Q K_{j+1}^T, which is ready long before softmax(j)
//     ends.  P.V(j) is issued when P(j) has been written and runs under softmax(j+1).
//   * The softmax reference is checked per 32-column chunk on the raw scores (one FMNMX per score; the block can no
//     longer be redone from S): a chunk whose maximum exceeds the reference by more than 2^80 moves it; the rare
//     path rescales the running sum, the P values of the block computed so far (bf16 x 2^-k is exact) and O in TMEM.
//
// TMEM columns: S_A [0,BKV) S_B [BKV,2BKV) P_A, P_B (BKV/2 rounded up to 16 each) O_A, O_B (head_dim each).
// Warps: 0-3 softmax/epilogue of tile A, 4-7 of tile B (warp w owns TMEM lanes [32(w%4), +32)), 8 TMA producer,
// 9 / 10 MMA issuers of tiles A / B (9 also allocates TMEM).
#include "host_util.h"
#include "kernels.cuh"
#include "attention_common.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <map>
#include <mutex>

namespace oasr {
namespace {

using namespace att;

constexpr int ATT_THREADS = 352;
constexpr int BQ = 128;
constexpr int MAX_KV_STAGES = 4;
constexpr int TMEM_COLS = 512;

__host__ __device__ constexpr int att_bkv(int hd) { return hd <= 64 ? 128 : (hd <= 96 ? 96 : 80); }

struct Attn4Params {
  int kv_stages;
  int T, H, d;
  float scale_log2e;
  const int* n_frames;
  __nv_bfloat16* out;
  long long* trace;   // debug: SM-clock timestamps of CTA (0,0,0), [role][event] (OASR_ATT_TRACE=file)
  int start_offset;   // tile B issues its first S this many cycles after tile A
};
constexpr int TRACE_EVENTS = 128;   // per role: 0 MMA warp, 1 softmax warp 0 (tile A), 2 softmax warp 4 (tile B)
// Tracing is a compile-time option (-DOASR_ATT_TRACING): even a never-taken stamp costs the softmax warps issue
// slots and a dependent parameter load, six times per key block.
#ifdef OASR_ATT_TRACING
#define ATT_TRACE(role, ev)                                                                                 \
  do {                                                                                                      \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (ev) < TRACE_EVENTS) \
      p.trace[(role) * TRACE_EVENTS + (ev)] = clock64();                                                    \
  } while (0)
#else
#define ATT_TRACE(role, ev) \
  do {                      \
  } while (0)
#endif

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_v4_kernel(const __grid_constant__ CUtensorMap tmq64, const __grid_constant__ CUtensorMap tmq32,
                    const __grid_constant__ CUtensorMap tmq16, const __grid_constant__ CUtensorMap tmk64,
                    const __grid_constant__ CUtensorMap tmk32, const __grid_constant__ CUtensorMap tmk16,
                    const __grid_constant__ CUtensorMap tmo, const Attn4Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int BKV = att_bkv(HD);
  constexpr int PSLOT = round16(BKV / 2);
  constexpr int TM_S = 0, TM_P = 2 * BKV, TM_O = 2 * BKV + 2 * PSLOT;
  static_assert(TM_O + 2 * HD <= TMEM_COLS, "TMEM budget");
  constexpr int NQK = qk_nchunks(HD);
  constexpr int VW = v_w(HD);
  constexpr int NV = HD / VW;
  constexpr int q_tile_bytes = BQ * HD * 2;
  constexpr int kv_tile_bytes = BKV * HD * 2;
  const int KS = p.kv_stages;
  uint8_t* sQ = smem;                    // [2 tiles]
  uint8_t* sKV = sQ + 2 * q_tile_bytes;  // [stage][K | V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + KS * 2 * kv_tile_bytes);
  uint64_t* q_full = bars;                       // 1
  uint64_t* kv_full = bars + 1;                  // MAX_KV_STAGES
  uint64_t* kv_empty = kv_full + MAX_KV_STAGES;  // MAX_KV_STAGES
  uint64_t* s_full = kv_empty + MAX_KV_STAGES;   // 2 (per query tile): S_X(j) is in TMEM
  uint64_t* s_free = s_full + 2;                 // 2: S_X(j) has been read into registers
  uint64_t* p_full = s_free + 2;                 // 2: P_X(j) is in TMEM
  uint64_t* o_done = p_full + 2;                 // 2: P.V_X(j) has retired (O updated, P buffer free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * BQ);
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_keys = min(p.n_frames ? p.n_frames[b] : p.T, p.T);
  const int nblk = (n_keys + BKV - 1) / BKV;

  if (nblk == 0) {  // fully padded window: attention output is defined as zero
    for (int i = threadIdx.x; i < 2 * BQ * (HD / 8); i += blockDim.x) {
      const int r = i / (HD / 8), c8 = i % (HD / 8);
      if (q0 + r < p.T)
        reinterpret_cast<uint4*>(p.out + ((long long)b * p.T + q0 + r) * p.d + h * HD)[c8] = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmq64);
    tma_prefetch_desc(&tmk64);
    tma_prefetch_desc(&tmq16);
    tma_prefetch_desc(&tmk16);
    mbar_init(q_full, 1);
    for (int i = 0; i < MAX_KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);   // one commit per tile's MMA issuer
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      const int qcol = h * HD, kcol = p.d + h * HD, vcol = 2 * p.d + h * HD;
      auto qmap = [&](int w) { return w == 64 ? &tmq64 : (w == 32 ? &tmq32 : &tmq16); };
      auto kmap = [&](int w) { return w == 64 ? &tmk64 : (w == 32 ? &tmk32 : &tmk16); };
      mbar_arrive_expect_tx(q_full, 2 * q_tile_bytes);
#pragma unroll
      for (int X = 0; X < 2; ++X)
#pragma unroll
        for (int c = 0; c < NQK; ++c)
          tma_load_3d(sQ + X * q_tile_bytes + 2 * BQ * qk_col(HD, c), qmap(qk_w(HD, c)), q_full, qcol + qk_col(HD, c),
                      q0 + X * BQ, b);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
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
      }
    }
  } else if (warp >= 9) {
    // ---------------------------------------------------------------- MMA issuers: one warp per query tile
    // Issuing a tcgen05.mma costs the issuing thread ~55 cycles and a blocking mbarrier wait ~150 even when its phase
    // has completed (stamp trace of the issuing thread, profiles/r1_notes.md).  A tile needs HD/16 + BKV/16 MMAs, two
    // commits and three waits per key block: ~1100-1200 cycles, so ONE thread serving both tiles took ~2400 cycles
    // per block - more than the exponentials (1280-1536) - and both tiles starved.  Each tile has its own issuing
    // warp; they meet only at the K/V ring (a stage is released by both commits).  Tile B starts half a block period
    // after tile A so that a softmax warp's non-exponential work falls into the other tile's exponentials.  S_X(j+1)
    // precedes P.V_X(j) on the same thread, so "s_full(j+2) implies P.V(j) retired" holds for the P buffer.
    const int X = warp - 9;
    const bool issuer = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0, 1);  // B = V is MN-major
    mbar_wait(q_full, 0);
    const uint32_t sq_lo = (smem_u32(sQ) & 0x3FFFF) >> 4;        // descriptor start-address fields (16-byte units)
    const uint32_t skv_lo = (smem_u32(sKV) & 0x3FFFF) >> 4;
    auto issue_s = [&](int X_, int st) {   // S_X = Q_X K^T for the K tile in stage st: HD/16 MMAs
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
      if (issuer) umma_commit(&s_full[X]);
    };
    // O_X += P_X V for the V tile in stage st.  V is MN-major: kv rows of 2*VW bytes, 8-row groups SBO = 16*VW
    // apart, the NV column chunks LBO = 2*BKV*VW apart.
    auto issue_pv = [&](int X_, int st, int j) {
      constexpr uint32_t hi = desc_hi(16 * VW, swz_of(VW));
      const uint32_t v_lo = skv_lo + ((st * 2 * kv_tile_bytes + kv_tile_bytes) >> 4) + (uint32_t((2 * BKV * VW) >> 4) << 16);
      const uint32_t d_tmem = tmem_base + TM_O + X * HD;
      const uint32_t p_tmem = tmem_base + TM_P + X * PSLOT;
#pragma unroll
      for (int kk = 0; kk < BKV / 16; ++kk)
        if (issuer)
          umma_ts(d_tmem, p_tmem + kk * 8, desc64(hi, v_lo + ((kk * 32 * VW) >> 4)), idesc_o, (j | kk) != 0 ? 1u : 0u);
      if (issuer) umma_commit(&o_done[X]);
    };
    mbar_wait(&kv_full[0], 0);
    if (X > 0) {
      const long long t_start = clock64();
      while (clock64() - t_start < (long long)p.start_offset) {
      }
    }
    tc_fence_after();
    issue_s(X, 0);
    int st = 0, st_next = KS > 1 ? 1 : 0;
    uint32_t ph_next = KS > 1 ? 0u : 1u;   // kv_full parity of block j+1
    for (int j = 0; j < nblk; ++j) {
      if (j + 1 < nblk) {
        mbar_wait(&kv_full[st_next], ph_next);
        mbar_wait(&s_free[X], j & 1);
        tc_fence_after();
        issue_s(X, st_next);
        if (lane == 0 && X == 0) ATT_TRACE(0, j * 2);
      }
      mbar_wait(&p_full[X], j & 1);
      tc_fence_after();
      issue_pv(X, st, j);
      if (lane == 0 && X == 0) ATT_TRACE(0, j * 2 + 1);
      if (issuer) umma_commit(&kv_empty[st]);   // K/V of block j: this tile's MMAs reading them have been issued
      __syncwarp();
      st = st_next;
      if (++st_next == KS) {
        st_next = 0;
        ph_next ^= 1;
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue (warps 0-7)
    const int X = warp >> 2;                     // query tile of this warpgroup
    const int r = (warp & 3) * 32 + lane;        // row within the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t t_s = t_lane + TM_S + X * BKV;
    const uint32_t t_p = t_lane + TM_P + X * PSLOT;
    const float c = p.scale_log2e;
    RowState<HD> rs;
    rs.m_ref = 0.f;
    rs.sum = 0.f;
    rs.t_o = t_lane + TM_O + X * HD;
    rs.o_done = &o_done[X];
    auto signal_s_free = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[X]);
    };
    uint32_t va[32], vb[32];
    mbar_wait(&s_full[X], 0);
    tc_fence_after();
    tmem_ld32(t_s, va);
    tmem_ld32(t_s + 32, vb);
    for (int j = 0; j < nblk; ++j) {
      const int ncols = min(BKV, n_keys - j * BKV);  // valid keys in this block
      rs.j = j;
      rs.sm[0] = make_float2(0.f, 0.f);
      rs.sm[1] = make_float2(0.f, 0.f);
      uint32_t pk[BKV / 2];
      const bool tr = (warp & 3) == 0 && lane == 0;
      if (tr) ATT_TRACE(1 + X, j * 6);
      // va / vb: columns [0,32) / [32,64) of S(j), requested at the end of the previous iteration.
      // The block is walked in chunks of 32 (16) columns; `full` blocks skip the key mask.
#define OASR_CHUNK(BASE, W, V)                                                         \
  do {                                                                                 \
    if (ncols == BKV) softmax_chunk<HD, BASE, W, false>(V, ncols, c, rs, pk);    \
    else softmax_chunk<HD, BASE, W, true>(V, ncols, c, rs, pk);                  \
  } while (0)
      if constexpr (BKV == 128) {
        tmem_ld_wait_on(va);
        tmem_ld_wait_on(vb);
        OASR_CHUNK(0, 32, va);
        tmem_ld32(t_s + 64, va);
        tmem_ld_wait_on(va);
        OASR_CHUNK(32, 32, vb);
        tmem_ld32(t_s + 96, vb);
        tmem_ld_wait_on(vb);
        signal_s_free();
        OASR_CHUNK(64, 32, va);
        OASR_CHUNK(96, 32, vb);
      } else if constexpr (BKV == 96) {
        tmem_ld_wait_on(va);
        tmem_ld_wait_on(vb);
        if (tr) ATT_TRACE(1 + X, j * 6 + 1);
        OASR_CHUNK(0, 32, va);
        if (tr) ATT_TRACE(1 + X, j * 6 + 2);
        tmem_ld32(t_s + 64, va);
        tmem_ld_wait_on(va);
        signal_s_free();
        if (tr) ATT_TRACE(1 + X, j * 6 + 3);
        OASR_CHUNK(32, 32, vb);
        OASR_CHUNK(64, 32, va);
        if (tr) ATT_TRACE(1 + X, j * 6 + 4);
      } else {
        static_assert(BKV == 80, "key block");
        uint32_t vc[16];
        tmem_ld16(t_s + 64, vc);
        tmem_ld_wait_on(va);
        tmem_ld_wait_on(vb);
        tmem_ld_wait_on16(vc);
        signal_s_free();
        OASR_CHUNK(0, 32, va);
        OASR_CHUNK(32, 32, vb);
        OASR_CHUNK(64, 16, vc);
      }
#undef OASR_CHUNK
      // S(j+1) was issued when s_free(j) arrived, i.e. long ago: request its first two chunks now so that the TMEM
      // read latency hides under the P hand-off below
      if (j + 1 < nblk) {
        mbar_wait(&s_full[X], (j + 1) & 1);
        tc_fence_after();
        tmem_ld32(t_s, va);
        tmem_ld32(t_s + 32, vb);
      }
      {
        const float2 t = fadd2(rs.sm[0], rs.sm[1]);
        rs.sum += t.x + t.y;
      }
      // The P buffer is free once P.V_X(j-1) has retired.  S_X(j+1) was issued after P.V_X(j-1) by the same thread and
      // tcgen05.commit covers every earlier MMA, so the s_full(j+1) wait above already implies it; only the last
      // block has to ask o_done (an mbarrier round trip costs ~150 cycles on this critical path).
      if (j > 0 && j + 1 >= nblk) {
        mbar_wait(&o_done[X], (j - 1) & 1);
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
      if (lane == 0) mbar_arrive(&p_full[X]);
      if (tr) ATT_TRACE(1 + X, j * 6 + 5);
    }
    // epilogue: O / rowsum -> bf16
    mbar_wait(&o_done[X], (nblk - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / rs.sum;
    const int qrow = q0 + X * BQ + r;
    // A lane owns a row, and rows are d*2 bytes apart in `out`: direct stores would put 16 bytes into each of 32 lines
    // per instruction.  The rows go to shared memory instead - into this tile's own Q buffer, which nothing reads any
    // more once the tile's last P.V has retired (tcgen05.commit covers every earlier MMA of the issuing thread) - and
    // one TMA store per warp writes its [32 x HD] box; rows >= T are clipped by the tensor map.
    uint8_t* stage = sQ + X * q_tile_bytes + (warp & 3) * (32 * HD * 2);
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
    if (lane == 0 && qrow < p.T) {   // lane 0 holds the first row of the warp's box
      tma_store_3d(&tmo, stage, h * HD, qrow, b);
      bulk_commit_group();
      bulk_wait_group_read<0>();     // shared memory must outlive the store's read
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 9) {
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
std::map<AttKey, AttMaps> g_att4_tmaps;
std::mutex g_att4_mu;

}  // namespace

int attention_bf16_v4(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
                      cudaStream_t stream) {
  OASR_REQUIRE(qkv && out && B > 0 && T > 0 && H > 0, "attention: bad arguments");
  OASR_REQUIRE(hd % 16 == 0 && hd >= 16 && hd <= 128, "attention: head_dim must be a multiple of 16 in [16, 128]");
  OASR_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "attention: buffers must be 16-byte aligned");
  const int d = H * hd;
  const int bkv = att_bkv(hd);
  AttMaps m;
  {
    std::lock_guard<std::mutex> g(g_att4_mu);
    AttKey key{qkv, out, B, T, 3 * d, bkv};
    auto it = g_att4_tmaps.find(key);
    if (it == g_att4_tmaps.end()) {
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
      if (g_att4_tmaps.size() > 1024) g_att4_tmaps.clear();
      g_att4_tmaps[key] = m;
    } else {
      m = it->second;
    }
  }
  Attn4Params p;
  const int q_tile_bytes = BQ * hd * 2, kv_tile_bytes = bkv * hd * 2;
  int kv_stages = (227 * 1024 - 2048 - 2 * q_tile_bytes) / (2 * kv_tile_bytes);
  kv_stages = kv_stages > MAX_KV_STAGES ? MAX_KV_STAGES : kv_stages;
  OASR_REQUIRE(kv_stages >= 2, "attention: tile does not fit shared memory");
  p.kv_stages = kv_stages;
  const int smem_bytes = 2 * q_tile_bytes + 2 * kv_tile_bytes * kv_stages + 256 + 1024;
  p.T = T;
  p.H = H;
  p.d = d;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.n_frames = n_frames;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.trace = nullptr;
  p.start_offset = 1200;   // half a block period (measured best of 0 / 400 / 800 / 1200)
  const char* trace_path = std::getenv("OASR_ATT_TRACE");
  if (trace_path != nullptr) {
    OASR_CUDA_CHECK(cudaMalloc(&p.trace, 3 * TRACE_EVENTS * sizeof(long long)));
    OASR_CUDA_CHECK(cudaMemset(p.trace, 0, 3 * TRACE_EVENTS * sizeof(long long)));
  }
  dim3 grid((T + 2 * BQ - 1) / (2 * BQ), H, B);
  cudaError_t attr_err = cudaSuccess;
#define OASR_ATT_CASE(HDV)                                                                                      \
  case HDV: {                                                                                                   \
    static unsigned long long attr_mask = 0;                                                                    \
    if (first_use_on_this_device(&attr_mask)) {                                                                 \
      attr_err = cudaFuncSetAttribute(attention_v4_kernel<HDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      227 * 1024);                                                              \
    }                                                                                                           \
    if (attr_err == cudaSuccess)                                                                                \
      attention_v4_kernel<HDV><<<grid, ATT_THREADS, smem_bytes, stream>>>(m.tm[0], m.tm[1], m.tm[2], m.tm[3], \
                                                                             m.tm[4], m.tm[5], m.tm[6], p);              \
    break;                                                                                                      \
  }
  switch (hd) {
    OASR_ATT_CASE(16)
    OASR_ATT_CASE(32)
    OASR_ATT_CASE(48)
    OASR_ATT_CASE(64)
    OASR_ATT_CASE(80)
    OASR_ATT_CASE(96)
    OASR_ATT_CASE(112)
    OASR_ATT_CASE(128)
    default: return fail(OASR_ERR_UNSUPPORTED, "attention: head_dim must be a multiple of 16 in [16, 128]");
  }
#undef OASR_ATT_CASE
  OASR_CUDA_CHECK(attr_err);
  OASR_CUDA_CHECK(cudaGetLastError());
  if (p.trace != nullptr) {
    long long host[3 * TRACE_EVENTS];
    OASR_CUDA_CHECK(cudaStreamSynchronize(stream));
    OASR_CUDA_CHECK(cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int r = 0; r < 3; ++r) {
        for (int e = 0; e < TRACE_EVENTS; ++e) fprintf(f, "%lld ", host[r * TRACE_EVENTS + e]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return OASR_OK;
}

}  // namespace oasr
