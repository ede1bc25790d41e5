// Bidirectional self-attention on tcgen05, fifth version (a14): v4's pipeline (P in its own TMEM columns, S(j+1)
// issued as soon as S(j) has been read, one pass over the keys, integer log2-domain softmax reference) with SIXTEEN
// softmax warps instead of eight: four per SM sub-partition.  Measured on B200 (scripts/ubench/softmax_pattern.cu) the
// softmax instruction pattern sustains one MUFU.EX2 per 10.6 / 9.2 / 8.3 cycles with 1 / 2 / 4 warps per sub-partition
// (pipe limit 8.0), and with two warps the fixed per-block work of a warp (TMEM loads and stores, fences, barrier
// round trips: ~450 cycles) leaves the MUFU pipe idle a third of the time (v4 timeline, profiles/r1_notes.md).
//
// Every query row is therefore shared by TWO warps: warp "lo" takes the first half of the key block's columns, warp
// "hi" the second half (both address the same TMEM lanes: warp id % 4 is the row group).  Per block the pair
//   1. loads its half of S(j) into registers and hands S back (s_free),
//   2. exchanges the per-row maximum of its half through shared memory (named barrier of the 64 threads) and so
//      derives the SAME reference decision: the reference moves when the block maximum exceeds it by more than 2^80
//      (both scale their partial row sums, the lo warp rescales O in TMEM),
//   3. computes P = 2^(s c - m_ref) for its columns, writes it to its half of the P columns and arrives on p_full.
// The row sum is the sum of the two partial sums (exchanged once, at the end); the epilogue splits the head
// dimension between the two warps.
//
// Key block: BKV = 96 (head_dim <= 96) or 80, so that 2 S + 2 P + 2 O fit 512 TMEM columns and a warp's half block
// (48 / 40 scores + 24 / 20 packed P words) fits the 112 registers a 576-thread CTA leaves per thread.
// Warps: 0-3 tile A lo, 4-7 tile A hi, 8-11 tile B lo, 12-15 tile B hi, 16 TMA producer, 17 MMA issuer / TMEM allocator.
#include "host_util.h"
#include "kernels.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <map>
#include <mutex>

namespace oasr {
namespace {

constexpr int ATT_THREADS = 576;
constexpr int SOFTMAX_WARPS = 16;
constexpr int BQ = 128;
constexpr int MAX_KV_STAGES = 4;
constexpr int TMEM_COLS = 512;
constexpr float REF_MARGIN = 80.f;   // a block maximum more than 2^80 above the reference moves the reference

__host__ __device__ constexpr int att_bkv(int hd) { return hd <= 96 ? 96 : 80; }
__host__ __device__ constexpr int round16(int v) { return (v + 15) & ~15; }

// Column chunks of a [rows][HD] bf16 K-major tile: greedy 64 / 32 / 16 (128B / 64B / 32B swizzle); see v3.
__host__ __device__ constexpr int qk_nchunks(int hd) {
  int n = 0;
  for (int w = 64; w >= 16; w >>= 1)
    while (hd >= w) {
      hd -= w;
      ++n;
    }
  return n;
}
__host__ __device__ constexpr int qk_w(int hd, int i) {
  int n = 0;
  for (int w = 64; w >= 16; w >>= 1)
    while (hd >= w) {
      if (n == i) return w;
      hd -= w;
      ++n;
    }
  return 0;
}
__host__ __device__ constexpr int qk_col(int hd, int i) {
  int c = 0;
  for (int j = 0; j < i; ++j) c += qk_w(hd, j);
  return c;
}
__host__ __device__ constexpr int v_w(int hd) { return hd % 64 == 0 ? 64 : (hd % 32 == 0 ? 32 : 16); }
__host__ __device__ constexpr uint32_t swz_of(int w) { return w == 64 ? SWZ_128B : (w == 32 ? SWZ_64B : SWZ_32B); }
__host__ __device__ constexpr uint32_t desc_hi(int sbo_bytes, uint32_t layout) {
  return uint32_t((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | ((layout & 7u) << 29);
}
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return (uint64_t(hi) << 32) | lo; }

struct Attn5Params {
  int kv_stages;
  int T, H, d;
  float scale_log2e;
  const int* n_frames;
  __nv_bfloat16* out;
  long long* trace;   // debug: SM-clock timestamps of CTA (0,0,0), [role][event] (OASR_ATT_TRACE=file)
};
constexpr int TRACE_EVENTS = 128;   // per role: 0 MMA warp, 1 softmax warp 0 (tile A lo), 2 softmax warp 8 (tile B lo)
#define ATT_TRACE(role, ev)                                                                                 \
  do {                                                                                                      \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (ev) < TRACE_EVENTS) \
      p.trace[(role) * TRACE_EVENTS + (ev)] = clock64();                                                    \
  } while (0)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// maximum of the first nv of W scores (four independent chains)
template <int W, bool MASKED>
__device__ __forceinline__ float chunk_max(const uint32_t* v, int nv) {
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < W; ++i)
    if (!MASKED || i < nv) m4[(i >> 1) & 3] = fmaxf(m4[(i >> 1) & 3], __uint_as_float(v[i]));
  return fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
}

// W scores -> P = 2^(s c - m_ref), packed to bf16 into pk[PK0 ...], unrounded P accumulated in sm
template <int W, int PK0, bool MASKED, int PKN>
__device__ __forceinline__ void chunk_exp(const uint32_t* v, int nv, float2 c2, float2 nm2, float2 (&sm)[2],
                                          uint32_t (&pk)[PKN]) {
#pragma unroll
  for (int i = 0; i < W; i += 2) {
    const float2 x = ffma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, nm2);
    float p0 = ex2(x.x), p1 = ex2(x.y);
    if (MASKED) {
      if (i >= nv) p0 = 0.f;
      if (i + 1 >= nv) p1 = 0.f;
    }
    sm[(i >> 1) & 1] = fadd2(sm[(i >> 1) & 1], make_float2(p0, p1));
    pk[PK0 + (i >> 1)] = pack_bf16x2(p0, p1);
  }
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_v5_kernel(const __grid_constant__ CUtensorMap tmq64, const __grid_constant__ CUtensorMap tmq32,
                    const __grid_constant__ CUtensorMap tmq16, const __grid_constant__ CUtensorMap tmk64,
                    const __grid_constant__ CUtensorMap tmk32, const __grid_constant__ CUtensorMap tmk16,
                    const Attn5Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int BKV = att_bkv(HD);
  constexpr int HW = BKV / 2;          // score columns per softmax warp: 48 or 40
  constexpr int PW = HW / 2;           // packed P words per softmax warp: 24 or 20
  constexpr int TAILW = HW - 32;       // 16 or 8
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
  uint64_t* s_free = s_full + 2;                 // 2: S_X(j) has been read into registers (8 warps)
  uint64_t* p_full = s_free + 2;                 // 2: P_X(j) is in TMEM (8 warps)
  uint64_t* o_done = p_full + 2;                 // 2: P.V_X(j) has retired (O updated, P buffer free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);
  float* xch_max = reinterpret_cast<float*>(bars + 32);   // [2 parities][2 tiles][2 halves][128 rows]
  float* xch_sum = xch_max + 2 * 2 * 2 * BQ;              // [2 tiles][2 halves][128 rows]

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

  if (warp == SOFTMAX_WARPS && lane == 0) {
    tma_prefetch_desc(&tmq64);
    tma_prefetch_desc(&tmk64);
    tma_prefetch_desc(&tmq16);
    tma_prefetch_desc(&tmk16);
    mbar_init(q_full, 1);
    for (int i = 0; i < MAX_KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 8);
      mbar_init(&p_full[i], 8);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == SOFTMAX_WARPS + 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == SOFTMAX_WARPS) {
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
  } else if (warp == SOFTMAX_WARPS + 1) {
    // ---------------------------------------------------------------- MMA issuer (as v4)
    // Warp-uniform control flow; one elected lane issues.  Static order per key block j:
    //   p_full_B(j-1) -> P.V_B(j-1), release K/V(j-1);  s_free_A(j) -> S_A(j+1);  s_free_B(j) -> S_B(j+1);
    //   p_full_A(j) -> P.V_A(j)
    const bool issuer = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0, 1);  // B = V is MN-major
    mbar_wait(q_full, 0);
    const uint32_t sq_lo = (smem_u32(sQ) & 0x3FFFF) >> 4;        // descriptor start-address fields (16-byte units)
    const uint32_t skv_lo = (smem_u32(sKV) & 0x3FFFF) >> 4;
    auto issue_s = [&](int X, int st) {   // S_X = Q_X K^T for the K tile in stage st: HD/16 MMAs
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
    auto issue_pv = [&](int X, int st, int j) {
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
    tc_fence_after();
    issue_s(0, 0);
    issue_s(1, 0);
    int st_prev = 0, st = 0, st_next = KS > 1 ? 1 : 0;
    uint32_t ph_next = KS > 1 ? 0u : 1u;   // kv_full parity of block j+1
    for (int j = 0; j < nblk; ++j) {
      const bool more = j + 1 < nblk;
      if (j > 0) {
        mbar_wait(&p_full[1], (j - 1) & 1);
        tc_fence_after();
        issue_pv(1, st_prev, j - 1);
        if (issuer) umma_commit(&kv_empty[st_prev]);   // K/V of block j-1: every MMA reading them has been issued
        __syncwarp();
      }
      if (more) {
        mbar_wait(&kv_full[st_next], ph_next);
        mbar_wait(&s_free[0], j & 1);
        tc_fence_after();
        issue_s(0, st_next);
        mbar_wait(&s_free[1], j & 1);
        tc_fence_after();
        issue_s(1, st_next);
      }
      if (lane == 0) ATT_TRACE(0, j * 2);
      mbar_wait(&p_full[0], j & 1);
      tc_fence_after();
      issue_pv(0, st, j);
      if (lane == 0) ATT_TRACE(0, j * 2 + 1);
      st_prev = st;
      st = st_next;
      if (++st_next == KS) {
        st_next = 0;
        ph_next ^= 1;
      }
    }
    mbar_wait(&p_full[1], (nblk - 1) & 1);
    tc_fence_after();
    issue_pv(1, st_prev, nblk - 1);
  } else {
    // ---------------------------------------------------------------- softmax + epilogue (warps 0-15)
    const int X = warp >> 3;                     // query tile
    const int half = (warp >> 2) & 1;            // column half of the key block
    const int rg = warp & 3;                     // row group == TMEM lane quarter
    const int r = rg * 32 + lane;                // row within the tile == TMEM lane
    const int pair_bar = 1 + X * 4 + rg;         // named barrier of the two warps sharing these rows
    const uint32_t t_lane = tmem_base + (uint32_t(rg * 32) << 16);
    const uint32_t t_s = t_lane + TM_S + X * BKV + half * HW;
    const uint32_t t_p = t_lane + TM_P + X * PSLOT + half * PW;
    const uint32_t t_o = t_lane + TM_O + X * HD;
    const float c = p.scale_log2e;
    float m_ref = 0.f;   // integer-valued reference in the log2 domain; identical in both warps of the pair
    float sum = 0.f;     // partial row sum over this warp's columns
    const bool tr = (warp & 7) == 0 && lane == 0;
    uint32_t va[32], vt[TAILW];
    auto load_s = [&]() {
      tmem_ld32(t_s, va);
      if constexpr (TAILW == 16) tmem_ld16(t_s + 32, vt);
      else tmem_ld8(t_s + 32, vt);
    };
    mbar_wait(&s_full[X], 0);
    tc_fence_after();
    load_s();
    for (int j = 0; j < nblk; ++j) {
      const int ncols = min(BKV, n_keys - j * BKV);      // valid keys in this block
      const int nv = max(0, min(HW, ncols - half * HW)); // valid columns among this warp's HW
      const bool full = nv == HW;
      if (tr) ATT_TRACE(1 + X, j * 6);
      tmem_ld_wait_on(va);
      if constexpr (TAILW == 16) tmem_ld_wait_on16(vt);
      else tmem_ld_wait_on8(vt);
      // S(j) is in registers: hand it back so that S(j+1) can be computed under this block's exponentials
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[X]);
      if (tr) ATT_TRACE(1 + X, j * 6 + 1);
      // row maximum of the block = max over the two halves
      float cm = full ? fmaxf(chunk_max<32, false>(va, nv), chunk_max<TAILW, false>(vt, nv - 32))
                      : fmaxf(chunk_max<32, true>(va, nv), chunk_max<TAILW, true>(vt, nv - 32));
      float* xm = xch_max + (((j & 1) * 2 + X) * 2) * BQ;
      xm[half * BQ + r] = cm;
      named_bar_sync(pair_bar, 64);
      cm = fmaxf(cm, xm[(half ^ 1) * BQ + r]);
      if (tr) ATT_TRACE(1 + X, j * 6 + 2);
      if (j == 0) {
        m_ref = ceilf(cm * c);   // column 0 of block 0 is always a valid key
      } else {
        const float need = fmaf(cm, c, -m_ref);
        if (__any_sync(0xffffffffu, need > REF_MARGIN)) {
          // rare: move the reference (same decision in both warps of the pair)
          const float k = need > REF_MARGIN ? ceilf(need) : 0.f;
          const float f = ex2(-k);   // exact (k is an integer); 0 when the old reference was hopelessly low
          m_ref += k;
          sum *= f;
          if (half == 0) {   // the lo warp rescales O; P.V(j) cannot start before both warps' p_full
            mbar_wait(&o_done[X], (j - 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < HD; cc += 16) {
              uint32_t v[16];
              tmem_ld16(t_o + cc, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
              tmem_st16(t_o + cc, v);
            }
            tmem_st_wait();
          }
        }
      }
      uint32_t pk[PW];
      float2 sm[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_ref, -m_ref);
      if (full) {
        chunk_exp<32, 0, false>(va, nv, c2, nm2, sm, pk);
        chunk_exp<TAILW, 16, false>(vt, nv - 32, c2, nm2, sm, pk);
      } else {
        chunk_exp<32, 0, true>(va, nv, c2, nm2, sm, pk);
        chunk_exp<TAILW, 16, true>(vt, nv - 32, c2, nm2, sm, pk);
      }
      if (tr) ATT_TRACE(1 + X, j * 6 + 3);
      // S(j+1) was issued when s_free(j) completed: request it now so that the TMEM read hides under the P hand-off
      if (j + 1 < nblk) {
        mbar_wait(&s_full[X], (j + 1) & 1);
        tc_fence_after();
        load_s();
      }
      {
        const float2 t = fadd2(sm[0], sm[1]);
        sum += t.x + t.y;
      }
      // The P buffer is free once P.V_X(j-1) has retired; the s_full(j+1) wait above implies it (see v4)
      if (j > 0 && j + 1 >= nblk) {
        mbar_wait(&o_done[X], (j - 1) & 1);
        tc_fence_after();
      }
      if (tr) ATT_TRACE(1 + X, j * 6 + 4);
      {
        uint32_t w16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w16[i] = pk[i];
        tmem_st16(t_p, w16);
        if constexpr (PW == 24) {
          uint32_t w8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w8[i] = pk[16 + i];
          tmem_st8(t_p + 16, w8);
        } else {
          static_assert(PW == 20, "packed words per warp");
          uint32_t w4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) w4[i] = pk[16 + i];
          tmem_st4(t_p + 16, w4);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[X]);
      if (tr) ATT_TRACE(1 + X, j * 6 + 5);
    }
    // row sum = sum of the two partial sums
    xch_sum[(X * 2 + half) * BQ + r] = sum;
    named_bar_sync(pair_bar, 64);
    sum += xch_sum[(X * 2 + (half ^ 1)) * BQ + r];
    // epilogue: O / rowsum -> bf16; the head dimension is split between the two warps in groups of 16 columns
    mbar_wait(&o_done[X], (nblk - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / sum;
    const int qrow = q0 + X * BQ + r;
    const bool row_ok = qrow < p.T;
    __nv_bfloat16* orow = p.out + ((long long)b * p.T + qrow) * p.d + h * HD;
    constexpr int NG = HD / 16, NG_LO = (NG + 1) / 2;
    const int g0 = half == 0 ? 0 : NG_LO, g1 = half == 0 ? NG_LO : NG;
#pragma unroll 1
    for (int g = g0; g < g1; ++g) {
      const int cc = g * 16;
      uint32_t v[16];
      tmem_ld16(t_o + cc, v);
      tmem_ld_wait();
      if (row_ok) {
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2)
          o[i >> 1] = pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
        uint4* dst = reinterpret_cast<uint4*>(orow + cc);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == SOFTMAX_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

struct AttKey {
  const void* base;
  int B, T, d3, bkv;
  bool operator<(const AttKey& o) const {
    if (base != o.base) return base < o.base;
    if (B != o.B) return B < o.B;
    if (T != o.T) return T < o.T;
    if (d3 != o.d3) return d3 < o.d3;
    return bkv < o.bkv;
  }
};
struct AttMaps {
  CUtensorMap tm[6];
};
std::map<AttKey, AttMaps> g_att5_tmaps;
std::mutex g_att5_mu;

}  // namespace

int attention_bf16_v5(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
                      cudaStream_t stream) {
  OASR_REQUIRE(qkv && out && B > 0 && T > 0 && H > 0, "attention: bad arguments");
  OASR_REQUIRE(hd % 16 == 0 && hd >= 16 && hd <= 128, "attention: head_dim must be a multiple of 16 in [16, 128]");
  OASR_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "attention: buffers must be 16-byte aligned");
  const int d = H * hd;
  const int bkv = att_bkv(hd);
  AttMaps m;
  {
    std::lock_guard<std::mutex> g(g_att5_mu);
    AttKey key{qkv, B, T, 3 * d, bkv};
    auto it = g_att5_tmaps.find(key);
    if (it == g_att5_tmaps.end()) {
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
      if (g_att5_tmaps.size() > 1024) g_att5_tmaps.clear();
      g_att5_tmaps[key] = m;
    } else {
      m = it->second;
    }
  }
  Attn5Params p;
  const int q_tile_bytes = BQ * hd * 2, kv_tile_bytes = bkv * hd * 2;
  const int tail_bytes = 256 + 2 * 2 * 2 * BQ * 4 + 2 * 2 * BQ * 4;   // barriers + maximum / sum exchange
  int kv_stages = (227 * 1024 - 1024 - tail_bytes - 2 * q_tile_bytes) / (2 * kv_tile_bytes);
  kv_stages = kv_stages > MAX_KV_STAGES ? MAX_KV_STAGES : kv_stages;
  OASR_REQUIRE(kv_stages >= 2, "attention: tile does not fit shared memory");
  p.kv_stages = kv_stages;
  const int smem_bytes = 2 * q_tile_bytes + 2 * kv_tile_bytes * kv_stages + tail_bytes + 1024;
  p.T = T;
  p.H = H;
  p.d = d;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.n_frames = n_frames;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.trace = nullptr;
  const char* trace_path = std::getenv("OASR_ATT_TRACE");
  if (trace_path != nullptr) {
    OASR_CUDA_CHECK(cudaMalloc(&p.trace, 3 * TRACE_EVENTS * sizeof(long long)));
    OASR_CUDA_CHECK(cudaMemset(p.trace, 0, 3 * TRACE_EVENTS * sizeof(long long)));
  }
  dim3 grid((T + 2 * BQ - 1) / (2 * BQ), H, B);
  cudaError_t attr_err = cudaSuccess;
#define OASR_ATT_CASE(HDV)                                                                                      \
  case HDV: {                                                                                                   \
    static bool attr_done = false;                                                                              \
    if (!attr_done) {                                                                                           \
      attr_err = cudaFuncSetAttribute(attention_v5_kernel<HDV>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                      227 * 1024);                                                              \
      attr_done = attr_err == cudaSuccess;                                                                      \
    }                                                                                                           \
    if (attr_err == cudaSuccess)                                                                                \
      attention_v5_kernel<HDV><<<grid, ATT_THREADS, smem_bytes, stream>>>(m.tm[0], m.tm[1], m.tm[2], m.tm[3],   \
                                                                          m.tm[4], m.tm[5], p);                 \
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
