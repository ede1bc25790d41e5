// Bidirectional self-attention on tcgen05, third version (a14): one (sequence, head, PAIR of 128-query tiles) per
// CTA, ONE pass over the keys, everything between the two MMAs stays in tensor memory.  Two softmax warpgroups
// (one per query tile) ping-pong against a single MMA-issuing thread, so every SM sub-partition always has a
// second warp to issue while the first waits on TMEM, MUFU or an mbarrier, and K/V tiles are fetched once per
// 256 queries.
//
//   S_j = Q K_j^T             tcgen05.mma SS, both operands K-major bf16; the head dimension is cut into column
//                             chunks of 64 / 32 / 16 (128B / 64B / 32B swizzle) so any head_dim % 16 == 0 works
//                             with the widest possible TMA boxes (80 = 64 + 16);
//                             S is double buffered in TMEM so S_{j+1} runs under the softmax of block j
//   P_j = 2^(S_j c - m_ref)   one softmax thread per query row (= TMEM lane); m_ref is an INTEGER in the log2
//                             domain, so a change of reference rescales P, the row sum and O by an exact power
//                             of two: the result is bit-identical to a two-pass softmax whose reference is
//                             ceil(rowmax * c) - the numerics contract the oracle's emulate_bf16 mode restates
//   O  += P_j V_j             tcgen05.mma TS: A = bf16(P_j) read from TMEM (tcgen05.st over the first 64 columns of
//                             the S it was computed from), B = V tile as an MN-major smem operand; MMAs of one
//                             thread retire in order, so S_{j+1} may overwrite P_j without a barrier
// The reference only moves when a block's scores exceed it by ~2^80 (detected on the block sum, no per-score
// maximum in the hot loop); then the thread rescales its O row in TMEM and redoes the block (S is still intact:
// P is written last).
//
// TMEM columns: S_A [0,128) S_B [128,256) O_A [256,384) O_B [384,512); P_X aliases S_X[0,64).
// Warps: 0-3 softmax/epilogue of tile A, 4-7 of tile B (warp w owns TMEM lanes [32(w%4), +32)), 8 TMA producer,
// 9 MMA issuer / TMEM allocator.
#include "host_util.h"
#include "kernels.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <map>
#include <mutex>

namespace oasr {
namespace {

constexpr int ATT_THREADS = 320;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int MAX_CHUNKS = 3;  // head_dim is cut into column chunks of 64 / 32 / 16 (128B / 64B / 32B swizzle)
constexpr int MAX_KV_STAGES = 4;
constexpr int TMEM_COLS = 512;
constexpr int TM_S = 0, TM_O = 256;
constexpr float RESCALE_SUM_LIMIT = 1.2089258e24f;  // 2^80: a block sum beyond this moves the softmax reference

// Compile-time tile layouts.  Q and K tiles ([128 rows][HD] bf16, K-major operands) are cut greedily into column
// chunks of 64 / 32 / 16 (128B / 64B / 32B swizzle): few TMA boxes and HD/16 MMAs per S.  V tiles (MN-major B
// operand of P.V) use the widest chunk width that divides HD, so that ONE tcgen05.mma per 16 keys covers all HD
// output columns (the chunks are the N-repeat of the descriptor, LBO apart): 8 MMAs per P.V instead of 8 per chunk.
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
__host__ __device__ constexpr int map_of(int w) { return w == 64 ? 0 : (w == 32 ? 1 : 2); }
// upper 32 bits of an smem matrix descriptor: SBO, version 1, layout
__host__ __device__ constexpr uint32_t desc_hi(int sbo_bytes, uint32_t layout) {
  return uint32_t((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | ((layout & 7u) << 29);
}
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return (uint64_t(hi) << 32) | lo; }

struct AttnParams {
  int kv_stages;
  int T, H, d;
  float scale_log2e;
  const int* n_frames;
  __nv_bfloat16* out;
  long long* trace;   // debug: SM-clock timestamps of CTA (0,0,0), [role][event] (OASR_ATT_TRACE=file)
};
constexpr int TRACE_EVENTS = 64;   // per role
#define ATT_TRACE(role, ev)                                                                            \
  do {                                                                                                 \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (ev) < TRACE_EVENTS) \
      p.trace[(role) * TRACE_EVENTS + (ev)] = clock64();                                               \
  } while (0)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// 32 scores of a row: x = s*c - m_ref, P = 2^x packed to bf16 into pk[base/2 ...]; packed FFMA2 / FADD2, two
// independent pair-accumulators (no serial add chain).  Per two scores: FFMA2, 2 x MUFU.EX2, FADD2, F2FP.
template <bool MASKED>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&v)[32], int base, float2 c2, float2 nm2, int ncols,
                                              uint32_t (&pk)[64], float2 (&sm)[2]) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float2 x = ffma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, nm2);
    float p0 = ex2(x.x), p1 = ex2(x.y);
    if (MASKED) {
      if (base + i >= ncols) p0 = 0.f;
      if (base + i + 1 >= ncols) p1 = 0.f;
    }
    sm[(i >> 1) & 1] = fadd2(sm[(i >> 1) & 1], make_float2(p0, p1));
    pk[(base + i) >> 1] = pack_bf16x2(p0, p1);
  }
}

// One pass over the 128 scores of a row (fully unrolled so that pk[] stays in registers).  The TMEM read of chunk
// k+1 is in flight while chunk k is processed.  Returns the sum of the unrounded P.
template <bool MASKED>
__device__ __forceinline__ float softmax_block(uint32_t t_s, float c, float m_ref, int ncols, uint32_t (&pk)[64]) {
  const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_ref, -m_ref);
  float2 sm[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  uint32_t va[32], vb[32];
  tmem_ld32(t_s, va);
  tmem_ld_wait_on(va);
  tmem_ld32(t_s + 32, vb);
  softmax_chunk<MASKED>(va, 0, c2, nm2, ncols, pk, sm);
  tmem_ld_wait_on(vb);
  tmem_ld32(t_s + 64, va);
  softmax_chunk<MASKED>(vb, 32, c2, nm2, ncols, pk, sm);
  tmem_ld_wait_on(va);
  tmem_ld32(t_s + 96, vb);
  softmax_chunk<MASKED>(va, 64, c2, nm2, ncols, pk, sm);
  tmem_ld_wait_on(vb);
  softmax_chunk<MASKED>(vb, 96, c2, nm2, ncols, pk, sm);
  const float2 t = fadd2(sm[0], sm[1]);
  return t.x + t.y;
}

// Exact maximum of the valid scores of a row block (first block, and the rare reference move).
__device__ __forceinline__ float row_block_max(uint32_t t_s, int ncols) {
  float mx = -INFINITY;
#pragma unroll 1
  for (int cc = 0; cc < BKV; cc += 32) {
    uint32_t v[32];
    tmem_ld32(t_s + cc, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (cc + i < ncols) mx = fmaxf(mx, __uint_as_float(v[i]));
  }
  return mx;
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_v3_kernel(const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm32,
                    const __grid_constant__ CUtensorMap tm16, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int NQK = qk_nchunks(HD);
  constexpr int VW = v_w(HD);
  constexpr int NV = HD / VW;
  constexpr int tile_bytes = 128 * HD * 2;
  const int KS = p.kv_stages;
  uint8_t* sQ = smem;                  // [2 tiles]
  uint8_t* sKV = sQ + 2 * tile_bytes;  // [stage][K | V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + KS * 2 * tile_bytes);
  uint64_t* q_full = bars;                       // 1
  uint64_t* kv_full = bars + 1;                  // MAX_KV_STAGES
  uint64_t* kv_empty = kv_full + MAX_KV_STAGES;  // MAX_KV_STAGES
  uint64_t* s_full = kv_empty + MAX_KV_STAGES;   // 2 (per query tile)
  uint64_t* p_full = s_full + 2;                 // 2
  uint64_t* o_done = p_full + 2;                 // 2: P.V_X(j) has retired
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
    tma_prefetch_desc(&tm64);
    tma_prefetch_desc(&tm32);
    tma_prefetch_desc(&tm16);
    mbar_init(q_full, 1);
    for (int i = 0; i < MAX_KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
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
      auto map = [&](int w) { return w == 64 ? &tm64 : (w == 32 ? &tm32 : &tm16); };
      auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col, int row) {   // Q / K layout
#pragma unroll
        for (int c = 0; c < NQK; ++c)
          tma_load_3d(dst + 256 * qk_col(HD, c), map(qk_w(HD, c)), bar, col + qk_col(HD, c), row, b);
      };
      auto load_v = [&](uint8_t* dst, uint64_t* bar, int col, int row) {      // V layout: NV uniform chunks
#pragma unroll
        for (int c = 0; c < NV; ++c) tma_load_3d(dst + c * (256 * VW), map(VW), bar, col + c * VW, row, b);
      };
      mbar_arrive_expect_tx(q_full, 2 * tile_bytes);
      load_tile(sQ, q_full, qcol, q0);
      load_tile(sQ + tile_bytes, q_full, qcol, q0 + BQ);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* sK = sKV + s * 2 * tile_bytes;
        mbar_arrive_expect_tx(&kv_full[s], 2 * tile_bytes);
        load_tile(sK, &kv_full[s], kcol, j * BKV);
        load_v(sK + tile_bytes, &kv_full[s], vcol, j * BKV);
        ATT_TRACE(0, j);
        if (++s == KS) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issuer
    // The whole warp runs the (warp-uniform) control flow so that descriptor arithmetic stays in the uniform
    // datapath; one elected lane issues the tcgen05 instructions.
    const bool issuer = elect_one();
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0, 1);  // B = V is MN-major
      mbar_wait(q_full, 0);
      const uint32_t sq_lo = (smem_u32(sQ) & 0x3FFFF) >> 4;        // descriptor start-address fields (16-byte units)
      const uint32_t skv_lo = (smem_u32(sKV) & 0x3FFFF) >> 4;
      // S_X = Q_X K^T for the K tile in stage `st`: HD/16 MMAs, descriptors differ by compile-time constants
      auto issue_s = [&](int X, int st) {
        const uint32_t q_lo = sq_lo + X * (tile_bytes >> 4) + (1u << 16);            // LBO field = 1 (unused)
        const uint32_t k_lo = skv_lo + st * (2 * tile_bytes >> 4) + (1u << 16);
        const uint32_t d_tmem = tmem_base + TM_S + X * BKV;
        bool first = true;
#pragma unroll
        for (int c = 0; c < NQK; ++c) {
          constexpr int dummy = 0;
          (void)dummy;
          const int w = qk_w(HD, c);
          const uint32_t hi = desc_hi(16 * w, swz_of(w));   // K-major: rows of 2w bytes, 8-row groups of 16w bytes
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kk < w / 16) {
              const uint32_t off = (256 * qk_col(HD, c) + kk * 32) >> 4;
              if (issuer) umma_ss(d_tmem, desc64(hi, q_lo + off), desc64(hi, k_lo + off), idesc_s, first ? 0u : 1u);
              first = false;
            }
          }
        }
        if (issuer) umma_commit(&s_full[X]);
      };
      // O_X += P_X V for the V tile in stage `st`; P_X sits in the first 64 columns of S_X.  V is MN-major: kv
      // rows of 2*VW bytes, 8-row groups SBO = 16*VW apart, the NV column chunks LBO = 256*VW apart.
      auto issue_pv = [&](int X, int st, int j) {
        constexpr uint32_t hi = desc_hi(16 * VW, swz_of(VW));
        const uint32_t v_lo = skv_lo + ((st * 2 * tile_bytes + tile_bytes) >> 4) + (uint32_t((256 * VW) >> 4) << 16);
        const uint32_t d_tmem = tmem_base + TM_O + X * BQ;
        const uint32_t p_tmem = tmem_base + TM_S + X * BKV;
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk)
          if (issuer)
            umma_ts(d_tmem, p_tmem + kk * 8, desc64(hi, v_lo + ((kk * 32 * VW) >> 4)), idesc_o, (j | kk) != 0 ? 1u : 0u);
        if (issuer) umma_commit(&o_done[X]);
      };
      int st = 0, st_next = KS > 1 ? 1 : 0;
      uint32_t ph = 0, ph_next = KS > 1 ? 0 : 1;   // phase of stage st / st_next
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < nblk; ++j) {
        const bool more = j + 1 < nblk;
        for (int X = 0; X < 2; ++X) {
          mbar_wait(&p_full[X], j & 1);
          tc_fence_after();
          if (lane == 0) ATT_TRACE(1, j * 4 + X * 2);
          issue_pv(X, st, j);
          if (more) {
            if (X == 0) {
              mbar_wait(&kv_full[st_next], ph_next);
              tc_fence_after();
            }
            issue_s(X, st_next);  // retires after P.V_X(j): same thread, in order
          }
          if (lane == 0) ATT_TRACE(1, j * 4 + X * 2 + 1);
        }
        if (issuer) umma_commit(&kv_empty[st]);
        __syncwarp();
        st = st_next;
        ph = ph_next;
        if (++st_next == KS) {
          st_next = 0;
        }
        ph_next = (st_next == 0) ? (ph ^ 1) : ph;
        if (KS == 1) ph_next = ph ^ 1;
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue (warps 0-7)
    const int X = warp >> 2;                     // query tile of this warpgroup
    const int r = (warp & 3) * 32 + lane;        // row within the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t t_s = t_lane + TM_S + X * BKV;
    const uint32_t t_o = t_lane + TM_O + X * BQ;
    const float c = p.scale_log2e;
    float m_ref = 0.f;  // integer-valued reference in the log2 domain
    float sum = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int ncols = min(BKV, n_keys - j * BKV);  // valid keys in this block
      mbar_wait(&s_full[X], j & 1);
      tc_fence_after();
      if ((warp & 3) == 0 && lane == 0) ATT_TRACE(2 + X, j * 3);
      if (j == 0) m_ref = ceilf(row_block_max(t_s, ncols) * c);   // reference before any P is produced
      uint32_t pk[64];
      float bsum = ncols == BKV ? softmax_block<false>(t_s, c, m_ref, ncols, pk)
                                : softmax_block<true>(t_s, c, m_ref, ncols, pk);
      // P is computed against a possibly stale (too low) reference; that is exact as long as nothing overflows,
      // because references differ by integers in the log2 domain.  A block sum beyond 2^80 (or inf) means some
      // score outgrew the reference by more than ~80: move it to this block's maximum and rescale O and the
      // running sum by the exact power of two.  S is still intact (P is written last), so the block is redone.
      const bool grow = !(bsum < RESCALE_SUM_LIMIT);
      if (__any_sync(0xffffffffu, grow)) {
        const float bm = row_block_max(t_s, ncols);   // warp-collective TMEM reads: every lane takes part
        const float k = grow ? fmaxf(ceilf(fmaf(bm, c, -m_ref)), 0.f) : 0.f;
        const float f = ex2(-k);  // exact: k is an integer
        m_ref += k;
        sum *= f;
        if (j > 0) {
          mbar_wait(&o_done[X], (j - 1) & 1);  // P.V_X(j-1) has finished updating O_X
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
        }
        bsum = ncols == BKV ? softmax_block<false>(t_s, c, m_ref, ncols, pk)
                            : softmax_block<true>(t_s, c, m_ref, ncols, pk);
      }
      sum += bsum;
      if ((warp & 3) == 0 && lane == 0) ATT_TRACE(2 + X, j * 3 + 1);
      // P overwrites the first 64 columns of S (every score of the row has been consumed by now)
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint32_t w16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w16[i] = pk[q4 * 16 + i];
        tmem_st16(t_s + q4 * 16, w16);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[X]);
      if ((warp & 3) == 0 && lane == 0) ATT_TRACE(2 + X, j * 3 + 2);
    }
    // epilogue: O / rowsum -> bf16
    mbar_wait(&o_done[X], (nblk - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / sum;
    const int qrow = q0 + X * BQ + r;
    const bool row_ok = qrow < p.T;
    __nv_bfloat16* orow = p.out + ((long long)b * p.T + qrow) * p.d + h * HD;
#pragma unroll 1
    for (int cc = 0; cc < HD; cc += 16) {
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
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

struct AttKey {
  const void* base;
  int B, T, d3;
  bool operator<(const AttKey& o) const {
    if (base != o.base) return base < o.base;
    if (B != o.B) return B < o.B;
    if (T != o.T) return T < o.T;
    return d3 < o.d3;
  }
};
struct AttMaps {
  CUtensorMap tm[3];
};
std::map<AttKey, AttMaps> g_att_tmaps;
std::mutex g_att_mu;

}  // namespace

int attention_bf16_v3(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
                      cudaStream_t stream) {
  OASR_REQUIRE(qkv && out && B > 0 && T > 0 && H > 0, "attention: bad arguments");
  OASR_REQUIRE(hd % 16 == 0 && hd >= 16 && hd <= 128, "attention: head_dim must be a multiple of 16 in [16, 128]");
  OASR_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "attention: buffers must be 16-byte aligned");
  const int d = H * hd;
  CUtensorMap tms[3];
  {
    std::lock_guard<std::mutex> g(g_att_mu);
    AttKey key{qkv, B, T, 3 * d};
    auto it = g_att_tmaps.find(key);
    if (it == g_att_tmaps.end()) {
      uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)B};
      uint64_t strides[2] = {(uint64_t)3 * d * 2, (uint64_t)T * 3 * d * 2};
      AttMaps m;
      const uint32_t widths[3] = {64, 32, 16};
      const CUtensorMapSwizzle swz[3] = {CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_32B};
      for (int i = 0; i < 3; ++i) {
        uint32_t box[3] = {widths[i], 128, 1};
        OASR_TRY(make_tmap_bf16(&m.tm[i], qkv, 3, dims, strides, box, swz[i]));
      }
      if (g_att_tmaps.size() > 1024) g_att_tmaps.clear();
      g_att_tmaps[key] = m;
      for (int i = 0; i < 3; ++i) tms[i] = m.tm[i];
    } else {
      for (int i = 0; i < 3; ++i) tms[i] = it->second.tm[i];
    }
  }
  AttnParams p;
  const int tile_bytes = 128 * hd * 2;
  int kv_stages = (227 * 1024 - 2048 - 2 * tile_bytes) / (2 * tile_bytes);
  kv_stages = kv_stages > MAX_KV_STAGES ? MAX_KV_STAGES : kv_stages;
  OASR_REQUIRE(kv_stages >= 2, "attention: tile does not fit shared memory");
  p.kv_stages = kv_stages;
  const int smem_bytes = tile_bytes * (2 + 2 * kv_stages) + 256 + 1024;
  p.T = T;
  p.H = H;
  p.d = d;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.n_frames = n_frames;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.trace = nullptr;
  const char* trace_path = std::getenv("OASR_ATT_TRACE");
  if (trace_path != nullptr) {
    OASR_CUDA_CHECK(cudaMalloc(&p.trace, 4 * TRACE_EVENTS * sizeof(long long)));
    OASR_CUDA_CHECK(cudaMemset(p.trace, 0, 4 * TRACE_EVENTS * sizeof(long long)));
  }
  dim3 grid((T + 2 * BQ - 1) / (2 * BQ), H, B);
  cudaError_t attr_err = cudaSuccess;
#define OASR_ATT_CASE(HDV)                                                                                      \
  case HDV: {                                                                                                   \
    static bool attr_done = false;                                                                              \
    if (!attr_done) {                                                                                           \
      attr_err = cudaFuncSetAttribute(attention_v3_kernel<HDV>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                      227 * 1024);                                                              \
      attr_done = attr_err == cudaSuccess;                                                                      \
    }                                                                                                           \
    if (attr_err == cudaSuccess)                                                                                \
      attention_v3_kernel<HDV><<<grid, ATT_THREADS, smem_bytes, stream>>>(tms[0], tms[1], tms[2], p);           \
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
    long long host[4 * TRACE_EVENTS];
    OASR_CUDA_CHECK(cudaStreamSynchronize(stream));
    OASR_CUDA_CHECK(cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int r = 0; r < 4; ++r) {
        for (int e = 0; e < TRACE_EVENTS; ++e) fprintf(f, "%lld ", host[r * TRACE_EVENTS + e]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return OASR_OK;
}

int attention_bf16(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
                   cudaStream_t stream) {
  static const int version = [] {
    const char* e = std::getenv("OASR_ATTN");
    return e != nullptr ? std::atoi(e) : 6;
  }();
  if (version == 1) return attention_bf16_v1(qkv, out, n_frames, B, T, H, hd, scale, stream);
  if (version == 2) return attention_bf16_v2(qkv, out, n_frames, B, T, H, hd, scale, stream);
  if (version == 3) return attention_bf16_v3(qkv, out, n_frames, B, T, H, hd, scale, stream);
  if (version == 5) return attention_bf16_v5(qkv, out, n_frames, B, T, H, hd, scale, stream);   // experiment, see v5
  if (version == 6 && hd <= 80) return attention_bf16_v6(qkv, out, n_frames, B, T, H, hd, scale, stream);
  return attention_bf16_v4(qkv, out, n_frames, B, T, H, hd, scale, stream);
}

}  // namespace oasr
