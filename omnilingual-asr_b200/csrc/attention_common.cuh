// Pieces shared by the attention kernels (attention_v4.cu: two query tiles per CTA; attention_v6.cu: three): operand
// chunking for any head_dim % 16 == 0, UMMA descriptor halves, and the per-row softmax with an integer log2-domain
// reference (a reference move rescales P, the row sum and O by an exact power of two).
// Moving a share of the exponentials to the FMA pipe (Cody-Waite range reduction folded into the score FMA + a degree-4
// polynomial, 2.6e-6 relative) was measured in both rounds - round 2 in four kernel shapes - and made the kernel slower
// in proportion to the instructions it added (profiles/r2_notes.md); it is not in the tree.
#pragma once
#include "ptx.cuh"

namespace oasr {
namespace att {

constexpr float REF_MARGIN = 80.f;   // a chunk maximum more than 2^80 above the reference moves the reference

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

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t bf16x2_scale(uint32_t v, uint32_t f2) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(f2));
  return r;
}

// Everything a reference move has to touch.
template <int HD>
struct RowState {
  float m_ref;       // integer-valued reference in the log2 domain
  float sum;         // running sum of the unrounded P of finished blocks
  float2 sm[2];      // pair-accumulators of the block in flight
  uint32_t t_o;      // TMEM address of this row's O
  uint64_t* o_done;  // P.V_X(j) has retired
  int j;             // key block in flight
  uint32_t gb = 0;   // o_done phases completed before block 0 (persistent kernels carry the barrier across work items)
};

// Moves the reference of the rows whose `need` = chunk maximum (log2 domain) - m_ref exceeds REF_MARGIN.  NPK =
// packed P words of the block computed so far.  Warp-collective (TMEM accesses): called under a warp-uniform branch.
template <int HD, int NPK, int PKN>
__device__ __forceinline__ void move_reference(float need, RowState<HD>& rs, uint32_t (&pk)[PKN]) {
  const float k = need > REF_MARGIN ? ceilf(need) : 0.f;
  const float f = ex2(-k);   // exact (k is an integer); 0 when the old reference was hopelessly low
  rs.m_ref += k;
  rs.sum *= f;
  rs.sm[0].x *= f; rs.sm[0].y *= f; rs.sm[1].x *= f; rs.sm[1].y *= f;
  const uint32_t f2 = pack_bf16x2(f, f);
#pragma unroll
  for (int i = 0; i < NPK; ++i) pk[i] = bf16x2_scale(pk[i], f2);
  if (rs.j > 0) {
    mbar_wait(rs.o_done, (rs.gb + rs.j - 1) & 1);   // P.V(j-1) has finished updating O; P.V(j) cannot start before our p_full
    tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < HD; cc += 16) {
      uint32_t v[16];
      tmem_ld16(rs.t_o + cc, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
      tmem_st16(rs.t_o + cc, v);
    }
    tmem_st_wait();
  }
}

// W scores of a row (columns [BASE, BASE+W) of the block): reference check, P = 2^(s c - m_ref) -> pk, sums.
template <int HD, int BASE, int W, bool MASKED, int PKN>
__device__ __forceinline__ void softmax_chunk(const uint32_t* v, int ncols, float c, RowState<HD>& rs,
                                              uint32_t (&pk)[PKN]) {
  float cm4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
#pragma unroll
  for (int i = 0; i < W; ++i)
    if (!MASKED || BASE + i < ncols) cm4[(i >> 1) & 3] = fmaxf(cm4[(i >> 1) & 3], __uint_as_float(v[i]));
  const float cm = fmaxf(fmaxf(cm4[0], cm4[1]), fmaxf(cm4[2], cm4[3]));
  if (BASE == 0 && rs.j == 0) {
    rs.m_ref = ceilf(cm * c);   // first chunk of the row (column 0 is always a valid key)
  } else {
    const float need = fmaf(cm, c, -rs.m_ref);
    if (__any_sync(0xffffffffu, need > REF_MARGIN)) move_reference<HD, BASE / 2>(need, rs, pk);
  }
  const float2 c2 = make_float2(c, c), nm2 = make_float2(-rs.m_ref, -rs.m_ref);
#pragma unroll
  for (int i = 0; i < W; i += 2) {
    const float2 x = ffma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, nm2);
    float p0 = ex2(x.x), p1 = ex2(x.y);
    if (MASKED) {
      if (BASE + i >= ncols) p0 = 0.f;
      if (BASE + i + 1 >= ncols) p1 = 0.f;
    }
    rs.sm[(i >> 1) & 1] = fadd2(rs.sm[(i >> 1) & 1], make_float2(p0, p1));
    pk[(BASE + i) >> 1] = pack_bf16x2(p0, p1);
  }
}

// The two halves of softmax_chunk for kernels that hold a whole key block in registers (attention_v7.cu): the
// maximum of W scores, and P = 2^(s c - m_ref) -> pk / pair sums with NO reference check in front of the exponentials.
template <int BASE, int W, bool MASKED>
__device__ __forceinline__ float chunk_max(const uint32_t* v, int ncols) {
  float cm4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
#pragma unroll
  for (int i = 0; i < W; ++i)
    if (!MASKED || BASE + i < ncols) cm4[(i >> 1) & 3] = fmaxf(cm4[(i >> 1) & 3], __uint_as_float(v[i]));
  return fmaxf(fmaxf(cm4[0], cm4[1]), fmaxf(cm4[2], cm4[3]));
}
template <int HD, int BASE, int W, bool MASKED, int PKN>
__device__ __forceinline__ void exp_chunk(const uint32_t* v, int ncols, float c, float neg_ref, RowState<HD>& rs,
                                          uint32_t (&pk)[PKN]) {
  const float2 c2 = make_float2(c, c), nm2 = make_float2(neg_ref, neg_ref);
#pragma unroll
  for (int i = 0; i < W; i += 2) {
    const float2 x = ffma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, nm2);
    float p0 = ex2(x.x), p1 = ex2(x.y);
    if (MASKED) {
      if (BASE + i >= ncols) p0 = 0.f;
      if (BASE + i + 1 >= ncols) p1 = 0.f;
    }
    rs.sm[(i >> 1) & 1] = fadd2(rs.sm[(i >> 1) & 1], make_float2(p0, p1));
    pk[(BASE + i) >> 1] = pack_bf16x2(p0, p1);
  }
}

}  // namespace att
}  // namespace oasr
