// Interface of the tcgen05 GEMM family (gemm_tcgen05.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oasr {

enum GemmEpilogue : int {
  EPI_BF16 = 0,          // out bf16 = acc + bias
  EPI_BF16_GELU = 1,     // out bf16 = gelu(acc + bias)
  EPI_F32 = 2,           // out f32  = acc + bias          (optionally zeroing padded frames)
  EPI_F32_RESID = 3,     // out f32  = resid + acc + bias  (resid may alias out)
  EPI_ARGMAX = 4,        // packed (value, ~index) atomicMax per row; logits are never stored
  EPI_LN_GELU_BF16 = 5,  // out bf16 = gelu(LN_N(acc + bias) * g + b); needs N == 512 (one tile spans the row)
  EPI_F32_GELU_RESID = 6 // out f32  = resid + gelu(acc + bias)   (positional conv)
};

// D[(b*rows + r), g*N + n] = sum_{j<taps} sum_{c<a_inner} A[g][b][(r + j/P)][j%P][c] * W[g][n][j*k_pad + c]
//
// A is described to TMA as a 5-D bf16 tensor {a_inner, P, U, batches, groups}; tap j of output row r reads
// the a_inner contiguous elements at (parity j%P, position r + j/P).  This one description covers
//   * plain GEMMs              taps=1, P=1, groups=1                        (QKV, out-proj, FFN, CTC head)
//   * stride-2 FE conv layers  taps=k, P=2 (rows split into even/odd), a_inner=512   (implicit GEMM)
//   * grouped positional conv  taps=128, P=1, groups=16, a_inner=d/16 over a zero-padded copy of x
// k_pad (multiple of 64) is the per-tap K extent in W; columns [a_inner, k_pad) of every tap must be zero
// in W (TMA zero-fills the matching A columns).
struct GemmArgs {
  const void* A = nullptr;        // bf16
  int a_inner = 0;                // valid contiguous elements per tap (== K for a plain GEMM)
  int taps = 1;
  int k_pad = 0;                  // per-tap K extent in W, multiple of 64 (0 -> round_up(a_inner, 64))
  int P = 1;                      // parity split of the position axis (conv stride)
  long long a_p_stride = 0;       // elements between parities
  long long a_pos_stride = 0;     // elements between positions
  long long a_batch_stride = 0;   // elements between batches
  long long a_group_stride = 0;   // elements between groups
  int a_positions = 0;            // U: extent of the position axis visible to TMA (rows beyond read as 0)
  int rows_per_batch = 0;         // output rows per batch
  int batches = 1;
  int groups = 1;
  const void* W = nullptr;        // bf16 [groups][N][taps*k_pad]
  int N = 0;                      // output columns per group
  const float* bias = nullptr;    // [groups*N] or null
  void* out = nullptr;            // [batches*rows_per_batch, ldo], group g writes columns [g*N, (g+1)*N)
  int ldo = 0;
  long long out_batch_rows = 0;  // rows between batches in `out` (0 -> rows_per_batch)
  const float* resid = nullptr;   // EPI_F32_RESID / EPI_F32_GELU_RESID (may alias out)
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  unsigned long long* argmax = nullptr;  // [rows] EPI_ARGMAX (zero-initialised by the caller)
  const int* n_valid = nullptr;   // EPI_F32: rows with (row % frames_per_seq) >= n_valid[row / frames_per_seq] -> 0
  int frames_per_seq = 0;
  int epilogue = EPI_BF16;
  // EPI_BF16 on a plain GEMM only - ROUTED rows (tensor parallelism, tp_fused.cu): output row m does not go to `out`
  // but to its owner o = min(m / route_per, route_n - 1), at route_base[o] + (m - o * route_per) * ldo.  The bases are
  // peer-mapped (CUDA IPC) receive buffers: a rank's partial sums leave the epilogue straight for the ranks that reduce
  // them, as posted NVLink stores spread over the GEMM's duration - no separate transfer, nothing to pull.
  int route_n = 0;
  int route_per = 0;
  void* route_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

  // plain row-major GEMM: A [M, K] with leading dimension lda
  static GemmArgs plain(const void* A, int M, int K, long long lda, const void* W, int N) {
    GemmArgs g;
    g.A = A; g.a_inner = K; g.a_pos_stride = lda; g.a_positions = M; g.rows_per_batch = M; g.W = W; g.N = N;
    return g;
  }
};

int gemm_bf16_tcgen05(const GemmArgs& a, cudaStream_t stream);

// (value, index) <-> order-preserving 64-bit key used by EPI_ARGMAX: larger value wins, then lower index.
__host__ __device__ inline unsigned long long argmax_pack(float v, int idx) {
  unsigned int b;
#ifdef __CUDA_ARCH__
  b = __float_as_uint(v);
#else
  union { float f; unsigned int u; } cvt; cvt.f = v; b = cvt.u;
#endif
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return (static_cast<unsigned long long>(b) << 32) | static_cast<unsigned int>(0xFFFFFFFFu - static_cast<unsigned int>(idx));
}
__host__ __device__ inline int argmax_unpack_index(unsigned long long key) {
  return static_cast<int>(0xFFFFFFFFu - static_cast<unsigned int>(key & 0xFFFFFFFFull));
}

}  // namespace oasr
