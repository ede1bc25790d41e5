// Peer-memory view of a tensor-parallel group (tp_fused.cu): the same arena layout on every rank, mapped with CUDA IPC.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace oasr {

constexpr int TP_MAX_WORLD = 8;

struct TpPeerView {
  int rank, world;
  float* x[TP_MAX_WORLD];                  // fp32 residual stream [M, d] of every rank
  __nv_bfloat16* ln[TP_MAX_WORLD];         // LayerNorm output [M, d] of every rank (next GEMM's A operand)
  __nv_bfloat16* recv[TP_MAX_WORLD];       // receive region of every rank: [source rank][rows of its share][d] bf16.  A
                                           // rank's row-parallel GEMM writes its partial sums of the rows rank q owns
                                           // into recv[q] at slot `rank` (GemmArgs::route_*), so what crosses NVLink in
                                           // the reduce-scatter half is 2 bytes per element, pushed from the epilogue
  unsigned long long* ready[TP_MAX_WORLD]; // [world] flags in rank q's arena, slot = source rank
  unsigned long long* done[TP_MAX_WORLD];
  unsigned int* error;                     // host-mapped word of this rank: set (never trapped on) when a peer's flag
                                           // did not arrive within timeout_ns; the host turns it into OASR_ERR_STATE
  unsigned long long timeout_ns;
};

// Row share of a rank: rows [first + per * r, ...) with per = rows / world, the last rank taking the remainder.
struct TpShare {
  long long per, row0, nrows, slot_rows;   // slot_rows: rows one source's slot holds (the largest share)
};
inline TpShare tp_share(long long rows_total, int rank, int world) {
  TpShare s;
  s.per = rows_total / world;
  if (s.per < 1) s.per = 1;   // fewer rows than ranks: the last rank owns whatever lies beyond (world - 1) * 1
  s.row0 = s.per * rank;
  s.nrows = rank == world - 1 ? rows_total - s.row0 : s.per;
  if (s.nrows < 0) s.nrows = 0;
  if (s.row0 > rows_total) s.row0 = rows_total;
  s.slot_rows = rows_total - s.per * (world - 1);
  if (s.slot_rows < s.per) s.slot_rows = s.per;
  return s;
}

// The tail of a row-parallel GEMM whose epilogue has PUSHED the partial sums to their owners (all ranks call it with
// the same arguments, on `stream`, after that GEMM):
//   flag hand-shake "my pushes are out / everybody's have landed here"
//   -> one kernel on this rank's row share: x += sum over sources (rank order, fp32), LayerNorm, bf16 LN rows stored
//      into EVERY rank's ln buffer (posted NVLink stores; x too when bcast_x)
//   -> flag hand-shake "my LN rows are out / everybody's have landed here".
// recv_off: element offset of this reduction's region inside every rank's receive buffer (half-batches use disjoint
// regions); first_row: first row of the reduction in x / ln.
int tp_push_reduce_layernorm(const TpPeerView& P, long long recv_off, long long first_row, long long rows_total, int D,
                             const float* gamma, const float* beta, unsigned long long epoch, bool bcast_x,
                             cudaStream_t stream);

// The reduction's arithmetic alone, everything local: x[rows] += part_0 + part_1 + ... (rank order), LayerNorm -> ln.
// Used by the single-GPU emulation of the split (oasr_tp_emulate): same kernel, same summation order, no flags.
int tp_local_reduce_layernorm(float* x, const __nv_bfloat16* const* parts, int nparts, long long rows, int D,
                              const float* gamma, const float* beta, __nv_bfloat16* ln, cudaStream_t stream);

// OASR_TP_TIMEOUT_MS (default 60 000): how long a rank waits for a peer's flag before it gives up
unsigned long long tp_timeout_ns();

}  // namespace oasr
