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
  const __nv_bfloat16* part[TP_MAX_WORLD]; // partial sums [M, d] of every rank, rounded to bf16 by the GEMM epilogue:
                                           // what crosses NVLink in the reduce-scatter half is 2 bytes per element
  unsigned long long* ready[TP_MAX_WORLD]; // [world] flags in rank q's arena, slot = source rank
  unsigned long long* done[TP_MAX_WORLD];
  unsigned int* cta_counter;               // local
  unsigned int* error;                     // host-mapped word of this rank: set (never trapped on) when a peer's flag
                                           // did not arrive within timeout_ns; the host turns it into OASR_ERR_STATE
  unsigned long long timeout_ns;
};

// x += sum_q part_q on this rank's rows (x stays row-sharded unless bcast_x), LayerNorm of those rows written to every
// rank; two launches (kernel + flag wait)
int tp_fused_reduce_layernorm(const TpPeerView& P, long long rows_total, int D, const float* gamma, const float* beta,
                              unsigned long long epoch, bool bcast_x, cudaStream_t stream);

// Same result with the transfers on the copy engines (see tp_fused.cu); recv: (world - 1) * ceil-share rows * D bf16.
int tp_dma_reduce_layernorm(const TpPeerView& P, long long first_row, long long rows_total, int D, const float* gamma,
                            const float* beta, unsigned long long epoch, bool bcast_x, __nv_bfloat16* recv,
                            cudaStream_t stream, cudaEvent_t* trace = nullptr);   // trace: 6 events between the steps

// The reduction's arithmetic alone, everything local: x[rows] += part_0 + part_1 + ... (rank order), LayerNorm -> ln.
// Used by the single-GPU emulation of the split (oasr_tp_emulate): same kernel, same summation order, no flags.
int tp_local_reduce_layernorm(float* x, const __nv_bfloat16* const* parts, int nparts, long long rows, int D,
                              const float* gamma, const float* beta, __nv_bfloat16* ln, cudaStream_t stream);

// OASR_TP_TIMEOUT_MS (default 60 000): how long a rank waits for a peer's flag before it gives up
unsigned long long tp_timeout_ns();

}  // namespace oasr
