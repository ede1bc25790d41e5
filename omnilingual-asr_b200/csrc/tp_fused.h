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
  const float* part[TP_MAX_WORLD];         // partial sums [M, d] of every rank
  unsigned long long* ready[TP_MAX_WORLD]; // [world] flags in rank q's arena, slot = source rank
  unsigned long long* done[TP_MAX_WORLD];
  unsigned int* cta_counter;               // local
};

// x += sum_q part_q on this rank's rows (x stays row-sharded unless bcast_x), LayerNorm of those rows written to every
// rank; two launches (kernel + flag wait)
int tp_fused_reduce_layernorm(const TpPeerView& P, long long rows_total, int D, const float* gamma, const float* beta,
                              unsigned long long epoch, bool bcast_x, cudaStream_t stream);

// Same result with the transfers on the copy engines (see tp_fused.cu); recv: (world - 1) * ceil-share rows * D floats.
int tp_dma_reduce_layernorm(const TpPeerView& P, long long first_row, long long rows_total, int D, const float* gamma,
                            const float* beta, unsigned long long epoch, bool bcast_x, float* recv, cudaStream_t stream,
                            cudaEvent_t* trace = nullptr);   // trace: 6 events recorded between the steps (timing aid)

}  // namespace oasr
