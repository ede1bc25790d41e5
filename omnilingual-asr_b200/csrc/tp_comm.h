// Tensor-parallel plumbing of liboasr: NCCL reached through dlopen (no link-time dependency: a single-GPU host never
// loads it), one communicator per engine handle.  Only the 7B encoder uses it (BASELINE config 4).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace oasr {

constexpr int TP_UNIQUE_ID_BYTES = 128;   // sizeof(ncclUniqueId)

// fills 128 bytes with a fresh ncclUniqueId (rank 0 calls it, the host broadcasts the bytes)
int tp_unique_id(void* out128);
// collective over all ranks: creates the communicator of this rank on the current device
int tp_comm_create(void** comm, int rank, int world, const void* id128);
void tp_comm_destroy(void* comm);
// in-place sum of `count` fp32 values across ranks, on `stream`
int tp_allreduce_f32(void* comm, float* buf, size_t count, cudaStream_t stream);

}  // namespace oasr
