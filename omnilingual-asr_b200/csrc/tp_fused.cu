// Tensor-parallel row-parallel GEMM tail in ONE kernel over NVLink peer memory (7B encoder, BASELINE config 4):
//   all-reduce of the partial sums  +  residual add  +  LayerNorm  +  redistribution of x and LN(x) to every rank.
// Replaces ncclAllReduce(part) followed by the add+LayerNorm pass.  Rank r owns the rows [r M/W, (r+1) M/W): for each
// of them it reads x (local) and the W partial rows (one local, W-1 over NVLink: plain ld.global on peer pointers
// mapped with CUDA IPC), and writes the new fp32 residual row and the bf16 LayerNorm row into EVERY rank's buffers
// (peer st.global).  The fp32 residual stream itself stays row-sharded (only a row's owner ever adds to it), and the
// partial sums are rounded to bf16 by the GEMM epilogue that produces them (the oracle rounds at the same point:
// ctc_oracle._row_parallel_linear), so per rank and call (W-1)/W of |part| (bf16) comes in and (W-1)/W of |ln| (bf16)
// goes out over NVLink: 196 MB for the 7B shape at W = 2, against 786 MB for an fp32 ring all-reduce - and the separate
// add + LayerNorm pass over HBM disappears.  The sum is taken in fp32 in rank order on every rank: deterministic.
//
// Cross-GPU ordering uses two monotonically increasing flags per (rank, peer) in peer memory:
//   ready[src] = e   "src's partial sums of call e are complete"   signalled by CTA 0 at kernel start (the GEMM that
//                    produced them precedes this kernel on src's stream), awaited by every CTA before its first peer read;
//   done[src]  = e   "src has written its rows of call e everywhere" signalled by the last CTA to finish after a
//                    system-scope fence, awaited by tp_wait_kernel before the next GEMM of the stream reads LN(x).
// No kernel waits for a kernel of the SAME GPU.  Every wait is bounded by OASR_TP_TIMEOUT_MS of %globaltimer (default
// 60 s: ranks are separate processes and may be skewed by a first-call allocation, host-side audio loading or a
// profiler attaching); on expiry the kernel sets this rank's host-mapped error word and carries on with whatever is
// there - the engine reports OASR_ERR_STATE at its next check instead of losing the context to a trap.  The host is
// expected to start the forward of a batch on all ranks of the group together (same batch, same order).
//
// tp_dma_reduce_layernorm (end of this file) is the same reduction with the NVLink transfers on the copy engines and
// only the add + LayerNorm of this rank's rows on SMs (tp_reduce_ln_kernel on all-local pointers, barriers = 0); the
// engine runs it on a stream of its own beside the other half-batch's GEMMs (engine.cu: tp_layers_overlapped).
#include "host_util.h"
#include "tp_fused.h"

#include <cuda_bf16.h>

#include <cstdlib>

namespace oasr {
namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// false: the peer did not arrive in time (the error word is set; the caller goes on so that the stream drains)
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long epoch, const TpPeerView& P) {
  if (ld_acquire_sys(flag) >= epoch) return true;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) < epoch) {
    if (*reinterpret_cast<volatile unsigned int*>(P.error) != 0) return false;   // this rank has given up already
    if (global_ns() - t0 > P.timeout_ns) {
      *reinterpret_cast<volatile unsigned int*>(P.error) = 1u;
      __threadfence_system();
      return false;
    }
  }
  return true;
}
__device__ __forceinline__ float4 bf16x4_to_f32(uint2 u) {
  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// MAXJ: float4 groups of a row per lane (row in registers); AJ: 4-element groups of a bf16 partial row requested at a
// time (8 bytes per lane and group).  Whole partial rows are requested at once: most bytes in flight over NVLink.
template <int MAXJ, int AJ>
__global__ void __launch_bounds__(256, 2)
tp_reduce_ln_kernel(const TpPeerView P, long long row0, long long nrows, int D, const float* __restrict__ gamma,
                    const float* __restrict__ beta, unsigned long long epoch, int bcast_x, int barriers) {
  // ---- barrier 1: every rank's partial sums are complete
  if (barriers && threadIdx.x == 0) {
    if (blockIdx.x == 0)
      for (int q = 0; q < P.world; ++q) st_release_sys(P.ready[q] + P.rank, epoch);
    for (int q = 0; q < P.world; ++q) spin_until(P.ready[P.rank] + q, epoch, P);
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a warp per row; with a small grid (overlap mode: the CTAs sit on the few SMs the GEMMs leave free) a warp walks
  // over several rows
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + warp; r < nrows;
       r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const long long row = row0 + r;
    const int ngroups = D >> 2;
    float4 v[MAXJ];
    {
      const float4* xs = reinterpret_cast<const float4*>(P.x[P.rank] + row * D);
#pragma unroll
      for (int j = 0; j < MAXJ; ++j)
        if (lane + 32 * j < ngroups) v[j] = xs[lane + 32 * j];
    }
    for (int q = 0; q < P.world; ++q) {   // partial sums in rank order: local and peer rows, bf16
      const uint2* ps = reinterpret_cast<const uint2*>(P.part[q] + row * D);
#pragma unroll
      for (int j0 = 0; j0 < MAXJ; j0 += AJ) {
        uint2 a[AJ];
#pragma unroll
        for (int j = 0; j < AJ; ++j)
          if (j0 + j < MAXJ && lane + 32 * (j0 + j) < ngroups) a[j] = ps[lane + 32 * (j0 + j)];
#pragma unroll
        for (int j = 0; j < AJ; ++j)
          if (j0 + j < MAXJ && lane + 32 * (j0 + j) < ngroups) {
            const float4 f = bf16x4_to_f32(a[j]);
            v[j0 + j].x += f.x; v[j0 + j].y += f.y; v[j0 + j].z += f.z; v[j0 + j].w += f.w;
          }
      }
    }
    // The fp32 residual stream stays ROW-SHARDED: only the owner of a row ever adds to it, the other ranks need
    // LN(x) alone (next GEMM's operand), so the new residual row is written locally - unless the caller wants the
    // whole x on every rank (bcast_x: final layer with a hidden-state output).
    for (int q = 0; q < P.world; ++q) {
      if (!bcast_x && q != P.rank) continue;
      float4* xd = reinterpret_cast<float4*>(P.x[q] + row * D);
#pragma unroll
      for (int j = 0; j < MAXJ; ++j)
        if (lane + 32 * j < ngroups) xd[lane + 32 * j] = v[j];
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (lane + 32 * j < ngroups) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum(s) / (float)D;
    float qq = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (lane + 32 * j < ngroups) {
        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
        qq += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
      }
    const float rstd = rsqrtf(warp_sum(qq) / (float)D + 1e-5f);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int g = lane + 32 * j;
      if (g < ngroups) {
        const float4 ga = __ldg(g4 + g), be = __ldg(b4 + g);
        uint2 o;
        o.x = pack2(v[j].x * rstd * ga.x + be.x, v[j].y * rstd * ga.y + be.y);
        o.y = pack2(v[j].z * rstd * ga.z + be.z, v[j].w * rstd * ga.w + be.w);
        for (int q = 0; q < P.world; ++q) reinterpret_cast<uint2*>(P.ln[q] + row * D)[g] = o;
      }
    }
  }

  // ---- barrier 2: the last CTA of this rank tells every rank that this rank's rows have landed
  if (!barriers) return;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(P.cta_counter, 1u);
    if (prev + 1 == gridDim.x) {
      *P.cta_counter = 0;
      __threadfence_system();
      for (int q = 0; q < P.world; ++q) st_release_sys(P.done[q] + P.rank, epoch);
    }
  }
}

__global__ void tp_wait_kernel(const TpPeerView P, unsigned long long epoch) {
  if ((int)threadIdx.x < P.world) spin_until(P.done[P.rank] + threadIdx.x, epoch, P);
}

// thread q tells rank q that this rank has reached `epoch` (which = 0: P.ready, 1: P.done), then waits until rank q
// has said the same
__global__ void tp_signal_wait_kernel(const TpPeerView P, int which, unsigned long long epoch) {
  if ((int)threadIdx.x < P.world) {
    __threadfence_system();
    st_release_sys((which == 0 ? P.ready[threadIdx.x] : P.done[threadIdx.x]) + P.rank, epoch);
    spin_until((which == 0 ? P.ready[P.rank] : P.done[P.rank]) + threadIdx.x, epoch, P);
  }
}

}  // namespace

int tp_fused_reduce_layernorm(const TpPeerView& P, long long rows_total, int D, const float* gamma, const float* beta,
                              unsigned long long epoch, bool bcast_x, cudaStream_t stream) {
  OASR_REQUIRE(P.world >= 2 && P.world <= TP_MAX_WORLD && D % 4 == 0 && D <= 2048, "tp_fused: bad arguments");
  const long long per = rows_total / P.world;
  const long long row0 = per * P.rank;
  const long long nrows = P.rank == P.world - 1 ? rows_total - row0 : per;
  const unsigned grid = (unsigned)((nrows + 7) / 8 > 0 ? (nrows + 7) / 8 : 1);
  if (D <= 512) tp_reduce_ln_kernel<4, 4><<<grid, 256, 0, stream>>>(P, row0, nrows, D, gamma, beta, epoch, bcast_x ? 1 : 0, 1);
  else if (D <= 1280) tp_reduce_ln_kernel<10, 10><<<grid, 256, 0, stream>>>(P, row0, nrows, D, gamma, beta, epoch, bcast_x ? 1 : 0, 1);
  else tp_reduce_ln_kernel<16, 16><<<grid, 256, 0, stream>>>(P, row0, nrows, D, gamma, beta, epoch, bcast_x ? 1 : 0, 1);
  // the next kernel on this stream reads LN(x) written by every rank
  tp_wait_kernel<<<1, 32, 0, stream>>>(P, epoch);
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

// The same reduction with the NVLink transfers on the copy engines, so that no SM time is spent on them and the whole
// call can run beside GEMMs on another stream (engine.cu: tp_layers_overlapped).  On `stream`, in order:
//   signal ready + wait for every rank's ready (one kernel) -> DMA the peers' partial rows of MY row share into `recv`
//   -> local kernel: x += sum of the partials, LayerNorm -> DMA my LayerNorm rows (and x rows if bcast_x) to every
//   peer -> signal done -> wait for every rank's done.
// recv: [(world - 1)][rows_share_max * D] bf16, local.  A peer reads this rank's partial rows between its `ready`
// wait and its `done` signal, so the caller may overwrite them once this call's stream work has completed.
int tp_dma_reduce_layernorm(const TpPeerView& P, long long first_row, long long rows_total, int D, const float* gamma,
                            const float* beta, unsigned long long epoch, bool bcast_x, __nv_bfloat16* recv,
                            cudaStream_t stream, cudaEvent_t* trace) {
  OASR_REQUIRE(P.world >= 2 && P.world <= TP_MAX_WORLD && D % 4 == 0 && D <= 2048 && recv != nullptr,
               "tp_dma: bad arguments");
  int trace_i = 0;
  auto stamp = [&]() {   // OASR_TP_TIMING: events between the steps of one call
    if (trace != nullptr) cudaEventRecord(trace[trace_i++], stream);
  };
  stamp();
  const long long per = rows_total / P.world;
  const long long row0 = first_row + per * P.rank;
  const long long nrows = P.rank == P.world - 1 ? rows_total - per * P.rank : per;
  tp_signal_wait_kernel<<<1, 32, 0, stream>>>(P, 0, epoch);
  stamp();
  // a view in which every "peer" partial is the local copy the DMA brings in, and LN / x go to local rows only
  TpPeerView L = P;
  L.world = P.world;
  int slot = 0;
  for (int q = 0; q < P.world; ++q) {
    L.x[q] = P.x[P.rank];
    L.ln[q] = P.ln[P.rank];
    if (q == P.rank) continue;
    __nv_bfloat16* dst = recv + (size_t)slot * (size_t)(rows_total - per * (P.world - 1)) * D;   // a slot holds the largest share
    if (nrows > 0)
      OASR_CUDA_CHECK(cudaMemcpyAsync(dst, P.part[q] + row0 * D, (size_t)nrows * D * 2, cudaMemcpyDeviceToDevice, stream));
    L.part[q] = dst - row0 * D;   // the kernel indexes partials by absolute row
    ++slot;
  }
  stamp();
  if (nrows > 0) {
    const unsigned grid = (unsigned)((nrows + 7) / 8);
    // the local view writes each LN row `world` times to the same place; world = 1 for the stores is expressed by
    // pointing every ln / x entry at the local buffers (idempotent)
    if (D <= 512) tp_reduce_ln_kernel<4, 4><<<grid, 256, 0, stream>>>(L, row0, nrows, D, gamma, beta, epoch, 0, 0);
    else if (D <= 1280) tp_reduce_ln_kernel<10, 10><<<grid, 256, 0, stream>>>(L, row0, nrows, D, gamma, beta, epoch, 0, 0);
    else tp_reduce_ln_kernel<16, 16><<<grid, 256, 0, stream>>>(L, row0, nrows, D, gamma, beta, epoch, 0, 0);
    stamp();
    for (int q = 0; q < P.world; ++q) {
      if (q == P.rank) continue;
      OASR_CUDA_CHECK(cudaMemcpyAsync(P.ln[q] + row0 * D, P.ln[P.rank] + row0 * D, (size_t)nrows * D * 2,
                                      cudaMemcpyDeviceToDevice, stream));
      if (bcast_x)
        OASR_CUDA_CHECK(cudaMemcpyAsync(P.x[q] + row0 * D, P.x[P.rank] + row0 * D, (size_t)nrows * D * 4,
                                        cudaMemcpyDeviceToDevice, stream));
    }
  }
  stamp();
  tp_signal_wait_kernel<<<1, 32, 0, stream>>>(P, 1, epoch);
  stamp();
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

int tp_local_reduce_layernorm(float* x, const __nv_bfloat16* const* parts, int nparts, long long rows, int D,
                              const float* gamma, const float* beta, __nv_bfloat16* ln, cudaStream_t stream) {
  OASR_REQUIRE(x && parts && ln && nparts >= 1 && nparts <= TP_MAX_WORLD && D % 4 == 0 && D <= 2048, "tp_local: bad arguments");
  if (rows <= 0) return OASR_OK;
  TpPeerView L{};
  L.rank = 0;
  L.world = nparts;
  for (int q = 0; q < nparts; ++q) {
    L.x[q] = x;
    L.ln[q] = ln;
    L.part[q] = parts[q];
  }
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (D <= 512) tp_reduce_ln_kernel<4, 4><<<grid, 256, 0, stream>>>(L, 0, rows, D, gamma, beta, 0, 0, 0);
  else if (D <= 1280) tp_reduce_ln_kernel<10, 10><<<grid, 256, 0, stream>>>(L, 0, rows, D, gamma, beta, 0, 0, 0);
  else tp_reduce_ln_kernel<16, 16><<<grid, 256, 0, stream>>>(L, 0, rows, D, gamma, beta, 0, 0, 0);
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

unsigned long long tp_timeout_ns() {
  static const unsigned long long v = [] {
    const char* e = std::getenv("OASR_TP_TIMEOUT_MS");
    const long long ms = e != nullptr ? std::atoll(e) : 60000;
    return (unsigned long long)(ms > 0 ? ms : 60000) * 1000000ull;
  }();
  return v;
}

}  // namespace oasr
