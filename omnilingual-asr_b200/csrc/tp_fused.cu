// Tensor-parallel row-parallel GEMM tail over NVLink peer memory (7B encoder, BASELINE config 4):
//   all-reduce of the partial sums  +  residual add  +  LayerNorm  +  redistribution of LN(x) to every rank,
// PUSH-based.  Rank r owns the rows [r M/W, (r+1) M/W).
//   * reduce-scatter half: the GEMM that produces the partial sums (out-proj, FFN2) writes each output row, rounded to
//     bf16, straight into the receive buffer of the rank that owns it (GemmArgs::route_*: plain st.global on pointers
//     mapped with CUDA IPC).  The transfer is part of the epilogue, tile by tile under the MMAs of the next tile;
//     nothing is pulled and no copy engine is involved.
//   * all-gather half: the owner adds the W partial rows to its slice of the fp32 residual stream (which stays
//     row-sharded: only a row's owner ever adds to it) in rank order, applies LayerNorm and stores the bf16 row into
//     EVERY rank's ln buffer - posted stores again.
// Per rank and reduction (W-1)/W of |part| (bf16) goes out and comes in, and the same for |ln|: 196 MB for the 7B
// shape at W = 2, against 786 MB for an fp32 ring all-reduce - and the separate add + LayerNorm pass over HBM
// disappears.  The oracle rounds the partial sums at the same point (ctc_oracle._row_parallel_linear).
// Round 1 PULLED the partial rows (peer ld.global from SMs: an SM keeps too few bytes in flight over NVLink; then
// per-peer cudaMemcpyAsync on the copy engines: 14 dependent copies per reduction at 8 ranks, ~0.7 ms each time,
// 351 ms per 7B step at TP8 against 100 ms of compute - profiles/r2_notes.md).
//
// Cross-GPU ordering: two monotonically increasing flags per (rank, peer) in peer memory, written with system-scope
// release after a kernel boundary and awaited with acquire:
//   ready[src] = e   "src's pushes of reduction e are out"      -> the owner may read its receive buffer
//   done[src]  = e   "src's LN rows of reduction e are out"     -> the next GEMM of the stream may read LN(x)
// No kernel waits for a kernel of the SAME GPU.  Every wait is bounded by OASR_TP_TIMEOUT_MS of %globaltimer (default
// 60 s: ranks are separate processes and may be skewed by a first-call allocation, host-side audio loading or a
// profiler attaching); on expiry the kernel sets this rank's host-mapped error word and carries on with whatever is
// there - the engine reports OASR_ERR_STATE at its next check instead of losing the context to a trap.  The host is
// expected to start the forward of a batch on all ranks of the group together (same batch, same order).
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

// The W partial rows of a row, all LOCAL by the time this runs: parts.p[q] + row * D with `row` the absolute row index.
struct TpParts {
  const __nv_bfloat16* p[TP_MAX_WORLD];
};

// x[row] += sum of the partial rows (rank order), LayerNorm, LN row -> every rank.  A warp per row, the row in
// registers.  MAXJ: float4 groups of a row per lane; AJ: 4-element groups of a bf16 partial row requested at a time.
template <int MAXJ, int AJ>
__global__ void __launch_bounds__(256, 2)
tp_reduce_ln_kernel(const TpPeerView P, const TpParts parts, long long row0, long long nrows, int D,
                    const float* __restrict__ gamma, const float* __restrict__ beta, int bcast_x) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    for (int q = 0; q < P.world; ++q) {   // partial sums in rank order, bf16
      const uint2* ps = reinterpret_cast<const uint2*>(parts.p[q] + row * D);
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
        for (int q = 0; q < P.world; ++q) reinterpret_cast<uint2*>(P.ln[q] + row * D)[g] = o;   // posted peer stores
      }
    }
  }
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

template <typename... Args>
static void launch_reduce(int D, unsigned grid, cudaStream_t stream, Args... args) {
  if (D <= 512) tp_reduce_ln_kernel<4, 4><<<grid, 256, 0, stream>>>(args...);
  else if (D <= 1280) tp_reduce_ln_kernel<10, 10><<<grid, 256, 0, stream>>>(args...);
  else tp_reduce_ln_kernel<16, 16><<<grid, 256, 0, stream>>>(args...);
}

int tp_push_reduce_layernorm(const TpPeerView& P, long long recv_off, long long first_row, long long rows_total, int D,
                             const float* gamma, const float* beta, unsigned long long epoch, bool bcast_x,
                             cudaStream_t stream) {
  OASR_REQUIRE(P.world >= 2 && P.world <= TP_MAX_WORLD && D % 4 == 0 && D <= 2048, "tp_push: bad arguments");
  const TpShare sh = tp_share(rows_total, P.rank, P.world);
  // every rank's pushes of this reduction have landed in my receive region
  tp_signal_wait_kernel<<<1, 32, 0, stream>>>(P, 0, epoch);
  if (sh.nrows > 0) {
    TpParts parts{};
    const long long row0 = first_row + sh.row0;
    for (int q = 0; q < P.world; ++q)   // slot q holds source q's rows [0, nrows) of my share: index by absolute row
      parts.p[q] = P.recv[P.rank] + recv_off + ((long long)q * sh.slot_rows - row0) * D;
    launch_reduce(D, (unsigned)((sh.nrows + 7) / 8), stream, P, parts, row0, sh.nrows, D, gamma, beta, bcast_x ? 1 : 0);
  }
  // my LN rows are out (kernel boundary + system fence in the signal kernel); wait for everybody's
  tp_signal_wait_kernel<<<1, 32, 0, stream>>>(P, 1, epoch);
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
  TpParts pp{};
  for (int q = 0; q < nparts; ++q) {
    L.x[q] = x;      // every "rank" of the local view is this buffer: the stores below are idempotent
    L.ln[q] = ln;
    pp.p[q] = parts[q];
  }
  launch_reduce(D, (unsigned)((rows + 7) / 8), stream, L, pp, 0ll, rows, D, gamma, beta, 0);
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
