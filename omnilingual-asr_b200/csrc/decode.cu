// a15 tail + a16: arg-max keys -> frame ids, then greedy CTC collapse (drop repeats, then drop blank).
// Integer work, bit-exact against oracle.greedy_collapse.  One CTA per sequence, ordered stream
// compaction with a ballot/popc scan so kept ids keep their frame order.
#include "gemm.cuh"
#include "host_util.h"
#include "kernels.cuh"

namespace oasr {
namespace {

constexpr int DEC_THREADS = 256;

template <bool FROM_KEYS>
__global__ void __launch_bounds__(DEC_THREADS)
ctc_decode_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ ids_in,
                  const int* __restrict__ n_frames, int T, int blank, int* __restrict__ frame_ids,
                  int* __restrict__ out_ids, int* __restrict__ out_frames, int* __restrict__ out_lens) {
  const int b = blockIdx.x;
  const int n = min(n_frames[b], T);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int warp_counts[DEC_THREADS / 32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();

  auto id_at = [&](int t) -> int {
    if (t >= n) return blank;
    if (FROM_KEYS) return argmax_unpack_index(keys[(long long)b * T + t]);
    return ids_in[(long long)b * T + t];
  };

  for (int t0 = 0; t0 < T; t0 += DEC_THREADS) {
    const int t = t0 + threadIdx.x;
    int id = blank;
    bool keep = false;
    if (t < T) {
      id = id_at(t);
      if (FROM_KEYS || frame_ids != ids_in) {
        if (frame_ids) frame_ids[(long long)b * T + t] = id;
      }
      if (t < n) {
        const int prev = t > 0 ? id_at(t - 1) : -1;
        keep = (t == 0 || id != prev) && id != blank;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_counts[warp] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_counts[w];
    off += __popc(m & ((1u << lane) - 1u));
    if (keep) {
      out_ids[(long long)b * T + off] = id;
      out_frames[(long long)b * T + off] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < DEC_THREADS / 32; ++w) tot += warp_counts[w];
      base += tot;
    }
    __syncthreads();
  }
  const int len = base;
  for (int i = len + threadIdx.x; i < T; i += DEC_THREADS) {
    out_ids[(long long)b * T + i] = -1;
    out_frames[(long long)b * T + i] = -1;
  }
  if (threadIdx.x == 0) out_lens[b] = len;
}

}  // namespace

int ctc_decode(const unsigned long long* keys, const int* n_frames, int B, int T, int blank, int* frame_ids,
               int* out_ids, int* out_frames, int* out_lens, cudaStream_t stream) {
  OASR_REQUIRE(keys && n_frames && out_ids && out_frames && out_lens && B >= 0 && T >= 0, "ctc_decode: bad arguments");
  if (B == 0) return OASR_OK;
  if (T == 0) {
    OASR_CUDA_CHECK(cudaMemsetAsync(out_lens, 0, sizeof(int) * B, stream));
    return OASR_OK;
  }
  ctc_decode_kernel<true><<<B, DEC_THREADS, 0, stream>>>(keys, nullptr, n_frames, T, blank, frame_ids, out_ids,
                                                         out_frames, out_lens);
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

int ctc_collapse(const int* frame_ids, const int* n_frames, int B, int T, int blank, int* out_ids, int* out_frames,
                 int* out_lens, cudaStream_t stream) {
  OASR_REQUIRE(frame_ids && n_frames && out_ids && out_frames && out_lens && B >= 0 && T >= 0,
               "ctc_collapse: bad arguments");
  if (B == 0) return OASR_OK;
  if (T == 0) {
    OASR_CUDA_CHECK(cudaMemsetAsync(out_lens, 0, sizeof(int) * B, stream));
    return OASR_OK;
  }
  ctc_decode_kernel<false><<<B, DEC_THREADS, 0, stream>>>(nullptr, frame_ids, n_frames, T, blank, nullptr, out_ids,
                                                          out_frames, out_lens);
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

}  // namespace oasr
