// Device-side audio front end, step before the path: channel mean + band-limited polyphase
// resampling to the model rate, one HBM-bound kernel.  The filter restates torchaudio.functional.resample's default
// (what the host path uses, audio.to_mono_16k): Hann-windowed sinc, lowpass_filter_width = 6, rolloff = 0.99, the
// rates reduced by their gcd to orig : new; output sample m = frame m / new, phase m % new,
//   y[m] = sum_k xpad[frame * orig + k] * h[phase][k],   xpad = x shifted by `width` zeros,   K = 2 width + orig.
// torchaudio evaluates all K taps although the window zeroes all but ~2 * 6 * orig / base_freq of them; only the
// non-zero span per phase is kept here (17 taps for 22.05 -> 16 kHz instead of 459).  Kernel coefficients are computed
// in fp64 and rounded to fp32 as torchaudio does; fp32 accumulation order differs (parity to ~1e-6).
#include "host_util.h"
#include "kernels.cuh"

#include <cmath>
#include <map>
#include <mutex>
#include <numeric>
#include <vector>

namespace oasr {
namespace {

struct ResamplePlan {
  int orig = 0, nw = 0, width = 0, ntaps = 0;
  float* taps = nullptr;   // [ntaps][nw]
  int* first = nullptr;    // [nw] first input index of the span, relative to frame * orig
};
std::map<std::tuple<int, int, int>, ResamplePlan> g_plans;   // (device, orig, new)
std::mutex g_plan_mu;

// mono sample i of an interleaved [n, CH] signal (CH = 0: run-time channel count)
template <int CH>
__device__ __forceinline__ float mono_at(const float* p, long long i, int channels) {
  if (CH == 1) return __ldg(p + i);
  if (CH == 2) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p) + i);
    return (v.x + v.y) * 0.5f;
  }
  float s = 0.f;
  for (int c = 0; c < channels; ++c) s += __ldg(p + i * channels + c);
  return s / (float)channels;
}
template <int CH>
__device__ __forceinline__ float mono_at(const short* p, long long i, int channels) {
  if (CH == 1) return (float)__ldg(p + i) * (1.0f / 32768.0f);
  if (CH == 2) {
    const short2 v = __ldg(reinterpret_cast<const short2*>(p) + i);
    return ((float)v.x * (1.0f / 32768.0f) + (float)v.y * (1.0f / 32768.0f)) * 0.5f;
  }
  float s = 0.f;
  for (int c = 0; c < channels; ++c) s += (float)__ldg(p + i * channels + c) * (1.0f / 32768.0f);
  return s / (float)channels;
}

// One output sample per thread.  Neighbouring threads read overlapping input spans (L1 hits); interior samples take
// the unchecked loop.
template <typename TIn, int CH>
__global__ void __launch_bounds__(256)
resample_kernel(const TIn* __restrict__ in, long long n_in, int channels, const float* __restrict__ taps,
                const int* __restrict__ first, int ntaps, int orig, int nw, float* __restrict__ out, long long n_out) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_out) return;
  const long long frame = m / nw;
  const int phase = int(m - frame * nw);
  const long long base = frame * orig + first[phase];
  const float* h = taps + phase;   // taps are stored [ntaps][nw]: neighbouring threads = neighbouring phases
  float acc = 0.f;
  if (base >= 0 && base + ntaps <= n_in) {
#pragma unroll 4
    for (int k = 0; k < ntaps; ++k) acc = fmaf(mono_at<CH>(in, base + k, channels), __ldg(h + (long long)k * nw), acc);
  } else {
    for (int k = 0; k < ntaps; ++k) {
      const long long i = base + k;
      if (i >= 0 && i < n_in) acc = fmaf(mono_at<CH>(in, i, channels), __ldg(h + (long long)k * nw), acc);
    }
  }
  out[m] = acc;
}

int get_plan(int sr_in, int sr_out, ResamplePlan* out) {
  int dev = 0;
  OASR_CUDA_CHECK(cudaGetDevice(&dev));
  const int g = std::gcd(sr_in, sr_out);
  const int orig = sr_in / g, nw = sr_out / g;
  std::lock_guard<std::mutex> lk(g_plan_mu);
  auto key = std::make_tuple(dev, orig, nw);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) {
    *out = it->second;
    return OASR_OK;
  }
  if (orig == nw) {   // same rate: torchaudio returns the input unchanged; only the channel mean remains
    ResamplePlan id;
    id.orig = id.nw = 1;
    id.ntaps = 1;
    const float one = 1.0f;
    const int zero = 0;
    OASR_CUDA_CHECK(cudaMalloc(&id.taps, 4));
    OASR_CUDA_CHECK(cudaMalloc(&id.first, 4));
    OASR_CUDA_CHECK(cudaMemcpy(id.taps, &one, 4, cudaMemcpyHostToDevice));
    OASR_CUDA_CHECK(cudaMemcpy(id.first, &zero, 4, cudaMemcpyHostToDevice));
    g_plans[key] = id;
    *out = id;
    return OASR_OK;
  }
  const double lowpass_width = 6.0, rolloff = 0.99;
  const double base_freq = std::min(orig, nw) * rolloff;
  const int width = (int)std::ceil(lowpass_width * orig / base_freq);
  const int K = 2 * width + orig;
  const double pi = 3.14159265358979323846;
  std::vector<std::vector<float>> rows(nw, std::vector<float>(K));
  std::vector<int> lo(nw, K), hi(nw, -1);
  for (int p = 0; p < nw; ++p)
    for (int k = 0; k < K; ++k) {
      double t = (-(double)p / nw + (double)(k - width) / orig) * base_freq;
      t = std::min(std::max(t, -lowpass_width), lowpass_width);
      const double c = std::cos(t * pi / lowpass_width / 2.0);
      const double window = c * c;
      const double tp = t * pi;
      const double v = (tp == 0.0 ? 1.0 : std::sin(tp) / tp) * window * (base_freq / orig);
      rows[p][k] = (float)v;
      if (rows[p][k] != 0.0f) {
        lo[p] = std::min(lo[p], k);
        hi[p] = std::max(hi[p], k);
      }
    }
  int ntaps = 1;
  for (int p = 0; p < nw; ++p) ntaps = std::max(ntaps, hi[p] - lo[p] + 1);
  std::vector<float> taps((size_t)nw * ntaps, 0.f);
  std::vector<int> first(nw, 0);
  for (int p = 0; p < nw; ++p) {
    if (hi[p] < lo[p]) continue;
    first[p] = lo[p] - width;
    for (int k = lo[p]; k <= hi[p]; ++k) taps[(size_t)(k - lo[p]) * nw + p] = rows[p][k];
  }
  ResamplePlan pl;
  pl.orig = orig;
  pl.nw = nw;
  pl.width = width;
  pl.ntaps = ntaps;
  OASR_CUDA_CHECK(cudaMalloc(&pl.taps, taps.size() * 4));
  OASR_CUDA_CHECK(cudaMalloc(&pl.first, first.size() * 4));
  OASR_CUDA_CHECK(cudaMemcpy(pl.taps, taps.data(), taps.size() * 4, cudaMemcpyHostToDevice));
  OASR_CUDA_CHECK(cudaMemcpy(pl.first, first.data(), first.size() * 4, cudaMemcpyHostToDevice));
  g_plans[key] = pl;
  *out = pl;
  return OASR_OK;
}

}  // namespace

long long resample_length(long long n_in, int sr_in, int sr_out) {
  const int g = std::gcd(sr_in, sr_out);
  const long long orig = sr_in / g, nw = sr_out / g;
  return (nw * n_in + orig - 1) / orig;
}

int resample_mono(const void* in, int in_is_i16, long long n_in, int channels, int sr_in, int sr_out, float* out,
                  long long out_capacity, cudaStream_t stream) {
  OASR_REQUIRE(in && out && n_in >= 0 && channels >= 1 && channels <= 8 && sr_in > 0 && sr_out > 0,
               "resample: bad arguments");
  const long long n_out = resample_length(n_in, sr_in, sr_out);
  OASR_REQUIRE(out_capacity >= n_out, "resample: output buffer too small");
  if (n_out == 0) return OASR_OK;
  ResamplePlan pl;
  OASR_TRY(get_plan(sr_in, sr_out, &pl));
  const unsigned grid = (unsigned)((n_out + 255) / 256);
#define OASR_RS_LAUNCH(T, CHV)                                                                                       \
  resample_kernel<T, CHV><<<grid, 256, 0, stream>>>(reinterpret_cast<const T*>(in), n_in, channels, pl.taps, pl.first, \
                                                    pl.ntaps, pl.orig, pl.nw, out, n_out)
  if (in_is_i16) {
    if (channels == 1) OASR_RS_LAUNCH(short, 1);
    else if (channels == 2) OASR_RS_LAUNCH(short, 2);
    else OASR_RS_LAUNCH(short, 0);
  } else {
    if (channels == 1) OASR_RS_LAUNCH(float, 1);
    else if (channels == 2) OASR_RS_LAUNCH(float, 2);
    else OASR_RS_LAUNCH(float, 0);
  }
#undef OASR_RS_LAUNCH
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

}  // namespace oasr
