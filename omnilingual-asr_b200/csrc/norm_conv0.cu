// HBM-bound front of the path: waveform normalisation (a8), FE layer 0 (a9), row LayerNorm (a12/a14),
// and the zero-padded bf16 copy that feeds the positional conv (a13).
#include "host_util.h"
#include "kernels.cuh"
#include "ptx.cuh"

namespace oasr {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// a8 wave normalisation: pass 1 = per-slice (sum, sum of squares) in fp64, pass 2 = normalise.
// grid (WAVE_NORM_SLICES, B), 256 threads.
// ---------------------------------------------------------------------------------------------
// Sample loads: fp32 as is, PCM16 scaled by 1/32768 (exact in fp32: identical to a host-side conversion).
__device__ __forceinline__ float load_sample(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_sample(const short* p) { return (float)__ldg(p) * (1.0f / 32768.0f); }

template <typename TIn>
__global__ void wave_stats_kernel(const TIn* __restrict__ in, const int* __restrict__ n_samples, int L,
                                  long long in_stride, double* __restrict__ partials) {
  const int b = blockIdx.y, slice = blockIdx.x;
  const int n = min(n_samples[b], L);
  const int per = (n + WAVE_NORM_SLICES - 1) / WAVE_NORM_SLICES;
  const int lo = slice * per, hi = min(n, lo + per);
  const TIn* x = in + (long long)b * in_stride;
  double s = 0.0, ss = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const double v = (double)load_sample(x + i);
    s += v;
    ss += v * v;
  }
  __shared__ double sh[2][8];
  s = warp_sum_d(s);
  ss = warp_sum_d(ss);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sh[0][w] = s;
    sh[1][w] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, c = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      a += sh[0][i];
      c += sh[1][i];
    }
    partials[((long long)b * WAVE_NORM_SLICES + slice) * 2 + 0] = a;
    partials[((long long)b * WAVE_NORM_SLICES + slice) * 2 + 1] = c;
  }
}

template <typename TIn>
__global__ void wave_apply_kernel(const TIn* __restrict__ in, float* __restrict__ out,
                                  const int* __restrict__ n_samples, int L, long long in_stride, long long out_stride,
                                  const double* __restrict__ partials) {
  const int b = blockIdx.y, slice = blockIdx.x;
  const int n = min(n_samples[b], L);
  __shared__ float sh_mean, sh_rstd;
  if (threadIdx.x == 0) {
    double s = 0, ss = 0;
    for (int i = 0; i < WAVE_NORM_SLICES; ++i) {
      s += partials[((long long)b * WAVE_NORM_SLICES + i) * 2 + 0];
      ss += partials[((long long)b * WAVE_NORM_SLICES + i) * 2 + 1];
    }
    const double mean = n > 0 ? s / n : 0.0;
    double var = n > 0 ? ss / n - mean * mean : 0.0;
    if (var < 0) var = 0;
    sh_mean = (float)mean;
    sh_rstd = (float)(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  const float mean = sh_mean, rstd = sh_rstd;
  const int per = (L + WAVE_NORM_SLICES - 1) / WAVE_NORM_SLICES;
  const int lo = slice * per, hi = min(L, lo + per);
  const TIn* x = in + (long long)b * in_stride;
  float* y = out + (long long)b * out_stride;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) y[i] = i < n ? (load_sample(x + i) - mean) * rstd : 0.f;
}

// ---------------------------------------------------------------------------------------------
// a9 FE layer 0.  One warp per output frame pair; lane owns channels {2*lane + 64*j, +1 : j < 8}.
// Weights [10][512] live in shared memory; stores are 128-byte coalesced bf16x2 rows.
// ---------------------------------------------------------------------------------------------
constexpr int L0_K = 10, L0_S = 5, L0_C = 512;
constexpr int L0_WARPS = 8;
constexpr int L0_FRAMES_PER_WARP_ITER = 2;

__global__ void __launch_bounds__(L0_WARPS * 32)
fe_layer0_kernel(const float* __restrict__ wave, long long in_stride, int T0, const float* __restrict__ w,
                 const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ out, long long out_batch_stride, int frames_per_block) {
  __shared__ float2 sw[L0_K][L0_C / 2];
  for (int i = threadIdx.x; i < L0_K * L0_C / 2; i += blockDim.x)
    sw[i / (L0_C / 2)][i % (L0_C / 2)] = reinterpret_cast<const float2*>(w)[i];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const float* x = wave + (long long)b * in_stride;
  __nv_bfloat16* o = out + (long long)b * out_batch_stride;

  float2 bi[8], ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c2 = lane + 32 * j;  // float2 index: channels 2*c2, 2*c2+1
    bi[j] = __ldg(reinterpret_cast<const float2*>(bias) + c2);
    ga[j] = __ldg(reinterpret_cast<const float2*>(gamma) + c2);
    be[j] = __ldg(reinterpret_cast<const float2*>(beta) + c2);
  }

  const int t_begin = blockIdx.x * frames_per_block;
  const int t_end = min(T0, t_begin + frames_per_block);
  for (int t = t_begin + warp * L0_FRAMES_PER_WARP_ITER; t < t_end; t += L0_WARPS * L0_FRAMES_PER_WARP_ITER) {
    const bool two = (t + 1 < t_end);
    float xs[L0_K + L0_S];
#pragma unroll
    for (int k = 0; k < L0_K + L0_S; ++k) xs[k] = (k < L0_K || two) ? __ldg(x + (long long)t * L0_S + k) : 0.f;

    // Packed fp32 pairs throughout (channels 2c, 2c+1 of a lane): the kernel is issue-bound (35 scalar instructions per
    // output element: 10 FMA conv + LayerNorm + erf-GELU), FFMA2 / FADD2 halve the instruction count of all three.
    f32x2 a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0[j] = f2_pack(bi[j].x, bi[j].y);
      a1[j] = a0[j];
    }
#pragma unroll
    for (int k = 0; k < L0_K; ++k) {
      const f32x2 x0 = f2_splat(xs[k]), x1 = f2_splat(xs[k + L0_S]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 wv = sw[k][lane + 32 * j];
        const f32x2 w2 = f2_pack(wv.x, wv.y);
        a0[j] = f2_fma(w2, x0, a0[j]);
        a1[j] = f2_fma(w2, x1, a1[j]);
      }
    }
    // LayerNorm over 512 channels (two-pass, fp32) for both frames
    f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s0 = f2_add(s0, a0[j]);
      s1 = f2_add(s1, a1[j]);
    }
    float sa, sb;
    f2_unpack(s0, sa, sb);
    const float m0 = warp_sum(sa + sb) * (1.0f / L0_C);
    f2_unpack(s1, sa, sb);
    const float m1 = warp_sum(sa + sb) * (1.0f / L0_C);
    const f32x2 nm0 = f2_splat(-m0), nm1 = f2_splat(-m1);
    f32x2 v0 = 0ull, v1 = 0ull;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0[j] = f2_add(a0[j], nm0);
      a1[j] = f2_add(a1[j], nm1);
      v0 = f2_fma(a0[j], a0[j], v0);
      v1 = f2_fma(a1[j], a1[j], v1);
    }
    f2_unpack(v0, sa, sb);
    const float r0 = rsqrtf(warp_sum(sa + sb) * (1.0f / L0_C) + 1e-5f);
    f2_unpack(v1, sa, sb);
    const float r1 = rsqrtf(warp_sum(sa + sb) * (1.0f / L0_C) + 1e-5f);
    const f32x2 r02 = f2_splat(r0), r12 = f2_splat(r1);
    uint32_t* o0 = reinterpret_cast<uint32_t*>(o + (long long)t * L0_C);
    uint32_t* o1 = reinterpret_cast<uint32_t*>(o + (long long)(t + 1) * L0_C);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const f32x2 g2 = f2_pack(ga[j].x, ga[j].y), b2 = f2_pack(be[j].x, be[j].y);
      float y0, y1;
      gelu_erf_x2(f2_fma(f2_mul(a0[j], r02), g2, b2), y0, y1);
      o0[lane + 32 * j] = pack_bf16x2(y0, y1);
      if (two) {
        gelu_erf_x2(f2_fma(f2_mul(a1[j], r12), g2, b2), y0, y1);
        o1[lane + 32 * j] = pack_bf16x2(y0, y1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row LayerNorm: one warp per row, the row lives in registers as float4 groups (lane owns groups
// lane + 32 j).  D % 4 == 0, D <= 2048.  16-byte loads, 8/16-byte stores, two-pass fp32 statistics.
// ---------------------------------------------------------------------------------------------
// MAXJ = float4 groups per lane (D <= 128 * MAXJ): sized to the row so that the register count allows three or
// four 256-thread blocks per SM; the kernel is HBM-bound and needs the loads of many rows in flight.
template <bool IN_BF16, int MAXJ>
__global__ void __launch_bounds__(256, MAXJ <= 10 ? 3 : 2)
layernorm_kernel(const void* in, long long in_batch_stride, int batches, int rows_per_batch, int D,
                 const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ out_bf16,
                 float* __restrict__ out_f32, const float* __restrict__ add, float* sum_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long long total = (long long)batches * rows_per_batch;
  if (row >= total) return;
  const int b = int(row / rows_per_batch);
  const int t = int(row - (long long)b * rows_per_batch);
  const int ngroups = D >> 2;  // float4 groups in the row
  float4 v[MAXJ];
  float s = 0.f;
  if (IN_BF16) {
    const uint2* p = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(in) +
                                                    (long long)b * in_batch_stride + (long long)t * D);
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int g = lane + 32 * j;
      if (g < ngroups) {
        const uint2 u = __ldg(p + g);
        const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        v[j] = make_float4(lo.x, lo.y, hi.x, hi.y);
        s += (lo.x + lo.y) + (hi.x + hi.y);
      }
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) +
                                                      (long long)b * in_batch_stride + (long long)t * D);
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int g = lane + 32 * j;
      if (g < ngroups) v[j] = p[g];
    }
    if (add != nullptr) {   // tensor-parallel residual: row = in + add, written back to sum_out (may alias in)
      // all loads of the row are issued before the first store: sum_out may alias in, and a store inside the load
      // loop would order every later load behind it
      float4 a4[MAXJ];
#pragma unroll
      for (int j = 0; j < MAXJ; ++j)
        if (lane + 32 * j < ngroups) a4[j] = __ldg(reinterpret_cast<const float4*>(add + row * D) + lane + 32 * j);
#pragma unroll
      for (int j = 0; j < MAXJ; ++j)
        if (lane + 32 * j < ngroups) {
          v[j].x += a4[j].x; v[j].y += a4[j].y; v[j].z += a4[j].z; v[j].w += a4[j].w;
          reinterpret_cast<float4*>(sum_out + row * D)[lane + 32 * j] = v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (lane + 32 * j < ngroups) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < MAXJ; ++j)
    if (lane + 32 * j < ngroups) {
      v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
      q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    const int g = lane + 32 * j;
    if (g < ngroups) {
      const float4 ga = __ldg(g4 + g), be = __ldg(b4 + g);
      float4 y;
      y.x = v[j].x * rstd * ga.x + be.x;
      y.y = v[j].y * rstd * ga.y + be.y;
      y.z = v[j].z * rstd * ga.z + be.z;
      y.w = v[j].w * rstd * ga.w + be.w;
      if (out_bf16) {
        uint2 o;
        o.x = pack_bf16x2(y.x, y.y);
        o.y = pack_bf16x2(y.z, y.w);
        reinterpret_cast<uint2*>(out_bf16 + row * D)[g] = o;
      }
      if (out_f32) reinterpret_cast<float4*>(out_f32 + row * D)[g] = y;
    }
  }
}

__global__ void pad_cast_kernel(const float* __restrict__ x, int T, int d, int pad, __nv_bfloat16* __restrict__ out,
                                long long total_vec4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // index of a 4-element group
  if (i >= total_vec4) return;
  const int d4 = d >> 2;
  const long long row = i / d4;
  const int c4 = int(i - row * d4);
  const int b = int(row / T);
  const int t = int(row - (long long)b * T);
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  uint2 o;
  o.x = pack_bf16x2(v.x, v.y);
  o.y = pack_bf16x2(v.z, v.w);
  const long long orow = (long long)b * (T + 2 * pad) + pad + t;
  reinterpret_cast<uint2*>(out + orow * d)[c4] = o;
}

}  // namespace

int wave_norm(const float* in, float* out, const int* n_samples, int B, int L, long long in_stride,
              long long out_stride, double* partials, cudaStream_t stream) {
  OASR_REQUIRE(in && out && n_samples && partials && B > 0 && L > 0, "wave_norm: bad arguments");
  dim3 grid(WAVE_NORM_SLICES, B);
  wave_stats_kernel<float><<<grid, 256, 0, stream>>>(in, n_samples, L, in_stride, partials);
  wave_apply_kernel<float><<<grid, 256, 0, stream>>>(in, out, n_samples, L, in_stride, out_stride, partials);
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

int wave_norm_i16(const short* in, float* out, const int* n_samples, int B, int L, long long in_stride,
                  long long out_stride, double* partials, cudaStream_t stream) {
  OASR_REQUIRE(in && out && n_samples && partials && B > 0 && L > 0, "wave_norm: bad arguments");
  dim3 grid(WAVE_NORM_SLICES, B);
  wave_stats_kernel<short><<<grid, 256, 0, stream>>>(in, n_samples, L, in_stride, partials);
  wave_apply_kernel<short><<<grid, 256, 0, stream>>>(in, out, n_samples, L, in_stride, out_stride, partials);
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

int fe_layer0(const float* wave, long long in_stride, int B, int L, const float* w, const float* bias,
              const float* gamma, const float* beta, void* out_bf16, long long out_batch_stride_elems, int T0,
              cudaStream_t stream) {
  OASR_REQUIRE(wave && w && bias && gamma && beta && out_bf16 && B > 0, "fe_layer0: bad arguments");
  OASR_REQUIRE(T0 == (L >= L0_K ? (L - L0_K) / L0_S + 1 : 0), "fe_layer0: T0 does not match L");
  if (T0 == 0) return OASR_OK;
  // ~4 blocks per SM per batch row keeps the tail short without re-reading the 20 KB of weights too often
  int blocks_x = (device_sm_count() * 8 + B - 1) / B;
  if (blocks_x < 1) blocks_x = 1;
  int frames_per_block = (T0 + blocks_x - 1) / blocks_x;
  const int gran = L0_WARPS * L0_FRAMES_PER_WARP_ITER;
  frames_per_block = ((frames_per_block + gran - 1) / gran) * gran;
  blocks_x = (T0 + frames_per_block - 1) / frames_per_block;
  dim3 grid(blocks_x, B);
  fe_layer0_kernel<<<grid, L0_WARPS * 32, 0, stream>>>(wave, in_stride, T0, w, bias, gamma, beta,
                                                        reinterpret_cast<__nv_bfloat16*>(out_bf16),
                                                        out_batch_stride_elems, frames_per_block);
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

int layernorm_rows(const void* in, int in_is_bf16, long long in_batch_stride, int batches, int rows_per_batch, int D,
                   const float* gamma, const float* beta, void* out_bf16, float* out_f32, cudaStream_t stream) {
  return add_layernorm_rows(in, in_is_bf16, in_batch_stride, batches, rows_per_batch, D, gamma, beta, out_bf16, out_f32,
                            nullptr, nullptr, stream);
}

int add_layernorm_rows(const void* in, int in_is_bf16, long long in_batch_stride, int batches, int rows_per_batch, int D,
                       const float* gamma, const float* beta, void* out_bf16, float* out_f32, const float* add,
                       float* sum_out, cudaStream_t stream) {
  OASR_REQUIRE(in && gamma && beta && (out_bf16 || out_f32), "layernorm: bad arguments");
  OASR_REQUIRE(add == nullptr || (!in_is_bf16 && sum_out != nullptr && in_batch_stride == 0 && batches == 1),
               "layernorm: the add variant takes contiguous fp32 rows");
  OASR_REQUIRE(D % 4 == 0 && D <= 2048 && D > 0, "layernorm: D must be a multiple of 4 and <= 2048");
  const long long total = (long long)batches * rows_per_batch;
  if (total == 0) return OASR_OK;
  const int rows_per_block = 8;
  const unsigned grid = (unsigned)((total + rows_per_block - 1) / rows_per_block);
#define OASR_LN_LAUNCH(BF, MJ)                                                                                  \
  layernorm_kernel<BF, MJ><<<grid, 256, 0, stream>>>(in, in_batch_stride, batches, rows_per_batch, D, gamma, beta,   \
                                                     reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32, add, sum_out)
  if (in_is_bf16) {
    if (D <= 512) OASR_LN_LAUNCH(true, 4);
    else if (D <= 1280) OASR_LN_LAUNCH(true, 10);
    else OASR_LN_LAUNCH(true, 16);
  } else {
    if (D <= 512) OASR_LN_LAUNCH(false, 4);
    else if (D <= 1280) OASR_LN_LAUNCH(false, 10);
    else OASR_LN_LAUNCH(false, 16);
  }
#undef OASR_LN_LAUNCH
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

int pad_cast_bf16(const float* x, int B, int T, int d, int pad, void* out_bf16, cudaStream_t stream) {
  OASR_REQUIRE(x && out_bf16 && d % 4 == 0, "pad_cast: bad arguments");
  const long long total = (long long)B * T * (d / 4);
  if (total == 0) return OASR_OK;
  pad_cast_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, T, d, pad,
                                                                       reinterpret_cast<__nv_bfloat16*>(out_bf16), total);
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

}  // namespace oasr
