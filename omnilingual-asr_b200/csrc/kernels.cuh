// Launchers of the non-GEMM kernels (norm_conv0.cu, decode.cu, attention.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oasr {

// a8: per-window zero-mean / unit-variance over the first n_samples[b] samples; the tail is zero-filled.
// `partials` is a scratch buffer of 2 * B * WAVE_NORM_SLICES doubles.
constexpr int WAVE_NORM_SLICES = 64;
int wave_norm(const float* in, float* out, const int* n_samples, int B, int L, long long in_stride,
              long long out_stride, double* partials, cudaStream_t stream);

// same from PCM16 samples (device-side audio front end: int16 -> fp32 / 32768 fused into the normalisation)
int wave_norm_i16(const short* in, float* out, const int* n_samples, int B, int L, long long in_stride,
                  long long out_stride, double* partials, cudaStream_t stream);

// step before the path (resample.cu): interleaved [n_in, channels] fp32 / PCM16 at sr_in -> mono fp32 at sr_out
// (channel mean + Hann-windowed sinc polyphase filter, torchaudio.functional.resample's defaults)
long long resample_length(long long n_in, int sr_in, int sr_out);
int resample_mono(const void* in, int in_is_i16, long long n_in, int channels, int sr_in, int sr_out, float* out,
                  long long out_capacity, cudaStream_t stream);

// a9: Conv1d(1->512, k=10, s=5) + bias + LayerNorm(512) + GELU -> bf16 channels-last.
// in [B, L] fp32 (row stride in_stride); out [B, out_rows_stride rows, 512] bf16; T0 = (L-10)/5+1 rows written.
int fe_layer0(const float* wave, long long in_stride, int B, int L, const float* w /*[10][512]*/, const float* bias,
              const float* gamma, const float* beta, void* out_bf16, long long out_batch_stride_elems, int T0,
              cudaStream_t stream);

// LayerNorm over the last dimension D (multiple of 64, <= 2048).  Rows are addressed as
// in + b*in_batch_stride + t*D for t < rows_per_batch.  Writes bf16 (GEMM operand) and/or fp32.
int layernorm_rows(const void* in, int in_is_bf16, long long in_batch_stride, int batches, int rows_per_batch, int D,
                   const float* gamma, const float* beta, void* out_bf16, float* out_f32, cudaStream_t stream);

// Tensor-parallel residual + LayerNorm in one pass: row = in + add (fp32, contiguous rows), written to sum_out (may
// alias in), LayerNorm(row) -> out_bf16 / out_f32.
int add_layernorm_rows(const void* in, int in_is_bf16, long long in_batch_stride, int batches, int rows_per_batch, int D,
                       const float* gamma, const float* beta, void* out_bf16, float* out_f32, const float* add,
                       float* sum_out, cudaStream_t stream);

// x fp32 [B*T, d] -> bf16 copy with `pad` zero rows before and after every sequence: [B, T + 2*pad, d]
// (only the T middle rows are written; the caller zeroes the buffer once).
int pad_cast_bf16(const float* x, int B, int T, int d, int pad, void* out_bf16, cudaStream_t stream);

// a15 tail + a16: packed arg-max keys -> frame ids (padded frames -> blank), then greedy collapse.
int ctc_decode(const unsigned long long* keys, const int* n_frames, int B, int T, int blank, int* frame_ids,
               int* out_ids, int* out_frames, int* out_lens, cudaStream_t stream);
// a16 alone, on caller-provided frame ids
int ctc_collapse(const int* frame_ids, const int* n_frames, int B, int T, int blank, int* out_ids, int* out_frames,
                 int* out_lens, cudaStream_t stream);

// a14 attention: bidirectional softmax(q k^T * scale) v per (sequence, head), keys >= n_frames[b] masked.
// qkv bf16 [B*T, 3*d] (q | k | v column blocks, head h at columns h*hd), out bf16 [B*T, d].
int attention_bf16(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
                   cudaStream_t stream);
int attention_bf16_v4(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
                      cudaStream_t stream);
int attention_bf16_v7(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
                      cudaStream_t stream);

}  // namespace oasr
