/* liboasr C-ABI: the B200-native omniASR CTC inference path.
 *
 * Plain C, pointers and sizes only; no torch types.  Every entry point returns 0 (OASR_OK) or a negative
 * error code and never throws; oasr_last_error() returns the message of the calling thread's last failure.
 * All work is asynchronous on the given CUDA stream unless the name ends in _host.  One handle per device;
 * a handle is NOT thread-safe (the Python host serialises calls).
 *
 * What each entry point replaces in the reference (Nathan-Roll1/omnilingual-asr, paths relative to
 * /root/reference):
 *   - the per-chunk "transcribe" slot of the engine, i.e. the network call at
 *     src/omnilingual_asr/models/inference/gemini_pipeline.py:512-530 inside GeminiASRPipeline.transcribe
 *     (:474-539), reached through _transcribe_chunk (:541-575) and transcribe_chunked (:577-682)
 *         -> oasr_forward_ctc / oasr_transcribe_host
 *   - upstream units the fork removed (CONTRIBUTING.md:21 still names
 *     omnilingual_asr.models.inference.pipeline.ASRInferencePipeline; fairseq2.models.wav2vec2.*):
 *     wave layer_norm -> oasr_wave_norm; Wav2Vec2FeatureExtractor -> oasr_fe_layer0 + oasr_conv_ln_gelu;
 *     Wav2Vec2PositionEncoder -> oasr_posconv; encoder linears -> oasr_gemm; MHA -> oasr_attention;
 *     LayerNorm -> oasr_layernorm; final_proj + argmax -> oasr_gemm(OASR_EPI_ARGMAX) + oasr_ctc_decode;
 *     greedy collapse -> oasr_ctc_collapse.
 */
#ifndef OASR_H_
#define OASR_H_

#include <stdint.h>

#if defined(__GNUC__)
#define OASR_API __attribute__((visibility("default")))
#else
#define OASR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OasrEngine* OasrHandle;
typedef void* OasrStream; /* cudaStream_t; NULL = default stream */

enum {
  OASR_OK = 0,
  OASR_ERR_INVALID = -1,     /* bad argument / shape             (ValueError in the Python host)  */
  OASR_ERR_CUDA = -2,        /* CUDA runtime or driver failure   (RuntimeError)                   */
  OASR_ERR_STATE = -3,       /* wrong call order                 (RuntimeError)                   */
  OASR_ERR_UNSUPPORTED = -4  /* architecture outside the path    (ValueError)                     */
};

enum { OASR_DTYPE_F32 = 0, OASR_DTYPE_BF16 = 1 };

/* GEMM epilogues (oasr_gemm) */
enum {
  OASR_EPI_BF16 = 0,
  OASR_EPI_BF16_GELU = 1,
  OASR_EPI_F32 = 2,
  OASR_EPI_F32_RESID = 3,
  OASR_EPI_ARGMAX = 4,
  OASR_EPI_LN_GELU_BF16 = 5,
  OASR_EPI_F32_GELU_RESID = 6
};

/* forward flags */
enum {
  OASR_FLAG_INPUT_NORMALISED = 1, /* skip the per-window normalisation (a8) */
  OASR_FLAG_INPUT_I16 = 2         /* the waveform pointer holds PCM16 samples (int16, mono, 16 kHz): the conversion
                                     x / 32768 is fused into the normalisation kernels (device-side audio front end) */
};

typedef struct OasrConfig {
  int32_t d_model;
  int32_t n_layers;
  int32_t n_heads;
  int32_t d_ffn;
  int32_t vocab;
  int32_t fe_dim;        /* 512 */
  int32_t pos_kernel;    /* 128 */
  int32_t pos_groups;    /* 16 */
  int32_t n_fe_layers;   /* 7 */
  int32_t fe_kernel[8];  /* 10,3,3,3,3,2,2 */
  int32_t fe_stride[8];  /* 5,2,2,2,2,2,2 */
  int32_t blank_id;      /* 0 */
} OasrConfig;

OASR_API const char* oasr_version(void);
OASR_API const char* oasr_last_error(void);

/* ---- engine lifetime and weights ------------------------------------------------------------------- */
OASR_API int oasr_create(const OasrConfig* cfg, OasrHandle* out);
OASR_API void oasr_destroy(OasrHandle h);
/* `data` may be a host or a device pointer (fp32 or bf16, row-major, shape as in oracle.weight_shapes). */
OASR_API int oasr_load_weight(OasrHandle h, const char* name, const void* data, int dtype, const int64_t* shape, int ndim);
/* Folds weight-norm, repacks conv filters tap-major, converts tensor-core operands to bf16. */
OASR_API int oasr_finalize_weights(OasrHandle h);

/* ---- tensor parallelism (7B encoder, BASELINE config 4; upstream has no counterpart: a replica per GPU is the
 * default and needs none of this) --------------------------------------------------------------------------
 * The encoder layers are split Megatron-style over `world` ranks, one process and one handle per GPU: q/k/v and
 * FFN1 by output columns (whole heads), out-proj and FFN2 by input columns.  Every rank loads the FULL weights with
 * oasr_load_weight and keeps its slice at finalize.  Two ways to sum the partial outputs of out-proj / FFN2:
 *   nccl   one ncclAllReduce (fp32; NCCL loaded with dlopen) followed by one add + LayerNorm pass;
 *   peer   (after oasr_tp_ipc_export / _import) PUSH-based over NVLink peer memory: the GEMM epilogue writes each
 *          partial row, rounded to bf16, into the receive region of the rank that owns the row; the owner adds the
 *          `world` partial rows to its slice of the fp32 residual stream in rank order, applies LayerNorm and stores the
 *          bf16 row into every rank's buffer; flags in peer memory order the two steps.
 * oasr_tp_unique_id: rank 0 obtains the 128-byte NCCL id, the host broadcasts it.
 * oasr_tp_init: collective; call after oasr_create and before oasr_finalize_weights.
 * oasr_tp_emulate: one handle computes all `world` shards in turn and reduces them locally with the peer path's
 * kernel (no communicator) - the single-GPU parity check of the slicing and of the rounding points.
 * All ranks must call the forward with the SAME batch, in the same order, together: a rank that waits longer than
 * OASR_TP_TIMEOUT_MS (default 60 000) for a peer's flag gives up, and this and every later call on the handle returns
 * OASR_ERR_STATE (nothing traps; rebuild the group's engines). */
OASR_API int oasr_tp_unique_id(void* id_out_128_bytes);
OASR_API int oasr_tp_init(OasrHandle h, int32_t rank, int32_t world, const void* id_128_bytes);
OASR_API int oasr_tp_emulate(OasrHandle h, int32_t world);
/* Peer-memory path (after oasr_tp_init and before the first forward):
 * export: allocates this rank's arena (residual stream, LayerNorm rows, receive region, flags) for batches up to
 *         (B, L) and returns its 64-byte cudaIpcMemHandle_t;
 * import: takes the `world` handles in rank order (the host all-gathers them) and maps the peers. */
OASR_API int oasr_tp_ipc_export(OasrHandle h, int32_t B, int32_t L, void* handle_out_64_bytes);
OASR_API int oasr_tp_ipc_import(OasrHandle h, const void* handles_world_x_64_bytes);

/* Frames produced for n_samples input samples: chain of floor((L-k)/s)+1 over the FE layers. */
OASR_API int32_t oasr_feature_length(const OasrConfig* cfg, int64_t n_samples);

/* ---- the hot path ------------------------------------------------------------------------------------ */
/* wave_dev [B, L] fp32 (row stride wave_stride elements), zero padded beyond n_samples_host[b].
 * Outputs (device, each may be NULL): frame_ids [B,Tmax] int32; hidden [B,Tmax,d] fp32 (final LayerNorm);
 * out_ids / out_frames [B,Tmax] int32 (collapsed ids and their first frame, -1 padded); out_lens [B]. */
OASR_API int oasr_forward_ctc(OasrHandle h, const float* wave_dev, int64_t wave_stride, const int32_t* n_samples_host,
                     int32_t B, int32_t L, int32_t flags, int32_t* frame_ids_dev, float* hidden_dev,
                     int32_t* out_ids_dev, int32_t* out_frames_dev, int32_t* out_lens_dev, OasrStream stream);

/* Same from HOST buffers: H2D of the waveform, forward, D2H of the ids, stream synchronised on return.
 * Pinned host memory gives asynchronous copies.  Host outputs: out_ids / out_frames [B,Tmax], out_lens [B],
 * frame_ids [B,Tmax] (may be NULL). */
OASR_API int oasr_transcribe_host(OasrHandle h, const float* wave_host, int64_t wave_stride, const int32_t* n_samples_host,
                         int32_t B, int32_t L, int32_t flags, int32_t* out_ids_host, int32_t* out_frames_host,
                         int32_t* out_lens_host, int32_t* frame_ids_host, OasrStream stream);

/* The same call split in two, for a host loop that keeps the GPU fed (the reference keeps up to MAX_PARALLEL_CHUNKS = 4
 * requests in flight per recording, gemini_pipeline.py:217-219, 623-641: here the requests in flight are batches of
 * windows).  oasr_transcribe_host_async enqueues H2D + forward + D2H and returns a ticket without waiting; at most two
 * tickets may be outstanding per handle - each has its own device landing buffer, and the H2D copy runs on a copy
 * stream of the engine, so the waveform of batch k + 1 crosses PCIe under the forward of batch k.  The host buffers of
 * a ticket (PINNED memory, or the copies are not asynchronous) must stay untouched until oasr_wait(ticket) returns; the
 * outputs are valid after it.  stream = NULL: the engine's own non-blocking stream.  Tickets complete in order. */
OASR_API int oasr_transcribe_host_async(OasrHandle h, const float* wave_host, int64_t wave_stride,
                               const int32_t* n_samples_host, int32_t B, int32_t L, int32_t flags, int32_t* out_ids_host,
                               int32_t* out_frames_host, int32_t* out_lens_host, int32_t* frame_ids_host,
                               OasrStream stream, int64_t* ticket_out);
OASR_API int oasr_wait(OasrHandle h, int64_t ticket);

/* Debug/parity: run the forward up to `stop_stage` (1 = FE, 2 = projection, 3 = pos-conv, 4+l = encoder
 * layer l, 0 = everything) and expose internal device buffers by name: "fe" bf16 [B,Tpad,512],
 * "x" fp32 [B*T,d], "wave" fp32 [B,L].  shape gets up to 4 extents (0-terminated). */
OASR_API int oasr_debug_forward(OasrHandle h, const float* wave_dev, int64_t wave_stride, const int32_t* n_samples_host,
                       int32_t B, int32_t L, int32_t flags, int32_t stop_stage, OasrStream stream);
OASR_API int oasr_debug_buffer(OasrHandle h, const char* name, void** dev_ptr, int64_t* shape4, int32_t* dtype);
/* Device-to-device copy of `nbytes` of a named internal buffer into dst_dev (synchronous). */
OASR_API int oasr_debug_copy(OasrHandle h, const char* name, void* dst_dev, int64_t nbytes);

/* Kernel launches issued by this handle since creation (bench.py reports it as gpu_launches). */
OASR_API int64_t oasr_launch_count(OasrHandle h);

/* Per-stage device timing: when enabled, a CUDA event is recorded on the forward's stream at every stage
 * boundary; oasr_profile_read synchronises, returns accumulated milliseconds and stage counts per category
 * (arrays of OASR_PROF_NCAT entries) and resets the accumulators. */
enum {
  OASR_PROF_WAVE_NORM = 0, OASR_PROF_FE0, OASR_PROF_FE_CONV, OASR_PROF_LAYERNORM, OASR_PROF_PROJ,
  OASR_PROF_POSCONV, OASR_PROF_QKV, OASR_PROF_ATTENTION, OASR_PROF_OUTPROJ, OASR_PROF_FFN1, OASR_PROF_FFN2,
  OASR_PROF_CTC_HEAD, OASR_PROF_DECODE, OASR_PROF_ALLREDUCE, OASR_PROF_END, OASR_PROF_NCAT
};
OASR_API int oasr_profile_enable(OasrHandle h, int32_t on);
OASR_API int oasr_profile_read(OasrHandle h, double* ms, int64_t* counts, int32_t n);

/* ---- per-stage entry points (unit parity; all pointers are device pointers) ------------------------- */
OASR_API int oasr_wave_norm(const float* in, float* out, const int32_t* n_samples_dev, int32_t B, int32_t L, OasrStream stream);
/* Step before the path (device-side audio front end): interleaved [n_in, channels] fp32 or PCM16 samples at sr_in ->
 * mono fp32 at sr_out: channel mean + band-limited polyphase resampling with the filter of
 * torchaudio.functional.resample's defaults (what upstream's host pipeline applies before the model).
 * out_dev must hold oasr_resample_length(n_in, sr_in, sr_out) = ceil(n_in * sr_out / sr_in) samples. */
OASR_API int64_t oasr_resample_length(int64_t n_in, int32_t sr_in, int32_t sr_out);
OASR_API int oasr_resample(const void* in_dev, int32_t in_is_i16, int64_t n_in, int32_t channels, int32_t sr_in,
                  int32_t sr_out, float* out_dev, int64_t out_capacity, OasrStream stream);
OASR_API int oasr_fe_layer0(const float* wave, int32_t B, int32_t L, const float* w_10x512, const float* bias,
                   const float* gamma, const float* beta, void* out_bf16 /*[B,T0,512]*/, OasrStream stream);
/* Stride-2 Conv1d(512->512, k in {2,3}) + bias + LayerNorm(512) + GELU as an implicit GEMM.
 * in bf16 [B, L_in_pad, 512] (L_in_pad even, >= L_in + 2), w bf16 [512, k*512] tap-major, out bf16 [B, L_out, 512]. */
OASR_API int oasr_conv_ln_gelu(const void* in_bf16, int32_t B, int32_t L_in, int32_t L_in_pad, int32_t k, const void* w_bf16,
                      const float* bias, const float* gamma, const float* beta, void* out_bf16, OasrStream stream);
OASR_API int oasr_layernorm(const void* in, int32_t in_is_bf16, int64_t rows, int32_t D, const float* gamma, const float* beta,
                   void* out_bf16, float* out_f32, OasrStream stream);
/* out[M, ldo] = epilogue(A[M,K] . W[N,K]^T + bias) */
OASR_API int oasr_gemm(const void* A_bf16, const void* W_bf16, const float* bias, int32_t M, int32_t N, int32_t K,
              int32_t epilogue, void* out, int32_t ldo, const float* resid, const float* ln_gamma,
              const float* ln_beta, uint64_t* argmax_keys, OasrStream stream);
/* x[B*T, d] += gelu(grouped_conv(x) + bias); w bf16 [groups][d/groups][k * k_pad] (tap-major, zero padded). */
OASR_API int oasr_posconv(float* x, int32_t B, int32_t T, int32_t d, int32_t groups, int32_t k, const void* w_bf16,
                 const float* bias, void* scratch_bf16 /*[B, T+k, d]*/, OasrStream stream);
OASR_API int oasr_attention(const void* qkv_bf16, void* out_bf16, const int32_t* n_frames_dev, int32_t B, int32_t T, int32_t H,
                   int32_t hd, float scale, OasrStream stream);
OASR_API int oasr_ctc_decode(const uint64_t* keys, const int32_t* n_frames_dev, int32_t B, int32_t T, int32_t blank,
                    int32_t* frame_ids, int32_t* out_ids, int32_t* out_frames, int32_t* out_lens, OasrStream stream);
OASR_API int oasr_ctc_collapse(const int32_t* frame_ids, const int32_t* n_frames_dev, int32_t B, int32_t T, int32_t blank,
                      int32_t* out_ids, int32_t* out_frames, int32_t* out_lens, OasrStream stream);

#ifdef __cplusplus
}
#endif
#endif /* OASR_H_ */
