// liboasr engine: weight store, workspace, the a8-a16 forward schedule and the C-ABI (include/oasr.h).
#include "../../include/oasr.h"
#include "gemm.cuh"
#include "host_util.h"
#include "kernels.cuh"
#include "ptx.cuh"
#include "tp_comm.h"
#include "tp_fused.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace oasr;

namespace {

struct DevBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
  std::vector<int64_t> shape;
  int dtype = OASR_DTYPE_F32;
};

int dev_alloc(void** p, size_t bytes, bool zero) {
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess)
    return fail(OASR_ERR_CUDA, "cudaMalloc(" + std::to_string(bytes) + " bytes): " + cudaGetErrorString(e));
  if (zero) {
    e = cudaMemset(*p, 0, bytes);
    if (e != cudaSuccess) return fail(OASR_ERR_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(e));
  }
  return OASR_OK;
}

// ---------------------------------------------------------------------------------------------
// weight preparation kernels (run once in oasr_finalize_weights)
// ---------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16(in[i]);
}
// conv weight [N][C][J] fp32 -> tap-major bf16 [N][J*k_pad + c], zero for c >= C
__global__ void repack_conv_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int N, int C, int J,
                                   int k_pad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * J * k_pad;
  if (i >= total) return;
  const int c = int(i % k_pad);
  const int j = int((i / k_pad) % J);
  const int n = int(i / ((long long)k_pad * J));
  out[i] = c < C ? __float2bfloat16(in[((long long)n * C + c) * J + j]) : __float2bfloat16(0.f);
}
// layer-0 filter [512][1][10] -> [10][512] fp32
__global__ void transpose_l0_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * K) out[(i % K) * C + (i / K)] = in[i];
}
// weight-norm(dim=2): per tap j, norm over (out, in) of v[:, :, j]
__global__ void posnorm_kernel(const float* __restrict__ v, float* __restrict__ norms, long long rows, int J) {
  const int j = blockIdx.x;
  double s = 0.0;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    const double x = v[r * J + j];
    s += x * x;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norms[j] = (float)sqrt(sh[0]);
}
__global__ void posfold_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ norms,
                               float* __restrict__ w, long long n, int J) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int j = int(i % J);
    w[i] = v[i] * (g[j] / norms[j]);
  }
}

// out[r][c] = bf16(in[r][col0 + c]) for c < ncols: the input-column slice of a row-parallel weight
__global__ void slice_cols_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long rows,
                                       int in_ld, int col0, int ncols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ncols) return;
  const long long r = i / ncols;
  const int c = int(i - r * ncols);
  out[i] = __float2bfloat16(in[r * in_ld + col0 + c]);
}

inline unsigned blocks_for(long long n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

struct LayerW {
  float *attn_ln_g, *attn_ln_b, *ffn_ln_g, *ffn_ln_b;
  __nv_bfloat16 *wqkv, *wo, *w1, *w2;
  float *bqkv, *bo, *b1, *b2;
  // tensor parallelism: slices per local shard (one in a real run, `world` when emulated); bo / b2 stay whole and
  // are added by shard 0 only
  std::vector<__nv_bfloat16*> s_wqkv, s_wo, s_w1, s_w2;
  std::vector<float*> s_bqkv, s_b1;
};

}  // namespace

struct OasrEngine {
  OasrConfig cfg{};
  int device = 0;
  bool finalized = false;
  std::map<std::string, DevBuf> raw;       // fp32 masters as loaded
  std::vector<void*> owned;                // everything to cudaFree at destroy
  int64_t launches = 0;

  // finalised weights
  float* fe0_w = nullptr;                  // [10][512]
  std::vector<float*> fe_bias, fe_g, fe_b; // per FE layer
  std::vector<__nv_bfloat16*> fe_w;        // layers >= 1, [512][k*512]
  float *proj_ln_g = nullptr, *proj_ln_b = nullptr, *proj_bias = nullptr;
  __nv_bfloat16* proj_w = nullptr;
  __nv_bfloat16* pos_w = nullptr;          // [G][cg][K*k_pad]
  float* pos_bias = nullptr;
  int pos_kpad = 0;
  std::vector<LayerW> layers;
  float *final_ln_g = nullptr, *final_ln_b = nullptr, *ctc_bias = nullptr;
  __nv_bfloat16* ctc_w = nullptr;

  // workspace (grown on demand)
  int ws_B = 0, ws_L = 0;
  std::vector<void*> ws_owned;
  float* wave = nullptr;       // [B, L]
  double* wave_partials = nullptr;
  __nv_bfloat16* fe_buf[2] = {nullptr, nullptr};
  __nv_bfloat16* lnbuf = nullptr;   // [M, max(512, d)]
  float* x = nullptr;               // [M, d]
  __nv_bfloat16* xpad = nullptr;    // [B, T+K, d]
  __nv_bfloat16* qkv = nullptr;     // [M, 3d]
  __nv_bfloat16* att = nullptr;     // [M, d]
  __nv_bfloat16* ffn = nullptr;     // [M, F]
  unsigned long long* keys = nullptr;
  int *n_samples_dev = nullptr, *n_frames_dev = nullptr;
  int *frame_ids = nullptr, *out_ids = nullptr, *out_frames = nullptr, *out_lens = nullptr;
  // pinned staging for n_samples / n_frames: a small ring so that the host can run ahead of the stream
  static constexpr int STAGE_SLOTS = 8;
  int32_t* h_stage = nullptr;       // [STAGE_SLOTS][2 * ws_B]
  cudaEvent_t stage_ev[STAGE_SLOTS] = {};
  bool stage_used[STAGE_SLOTS] = {};
  int stage_next = 0;
  // oasr_transcribe_host(_async): ASYNC_SLOTS landing buffers for the waveform, so that the H2D copy of batch k + 1
  // (on h2d_stream) runs under the forward of batch k; a slot is busy from its submit to the oasr_wait of its ticket
  static constexpr int ASYNC_SLOTS = 2;
  float* wave_raw[ASYNC_SLOTS] = {nullptr, nullptr};   // [B, L] each
  cudaStream_t h2d_stream = nullptr, own_stream = nullptr;
  cudaEvent_t h2d_done[ASYNC_SLOTS] = {}, slot_done[ASYNC_SLOTS] = {};
  bool slot_busy[ASYNC_SLOTS] = {};
  int32_t* slot_lens_host[ASYNC_SLOTS] = {};           // empty-output case: zeroed at wait
  int slot_B[ASYNC_SLOTS] = {};
  int64_t tickets = 0;                                  // ticket t lives in slot t % ASYNC_SLOTS
  // optional per-stage timing (bench.py): an event at every stage boundary of the forward
  bool profiling = false;
  // CUDA graphs of the whole forward for small batches (latency path): one per (input buffer, B, L, flags), captured
  // the second time a shape is seen and replayed from then on, on the engine's own stream
  struct GraphEntry {
    const void* wave;
    long long stride;
    int B, L, flags;
    int state;   // 1: seen once (ran eagerly), 2: captured, -1: capture failed, stay eager
    cudaGraphExec_t exec;
    long long launches;
    int last_T, last_fe_idx;
    long long last_fe_pad;
  };
  std::vector<GraphEntry> graphs;
  cudaStream_t graph_stream = nullptr;
  cudaEvent_t graph_fork = nullptr, graph_join = nullptr;
  std::vector<std::pair<int, cudaEvent_t>> prof_marks;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[OASR_PROF_NCAT] = {};
  int64_t prof_n[OASR_PROF_NCAT] = {};
  // tensor parallelism (encoder layers only): this handle computes shards [tp_first, tp_first + tp_local) of tp_world
  int tp_world = 1, tp_first = 0, tp_local = 1;
  bool tp_emulated = false;
  void* tp_comm = nullptr;
  float* part = nullptr;            // [M, d] fp32 partial sums of the row-parallel GEMMs (NCCL mode only)
  __nv_bfloat16* part_bf16 = nullptr;   // emulation of the split on one GPU: [shards][M, d] bf16 partial sums
  long long part_shard_stride = 0;  // elements between the shards' partials (emulation)
  unsigned int* tp_err_host = nullptr;   // host-mapped error word the flag waits set on a timeout
  unsigned int* tp_err_dev = nullptr;
  // peer-memory path (tp_fused.cu): x | ln | part | flags of this rank live in one IPC-exported arena
  void* tp_arena = nullptr;
  size_t tp_off_ln = 0, tp_off_part = 0, tp_off_flags = 0, tp_x_bytes = 0;
  std::vector<void*> tp_peer_base;  // [world], own arena at [rank]
  TpPeerView tp_view{};
  TpPeerView tp_view2{};            // second flag set (half-batch 1 in overlap mode)
  unsigned long long tp_epoch2 = 0;
  cudaStream_t tp_comm_stream2 = nullptr;
  bool tp_fused = false;
  unsigned long long tp_epoch = 0;
  // overlap mode of the peer-memory path: the reduce kernels run on their own stream beside the other half-batch's GEMMs
  cudaStream_t tp_comm_stream = nullptr;
  cudaEvent_t tp_ev_compute[2] = {nullptr, nullptr}, tp_ev_reduce[2] = {nullptr, nullptr};
  // shapes of the last forward (debug buffers)
  int last_B = 0, last_L = 0, last_T = 0, last_fe_idx = 0;
  long long last_fe_pad = 0;
};

namespace {

int fe_len(const OasrConfig& c, int64_t n, int upto) {
  for (int i = 0; i < upto; ++i) n = n >= c.fe_kernel[i] ? (n - c.fe_kernel[i]) / c.fe_stride[i] + 1 : 0;
  return (int)n;
}
inline long long fe_pad_rows(int t) { return ((long long)t + 3) & ~1ll; }  // even and >= t + 2

int get_raw(OasrEngine* e, const std::string& name, std::initializer_list<int64_t> shape, float** out) {
  auto it = e->raw.find(name);
  if (it == e->raw.end()) return fail(OASR_ERR_STATE, "weight not loaded: " + name);
  if (it->second.shape != std::vector<int64_t>(shape)) return fail(OASR_ERR_INVALID, "weight has wrong shape: " + name);
  *out = reinterpret_cast<float*>(it->second.ptr);
  return OASR_OK;
}

int alloc_owned(OasrEngine* e, void** p, size_t bytes, bool zero = false) {
  OASR_TRY(dev_alloc(p, bytes, zero));
  e->owned.push_back(*p);
  return OASR_OK;
}

int to_bf16(OasrEngine* e, const float* src, long long n, __nv_bfloat16** dst) {
  OASR_TRY(alloc_owned(e, reinterpret_cast<void**>(dst), (size_t)n * 2));
  cast_bf16_kernel<<<blocks_for(n), 256>>>(src, *dst, n);
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

void free_raw(OasrEngine* e, const std::string& name) {
  auto it = e->raw.find(name);
  if (it != e->raw.end()) {
    cudaFree(it->second.ptr);
    e->raw.erase(it);
  }
}

int ensure_workspace(OasrEngine* e, int B, int L) {
  if (B <= e->ws_B && L <= e->ws_L) return OASR_OK;
  const int nB = std::max(B, e->ws_B), nL = std::max(L, e->ws_L);
  OASR_CUDA_CHECK(cudaDeviceSynchronize());
  for (auto& g : e->graphs)   // captured graphs hold pointers into the workspace that is about to be replaced
    if (g.exec) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
  for (void* p : e->ws_owned) cudaFree(p);
  e->ws_owned.clear();
  e->ws_B = e->ws_L = 0;
  const OasrConfig& c = e->cfg;
  const int d = c.d_model, F = c.d_ffn;
  auto A = [&](void** p, size_t bytes, bool zero) -> int {
    OASR_TRY(dev_alloc(p, bytes, zero));
    e->ws_owned.push_back(*p);
    return OASR_OK;
  };
  const long long T0 = fe_len(c, nL, 1), T1 = fe_len(c, nL, 2);
  const long long T = fe_len(c, nL, c.n_fe_layers);
  const long long M = (long long)nB * std::max<long long>(T, 1);
  OASR_TRY(A((void**)&e->wave, (size_t)nB * nL * 4, false));
  OASR_TRY(A((void**)&e->wave_partials, (size_t)nB * WAVE_NORM_SLICES * 2 * 8, false));
  OASR_TRY(A((void**)&e->fe_buf[0], (size_t)nB * fe_pad_rows((int)T0) * 512 * 2, true));
  OASR_TRY(A((void**)&e->fe_buf[1], (size_t)nB * fe_pad_rows((int)T1) * 512 * 2, true));
  if (e->tp_arena != nullptr) {
    if ((size_t)M * d * 4 > e->tp_x_bytes)
      return fail(OASR_ERR_INVALID, "batch exceeds the (B, L) the tensor-parallel peer arena was exported for");
    e->x = reinterpret_cast<float*>(e->tp_arena);
    e->lnbuf = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(e->tp_arena) + e->tp_off_ln);
  } else {
    OASR_TRY(A((void**)&e->lnbuf, (size_t)M * std::max(512, d) * 2, false));
    OASR_TRY(A((void**)&e->x, (size_t)M * d * 4, false));
  }
  OASR_TRY(A((void**)&e->xpad, (size_t)nB * (T + c.pos_kernel) * d * 2, true));
  OASR_TRY(A((void**)&e->qkv, (size_t)M * 3 * d * 2, false));
  OASR_TRY(A((void**)&e->att, (size_t)M * d * 2, false));
  OASR_TRY(A((void**)&e->ffn, (size_t)M * F * 2, false));
  if (e->tp_arena != nullptr) {
    // peer-memory path: the partial sums never sit in a buffer of their producer - the GEMM epilogue routes them into
    // the owners' receive regions (tp_view.recv, inside the arenas)
  } else if (e->tp_emulated) {
    e->part_shard_stride = M * d;
    OASR_TRY(A((void**)&e->part_bf16, (size_t)e->tp_local * M * d * 2, false));
  } else if (e->tp_world > 1) {
    OASR_TRY(A((void**)&e->part, (size_t)M * d * 4, false));
  }
  OASR_TRY(A((void**)&e->keys, (size_t)M * 8, true));
  OASR_TRY(A((void**)&e->n_samples_dev, (size_t)nB * 4, true));
  OASR_TRY(A((void**)&e->n_frames_dev, (size_t)nB * 4, true));
  OASR_TRY(A((void**)&e->frame_ids, (size_t)M * 4, true));
  OASR_TRY(A((void**)&e->out_ids, (size_t)M * 4, true));
  OASR_TRY(A((void**)&e->out_frames, (size_t)M * 4, true));
  OASR_TRY(A((void**)&e->out_lens, (size_t)nB * 4, true));
  if (e->h_stage) cudaFreeHost(e->h_stage);
  OASR_CUDA_CHECK(cudaMallocHost((void**)&e->h_stage, (size_t)OasrEngine::STAGE_SLOTS * nB * 2 * 4));
  for (int i = 0; i < OasrEngine::STAGE_SLOTS; ++i) e->stage_used[i] = false;
  for (int i = 0; i < OasrEngine::ASYNC_SLOTS; ++i) OASR_TRY(A((void**)&e->wave_raw[i], (size_t)nB * nL * 4, false));
  e->ws_B = nB;
  e->ws_L = nL;
  return OASR_OK;
}

void prof_mark(OasrEngine* e, int cat, cudaStream_t st) {
  if (!e->profiling) return;
  cudaEvent_t ev;
  if (!e->prof_pool.empty()) {
    ev = e->prof_pool.back();
    e->prof_pool.pop_back();
  } else if (cudaEventCreate(&ev) != cudaSuccess) {
    return;
  }
  cudaEventRecord(ev, st);
  e->prof_marks.emplace_back(cat, ev);
}

// One FE conv layer (i >= 1) as an implicit GEMM with fused LayerNorm + GELU.
int run_conv_layer(const void* in, int B, int T_in, long long in_pad_rows, int k, const void* w, const float* bias,
                   const float* g, const float* b, void* out, long long out_pad_rows, cudaStream_t st) {
  const int T_out = T_in >= k ? (T_in - k) / 2 + 1 : 0;
  if (T_out == 0) return OASR_OK;
  GemmArgs a;
  a.A = in;
  a.a_inner = 512;
  a.taps = k;
  a.k_pad = 512;
  a.P = 2;
  a.a_p_stride = 512;
  a.a_pos_stride = 1024;
  a.a_batch_stride = in_pad_rows * 512;
  a.a_positions = T_out + (k - 1) / 2;
  a.rows_per_batch = T_out;
  a.batches = B;
  a.W = w;
  a.N = 512;
  a.bias = bias;
  a.ln_gamma = g;
  a.ln_beta = b;
  a.out = out;
  a.ldo = 512;
  a.out_batch_rows = out_pad_rows;
  a.epilogue = EPI_LN_GELU_BF16;
  return gemm_bf16_tcgen05(a, st);
}

int run_posconv(float* x, int B, int T, int d, int groups, int k, int k_pad, const void* w, const float* bias,
                void* xpad, cudaStream_t st) {
  const int cg = d / groups;
  OASR_CUDA_CHECK(cudaMemsetAsync(xpad, 0, (size_t)B * (T + k) * d * 2, st));
  OASR_TRY(pad_cast_bf16(x, B, T, d, k / 2, xpad, st));
  GemmArgs a;
  a.A = xpad;
  a.a_inner = cg;
  a.taps = k;
  a.k_pad = k_pad;
  a.P = 1;
  a.a_pos_stride = d;
  a.a_batch_stride = (long long)(T + k) * d;
  a.a_group_stride = cg;
  a.a_positions = T + k;
  a.rows_per_batch = T;
  a.batches = B;
  a.groups = groups;
  a.W = w;
  a.N = cg;
  a.bias = bias;
  a.out = x;
  a.ldo = d;
  a.resid = x;
  a.epilogue = EPI_F32_GELU_RESID;
  return gemm_bf16_tcgen05(a, st);
}

// Row-parallel GEMM of the peer-memory path: output row m of the `rows` rows of this reduction goes to the rank that
// owns it, into slot `my rank` of that rank's receive region (tp_fused.h: TpPeerView::recv).
void tp_route_rows(OasrEngine* e, GemmArgs& a, long long recv_off, long long rows) {
  const int W = e->tp_world;
  const TpShare mine = tp_share(rows, e->tp_first, W);
  a.out = nullptr;
  a.route_n = W;
  a.route_per = (int)mine.per;
  for (int q = 0; q < W; ++q)
    a.route_base[q] = e->tp_view.recv[q] + recv_off + (long long)e->tp_first * mine.slot_rows * e->cfg.d_model;
}

// Tensor-parallel encoder layers, peer-memory path, B >= 2: the batch is cut into two half-batches of windows
// (attention never crosses a window).  All compute kernels stay on the caller's stream, in the order
// P1(h0) P1(h1) P2(h0) P2(h1) per layer (P1 = QKV, attention, out-proj; P2 = FFN1, FFN2).  The row-parallel GEMMs
// (out-proj, FFN2) PUSH their partial sums into the owners' receive regions from their epilogues (route_rows); the
// tail of a half-batch's reduction (tp_push_reduce_layernorm: flag hand-shake, add + LayerNorm + LN rows pushed to
// every rank, flag hand-shake) runs on that half-batch's own high-priority stream as soon as its GEMM has finished,
// i.e. beside the OTHER half-batch's next phase, and the phase that consumes its LayerNorm output waits for it through
// an event.  Each half-batch has its own flag set, epoch counter and receive region.
int tp_layers_overlapped(OasrEngine* e, int B, int T, float* hidden_out, int stop_stage, cudaStream_t st) {
  const OasrConfig& c = e->cfg;
  const int d = c.d_model, F = c.d_ffn, H = c.n_heads, hd = d / H;
  const int W = e->tp_world, d_loc = d / W, F_loc = F / W, H_loc = H / W;
  const float scale = 1.0f / sqrtf((float)hd);
  if (e->tp_comm_stream == nullptr) {
    // highest priority: when an SM frees up between two compute kernels, the (short) reduce work goes first
    int prio_lo = 0, prio_hi = 0;
    OASR_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    OASR_CUDA_CHECK(cudaStreamCreateWithPriority(&e->tp_comm_stream, cudaStreamNonBlocking, prio_hi));
    OASR_CUDA_CHECK(cudaStreamCreateWithPriority(&e->tp_comm_stream2, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < 2; ++i) {
      OASR_CUDA_CHECK(cudaEventCreateWithFlags(&e->tp_ev_compute[i], cudaEventDisableTiming));
      OASR_CUDA_CHECK(cudaEventCreateWithFlags(&e->tp_ev_reduce[i], cudaEventDisableTiming));
    }
  }
  const int Bh[2] = {(B + 1) / 2, B / 2};
  const int b0[2] = {0, Bh[0]};
  const long long r0[2] = {0, (long long)Bh[0] * T};
  const long long Mh[2] = {(long long)Bh[0] * T, (long long)Bh[1] * T};
  const long long M = Mh[0] + Mh[1];
  // receive regions of the two half-batches inside every rank's receive buffer: [source][slot_rows][d] each
  const long long recv_off[2] = {0, (long long)W * tp_share(Mh[0], 0, W).slot_rows * d};
  auto route = [&](GemmArgs& a, int h) { tp_route_rows(e, a, recv_off[h], Mh[h]); };
  auto reduce_ln = [&](int h, const float* g, const float* bta, bool bcast_x) -> int {
    // each half-batch has its own stream, flag set and receive region: the tail of one reduction runs beside the
    // compute of the other half-batch
    cudaStream_t cs = h == 1 ? e->tp_comm_stream2 : e->tp_comm_stream;
    OASR_CUDA_CHECK(cudaEventRecord(e->tp_ev_compute[h], st));
    OASR_CUDA_CHECK(cudaStreamWaitEvent(cs, e->tp_ev_compute[h], 0));
    if (h == 0) OASR_TRY(tp_push_reduce_layernorm(e->tp_view, recv_off[0], r0[0], Mh[0], d, g, bta, ++e->tp_epoch, bcast_x, cs));
    else OASR_TRY(tp_push_reduce_layernorm(e->tp_view2, recv_off[1], r0[1], Mh[1], d, g, bta, ++e->tp_epoch2, bcast_x, cs));
    OASR_CUDA_CHECK(cudaEventRecord(e->tp_ev_reduce[h], cs));
    return OASR_OK;
  };
  bool pending[2] = {false, false};   // a reduce of this half-batch is in flight: its consumer must wait for it
  auto wait_reduce = [&](int h) -> int {
    if (pending[h]) OASR_CUDA_CHECK(cudaStreamWaitEvent(st, e->tp_ev_reduce[h], 0));
    pending[h] = false;
    return OASR_OK;
  };
  prof_mark(e, OASR_PROF_LAYERNORM, st);
  if (c.n_layers > 0)
    OASR_TRY(layernorm_rows(e->x, 0, 0, 1, (int)M, d, e->layers[0].attn_ln_g, e->layers[0].attn_ln_b, e->lnbuf, nullptr, st));
  for (int l = 0; l < c.n_layers; ++l) {
    const LayerW& w = e->layers[l];
    const bool add_bias = e->tp_first == 0;
    for (int h = 0; h < 2; ++h) {
      OASR_TRY(wait_reduce(h));
      prof_mark(e, OASR_PROF_QKV, st);
      {
        GemmArgs a = GemmArgs::plain(e->lnbuf + r0[h] * d, (int)Mh[h], d, d, w.s_wqkv[0], 3 * d_loc);
        a.bias = w.s_bqkv[0];
        a.out = e->qkv + r0[h] * 3 * d_loc;
        a.ldo = 3 * d_loc;
        a.epilogue = EPI_BF16;
        OASR_TRY(gemm_bf16_tcgen05(a, st));
      }
      prof_mark(e, OASR_PROF_ATTENTION, st);
      OASR_TRY(attention_bf16(e->qkv + r0[h] * 3 * d_loc, e->att + r0[h] * d_loc, e->n_frames_dev + b0[h], Bh[h], T, H_loc,
                              hd, scale, st));
      prof_mark(e, OASR_PROF_OUTPROJ, st);
      {
        GemmArgs a = GemmArgs::plain(e->att + r0[h] * d_loc, (int)Mh[h], d_loc, d_loc, w.s_wo[0], d);
        a.bias = add_bias ? w.bo : nullptr;
        a.ldo = d;
        a.epilogue = EPI_BF16;   // this rank's partial sums, rounded to bf16, pushed to the ranks that own the rows
        route(a, h);
        OASR_TRY(gemm_bf16_tcgen05(a, st));
      }
      OASR_TRY(reduce_ln(h, w.ffn_ln_g, w.ffn_ln_b, false));
      pending[h] = true;
      e->launches += 6;   // 3 compute kernels + ready signal/wait, add + LayerNorm + push, done signal/wait
    }
    const bool last = l + 1 == c.n_layers;
    const float* g = last ? e->final_ln_g : e->layers[l + 1].attn_ln_g;
    const float* bta = last ? e->final_ln_b : e->layers[l + 1].attn_ln_b;
    const bool want_hidden = last && hidden_out != nullptr;   // parity runs: the fp32 LayerNorm output of all rows
    for (int h = 0; h < 2; ++h) {
      OASR_TRY(wait_reduce(h));
      prof_mark(e, OASR_PROF_FFN1, st);
      {
        GemmArgs a = GemmArgs::plain(e->lnbuf + r0[h] * d, (int)Mh[h], d, d, w.s_w1[0], F_loc);
        a.bias = w.s_b1[0];
        a.out = e->ffn + r0[h] * F_loc;
        a.ldo = F_loc;
        a.epilogue = EPI_BF16_GELU;
        OASR_TRY(gemm_bf16_tcgen05(a, st));
      }
      prof_mark(e, OASR_PROF_FFN2, st);
      {
        GemmArgs a = GemmArgs::plain(e->ffn + r0[h] * F_loc, (int)Mh[h], F_loc, F_loc, w.s_w2[0], d);
        a.bias = add_bias ? w.b2 : nullptr;
        a.ldo = d;
        a.epilogue = EPI_BF16;
        route(a, h);
        OASR_TRY(gemm_bf16_tcgen05(a, st));
      }
      OASR_TRY(reduce_ln(h, g, bta, want_hidden));
      pending[h] = true;
      e->launches += 5;
    }
    if (stop_stage == 4 + l) {
      OASR_TRY(wait_reduce(0));
      OASR_TRY(wait_reduce(1));
      return OASR_OK;
    }
  }
  OASR_TRY(wait_reduce(0));
  OASR_TRY(wait_reduce(1));
  if (hidden_out != nullptr) {
    prof_mark(e, OASR_PROF_LAYERNORM, st);
    OASR_TRY(layernorm_rows(e->x, 0, 0, 1, (int)M, d, e->final_ln_g, e->final_ln_b, e->lnbuf, hidden_out, st));
  }
  return OASR_OK;
}


// A flag wait of the peer-memory path that ran out of time leaves its mark in host-mapped memory (tp_fused.cu): the
// group is out of step from then on, so every later call fails until the engines are rebuilt.
int tp_check_error(OasrEngine* e) {
  if (e->tp_err_host != nullptr && *reinterpret_cast<volatile unsigned int*>(e->tp_err_host) != 0)
    return fail(OASR_ERR_STATE, "tensor parallelism: a peer rank did not reach a reduction within OASR_TP_TIMEOUT_MS "
                                "(ranks must run the same batches in the same order); rebuild the group's engines");
  return OASR_OK;
}

// window lengths -> device (pinned staging slots, so that the copies are asynchronous and a slot is not rewritten
// before its copy has run)
int stage_lengths(OasrEngine* e, const int32_t* n_samples_host, int B, int L, cudaStream_t st) {
  const OasrConfig& c = e->cfg;
  const int slot = e->stage_next;
  e->stage_next = (slot + 1) % OasrEngine::STAGE_SLOTS;
  if (e->stage_used[slot]) OASR_CUDA_CHECK(cudaEventSynchronize(e->stage_ev[slot]));
  int32_t* hs = e->h_stage + (size_t)slot * 2 * e->ws_B;
  for (int b = 0; b < B; ++b) {
    OASR_REQUIRE(n_samples_host[b] >= 0 && n_samples_host[b] <= L, "forward: n_samples out of range");
    hs[b] = n_samples_host[b];
    hs[B + b] = fe_len(c, n_samples_host[b], c.n_fe_layers);
  }
  OASR_CUDA_CHECK(cudaMemcpyAsync(e->n_samples_dev, hs, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  OASR_CUDA_CHECK(cudaMemcpyAsync(e->n_frames_dev, hs + B, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  if (!e->stage_ev[slot]) OASR_CUDA_CHECK(cudaEventCreateWithFlags(&e->stage_ev[slot], cudaEventDisableTiming));
  OASR_CUDA_CHECK(cudaEventRecord(e->stage_ev[slot], st));
  e->stage_used[slot] = true;
  return OASR_OK;
}

// staged: the window lengths are already on the device (graph capture / replay: only kernels and device-side
// memsets follow, nothing that reads host memory)
int forward_eager(OasrEngine* e, const float* wave_in, int64_t wave_stride, const int32_t* n_samples_host, int B, int L,
                  int flags, int stop_stage, float* hidden_out, bool staged, cudaStream_t st) {
  if (!e->finalized) return fail(OASR_ERR_STATE, "oasr_finalize_weights has not been called");
  OASR_REQUIRE(wave_in && n_samples_host && B > 0 && L > 0, "forward: bad arguments");
  OASR_TRY(tp_check_error(e));
  const OasrConfig& c = e->cfg;
  const int d = c.d_model, F = c.d_ffn, H = c.n_heads, hd = d / H;
  OASR_TRY(ensure_workspace(e, B, L));
  const int T = fe_len(c, L, c.n_fe_layers);
  const long long M = (long long)B * T;
  e->last_B = B;
  e->last_L = L;
  e->last_T = T;
  if (!staged) OASR_TRY(stage_lengths(e, n_samples_host, B, L, st));

  // a8
  prof_mark(e, OASR_PROF_WAVE_NORM, st);
  const float* wv = wave_in;
  long long wv_stride = wave_stride;
  if (flags & OASR_FLAG_INPUT_I16) {
    OASR_REQUIRE(!(flags & OASR_FLAG_INPUT_NORMALISED), "forward: PCM16 input cannot be flagged as normalised");
    OASR_TRY(wave_norm_i16(reinterpret_cast<const short*>(wave_in), e->wave, e->n_samples_dev, B, L, wave_stride, L,
                           e->wave_partials, st));
    e->launches += 2;
    wv = e->wave;
    wv_stride = L;
  } else if (!(flags & OASR_FLAG_INPUT_NORMALISED)) {
    OASR_TRY(wave_norm(wave_in, e->wave, e->n_samples_dev, B, L, wave_stride, L, e->wave_partials, st));
    e->launches += 2;
    wv = e->wave;
    wv_stride = L;
  }
  // a9
  prof_mark(e, OASR_PROF_FE0, st);
  int t_prev = fe_len(c, L, 1);
  long long pad_prev = fe_pad_rows(t_prev);
  OASR_TRY(fe_layer0(wv, wv_stride, B, L, e->fe0_w, e->fe_bias[0], e->fe_g[0], e->fe_b[0], e->fe_buf[0],
                     pad_prev * 512, t_prev, st));
  e->launches += 1;
  int cur = 0;
  // a10-a11
  prof_mark(e, OASR_PROF_FE_CONV, st);
  for (int i = 1; i < c.n_fe_layers; ++i) {
    const int t_out = t_prev >= c.fe_kernel[i] ? (t_prev - c.fe_kernel[i]) / 2 + 1 : 0;
    const long long pad_out = fe_pad_rows(t_out);
    OASR_TRY(run_conv_layer(e->fe_buf[cur], B, t_prev, pad_prev, c.fe_kernel[i], e->fe_w[i], e->fe_bias[i],
                            e->fe_g[i], e->fe_b[i], e->fe_buf[cur ^ 1], pad_out, st));
    e->launches += 1;
    cur ^= 1;
    t_prev = t_out;
    pad_prev = pad_out;
  }
  e->last_fe_idx = cur;
  e->last_fe_pad = pad_prev;
  if (stop_stage == 1 || T == 0) {
    prof_mark(e, OASR_PROF_END, st);
    return OASR_OK;
  }

  // a12
  prof_mark(e, OASR_PROF_LAYERNORM, st);
  OASR_TRY(layernorm_rows(e->fe_buf[cur], 1, pad_prev * 512, B, T, 512, e->proj_ln_g, e->proj_ln_b, e->lnbuf, nullptr, st));
  prof_mark(e, OASR_PROF_PROJ, st);
  {
    GemmArgs a = GemmArgs::plain(e->lnbuf, (int)M, 512, 512, e->proj_w, d);
    a.bias = e->proj_bias;
    a.out = e->x;
    a.ldo = d;
    a.n_valid = e->n_frames_dev;
    a.frames_per_seq = T;
    a.epilogue = EPI_F32;
    OASR_TRY(gemm_bf16_tcgen05(a, st));
  }
  e->launches += 2;
  if (stop_stage == 2) {
    prof_mark(e, OASR_PROF_END, st);
    return OASR_OK;
  }

  // a13
  prof_mark(e, OASR_PROF_POSCONV, st);
  OASR_TRY(run_posconv(e->x, B, T, d, c.pos_groups, c.pos_kernel, e->pos_kpad, e->pos_w, e->pos_bias, e->xpad, st));
  e->launches += 3;
  if (stop_stage == 3) {
    prof_mark(e, OASR_PROF_END, st);
    return OASR_OK;
  }

  // a14
  const float scale = 1.0f / sqrtf((float)hd);
  // OASR_TP_OVERLAP=off: no half-batch pipeline - every reduction's tail runs on the compute stream with nothing
  // beside it (what a single window takes anyway; scripts/tp_check.py covers both through the batch size)
  static const bool overlap = [] {
    const char* v = std::getenv("OASR_TP_OVERLAP");
    return v == nullptr || std::strcmp(v, "off") != 0;
  }();
  if (e->tp_world > 1 && e->tp_fused && e->tp_local == 1 && B >= 2 && c.n_layers > 0 && overlap) {
    OASR_TRY(tp_layers_overlapped(e, B, T, hidden_out, stop_stage, st));
    if (stop_stage >= 4 && stop_stage < 4 + c.n_layers) {
      prof_mark(e, OASR_PROF_END, st);
      return OASR_OK;
    }
  } else if (e->tp_world > 1) {
    // Tensor-parallel layers: q/k/v + attention + FFN1 on this rank's heads / hidden columns, out-proj and FFN2 as
    // partial sums over the local input columns -> all-reduce -> residual add fused into the next LayerNorm pass.
    const int W = e->tp_world, d_loc = d / W, F_loc = F / W, H_loc = H / W;
    // peer-memory runs and the one-GPU emulation of the split keep every shard's partial sum in bf16 and add them in
    // rank order (tp_fused.cu); the NCCL mode all-reduces fp32 partial sums
    const bool bf16_part = e->tp_fused || e->tp_emulated;
    auto emulated_reduce = [&](const float* g, const float* bta) -> int {
      const __nv_bfloat16* parts[TP_MAX_WORLD];
      for (int s = 0; s < e->tp_local; ++s) parts[s] = e->part_bf16 + s * e->part_shard_stride;
      return tp_local_reduce_layernorm(e->x, parts, e->tp_local, M, d, g, bta, e->lnbuf, st);
    };
    auto reduce_partial = [&]() -> int {
      if (e->tp_comm != nullptr) {
        prof_mark(e, OASR_PROF_ALLREDUCE, st);
        OASR_TRY(tp_allreduce_f32(e->tp_comm, e->part, (size_t)M * d, st));
      }
      return OASR_OK;
    };
    prof_mark(e, OASR_PROF_LAYERNORM, st);
    if (c.n_layers > 0)
      OASR_TRY(layernorm_rows(e->x, 0, 0, 1, (int)M, d, e->layers[0].attn_ln_g, e->layers[0].attn_ln_b, e->lnbuf, nullptr, st));
    for (int l = 0; l < c.n_layers; ++l) {
      const LayerW& w = e->layers[l];
      for (int s = 0; s < e->tp_local; ++s) {
        const bool first = s == 0;   // shard 0 of the handle starts the partial sum (and rank 0 adds the bias)
        const bool add_bias = e->tp_first + s == 0;
        prof_mark(e, OASR_PROF_QKV, st);
        {
          GemmArgs a = GemmArgs::plain(e->lnbuf, (int)M, d, d, w.s_wqkv[s], 3 * d_loc);
          a.bias = w.s_bqkv[s];
          a.out = e->qkv;
          a.ldo = 3 * d_loc;
          a.epilogue = EPI_BF16;
          OASR_TRY(gemm_bf16_tcgen05(a, st));
        }
        prof_mark(e, OASR_PROF_ATTENTION, st);
        OASR_TRY(attention_bf16(e->qkv, e->att, e->n_frames_dev, B, T, H_loc, hd, scale, st));
        prof_mark(e, OASR_PROF_OUTPROJ, st);
        {
          GemmArgs a = GemmArgs::plain(e->att, (int)M, d_loc, d_loc, w.s_wo[s], d);
          a.bias = add_bias ? w.bo : nullptr;
          a.ldo = d;
          if (e->tp_fused) {          // pushed to the owners' receive regions from the epilogue
            a.epilogue = EPI_BF16;
            tp_route_rows(e, a, 0, M);
          } else if (bf16_part) {     // every shard's partial sum on its own, rounded to bf16 as in a real run
            a.out = e->part_bf16 + s * e->part_shard_stride;
            a.epilogue = EPI_BF16;
          } else {
            a.out = e->part;
            a.resid = first ? nullptr : e->part;
            a.epilogue = first ? EPI_F32 : EPI_F32_RESID;
          }
          OASR_TRY(gemm_bf16_tcgen05(a, st));
        }
        e->launches += 3;
      }
      if (e->tp_fused) {   // the partial sums are in the owners' receive regions: add + LayerNorm + LN rows to every rank
        prof_mark(e, OASR_PROF_ALLREDUCE, st);
        OASR_TRY(tp_push_reduce_layernorm(e->tp_view, 0, 0, M, d, w.ffn_ln_g, w.ffn_ln_b, ++e->tp_epoch, false, st));
      } else if (e->tp_emulated) {
        prof_mark(e, OASR_PROF_ALLREDUCE, st);
        OASR_TRY(emulated_reduce(w.ffn_ln_g, w.ffn_ln_b));
      } else {
        OASR_TRY(reduce_partial());
        prof_mark(e, OASR_PROF_LAYERNORM, st);
        OASR_TRY(add_layernorm_rows(e->x, 0, 0, 1, (int)M, d, w.ffn_ln_g, w.ffn_ln_b, e->lnbuf, nullptr, e->part, e->x, st));
      }
      for (int s = 0; s < e->tp_local; ++s) {
        const bool first = s == 0;
        const bool add_bias = e->tp_first + s == 0;
        prof_mark(e, OASR_PROF_FFN1, st);
        {
          GemmArgs a = GemmArgs::plain(e->lnbuf, (int)M, d, d, w.s_w1[s], F_loc);
          a.bias = w.s_b1[s];
          a.out = e->ffn;
          a.ldo = F_loc;
          a.epilogue = EPI_BF16_GELU;
          OASR_TRY(gemm_bf16_tcgen05(a, st));
        }
        prof_mark(e, OASR_PROF_FFN2, st);
        {
          GemmArgs a = GemmArgs::plain(e->ffn, (int)M, F_loc, F_loc, w.s_w2[s], d);
          a.bias = add_bias ? w.b2 : nullptr;
          a.ldo = d;
          if (e->tp_fused) {
            a.epilogue = EPI_BF16;
            tp_route_rows(e, a, 0, M);
          } else if (bf16_part) {
            a.out = e->part_bf16 + s * e->part_shard_stride;
            a.epilogue = EPI_BF16;
          } else {
            a.out = e->part;
            a.resid = first ? nullptr : e->part;
            a.epilogue = first ? EPI_F32 : EPI_F32_RESID;
          }
          OASR_TRY(gemm_bf16_tcgen05(a, st));
        }
        e->launches += 2;
      }
      // residual add + the next LayerNorm (the next layer's attention norm, or the final norm) in one pass
      const bool last = l + 1 == c.n_layers;
      const float* g = last ? e->final_ln_g : e->layers[l + 1].attn_ln_g;
      const float* bta = last ? e->final_ln_b : e->layers[l + 1].attn_ln_b;
      if (e->tp_fused) {
        prof_mark(e, OASR_PROF_ALLREDUCE, st);
        const bool want_hidden = last && hidden_out != nullptr;   // parity runs: the fp32 LayerNorm output of all rows
        OASR_TRY(tp_push_reduce_layernorm(e->tp_view, 0, 0, M, d, g, bta, ++e->tp_epoch, want_hidden, st));
        if (want_hidden) {
          prof_mark(e, OASR_PROF_LAYERNORM, st);
          OASR_TRY(layernorm_rows(e->x, 0, 0, 1, (int)M, d, g, bta, e->lnbuf, hidden_out, st));
        }
      } else if (e->tp_emulated) {
        prof_mark(e, OASR_PROF_ALLREDUCE, st);
        OASR_TRY(emulated_reduce(g, bta));
        if (last && hidden_out != nullptr) {
          prof_mark(e, OASR_PROF_LAYERNORM, st);
          OASR_TRY(layernorm_rows(e->x, 0, 0, 1, (int)M, d, g, bta, e->lnbuf, hidden_out, st));
        }
      } else {
        OASR_TRY(reduce_partial());
        prof_mark(e, OASR_PROF_LAYERNORM, st);
        OASR_TRY(add_layernorm_rows(e->x, 0, 0, 1, (int)M, d, g, bta, e->lnbuf, last ? hidden_out : nullptr, e->part, e->x, st));
      }
      e->launches += 2;
      if (stop_stage == 4 + l) {
        prof_mark(e, OASR_PROF_END, st);
        return OASR_OK;
      }
    }
    if (c.n_layers == 0) {
      prof_mark(e, OASR_PROF_LAYERNORM, st);
      OASR_TRY(layernorm_rows(e->x, 0, 0, 1, (int)M, d, e->final_ln_g, e->final_ln_b, e->lnbuf, hidden_out, st));
    }
  } else {
  for (int l = 0; l < c.n_layers; ++l) {
    const LayerW& w = e->layers[l];
    prof_mark(e, OASR_PROF_LAYERNORM, st);
    OASR_TRY(layernorm_rows(e->x, 0, 0, 1, (int)M, d, w.attn_ln_g, w.attn_ln_b, e->lnbuf, nullptr, st));
    prof_mark(e, OASR_PROF_QKV, st);
    {
      GemmArgs a = GemmArgs::plain(e->lnbuf, (int)M, d, d, w.wqkv, 3 * d);
      a.bias = w.bqkv;
      a.out = e->qkv;
      a.ldo = 3 * d;
      a.epilogue = EPI_BF16;
      OASR_TRY(gemm_bf16_tcgen05(a, st));
    }
    prof_mark(e, OASR_PROF_ATTENTION, st);
    OASR_TRY(attention_bf16(e->qkv, e->att, e->n_frames_dev, B, T, H, hd, scale, st));
    prof_mark(e, OASR_PROF_OUTPROJ, st);
    {
      GemmArgs a = GemmArgs::plain(e->att, (int)M, d, d, w.wo, d);
      a.bias = w.bo;
      a.out = e->x;
      a.ldo = d;
      a.resid = e->x;
      a.epilogue = EPI_F32_RESID;
      OASR_TRY(gemm_bf16_tcgen05(a, st));
    }
    prof_mark(e, OASR_PROF_LAYERNORM, st);
    OASR_TRY(layernorm_rows(e->x, 0, 0, 1, (int)M, d, w.ffn_ln_g, w.ffn_ln_b, e->lnbuf, nullptr, st));
    prof_mark(e, OASR_PROF_FFN1, st);
    {
      GemmArgs a = GemmArgs::plain(e->lnbuf, (int)M, d, d, w.w1, F);
      a.bias = w.b1;
      a.out = e->ffn;
      a.ldo = F;
      a.epilogue = EPI_BF16_GELU;
      OASR_TRY(gemm_bf16_tcgen05(a, st));
    }
    prof_mark(e, OASR_PROF_FFN2, st);
    {
      GemmArgs a = GemmArgs::plain(e->ffn, (int)M, F, F, w.w2, d);
      a.bias = w.b2;
      a.out = e->x;
      a.ldo = d;
      a.resid = e->x;
      a.epilogue = EPI_F32_RESID;
      OASR_TRY(gemm_bf16_tcgen05(a, st));
    }
    e->launches += 7;
    if (stop_stage == 4 + l) {
      prof_mark(e, OASR_PROF_END, st);
      return OASR_OK;
    }
  }
  prof_mark(e, OASR_PROF_LAYERNORM, st);
  OASR_TRY(layernorm_rows(e->x, 0, 0, 1, (int)M, d, e->final_ln_g, e->final_ln_b, e->lnbuf, hidden_out, st));
  }
  prof_mark(e, OASR_PROF_CTC_HEAD, st);
  // a15: logits stay in TMEM; only a packed (max, index) key per frame reaches HBM
  OASR_CUDA_CHECK(cudaMemsetAsync(e->keys, 0, (size_t)M * 8, st));
  {
    GemmArgs a = GemmArgs::plain(e->lnbuf, (int)M, d, d, e->ctc_w, c.vocab);
    a.bias = e->ctc_bias;
    a.argmax = e->keys;
    a.epilogue = EPI_ARGMAX;
    OASR_TRY(gemm_bf16_tcgen05(a, st));
  }
  // a16
  prof_mark(e, OASR_PROF_DECODE, st);
  OASR_TRY(ctc_decode(e->keys, e->n_frames_dev, B, T, c.blank_id, e->frame_ids, e->out_ids, e->out_frames, e->out_lens, st));
  e->launches += 4;
  prof_mark(e, OASR_PROF_END, st);
  return OASR_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* oasr_version(void) { return "oasr-b200 0.1.0 (sm_100a)"; }
const char* oasr_last_error(void) { return last_error_cstr(); }

int32_t oasr_feature_length(const OasrConfig* cfg, int64_t n_samples) {
  if (!cfg) return 0;
  return fe_len(*cfg, n_samples, cfg->n_fe_layers);
}

int oasr_create(const OasrConfig* cfg, OasrHandle* out) {
  OASR_REQUIRE(cfg && out, "oasr_create: null argument");
  *out = nullptr;
  const OasrConfig& c = *cfg;
  if (c.fe_dim != 512 || c.n_fe_layers < 2 || c.n_fe_layers > 8 || c.fe_kernel[0] != 10 || c.fe_stride[0] != 5)
    return fail(OASR_ERR_UNSUPPORTED, "feature extractor must be 512 channels with a (k=10, s=5) first layer");
  for (int i = 1; i < c.n_fe_layers; ++i)
    if (c.fe_stride[i] != 2 || (c.fe_kernel[i] != 2 && c.fe_kernel[i] != 3))
      return fail(OASR_ERR_UNSUPPORTED, "feature extractor layers >= 1 must be stride 2 with kernel 2 or 3");
  OASR_REQUIRE(c.d_model > 0 && c.n_heads > 0 && c.d_model % c.n_heads == 0, "d_model must be divisible by n_heads");
  const int hd = c.d_model / c.n_heads;
  if (hd % 16 != 0 || hd > 128) return fail(OASR_ERR_UNSUPPORTED, "head_dim must be a multiple of 16 and <= 128");
  if (c.d_model % 64 != 0 || c.d_model > 2048 || c.d_ffn % 64 != 0)
    return fail(OASR_ERR_UNSUPPORTED, "d_model (<= 2048) and d_ffn must be multiples of 64");
  OASR_REQUIRE(c.pos_groups > 0 && c.d_model % c.pos_groups == 0, "d_model must be divisible by pos_groups");
  const int cg = c.d_model / c.pos_groups;
  if (cg % 16 != 0 || cg > 128) return fail(OASR_ERR_UNSUPPORTED, "pos-conv group width must be a multiple of 16 and <= 128");
  OASR_REQUIRE(c.pos_kernel >= 2 && c.pos_kernel % 2 == 0, "pos_kernel must be even");
  OASR_REQUIRE(c.vocab > 0 && c.n_layers >= 0 && c.blank_id >= 0 && c.blank_id < c.vocab, "bad vocab / layers / blank");
  int dev = 0;
  OASR_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  OASR_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(OASR_ERR_UNSUPPORTED, std::string("liboasr is built for sm_100a only; device is ") + prop.name);
  OasrEngine* e = new OasrEngine();
  e->cfg = c;
  e->device = dev;
  *out = e;
  return OASR_OK;
}

void oasr_destroy(OasrHandle h) {
  if (!h) return;
  cudaDeviceSynchronize();
  for (auto& kv : h->raw) cudaFree(kv.second.ptr);
  for (void* p : h->owned) cudaFree(p);
  for (void* p : h->ws_owned) cudaFree(p);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  for (cudaEvent_t ev : h->stage_ev) if (ev) cudaEventDestroy(ev);
  for (auto& m : h->prof_marks) cudaEventDestroy(m.second);
  for (cudaEvent_t ev : h->prof_pool) cudaEventDestroy(ev);
  tp_comm_destroy(h->tp_comm);
  for (size_t q = 0; q < h->tp_peer_base.size(); ++q)
    if ((int)q != h->tp_first && h->tp_peer_base[q]) cudaIpcCloseMemHandle(h->tp_peer_base[q]);
  if (h->tp_arena) cudaFree(h->tp_arena);
  if (h->tp_err_host) cudaFreeHost(h->tp_err_host);
  for (auto& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  for (int i = 0; i < OasrEngine::ASYNC_SLOTS; ++i) {
    if (h->h2d_done[i]) cudaEventDestroy(h->h2d_done[i]);
    if (h->slot_done[i]) cudaEventDestroy(h->slot_done[i]);
  }
  if (h->graph_stream) cudaStreamDestroy(h->graph_stream);
  if (h->graph_fork) cudaEventDestroy(h->graph_fork);
  if (h->graph_join) cudaEventDestroy(h->graph_join);
  if (h->tp_comm_stream) cudaStreamDestroy(h->tp_comm_stream);
  if (h->tp_comm_stream2) cudaStreamDestroy(h->tp_comm_stream2);
  for (int i = 0; i < 2; ++i) {
    if (h->tp_ev_compute[i]) cudaEventDestroy(h->tp_ev_compute[i]);
    if (h->tp_ev_reduce[i]) cudaEventDestroy(h->tp_ev_reduce[i]);
  }
  delete h;
}

// Small batches are launch-bound (one 30 s window: ~350 launches for ~8 ms of device time): the second time a
// (buffer, B, L, flags) combination is seen its forward is captured into a CUDA graph and replayed from then on.
// Capture and replay run on the engine's own stream (the caller's may be the legacy default stream, which cannot be
// captured), ordered after / before the caller's stream by events.  OASR_GRAPH_MAX_B: largest batch that takes this
// path (default 8; 0 disables).
static int forward_impl(OasrEngine* e, const float* wave_in, int64_t wave_stride, const int32_t* n_samples_host, int B,
                        int L, int flags, int stop_stage, float* hidden_out, cudaStream_t st) {
  static const int graph_max_b = [] {
    const char* v = std::getenv("OASR_GRAPH_MAX_B");
    return v != nullptr ? std::atoi(v) : 8;
  }();
  const bool eligible = B > 0 && B <= graph_max_b && L > 0 && stop_stage == 0 && hidden_out == nullptr && !e->profiling &&
                        e->tp_world == 1 && e->finalized && wave_in != nullptr && n_samples_host != nullptr;
  if (!eligible) return forward_eager(e, wave_in, wave_stride, n_samples_host, B, L, flags, stop_stage, hidden_out, false, st);
  OASR_TRY(ensure_workspace(e, B, L));   // before looking anything up: a reallocation drops the captured graphs
  int gi = -1;
  for (size_t i = 0; i < e->graphs.size(); ++i) {
    const auto& g = e->graphs[i];
    if (g.wave == wave_in && g.stride == wave_stride && g.B == B && g.L == L && g.flags == flags) gi = (int)i;
  }
  if (gi < 0) {   // first sight: run eagerly (sets kernel attributes, builds tensor maps)
    if (e->graphs.size() >= 16) {
      for (auto& g : e->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
      e->graphs.clear();
    }
    e->graphs.push_back({wave_in, (long long)wave_stride, B, L, flags, 1, nullptr, 0, 0, 0, 0});
    return forward_eager(e, wave_in, wave_stride, n_samples_host, B, L, flags, stop_stage, hidden_out, false, st);
  }
  if (e->graphs[gi].state < 0)
    return forward_eager(e, wave_in, wave_stride, n_samples_host, B, L, flags, stop_stage, hidden_out, false, st);
  if (e->graph_stream == nullptr) {
    OASR_CUDA_CHECK(cudaStreamCreateWithFlags(&e->graph_stream, cudaStreamNonBlocking));
    OASR_CUDA_CHECK(cudaEventCreateWithFlags(&e->graph_fork, cudaEventDisableTiming));
    OASR_CUDA_CHECK(cudaEventCreateWithFlags(&e->graph_join, cudaEventDisableTiming));
  }
  cudaStream_t gs = e->graph_stream;
  OASR_CUDA_CHECK(cudaEventRecord(e->graph_fork, st));
  OASR_CUDA_CHECK(cudaStreamWaitEvent(gs, e->graph_fork, 0));
  OASR_TRY(stage_lengths(e, n_samples_host, B, L, gs));
  if (e->graphs[gi].state == 1) {
    const long long launches0 = e->launches;
    cudaGraph_t graph = nullptr;
    OASR_CUDA_CHECK(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
    const int rc = forward_eager(e, wave_in, wave_stride, n_samples_host, B, L, flags, stop_stage, hidden_out, true, gs);
    const cudaError_t ce = cudaStreamEndCapture(gs, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc != OASR_OK || ce != cudaSuccess || graph == nullptr ||
        cudaGraphInstantiate(&exec, graph, nullptr, nullptr, 0) != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();   // clear the capture error; nothing has run yet: fall back to eager launches
      e->launches = launches0;
      e->graphs[gi].state = -1;
      OASR_TRY(forward_eager(e, wave_in, wave_stride, n_samples_host, B, L, flags, stop_stage, hidden_out, true, gs));
      OASR_CUDA_CHECK(cudaEventRecord(e->graph_join, gs));
      OASR_CUDA_CHECK(cudaStreamWaitEvent(st, e->graph_join, 0));
      return OASR_OK;
    }
    cudaGraphDestroy(graph);
    auto& g = e->graphs[gi];
    g.exec = exec;
    g.state = 2;
    g.launches = e->launches - launches0;
    g.last_T = e->last_T;
    g.last_fe_idx = e->last_fe_idx;
    g.last_fe_pad = e->last_fe_pad;
  } else {
    const auto& g = e->graphs[gi];
    e->launches += g.launches;
    e->last_B = B;
    e->last_L = L;
    e->last_T = g.last_T;
    e->last_fe_idx = g.last_fe_idx;
    e->last_fe_pad = g.last_fe_pad;
  }
  OASR_CUDA_CHECK(cudaGraphLaunch(e->graphs[gi].exec, gs));
  OASR_CUDA_CHECK(cudaEventRecord(e->graph_join, gs));
  OASR_CUDA_CHECK(cudaStreamWaitEvent(st, e->graph_join, 0));
  return OASR_OK;
}

static int tp_check_split(OasrHandle h, int world) {
  OASR_REQUIRE(h, "tensor parallelism: null handle");
  if (h->finalized) return fail(OASR_ERR_STATE, "tensor parallelism must be set up before oasr_finalize_weights");
  const OasrConfig& c = h->cfg;
  OASR_REQUIRE(world >= 1 && world <= 64, "tensor parallelism: world size out of range");
  if (c.n_heads % world != 0 || (c.d_model / world) % 64 != 0 || (c.d_ffn / world) % 64 != 0 || c.d_ffn % world != 0)
    return fail(OASR_ERR_UNSUPPORTED, "tensor parallelism: heads must divide by the world size and d_model / world, "
                                      "d_ffn / world must be multiples of 64");
  return OASR_OK;
}

int oasr_tp_unique_id(void* id_out) {
  OASR_REQUIRE(id_out, "oasr_tp_unique_id: null argument");
  return tp_unique_id(id_out);
}

int oasr_tp_init(OasrHandle h, int32_t rank, int32_t world, const void* id) {
  OASR_TRY(tp_check_split(h, world));
  OASR_REQUIRE(rank >= 0 && rank < world && id, "oasr_tp_init: bad rank / id");
  if (h->tp_comm != nullptr || h->tp_world != 1) return fail(OASR_ERR_STATE, "tensor parallelism already initialised");
  if (world > 1) OASR_TRY(tp_comm_create(&h->tp_comm, rank, world, id));
  h->tp_world = world;
  h->tp_first = rank;
  h->tp_local = 1;
  return OASR_OK;
}

int oasr_tp_emulate(OasrHandle h, int32_t world) {
  OASR_TRY(tp_check_split(h, world));
  if (h->tp_comm != nullptr || h->tp_world != 1) return fail(OASR_ERR_STATE, "tensor parallelism already initialised");
  if (world > TP_MAX_WORLD) return fail(OASR_ERR_UNSUPPORTED, "at most 8 tensor-parallel shards");
  h->tp_world = world;
  h->tp_first = 0;
  h->tp_local = world;
  h->tp_emulated = true;
  return OASR_OK;
}

int oasr_tp_ipc_export(OasrHandle h, int32_t B, int32_t L, void* handle_out) {
  OASR_REQUIRE(h && handle_out && B > 0 && L > 0, "oasr_tp_ipc_export: bad arguments");
  if (h->tp_world < 2 || h->tp_emulated || h->tp_comm == nullptr)
    return fail(OASR_ERR_STATE, "oasr_tp_ipc_export needs an initialised tensor-parallel group (oasr_tp_init)");
  if (h->tp_world > TP_MAX_WORLD) return fail(OASR_ERR_UNSUPPORTED, "peer-memory path supports at most 8 ranks");
  if (h->tp_arena != nullptr || h->ws_B != 0) return fail(OASR_ERR_STATE, "peer arena must be exported once, before the first forward");
  const OasrConfig& c = h->cfg;
  const long long T = std::max(1, fe_len(c, L, c.n_fe_layers));
  const size_t M = (size_t)B * T, d = (size_t)c.d_model;
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  h->tp_x_bytes = M * d * 4;
  h->tp_off_ln = up(h->tp_x_bytes);
  h->tp_off_part = h->tp_off_ln + up(M * std::max<size_t>(512, d) * 2);
  // receive region: [source][slot rows][d] bf16 per half-batch; the slots of a half-batch cover its rows plus at most
  // `world` rows of rounding each
  h->tp_off_flags = h->tp_off_part + up((M + 4 * TP_MAX_WORLD * TP_MAX_WORLD) * d * 2);
  const size_t total = h->tp_off_flags + 1024;   // two flag sets, 512 B apart (one per half-batch in overlap mode)
  OASR_TRY(dev_alloc(&h->tp_arena, total, true));
  OASR_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t mh;
  OASR_CUDA_CHECK(cudaIpcGetMemHandle(&mh, h->tp_arena));
  static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &mh, sizeof(mh));
  return OASR_OK;
}

int oasr_tp_ipc_import(OasrHandle h, const void* handles) {
  OASR_REQUIRE(h && handles, "oasr_tp_ipc_import: bad arguments");
  if (h->tp_arena == nullptr) return fail(OASR_ERR_STATE, "call oasr_tp_ipc_export first");
  if (h->tp_fused) return OASR_OK;
  const int W = h->tp_world, r = h->tp_first;
  h->tp_peer_base.assign(W, nullptr);
  for (int q = 0; q < W; ++q) {
    if (q == r) {
      h->tp_peer_base[q] = h->tp_arena;
      continue;
    }
    cudaIpcMemHandle_t mh;
    memcpy(&mh, reinterpret_cast<const uint8_t*>(handles) + (size_t)q * 64, 64);
    OASR_CUDA_CHECK(cudaIpcOpenMemHandle(&h->tp_peer_base[q], mh, cudaIpcMemLazyEnablePeerAccess));
  }
  TpPeerView& v = h->tp_view;
  v.rank = r;
  v.world = W;
  for (int q = 0; q < W; ++q) {
    uint8_t* base = reinterpret_cast<uint8_t*>(h->tp_peer_base[q]);
    v.x[q] = reinterpret_cast<float*>(base);
    v.ln[q] = reinterpret_cast<__nv_bfloat16*>(base + h->tp_off_ln);
    v.recv[q] = reinterpret_cast<__nv_bfloat16*>(base + h->tp_off_part);
    v.ready[q] = reinterpret_cast<unsigned long long*>(base + h->tp_off_flags);
    v.done[q] = v.ready[q] + TP_MAX_WORLD;
  }
  if (h->tp_err_host == nullptr) {
    OASR_CUDA_CHECK(cudaHostAlloc((void**)&h->tp_err_host, 64, cudaHostAllocMapped));
    *h->tp_err_host = 0;
    OASR_CUDA_CHECK(cudaHostGetDevicePointer((void**)&h->tp_err_dev, h->tp_err_host, 0));
  }
  v.error = h->tp_err_dev;
  v.timeout_ns = tp_timeout_ns();
  h->tp_view2 = v;   // the second flag set: same buffers, flags 512 B further on
  for (int q = 0; q < W; ++q) {
    h->tp_view2.ready[q] = v.ready[q] + 64;
    h->tp_view2.done[q] = v.done[q] + 64;
  }
  h->tp_fused = true;
  return OASR_OK;
}

int oasr_load_weight(OasrHandle h, const char* name, const void* data, int dtype, const int64_t* shape, int ndim) {
  OASR_REQUIRE(h && name && data && shape && ndim >= 1 && ndim <= 4, "oasr_load_weight: bad arguments");
  if (h->finalized) return fail(OASR_ERR_STATE, "weights already finalised");
  OASR_REQUIRE(dtype == OASR_DTYPE_F32 || dtype == OASR_DTYPE_BF16, "oasr_load_weight: dtype must be f32 or bf16");
  long long n = 1;
  for (int i = 0; i < ndim; ++i) {
    OASR_REQUIRE(shape[i] > 0, "oasr_load_weight: non-positive extent");
    n *= shape[i];
  }
  free_raw(h, name);
  DevBuf b;
  b.bytes = (size_t)n * 4;
  b.shape.assign(shape, shape + ndim);
  OASR_TRY(dev_alloc(&b.ptr, b.bytes, false));
  if (dtype == OASR_DTYPE_F32) {
    OASR_CUDA_CHECK(cudaMemcpy(b.ptr, data, b.bytes, cudaMemcpyDefault));
  } else {
    // widen bf16 -> fp32 master through a temporary device copy
    void* tmp = nullptr;
    OASR_TRY(dev_alloc(&tmp, (size_t)n * 2, false));
    cudaError_t ce = cudaMemcpy(tmp, data, (size_t)n * 2, cudaMemcpyDefault);
    if (ce == cudaSuccess) {
      std::vector<uint16_t> hb((size_t)n);
      std::vector<float> hf((size_t)n);
      ce = cudaMemcpy(hb.data(), tmp, (size_t)n * 2, cudaMemcpyDeviceToHost);
      for (long long i = 0; i < n && ce == cudaSuccess; ++i) {
        uint32_t u = (uint32_t)hb[(size_t)i] << 16;
        memcpy(&hf[(size_t)i], &u, 4);
      }
      if (ce == cudaSuccess) ce = cudaMemcpy(b.ptr, hf.data(), b.bytes, cudaMemcpyHostToDevice);
    }
    cudaFree(tmp);
    if (ce != cudaSuccess) {
      cudaFree(b.ptr);
      return fail(OASR_ERR_CUDA, std::string("oasr_load_weight copy: ") + cudaGetErrorString(ce));
    }
  }
  h->raw[name] = b;
  return OASR_OK;
}

int oasr_finalize_weights(OasrHandle h) {
  OASR_REQUIRE(h, "oasr_finalize_weights: null handle");
  if (h->finalized) return OASR_OK;
  OasrEngine* e = h;
  const OasrConfig& c = e->cfg;
  const int d = c.d_model, F = c.d_ffn;
  // --- feature extractor
  e->fe_bias.assign(c.n_fe_layers, nullptr);
  e->fe_g.assign(c.n_fe_layers, nullptr);
  e->fe_b.assign(c.n_fe_layers, nullptr);
  e->fe_w.assign(c.n_fe_layers, nullptr);
  for (int i = 0; i < c.n_fe_layers; ++i) {
    const std::string p = "fe." + std::to_string(i) + ".";
    float* w = nullptr;
    const int cin = i == 0 ? 1 : 512;
    OASR_TRY(get_raw(e, p + "conv.weight", {512, cin, c.fe_kernel[i]}, &w));
    OASR_TRY(get_raw(e, p + "conv.bias", {512}, &e->fe_bias[i]));
    OASR_TRY(get_raw(e, p + "ln.weight", {512}, &e->fe_g[i]));
    OASR_TRY(get_raw(e, p + "ln.bias", {512}, &e->fe_b[i]));
    if (i == 0) {
      OASR_TRY(alloc_owned(e, (void**)&e->fe0_w, 10 * 512 * 4));
      transpose_l0_kernel<<<blocks_for(5120), 256>>>(w, e->fe0_w, 512, 10);
    } else {
      const long long n = 512ll * c.fe_kernel[i] * 512;
      OASR_TRY(alloc_owned(e, (void**)&e->fe_w[i], (size_t)n * 2));
      repack_conv_kernel<<<blocks_for(n), 256>>>(w, e->fe_w[i], 512, 512, c.fe_kernel[i], 512);
    }
    OASR_CUDA_CHECK(cudaGetLastError());
  }
  // --- projection
  {
    float* w = nullptr;
    OASR_TRY(get_raw(e, "proj.ln.weight", {512}, &e->proj_ln_g));
    OASR_TRY(get_raw(e, "proj.ln.bias", {512}, &e->proj_ln_b));
    OASR_TRY(get_raw(e, "proj.linear.weight", {d, 512}, &w));
    OASR_TRY(get_raw(e, "proj.linear.bias", {d}, &e->proj_bias));
    OASR_TRY(to_bf16(e, w, (long long)d * 512, &e->proj_w));
  }
  // --- positional conv: fold weight-norm, repack per group tap-major, pad each tap to k_pad columns
  {
    const int G = c.pos_groups, cg = d / G, K = c.pos_kernel;
    float *g = nullptr, *v = nullptr;
    OASR_TRY(get_raw(e, "pos.weight_g", {1, 1, K}, &g));
    OASR_TRY(get_raw(e, "pos.weight_v", {d, cg, K}, &v));
    OASR_TRY(get_raw(e, "pos.bias", {d}, &e->pos_bias));
    float *norms = nullptr, *folded = nullptr;
    const long long n = (long long)d * cg * K;
    OASR_TRY(dev_alloc((void**)&norms, (size_t)K * 4, false));
    OASR_TRY(dev_alloc((void**)&folded, (size_t)n * 4, false));
    posnorm_kernel<<<K, 256>>>(v, norms, (long long)d * cg, K);
    posfold_kernel<<<blocks_for(n), 256>>>(v, g, norms, folded, n, K);
    e->pos_kpad = ((cg + 63) / 64) * 64;
    const long long nw = (long long)d * K * e->pos_kpad;
    OASR_TRY(alloc_owned(e, (void**)&e->pos_w, (size_t)nw * 2));
    // [d][cg][K] viewed as N=d rows: row n = g*cg + n_in_group, so one launch covers all groups
    repack_conv_kernel<<<blocks_for(nw), 256>>>(folded, e->pos_w, d, cg, K, e->pos_kpad);
    OASR_CUDA_CHECK(cudaGetLastError());
    OASR_CUDA_CHECK(cudaDeviceSynchronize());
    cudaFree(norms);
    cudaFree(folded);
  }
  // --- encoder layers
  e->layers.resize(c.n_layers);
  for (int l = 0; l < c.n_layers; ++l) {
    const std::string p = "enc." + std::to_string(l) + ".";
    LayerW& w = e->layers[l];
    float *wq, *wk, *wv, *wo, *w1, *w2, *bq, *bk, *bv;
    OASR_TRY(get_raw(e, p + "attn_ln.weight", {d}, &w.attn_ln_g));
    OASR_TRY(get_raw(e, p + "attn_ln.bias", {d}, &w.attn_ln_b));
    OASR_TRY(get_raw(e, p + "ffn_ln.weight", {d}, &w.ffn_ln_g));
    OASR_TRY(get_raw(e, p + "ffn_ln.bias", {d}, &w.ffn_ln_b));
    OASR_TRY(get_raw(e, p + "q.weight", {d, d}, &wq));
    OASR_TRY(get_raw(e, p + "k.weight", {d, d}, &wk));
    OASR_TRY(get_raw(e, p + "v.weight", {d, d}, &wv));
    OASR_TRY(get_raw(e, p + "o.weight", {d, d}, &wo));
    OASR_TRY(get_raw(e, p + "q.bias", {d}, &bq));
    OASR_TRY(get_raw(e, p + "k.bias", {d}, &bk));
    OASR_TRY(get_raw(e, p + "v.bias", {d}, &bv));
    OASR_TRY(get_raw(e, p + "o.bias", {d}, &w.bo));
    OASR_TRY(get_raw(e, p + "ffn1.weight", {F, d}, &w1));
    OASR_TRY(get_raw(e, p + "ffn1.bias", {F}, &w.b1));
    OASR_TRY(get_raw(e, p + "ffn2.weight", {d, F}, &w2));
    OASR_TRY(get_raw(e, p + "ffn2.bias", {d}, &w.b2));
    const long long dd = (long long)d * d;
    if (e->tp_world == 1) {
      OASR_TRY(alloc_owned(e, (void**)&w.wqkv, (size_t)dd * 3 * 2));
      cast_bf16_kernel<<<blocks_for(dd), 256>>>(wq, w.wqkv, dd);
      cast_bf16_kernel<<<blocks_for(dd), 256>>>(wk, w.wqkv + dd, dd);
      cast_bf16_kernel<<<blocks_for(dd), 256>>>(wv, w.wqkv + 2 * dd, dd);
      OASR_TRY(alloc_owned(e, (void**)&w.bqkv, (size_t)d * 3 * 4));
      OASR_CUDA_CHECK(cudaMemcpy(w.bqkv, bq, (size_t)d * 4, cudaMemcpyDeviceToDevice));
      OASR_CUDA_CHECK(cudaMemcpy(w.bqkv + d, bk, (size_t)d * 4, cudaMemcpyDeviceToDevice));
      OASR_CUDA_CHECK(cudaMemcpy(w.bqkv + 2 * d, bv, (size_t)d * 4, cudaMemcpyDeviceToDevice));
    }
    if (e->tp_world == 1) {
      OASR_TRY(to_bf16(e, wo, dd, &w.wo));
      OASR_TRY(to_bf16(e, w1, (long long)F * d, &w.w1));
      OASR_TRY(to_bf16(e, w2, (long long)F * d, &w.w2));
    } else {
      const int W = e->tp_world, d_loc = d / W, F_loc = F / W;
      for (int s = 0; s < e->tp_local; ++s) {
        const int r = e->tp_first + s;
        __nv_bfloat16 *sq = nullptr, *so = nullptr, *s1 = nullptr, *s2 = nullptr;
        float *sbq = nullptr, *sb1 = nullptr;
        // q | k | v rows of this shard's heads
        const long long blk = (long long)d_loc * d;
        OASR_TRY(alloc_owned(e, (void**)&sq, (size_t)blk * 3 * 2));
        cast_bf16_kernel<<<blocks_for(blk), 256>>>(wq + (long long)r * blk, sq, blk);
        cast_bf16_kernel<<<blocks_for(blk), 256>>>(wk + (long long)r * blk, sq + blk, blk);
        cast_bf16_kernel<<<blocks_for(blk), 256>>>(wv + (long long)r * blk, sq + 2 * blk, blk);
        OASR_TRY(alloc_owned(e, (void**)&sbq, (size_t)d_loc * 3 * 4));
        OASR_CUDA_CHECK(cudaMemcpy(sbq, bq + r * d_loc, (size_t)d_loc * 4, cudaMemcpyDeviceToDevice));
        OASR_CUDA_CHECK(cudaMemcpy(sbq + d_loc, bk + r * d_loc, (size_t)d_loc * 4, cudaMemcpyDeviceToDevice));
        OASR_CUDA_CHECK(cudaMemcpy(sbq + 2 * d_loc, bv + r * d_loc, (size_t)d_loc * 4, cudaMemcpyDeviceToDevice));
        // out-proj: input columns of this shard's heads
        OASR_TRY(alloc_owned(e, (void**)&so, (size_t)d * d_loc * 2));
        slice_cols_bf16_kernel<<<blocks_for((long long)d * d_loc), 256>>>(wo, so, d, d, r * d_loc, d_loc);
        // FFN1: rows, FFN2: input columns
        OASR_TRY(alloc_owned(e, (void**)&s1, (size_t)F_loc * d * 2));
        cast_bf16_kernel<<<blocks_for((long long)F_loc * d), 256>>>(w1 + (long long)r * F_loc * d, s1, (long long)F_loc * d);
        OASR_TRY(alloc_owned(e, (void**)&sb1, (size_t)F_loc * 4));
        OASR_CUDA_CHECK(cudaMemcpy(sb1, w.b1 + r * F_loc, (size_t)F_loc * 4, cudaMemcpyDeviceToDevice));
        OASR_TRY(alloc_owned(e, (void**)&s2, (size_t)d * F_loc * 2));
        slice_cols_bf16_kernel<<<blocks_for((long long)d * F_loc), 256>>>(w2, s2, d, F, r * F_loc, F_loc);
        OASR_CUDA_CHECK(cudaGetLastError());
        w.s_wqkv.push_back(sq); w.s_wo.push_back(so); w.s_w1.push_back(s1); w.s_w2.push_back(s2);
        w.s_bqkv.push_back(sbq); w.s_b1.push_back(sb1);
      }
    }
    OASR_CUDA_CHECK(cudaDeviceSynchronize());
    for (const char* n : {"q.weight", "k.weight", "v.weight", "o.weight", "ffn1.weight", "ffn2.weight"}) free_raw(e, p + n);
  }
  // --- head
  {
    float* w = nullptr;
    OASR_TRY(get_raw(e, "final_ln.weight", {d}, &e->final_ln_g));
    OASR_TRY(get_raw(e, "final_ln.bias", {d}, &e->final_ln_b));
    OASR_TRY(get_raw(e, "ctc.weight", {c.vocab, d}, &w));
    OASR_TRY(get_raw(e, "ctc.bias", {c.vocab}, &e->ctc_bias));
    OASR_TRY(to_bf16(e, w, (long long)c.vocab * d, &e->ctc_w));
  }
  OASR_CUDA_CHECK(cudaDeviceSynchronize());
  for (const char* n : {"proj.linear.weight", "pos.weight_v", "ctc.weight"}) free_raw(e, n);
  for (int i = 1; i < c.n_fe_layers; ++i) free_raw(e, "fe." + std::to_string(i) + ".conv.weight");
  e->finalized = true;
  return OASR_OK;
}

int oasr_forward_ctc(OasrHandle h, const float* wave_dev, int64_t wave_stride, const int32_t* n_samples_host,
                     int32_t B, int32_t L, int32_t flags, int32_t* frame_ids_dev, float* hidden_dev,
                     int32_t* out_ids_dev, int32_t* out_frames_dev, int32_t* out_lens_dev, OasrStream stream) {
  OASR_REQUIRE(h, "oasr_forward_ctc: null handle");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  OASR_TRY(forward_impl(h, wave_dev, wave_stride > 0 ? wave_stride : L, n_samples_host, B, L, flags, 0, hidden_dev, st));
  const size_t n = (size_t)B * h->last_T * 4;
  if (h->last_T > 0) {
    if (frame_ids_dev) OASR_CUDA_CHECK(cudaMemcpyAsync(frame_ids_dev, h->frame_ids, n, cudaMemcpyDeviceToDevice, st));
    if (out_ids_dev) OASR_CUDA_CHECK(cudaMemcpyAsync(out_ids_dev, h->out_ids, n, cudaMemcpyDeviceToDevice, st));
    if (out_frames_dev) OASR_CUDA_CHECK(cudaMemcpyAsync(out_frames_dev, h->out_frames, n, cudaMemcpyDeviceToDevice, st));
    if (out_lens_dev) OASR_CUDA_CHECK(cudaMemcpyAsync(out_lens_dev, h->out_lens, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  } else if (out_lens_dev) {
    OASR_CUDA_CHECK(cudaMemsetAsync(out_lens_dev, 0, (size_t)B * 4, st));
  }
  return OASR_OK;
}

int oasr_transcribe_host_async(OasrHandle h, const float* wave_host, int64_t wave_stride, const int32_t* n_samples_host,
                               int32_t B, int32_t L, int32_t flags, int32_t* out_ids_host, int32_t* out_frames_host,
                               int32_t* out_lens_host, int32_t* frame_ids_host, OasrStream stream, int64_t* ticket_out) {
  OASR_REQUIRE(h && wave_host && out_lens_host && ticket_out, "oasr_transcribe_host_async: null argument");
  OASR_REQUIRE(B > 0 && L > 0, "oasr_transcribe_host_async: empty batch");
  OasrEngine* e = h;
  const int slot = (int)(e->tickets % OasrEngine::ASYNC_SLOTS);
  if (e->slot_busy[slot])
    return fail(OASR_ERR_STATE, "oasr_transcribe_host_async: both slots in flight - oasr_wait the oldest ticket first");
  if (e->h2d_stream == nullptr) {
    OASR_CUDA_CHECK(cudaStreamCreateWithFlags(&e->h2d_stream, cudaStreamNonBlocking));
    OASR_CUDA_CHECK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    for (int i = 0; i < OasrEngine::ASYNC_SLOTS; ++i) {
      OASR_CUDA_CHECK(cudaEventCreateWithFlags(&e->h2d_done[i], cudaEventDisableTiming));
      OASR_CUDA_CHECK(cudaEventCreateWithFlags(&e->slot_done[i], cudaEventDisableTiming));
    }
  }
  // NULL = the engine's own non-blocking stream (the legacy default stream would serialise against h2d_stream's peers)
  cudaStream_t st = stream != nullptr ? reinterpret_cast<cudaStream_t>(stream) : e->own_stream;
  OASR_TRY(ensure_workspace(e, B, L));   // may synchronise the device and replace the landing buffers: nothing is in flight in this slot
  const int64_t ws = wave_stride > 0 ? wave_stride : L;
  float* dst = e->wave_raw[slot];
  const size_t esz = (flags & OASR_FLAG_INPUT_I16) ? 2 : 4;   // PCM16 windows land as they are: half the H2D bytes
  OASR_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)L * esz, wave_host, (size_t)ws * esz, (size_t)L * esz, B,
                                    cudaMemcpyHostToDevice, e->h2d_stream));
  OASR_CUDA_CHECK(cudaEventRecord(e->h2d_done[slot], e->h2d_stream));
  OASR_CUDA_CHECK(cudaStreamWaitEvent(st, e->h2d_done[slot], 0));
  OASR_TRY(forward_impl(e, dst, L, n_samples_host, B, L, flags, 0, nullptr, st));
  const size_t n = (size_t)B * e->last_T * 4;
  e->slot_lens_host[slot] = nullptr;
  if (e->last_T > 0) {
    if (out_ids_host) OASR_CUDA_CHECK(cudaMemcpyAsync(out_ids_host, e->out_ids, n, cudaMemcpyDeviceToHost, st));
    if (out_frames_host) OASR_CUDA_CHECK(cudaMemcpyAsync(out_frames_host, e->out_frames, n, cudaMemcpyDeviceToHost, st));
    if (frame_ids_host) OASR_CUDA_CHECK(cudaMemcpyAsync(frame_ids_host, e->frame_ids, n, cudaMemcpyDeviceToHost, st));
    OASR_CUDA_CHECK(cudaMemcpyAsync(out_lens_host, e->out_lens, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  } else {
    e->slot_lens_host[slot] = out_lens_host;   // windows too short for a single frame: no device work, lengths are zero
  }
  OASR_CUDA_CHECK(cudaEventRecord(e->slot_done[slot], st));
  e->slot_busy[slot] = true;
  e->slot_B[slot] = B;
  *ticket_out = e->tickets++;
  return OASR_OK;
}

int oasr_wait(OasrHandle h, int64_t ticket) {
  OASR_REQUIRE(h, "oasr_wait: null handle");
  OasrEngine* e = h;
  OASR_REQUIRE(ticket >= 0 && ticket < e->tickets && ticket + OasrEngine::ASYNC_SLOTS >= e->tickets,
               "oasr_wait: unknown or already recycled ticket");
  const int slot = (int)(ticket % OasrEngine::ASYNC_SLOTS);
  if (!e->slot_busy[slot]) return OASR_OK;   // waited already
  const cudaError_t ce = cudaEventSynchronize(e->slot_done[slot]);
  e->slot_busy[slot] = false;                // only now may a submit reuse the slot's landing buffer
  if (ce != cudaSuccess) return fail(OASR_ERR_CUDA, std::string("oasr_wait: ") + cudaGetErrorString(ce));
  if (e->slot_lens_host[slot] != nullptr) memset(e->slot_lens_host[slot], 0, (size_t)e->slot_B[slot] * 4);
  return tp_check_error(e);
}

int oasr_transcribe_host(OasrHandle h, const float* wave_host, int64_t wave_stride, const int32_t* n_samples_host,
                         int32_t B, int32_t L, int32_t flags, int32_t* out_ids_host, int32_t* out_frames_host,
                         int32_t* out_lens_host, int32_t* frame_ids_host, OasrStream stream) {
  OASR_REQUIRE(h && wave_host && out_lens_host, "oasr_transcribe_host: null argument");
  OASR_REQUIRE(B > 0 && L > 0, "oasr_transcribe_host: empty batch");
  // the synchronous form keeps its stream contract: NULL is the (legacy) default stream, as in round 1
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < OasrEngine::ASYNC_SLOTS; ++i)   // drain: the synchronous call may not overtake pending tickets
    if (h->slot_busy[(h->tickets + i) % OasrEngine::ASYNC_SLOTS]) {
      const int slot = (int)((h->tickets + i) % OasrEngine::ASYNC_SLOTS);
      OASR_CUDA_CHECK(cudaEventSynchronize(h->slot_done[slot]));
    }
  OASR_TRY(ensure_workspace(h, B, L));
  const int64_t ws = wave_stride > 0 ? wave_stride : L;
  float* dst = h->wave_raw[0];
  const size_t esz = (flags & OASR_FLAG_INPUT_I16) ? 2 : 4;
  cudaError_t ce = cudaMemcpy2DAsync(dst, (size_t)L * esz, wave_host, (size_t)ws * esz, (size_t)L * esz, B,
                                     cudaMemcpyHostToDevice, st);
  int rc = OASR_OK;
  if (ce != cudaSuccess) rc = fail(OASR_ERR_CUDA, std::string("H2D waveform: ") + cudaGetErrorString(ce));
  if (rc == OASR_OK) rc = forward_impl(h, dst, L, n_samples_host, B, L, flags, 0, nullptr, st);
  const size_t n = (size_t)B * h->last_T * 4;
  if (rc == OASR_OK && h->last_T > 0) {
    if (out_ids_host) ce = cudaMemcpyAsync(out_ids_host, h->out_ids, n, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess && out_frames_host) ce = cudaMemcpyAsync(out_frames_host, h->out_frames, n, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess && frame_ids_host) ce = cudaMemcpyAsync(frame_ids_host, h->frame_ids, n, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(out_lens_host, h->out_lens, (size_t)B * 4, cudaMemcpyDeviceToHost, st);
    if (ce != cudaSuccess) rc = fail(OASR_ERR_CUDA, std::string("D2H ids: ") + cudaGetErrorString(ce));
  } else if (rc == OASR_OK) {
    memset(out_lens_host, 0, (size_t)B * 4);
  }
  ce = cudaStreamSynchronize(st);
  if (rc == OASR_OK && ce != cudaSuccess) rc = fail(OASR_ERR_CUDA, std::string("stream sync: ") + cudaGetErrorString(ce));
  if (rc == OASR_OK) rc = tp_check_error(h);
  return rc;
}

int oasr_debug_forward(OasrHandle h, const float* wave_dev, int64_t wave_stride, const int32_t* n_samples_host,
                       int32_t B, int32_t L, int32_t flags, int32_t stop_stage, OasrStream stream) {
  OASR_REQUIRE(h, "oasr_debug_forward: null handle");
  return forward_impl(h, wave_dev, wave_stride > 0 ? wave_stride : L, n_samples_host, B, L, flags, stop_stage, nullptr,
                      reinterpret_cast<cudaStream_t>(stream));
}

int oasr_debug_buffer(OasrHandle h, const char* name, void** dev_ptr, int64_t* shape4, int32_t* dtype) {
  OASR_REQUIRE(h && name && dev_ptr && shape4 && dtype, "oasr_debug_buffer: null argument");
  const std::string n(name);
  for (int i = 0; i < 4; ++i) shape4[i] = 0;
  if (n == "fe") {
    *dev_ptr = h->fe_buf[h->last_fe_idx];
    shape4[0] = h->last_B; shape4[1] = h->last_fe_pad; shape4[2] = 512;
    *dtype = OASR_DTYPE_BF16;
  } else if (n == "x") {
    *dev_ptr = h->x;
    shape4[0] = (int64_t)h->last_B * h->last_T; shape4[1] = h->cfg.d_model;
    *dtype = OASR_DTYPE_F32;
  } else if (n == "wave") {
    *dev_ptr = h->wave;
    shape4[0] = h->last_B; shape4[1] = h->last_L;
    *dtype = OASR_DTYPE_F32;
  } else if (n == "att") {
    *dev_ptr = h->att;
    shape4[0] = (int64_t)h->last_B * h->last_T; shape4[1] = h->cfg.d_model;
    *dtype = OASR_DTYPE_BF16;
  } else if (n == "qkv") {
    *dev_ptr = h->qkv;
    shape4[0] = (int64_t)h->last_B * h->last_T; shape4[1] = 3 * h->cfg.d_model;
    *dtype = OASR_DTYPE_BF16;
  } else {
    return fail(OASR_ERR_INVALID, "unknown debug buffer: " + n);
  }
  return OASR_OK;
}

int oasr_debug_copy(OasrHandle h, const char* name, void* dst_dev, int64_t nbytes) {
  void* src = nullptr;
  int64_t shape[4];
  int32_t dt = 0;
  OASR_TRY(oasr_debug_buffer(h, name, &src, shape, &dt));
  OASR_REQUIRE(dst_dev && nbytes >= 0, "oasr_debug_copy: bad destination");
  int64_t have = dt == OASR_DTYPE_BF16 ? 2 : 4;
  for (int i = 0; i < 4 && shape[i] > 0; ++i) have *= shape[i];
  OASR_REQUIRE(nbytes <= have, "oasr_debug_copy: more bytes requested than the buffer holds");
  OASR_CUDA_CHECK(cudaDeviceSynchronize());
  OASR_CUDA_CHECK(cudaMemcpy(dst_dev, src, (size_t)nbytes, cudaMemcpyDeviceToDevice));
  return OASR_OK;
}

int64_t oasr_launch_count(OasrHandle h) { return h ? h->launches : 0; }

int oasr_profile_enable(OasrHandle h, int32_t on) {
  OASR_REQUIRE(h, "oasr_profile_enable: null handle");
  h->profiling = on != 0;
  return OASR_OK;
}

int oasr_profile_read(OasrHandle h, double* ms, int64_t* counts, int32_t n) {
  OASR_REQUIRE(h && ms && counts && n >= OASR_PROF_NCAT, "oasr_profile_read: need OASR_PROF_NCAT entries");
  OASR_CUDA_CHECK(cudaDeviceSynchronize());
  for (size_t i = 0; i + 1 < h->prof_marks.size(); ++i) {
    const int cat = h->prof_marks[i].first;
    if (cat == OASR_PROF_END) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, h->prof_marks[i].second, h->prof_marks[i + 1].second) == cudaSuccess) {
      h->prof_ms[cat] += t;
      h->prof_n[cat] += 1;
    }
  }
  for (auto& m : h->prof_marks) h->prof_pool.push_back(m.second);
  h->prof_marks.clear();
  for (int i = 0; i < OASR_PROF_NCAT; ++i) {
    ms[i] = h->prof_ms[i];
    counts[i] = h->prof_n[i];
    h->prof_ms[i] = 0;
    h->prof_n[i] = 0;
  }
  return OASR_OK;
}

// ---- per-stage entry points ---------------------------------------------------------------------------
int oasr_wave_norm(const float* in, float* out, const int32_t* n_samples_dev, int32_t B, int32_t L, OasrStream stream) {
  double* partials = nullptr;
  OASR_TRY(dev_alloc((void**)&partials, (size_t)B * WAVE_NORM_SLICES * 16, false));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = wave_norm(in, out, n_samples_dev, B, L, L, L, partials, st);
  cudaStreamSynchronize(st);
  cudaFree(partials);
  return rc;
}

int64_t oasr_resample_length(int64_t n_in, int32_t sr_in, int32_t sr_out) {
  if (n_in < 0 || sr_in <= 0 || sr_out <= 0) return 0;
  return resample_length(n_in, sr_in, sr_out);
}

int oasr_resample(const void* in_dev, int32_t in_is_i16, int64_t n_in, int32_t channels, int32_t sr_in, int32_t sr_out,
                  float* out_dev, int64_t out_capacity, OasrStream stream) {
  return resample_mono(in_dev, in_is_i16, n_in, channels, sr_in, sr_out, out_dev, out_capacity,
                       reinterpret_cast<cudaStream_t>(stream));
}

int oasr_fe_layer0(const float* wave, int32_t B, int32_t L, const float* w_10x512, const float* bias, const float* gamma,
                   const float* beta, void* out_bf16, OasrStream stream) {
  const int T0 = L >= 10 ? (L - 10) / 5 + 1 : 0;
  return fe_layer0(wave, L, B, L, w_10x512, bias, gamma, beta, out_bf16, (long long)T0 * 512, T0,
                   reinterpret_cast<cudaStream_t>(stream));
}

int oasr_conv_ln_gelu(const void* in_bf16, int32_t B, int32_t L_in, int32_t L_in_pad, int32_t k, const void* w_bf16,
                      const float* bias, const float* gamma, const float* beta, void* out_bf16, OasrStream stream) {
  OASR_REQUIRE(k == 2 || k == 3, "conv_ln_gelu: kernel must be 2 or 3");
  OASR_REQUIRE(L_in_pad % 2 == 0 && L_in_pad >= L_in + 2, "conv_ln_gelu: L_in_pad must be even and >= L_in + 2");
  const int T_out = L_in >= k ? (L_in - k) / 2 + 1 : 0;
  return run_conv_layer(in_bf16, B, L_in, L_in_pad, k, w_bf16, bias, gamma, beta, out_bf16, T_out,
                        reinterpret_cast<cudaStream_t>(stream));
}

int oasr_layernorm(const void* in, int32_t in_is_bf16, int64_t rows, int32_t D, const float* gamma, const float* beta,
                   void* out_bf16, float* out_f32, OasrStream stream) {
  return layernorm_rows(in, in_is_bf16, 0, 1, (int)rows, D, gamma, beta, out_bf16, out_f32,
                        reinterpret_cast<cudaStream_t>(stream));
}

int oasr_gemm(const void* A_bf16, const void* W_bf16, const float* bias, int32_t M, int32_t N, int32_t K,
              int32_t epilogue, void* out, int32_t ldo, const float* resid, const float* ln_gamma,
              const float* ln_beta, uint64_t* argmax_keys, OasrStream stream) {
  GemmArgs a = GemmArgs::plain(A_bf16, M, K, K, W_bf16, N);
  a.bias = bias;
  a.out = out;
  a.ldo = ldo;
  a.resid = resid;
  a.ln_gamma = ln_gamma;
  a.ln_beta = ln_beta;
  a.argmax = reinterpret_cast<unsigned long long*>(argmax_keys);
  a.epilogue = epilogue;
  return gemm_bf16_tcgen05(a, reinterpret_cast<cudaStream_t>(stream));
}

int oasr_posconv(float* x, int32_t B, int32_t T, int32_t d, int32_t groups, int32_t k, const void* w_bf16,
                 const float* bias, void* scratch_bf16, OasrStream stream) {
  OASR_REQUIRE(x && w_bf16 && bias && scratch_bf16 && groups > 0 && d % groups == 0, "posconv: bad arguments");
  const int cg = d / groups;
  OASR_REQUIRE(cg % 16 == 0 && cg <= 128 && k % 2 == 0, "posconv: unsupported group width / kernel");
  return run_posconv(x, B, T, d, groups, k, ((cg + 63) / 64) * 64, w_bf16, bias, scratch_bf16,
                     reinterpret_cast<cudaStream_t>(stream));
}

int oasr_attention(const void* qkv_bf16, void* out_bf16, const int32_t* n_frames_dev, int32_t B, int32_t T, int32_t H,
                   int32_t hd, float scale, OasrStream stream) {
  return attention_bf16(qkv_bf16, out_bf16, n_frames_dev, B, T, H, hd, scale, reinterpret_cast<cudaStream_t>(stream));
}

int oasr_ctc_decode(const uint64_t* keys, const int32_t* n_frames_dev, int32_t B, int32_t T, int32_t blank,
                    int32_t* frame_ids, int32_t* out_ids, int32_t* out_frames, int32_t* out_lens, OasrStream stream) {
  return ctc_decode(reinterpret_cast<const unsigned long long*>(keys), n_frames_dev, B, T, blank, frame_ids, out_ids,
                    out_frames, out_lens, reinterpret_cast<cudaStream_t>(stream));
}

int oasr_ctc_collapse(const int32_t* frame_ids, const int32_t* n_frames_dev, int32_t B, int32_t T, int32_t blank,
                      int32_t* out_ids, int32_t* out_frames, int32_t* out_lens, OasrStream stream) {
  return ctc_collapse(frame_ids, n_frames_dev, B, T, blank, out_ids, out_frames, out_lens,
                      reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
