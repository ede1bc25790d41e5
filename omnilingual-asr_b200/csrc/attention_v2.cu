// Bidirectional self-attention on tcgen05, second version (a14): one (sequence, head, 128-query tile) per CTA,
// ONE pass over the keys, everything between the two MMAs stays in tensor memory.
//
//   S_j = Q K_j^T             tcgen05.mma SS, both operands K-major bf16 in 32B-swizzled 16-column chunks;
//                             S is double buffered in TMEM so S_{j+1} runs under the softmax of block j
//   P_j = 2^(S_j c - m_ref)   one softmax thread per query row (= TMEM lane); m_ref is an INTEGER in the log2
//                             domain, so a change of reference rescales P, the row sum and O by an exact power
//                             of two: the result is bit-identical to a two-pass softmax whose reference is
//                             ceil(rowmax * c) - the numerics contract the oracle's emulate_bf16 mode restates
//   O  += P_j V_j             tcgen05.mma TS: A = bf16(P_j) read from TMEM (written with tcgen05.st, double
//                             buffered), B = V tile as an MN-major smem operand
// The reference only moves when a row's scores exceed it by more than 2^8 (rare after the first block); then
// the thread rescales its O row in TMEM and redoes the block.
//
// TMEM columns: S0 [0,128) S1 [128,256) O [256,384) P0 [384,448) P1 [448,512).
// Warps: 0-3 softmax + epilogue (warp w owns TMEM lanes [32w, 32w+32)), 4 TMA producer, 5 MMA issuer / TMEM.
#include "host_util.h"
#include "kernels.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <map>
#include <mutex>

namespace oasr {
namespace {

constexpr int ATT_THREADS = 192;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int CH = 16;                  // head-dim columns per smem chunk (32 bytes, SWIZZLE_32B)
constexpr int CH_BYTES = 128 * CH * 2;  // one [128 rows][16 cols] chunk
constexpr int KV_STAGES = 2;
constexpr int TMEM_COLS = 512;
constexpr int TM_S = 0, TM_O = 256, TM_P = 384;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units: P stays below 2^8 relative to the reference

struct AttnParams {
  int T, H, hd, d;
  float scale_log2e;
  const int* n_frames;
  __nv_bfloat16* out;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_v2_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nch = p.hd / CH;
  const int tile_bytes = nch * CH_BYTES;
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + tile_bytes;  // [stage][K | V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + KV_STAGES * 2 * tile_bytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // 2
  uint64_t* kv_empty = bars + 3;    // 2
  uint64_t* s_full = bars + 5;      // 2
  uint64_t* s_empty = bars + 7;     // 2
  uint64_t* p_full = bars + 9;      // 2
  uint64_t* p_free = bars + 11;     // 2: P.V that read P[b] has retired
  uint64_t* o_full = bars + 13;     // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_keys = min(p.n_frames ? p.n_frames[b] : p.T, p.T);
  const int nblk = (n_keys + BKV - 1) / BKV;

  if (nblk == 0) {  // fully padded window: attention output is defined as zero
    for (int i = threadIdx.x; i < BQ * (p.hd / 8); i += blockDim.x) {
      const int r = i / (p.hd / 8), c8 = i % (p.hd / 8);
      if (q0 + r < p.T)
        reinterpret_cast<uint4*>(p.out + ((long long)b * p.T + q0 + r) * p.d + h * p.hd)[c8] = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_free[i], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      const int qcol = h * p.hd, kcol = p.d + h * p.hd, vcol = 2 * p.d + h * p.hd;
      mbar_arrive_expect_tx(q_full, tile_bytes);
      for (int c = 0; c < nch; ++c) tma_load_3d(sQ + c * CH_BYTES, &tm, q_full, qcol + c * CH, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        uint8_t* sK = sKV + s * 2 * tile_bytes;
        uint8_t* sV = sK + tile_bytes;
        mbar_arrive_expect_tx(&kv_full[s], 2 * tile_bytes);
        for (int c = 0; c < nch; ++c) tma_load_3d(sK + c * CH_BYTES, &tm, &kv_full[s], kcol + c * CH, j * BKV, b);
        for (int c = 0; c < nch; ++c) tma_load_3d(sV + c * CH_BYTES, &tm, &kv_full[s], vcol + c * CH, j * BKV, b);
      }
    }
  } else if (warp == 5) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(BQ, p.hd, 0, 1);  // A = P from TMEM (K-major), B = V MN-major
      const uint32_t q_addr = smem_u32(sQ);
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        mbar_wait(&s_empty[s], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + s * 2 * tile_bytes);
        for (int c = 0; c < nch; ++c) {
          const uint64_t adesc = make_smem_desc(q_addr + c * CH_BYTES, 16, 256, SWZ_32B);
          const uint64_t bdesc = make_smem_desc(k_addr + c * CH_BYTES, 16, 256, SWZ_32B);
          umma_ss(tmem_base + TM_S + s * BKV, adesc, bdesc, idesc_s, c != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[s]);
      };
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);
        const int s = j & 1;
        mbar_wait(&p_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sKV + s * 2 * tile_bytes + tile_bytes);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t bdesc = make_smem_desc(v_addr + k * (16 * CH * 2), CH_BYTES, 256, SWZ_32B);
          umma_ts(tmem_base + TM_O, tmem_base + TM_P + s * (BKV / 2) + k * 8, bdesc, idesc_o, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(&p_free[s]);
      }
      umma_commit(o_full);
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue (warps 0-3)
    const int r = warp * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (uint32_t(warp * 32) << 16);
    const float c = p.scale_log2e;
    float m_ref = 0.f;  // integer-valued reference in the log2 domain
    float sum = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int s = j & 1;
      const uint32_t t_s = t_lane + TM_S + s * BKV;
      const uint32_t t_p = t_lane + TM_P + s * (BKV / 2);
      const int ncols = min(BKV, n_keys - j * BKV);  // valid keys in this block
      mbar_wait(&s_full[s], (j >> 1) & 1);
      tc_fence_after();
      if (j == 0) {  // first block: find the reference before any P is produced
        float mx = -INFINITY;
#pragma unroll 1
        for (int cc = 0; cc < BKV; cc += 32) {
          uint32_t v[32];
          tmem_ld32(t_s + cc, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc + i < ncols) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        m_ref = ceilf(mx * c);
      }
      mbar_wait(&p_free[s], ((j >> 1) & 1) ^ 1);  // P.V_{j-2} no longer reads this P buffer
      float bsum, bmax;
      auto produce_p = [&]() {
        bsum = 0.f;
        bmax = -INFINITY;
#pragma unroll 1
        for (int cc = 0; cc < BKV; cc += 32) {
          uint32_t v[32];
          tmem_ld32(t_s + cc, v);
          tmem_ld_wait();
          uint32_t pk[16];
          if (cc + 32 <= ncols) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float x0 = fmaf(__uint_as_float(v[i]), c, -m_ref);
              const float x1 = fmaf(__uint_as_float(v[i + 1]), c, -m_ref);
              bmax = fmaxf(bmax, fmaxf(x0, x1));
              const float p0 = ex2(x0), p1 = ex2(x1);
              bsum += p0 + p1;
              pk[i >> 1] = pack_bf16x2(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float x0 = fmaf(__uint_as_float(v[i]), c, -m_ref);
              const float x1 = fmaf(__uint_as_float(v[i + 1]), c, -m_ref);
              const bool ok0 = cc + i < ncols, ok1 = cc + i + 1 < ncols;
              if (ok0) bmax = fmaxf(bmax, x0);
              if (ok1) bmax = fmaxf(bmax, x1);
              const float p0 = ok0 ? ex2(x0) : 0.f, p1 = ok1 ? ex2(x1) : 0.f;
              bsum += p0 + p1;
              pk[i >> 1] = pack_bf16x2(p0, p1);
            }
          }
          tmem_st16(t_p + (cc >> 1), pk);
        }
      };
      produce_p();
      // a row whose scores outgrew the reference moves it by an integer and rescales by an exact power of two
      const bool grow = bmax > RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, grow)) {
        const float k = grow ? ceilf(bmax) : 0.f;
        const float f = ex2(-k);  // exact: k is an integer
        m_ref += k;
        sum *= f;
        if (j > 0) {
          mbar_wait(&p_free[s ^ 1], ((j - 1) >> 1) & 1);  // P.V_{j-1} has finished updating O
          tc_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < p.hd; cc += 16) {
            uint32_t v[16];
            tmem_ld16(t_lane + TM_O + cc, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st16(t_lane + TM_O + cc, v);
          }
        }
        produce_p();
      }
      sum += bsum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[s]);
        mbar_arrive(&p_full[s]);
      }
    }
    // epilogue: O / rowsum -> bf16
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    const bool row_ok = (q0 + r) < p.T;
    __nv_bfloat16* orow = p.out + ((long long)b * p.T + q0 + r) * p.d + h * p.hd;
#pragma unroll 1
    for (int cc = 0; cc < p.hd; cc += 16) {
      uint32_t v[16];
      tmem_ld16(t_lane + TM_O + cc, v);
      tmem_ld_wait();
      if (row_ok) {
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2)
          o[i >> 1] = pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
        uint4* dst = reinterpret_cast<uint4*>(orow + cc);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

struct AttKey {
  const void* base;
  int B, T, d3;
  bool operator<(const AttKey& o) const {
    if (base != o.base) return base < o.base;
    if (B != o.B) return B < o.B;
    if (T != o.T) return T < o.T;
    return d3 < o.d3;
  }
};
std::map<AttKey, CUtensorMap> g_att_tmaps;
std::mutex g_att_mu;

}  // namespace

int attention_bf16_v2(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
                      cudaStream_t stream) {
  OASR_REQUIRE(qkv && out && B > 0 && T > 0 && H > 0, "attention: bad arguments");
  OASR_REQUIRE(hd % 16 == 0 && hd >= 16 && hd <= 128, "attention: head_dim must be a multiple of 16 in [16, 128]");
  OASR_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "attention: buffers must be 16-byte aligned");
  const int d = H * hd;
  CUtensorMap tm;
  {
    std::lock_guard<std::mutex> g(g_att_mu);
    AttKey key{qkv, B, T, 3 * d};
    auto it = g_att_tmaps.find(key);
    if (it == g_att_tmaps.end()) {
      uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)B};
      uint64_t strides[2] = {(uint64_t)3 * d * 2, (uint64_t)T * 3 * d * 2};
      uint32_t box[3] = {CH, 128, 1};
      OASR_TRY(make_tmap_bf16(&tm, qkv, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_32B));
      if (g_att_tmaps.size() > 1024) g_att_tmaps.clear();
      g_att_tmaps[key] = tm;
    } else {
      tm = it->second;
    }
  }
  const int tile_bytes = (hd / CH) * CH_BYTES;
  const int smem_bytes = tile_bytes * (1 + 2 * KV_STAGES) + 256 + 1024;
  static int attr_smem = 0;
  if (smem_bytes > attr_smem) {
    OASR_CUDA_CHECK(cudaFuncSetAttribute(attention_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_smem = smem_bytes;
  }
  AttnParams p;
  p.T = T;
  p.H = H;
  p.hd = hd;
  p.d = d;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.n_frames = n_frames;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  dim3 grid((T + BQ - 1) / BQ, H, B);
  attention_v2_kernel<<<grid, ATT_THREADS, smem_bytes, stream>>>(tm, p);
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

int attention_bf16(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
                   cudaStream_t stream) {
  static const bool use_v1 = [] {
    const char* e = std::getenv("OASR_ATTN_V1");
    return e != nullptr && e[0] == '1';
  }();
  return use_v1 ? attention_bf16_v1(qkv, out, n_frames, B, T, H, hd, scale, stream)
                : attention_bf16_v2(qkv, out, n_frames, B, T, H, hd, scale, stream);
}

}  // namespace oasr
