// Attention, first version (two passes over the keys, P staged through shared memory).  Kept as the
// reference implementation of the numerics contract; select with OASR_ATTN_V1=1.  See attention_v2.cu.
// Bidirectional self-attention for one (sequence, head, 128-query tile) per CTA on tcgen05 (a14).
//
//   S = Q K^T   tcgen05.mma, A = Q tile, B = K tile, both K-major bf16 in 32B-swizzled 16-column chunks
//   P = exp2((S - rowmax) * scale * log2e)   one softmax thread per query row (TMEM lane), fp32
//   O += P V    tcgen05.mma, A = bf16(P) written to 128B-swizzled smem, B = V tile as an MN-major operand
//
// Two passes over the keys: pass 1 only finds the exact row maximum, pass 2 recomputes S and accumulates
// O with no rescaling.  T is at most a few thousand frames (1499 per 30 s window), the kernel is bound by
// the exp throughput and not by the tensor pipe, and the result is the plain "exp(s - max)" softmax whose
// bf16 rounding point the oracle's emulate_bf16 mode reproduces exactly.
//
// Warps: 0-3 softmax + epilogue (warp w owns TMEM lanes [32w, 32w+32)), 4 TMA producer, 5 MMA issuer/TMEM.
#include "host_util.h"
#include "kernels.cuh"
#include "ptx.cuh"

#include <map>
#include <mutex>

namespace oasr {
namespace {

constexpr int ATT_THREADS = 192;
constexpr int BQ = 128;   // query rows per CTA
constexpr int BKV = 128;  // keys per block
constexpr int CH = 16;    // head-dim columns per smem chunk (32 bytes, SWIZZLE_32B)
constexpr int CH_BYTES = 128 * CH * 2;  // one [128 rows][16 cols] chunk
constexpr int KV_STAGES = 2;
constexpr int P_BYTES = BQ * BKV * 2;
constexpr int TMEM_COLS = 512;
constexpr int TM_S = 0;    // S[0] at column 0, S[1] at column 128
constexpr int TM_O = 256;  // O at column 256

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
attention_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nch = p.hd / CH;
  const int tile_bytes = nch * CH_BYTES;  // one Q / K / V tile
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + tile_bytes;                    // [stage][K | V]
  uint8_t* sP = sKV + KV_STAGES * 2 * tile_bytes;    // 128B-swizzled [2 atoms][128 rows][128 bytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = bars + 3;
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 7;
  uint64_t* p_full = bars + 9;
  uint64_t* p_empty = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

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
    }
    mbar_init(p_full, 4);
    mbar_init(p_empty, 1);
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
      int it = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < nblk; ++j, ++it) {
          const int s = it & 1;
          mbar_wait(&kv_empty[s], ((it >> 1) & 1) ^ 1);
          uint8_t* sK = sKV + s * 2 * tile_bytes;
          uint8_t* sV = sK + tile_bytes;
          mbar_arrive_expect_tx(&kv_full[s], pass == 0 ? tile_bytes : 2 * tile_bytes);
          for (int c = 0; c < nch; ++c) tma_load_3d(sK + c * CH_BYTES, &tm, &kv_full[s], kcol + c * CH, j * BKV, b);
          if (pass == 1)
            for (int c = 0; c < nch; ++c) tma_load_3d(sV + c * CH_BYTES, &tm, &kv_full[s], vcol + c * CH, j * BKV, b);
        }
      }
    }
  } else if (warp == 5) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(BQ, p.hd, 0, 1);  // B = V is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      int it = 0;  // kv stage uses
      int sc = 0;  // S buffer uses
      mbar_wait(q_full, 0);
      auto issue_s = [&](int stage_it, bool release_kv) {
        const int s = stage_it & 1;
        mbar_wait(&kv_full[s], (stage_it >> 1) & 1);
        const int sb = sc & 1;
        mbar_wait(&s_empty[sb], ((sc >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + s * 2 * tile_bytes);
        for (int c = 0; c < nch; ++c) {
          const uint64_t adesc = make_smem_desc(q_addr + c * CH_BYTES, 16, 256, SWZ_32B);
          const uint64_t bdesc = make_smem_desc(k_addr + c * CH_BYTES, 16, 256, SWZ_32B);
          umma_ss(tmem_base + TM_S + sb * BKV, adesc, bdesc, idesc_s, c != 0 ? 1u : 0u);
        }
        if (release_kv) umma_commit(&kv_empty[s]);
        umma_commit(&s_full[sb]);
        ++sc;
      };
      // pass 1: S only (row maxima)
      for (int j = 0; j < nblk; ++j, ++it) issue_s(it, true);
      // pass 2: S one block ahead of P.V
      const int it2 = it;
      issue_s(it2, false);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(it2 + j + 1, false);
        const int s = (it2 + j) & 1;
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sKV + s * 2 * tile_bytes + tile_bytes);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t adesc = make_smem_desc(p_addr + (k >> 2) * (BQ * 128) + (k & 3) * 32, 16, 1024, SWZ_128B);
          const uint64_t bdesc = make_smem_desc(v_addr + k * (16 * CH * 2), CH_BYTES, 256, SWZ_32B);
          umma_ss(tmem_base + TM_O, adesc, bdesc, idesc_o, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(p_empty);
      }
      umma_commit(o_full);
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue (warps 0-3)
    const int r = warp * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (uint32_t(warp * 32) << 16);
    int sc = 0;
    float m = -INFINITY;
    for (int j = 0; j < nblk; ++j, ++sc) {
      const int sb = sc & 1;
      mbar_wait(&s_full[sb], (sc >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_lane + TM_S + sb * BKV + c, v);
        tmem_ld_wait();
        const int col0 = j * BKV + c;
        if (col0 + 32 <= n_keys) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i < n_keys) m = fmaxf(m, __uint_as_float(v[i]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
    }
    const float mneg = -m * p.scale_log2e;
    float sum = 0.f;
    uint8_t* prow = sP + r * 128;
    for (int j = 0; j < nblk; ++j, ++sc) {
      const int sb = sc & 1;
      mbar_wait(&s_full[sb], (sc >> 1) & 1);
      tc_fence_after();
      mbar_wait(p_empty, (j & 1) ^ 1);  // P.V of the previous block has finished reading sP
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_lane + TM_S + sb * BKV + c, v);
        tmem_ld_wait();
        const int col0 = j * BKV + c;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2(fmaf(__uint_as_float(v[i]), p.scale_log2e, mneg));
          float p1 = ex2(fmaf(__uint_as_float(v[i + 1]), p.scale_log2e, mneg));
          if (col0 + i >= n_keys) p0 = 0.f;
          if (col0 + i + 1 >= n_keys) p1 = 0.f;
          sum += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        // 128B swizzle: 16-byte chunk index XOR (row & 7) inside each 128-byte row of a [128][64] atom
        uint8_t* atom = prow + (c >> 6) * (BQ * 128);
        const int chunk0 = (c & 63) >> 3;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ch = (chunk0 + g) ^ (r & 7);
          *reinterpret_cast<uint4*>(atom + ch * 16) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[sb]);
        mbar_arrive(p_full);
      }
    }
    // epilogue: O / rowsum -> bf16
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    const bool row_ok = (q0 + r) < p.T;
    __nv_bfloat16* orow = p.out + ((long long)b * p.T + q0 + r) * p.d + h * p.hd;
#pragma unroll 1
    for (int c = 0; c < p.hd; c += 16) {
      uint32_t v[16];
      tmem_ld16(t_lane + TM_O + c, v);
      tmem_ld_wait();
      if (row_ok) {
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2)
          o[i >> 1] = pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
        uint4* dst = reinterpret_cast<uint4*>(orow + c);
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

int attention_bf16_v1(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
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
  const int smem_bytes = tile_bytes * (1 + 2 * KV_STAGES) + P_BYTES + 256 + 1024;
  static int attr_smem = 0;
  if (smem_bytes > attr_smem) {
    OASR_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
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
  attention_kernel<<<grid, ATT_THREADS, smem_bytes, stream>>>(tm, p);
  OASR_CUDA_CHECK(cudaGetLastError());
  return OASR_OK;
}

}  // namespace oasr
