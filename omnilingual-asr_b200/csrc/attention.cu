// a14 dispatcher: attention_v7 (persistent; three query tiles per CTA, 48-key blocks) for head_dim <= 80 (300M, 1B),
// attention_v4 (two tiles, 96 / 80-key blocks) above (3B, 7B).  OASR_ATTN=4 forces v4 everywhere (A/B runs; covered by
// tests/test_gpu_kernels.py::test_attention_env_switch_v4).
// Earlier generations (v1 two-pass reference, v2 single pass with P aliased onto S, v3 ping-pong warpgroups, v5 with
// two warps per query row, v6 = v7 with one work item per CTA) are in the history; what each taught is in
// profiles/r1_notes.md.
#include "kernels.cuh"

#include <cstdlib>

namespace oasr {

int attention_bf16(const void* qkv, void* out, const int* n_frames, int B, int T, int H, int hd, float scale,
                   cudaStream_t stream) {
  static const int version = [] {
    const char* e = std::getenv("OASR_ATTN");
    return e != nullptr ? std::atoi(e) : 7;
  }();
  if (version != 4 && hd <= 80) return attention_bf16_v7(qkv, out, n_frames, B, T, H, hd, scale, stream);
  return attention_bf16_v4(qkv, out, n_frames, B, T, H, hd, scale, stream);
}

}  // namespace oasr
