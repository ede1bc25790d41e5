"""Runs the attention kernel alone at the 1B bench shape (for ncu) and prints its CUDA-event time."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr import _native as N  # noqa: E402

B, T, H, hd = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 1499, 16, 80)))
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
lib = N.load()
d = H * hd
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * T, 3 * d, device="cuda", generator=g).bfloat16()
out = torch.empty(B * T, d, device="cuda", dtype=torch.bfloat16)
nf = torch.full((B,), T, dtype=torch.int32, device="cuda")
for _ in range(2):
    N.check(lib.oasr_attention(N.ptr(qkv), N.ptr(out), N.ptr(nf), B, T, H, hd, hd ** -0.5, N.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    N.check(lib.oasr_attention(N.ptr(qkv), N.ptr(out), N.ptr(nf), B, T, H, hd, hd ** -0.5, N.stream_ptr()))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 4.0 * B * H * T * T * hd
print(f"attention B={B} T={T} H={H} hd={hd}: {ms:.3f} ms/launch, {fl / ms / 1e9:.1f} TFLOP/s")
