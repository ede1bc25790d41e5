"""Runs one encoder GEMM shape alone (for ncu) and prints its CUDA-event time.
usage: prof_gemm.py [M N K epilogue reps]   (epilogue: 0 bf16, 1 bf16+gelu, 3 f32+resid)"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr import _native as N  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
M, Nn, K, epi, reps = (a + [47968, 5120, 1280, 1, 5][len(a):])[:5]
lib = N.load()
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
W = (torch.randn(Nn, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
bias = torch.randn(Nn, device="cuda", generator=g)
out = torch.empty(M, Nn, device="cuda", dtype=torch.float32 if epi in (2, 3) else torch.bfloat16)
resid = out if epi == 3 else None
if epi == 3:
    out.normal_()


def run():
    N.check(lib.oasr_gemm(N.ptr(A), N.ptr(W), N.ptr(bias), M, Nn, K, epi, N.ptr(out), Nn, N.ptr(resid), None, None, None,
                          N.stream_ptr()))


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"gemm M={M} N={Nn} K={K} epi={epi}: {ms:.3f} ms/launch, {2.0 * M * Nn * K / ms / 1e9:.1f} TFLOP/s")
