"""LayerNorm timed where it runs: right after the residual GEMM that wrote x (L2 holds the tail of x).
usage: prof_ln_insitu.py [D]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr import _native as N  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
lib = N.load()
M = 47968
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, D, device="cuda", generator=g).bfloat16()
W = (torch.randn(D, D, device="cuda", generator=g) / D ** 0.5).bfloat16()
bias = torch.randn(D, device="cuda", generator=g)
x = torch.randn(M, D, device="cuda", generator=g)
gam, bet = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
out = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
ts = []
for i in range(12):
    N.check(lib.oasr_gemm(N.ptr(A), N.ptr(W), N.ptr(bias), M, D, D, 3, N.ptr(x), D, N.ptr(x), None, None, None, N.stream_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    N.check(lib.oasr_layernorm(N.ptr(x), 0, M, D, N.ptr(gam), N.ptr(bet), N.ptr(out), None, N.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts[2:])[len(ts[2:]) // 2]
print(f"layernorm after the residual GEMM, D={D}: {ms * 1e3:.1f} us  {M * D * 6 / ms / 1e6:.0f} GB/s algorithmic")
