"""One GEMM shape through liboasr and through cuBLAS (torch.matmul), a few launches each: the target of an
ncu capture.  usage: prof_gemm_vs_cublas.py M N K epilogue"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr import _native as N  # noqa: E402

M, Nn, K, epi = (int(v) for v in sys.argv[1:5])
lib = N.load()
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
W = (torch.randn(Nn, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
bias = torch.randn(Nn, device="cuda", generator=g)
out = torch.empty(M, Nn, device="cuda", dtype=torch.float32 if epi in (2, 3) else torch.bfloat16)
resid = out if epi == 3 else None
if epi == 3:
    out.normal_()
ref = torch.empty(M, Nn, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    N.check(lib.oasr_gemm(N.ptr(A), N.ptr(W), N.ptr(bias), M, Nn, K, epi, N.ptr(out), Nn, N.ptr(resid), None, None, None,
                          N.stream_ptr()))
    torch.matmul(A, W.t(), out=ref)
torch.cuda.synchronize()
print("ok")
