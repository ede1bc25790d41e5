"""CUDA-event timings of the encoder GEMM shapes (1B, B = 32) and the FE conv layers, each run back to back
`reps` times.  usage: prof_gemm_shapes.py [reps]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr import _native as N  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
lib = N.load()
g = torch.Generator(device="cuda").manual_seed(0)
M = 47968


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gemm(name, Nn, K, epi):
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(Nn, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(Nn, device="cuda", generator=g)
    out = torch.empty(M, Nn, device="cuda", dtype=torch.float32 if epi in (2, 3) else torch.bfloat16)
    resid = out if epi == 3 else None
    if epi == 3:
        out.normal_()
    ms = timeit(lambda: N.check(lib.oasr_gemm(N.ptr(A), N.ptr(W), N.ptr(bias), M, Nn, K, epi, N.ptr(out), Nn, N.ptr(resid),
                                              None, None, None, N.stream_ptr())))
    print(f"{name:10s} M={M} N={Nn} K={K} epi={epi}: {ms * 1e3:8.1f} us  {2.0 * M * Nn * K / ms / 1e9:7.1f} TFLOP/s", flush=True)


def conv(name, B, L_in, k):
    L_pad = (L_in + 3) & ~1
    x = torch.randn(B, L_pad, 512, device="cuda", generator=g).bfloat16()
    w = (torch.randn(512, k * 512, device="cuda", generator=g) / (k * 512) ** 0.5).bfloat16()
    bias, gamma, beta = (torch.randn(512, device="cuda", generator=g) for _ in range(3))
    T = (L_in - k) // 2 + 1
    out = torch.empty(B, T, 512, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: N.check(lib.oasr_conv_ln_gelu(N.ptr(x), B, L_in, L_pad, k, N.ptr(w), N.ptr(bias), N.ptr(gamma),
                                                      N.ptr(beta), N.ptr(out), N.stream_ptr())))
    fl = 2.0 * B * T * 512 * 512 * k
    print(f"{name:10s} B={B} L_in={L_in} k={k}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)


gemm("qkv", 3840, 1280, 0)
gemm("ffn1", 5120, 1280, 1)
gemm("outproj", 1280, 1280, 3)
gemm("ffn2", 1280, 5120, 3)
conv("fe1", 32, 95999, 3)
conv("fe2", 32, 47999, 3)
conv("fe5", 32, 5999, 2)

# cuBLAS yard-stick on the same shapes (no epilogue)
for name, Nn, K in (("qkv", 3840, 1280), ("ffn1", 5120, 1280), ("outproj", 1280, 1280), ("ffn2", 1280, 5120)):
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = torch.randn(Nn, K, device="cuda", generator=g).bfloat16()
    out = torch.empty(M, Nn, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: torch.matmul(A, W.t(), out=out))
    print(f"cublas {name:8s} M={M} N={Nn} K={K}: {ms * 1e3:8.1f} us  {2.0 * M * Nn * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
