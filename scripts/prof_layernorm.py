"""CUDA-event timing of the row LayerNorm at the encoder shapes.  usage: prof_layernorm.py [reps]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr import _native as N  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lib = N.load()
M = 47968
for D in (1024, 1280, 2048):
    x = torch.randn(M, D, device="cuda")
    g = torch.randn(D, device="cuda")
    b = torch.randn(D, device="cuda")
    out = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def run():
        N.check(lib.oasr_layernorm(N.ptr(x), 0, M, D, N.ptr(g), N.ptr(b), N.ptr(out), None, N.stream_ptr()))

    run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()   # evict x from L2 (126 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"layernorm M={M} D={D}: {ms * 1e3:.1f} us  {M * D * 6 / ms / 1e6:.0f} GB/s (algorithmic 6 B/element)")
