"""Runs oasr_attention once on seeded inputs and saves the output (tests/test_gpu_kernels.py runs it under different
OASR_ATTN / OASR_ATT_POLY settings: the switches are read once per process).
    python scripts/attention_variant.py out.npy B T H hd"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr import _native as N  # noqa: E402

out_path = sys.argv[1]
B, T, H, hd = (int(v) for v in sys.argv[2:6])
d = H * hd
g = torch.Generator().manual_seed(11)
qkv = (torch.randn(B * T, 3 * d, generator=g) * 1.5).bfloat16().cuda()
nf = torch.tensor([T] + [max(1, T - 37 * (b + 1)) for b in range(B - 1)], dtype=torch.int32).cuda()   # ragged
out = torch.zeros(B * T, d, dtype=torch.bfloat16, device="cuda")
N.check(N.load().oasr_attention(N.ptr(qkv), N.ptr(out), N.ptr(nf), B, T, H, hd, hd ** -0.5, N.stream_ptr()), "attention")
torch.cuda.synchronize()
np.save(out_path, out.float().cpu().numpy())
