"""Times the pos-conv stage alone (memset + pad-cast + slab conv GEMM) at the 1B shape through the C-ABI.

    gpurun -- python scripts/prof_posconv.py
"""
import sys, math
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/omnilingual-asr_b200")
import torch
from omnilingual_asr import _native as N
lib = N.load()
d, groups, k, T, B = 1280, 16, 128, 1499, 32
cg = d // groups
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B * T, d, device="cuda", generator=g)
k_pad = ((cg + 63) // 64) * 64
wt = torch.zeros((d, k, k_pad), device="cuda")
wt[:, :, :cg] = torch.randn(d, k, cg, device="cuda", generator=g) / math.sqrt(cg * k)
wt = wt.reshape(d, k * k_pad).contiguous().bfloat16()
bias = torch.zeros(d, device="cuda")
scratch = torch.zeros((B, T + k, d), dtype=torch.bfloat16, device="cuda")
def run():
    N.check(lib.oasr_posconv(N.ptr(x), B, T, d, groups, k, N.ptr(wt), N.ptr(bias), N.ptr(scratch), N.stream_ptr()), "posconv")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"posconv (memset + pad-cast + conv) d={d}: {e0.elapsed_time(e1) / 10:.3f} ms")
