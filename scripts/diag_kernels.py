"""How close is each kernel to 'reference value rounded once to bf16'?  Prints, per kernel, the relative error
against the unrounded fp32 reference (pure bf16 rounding gives ~1.1e-3) and the fraction of outputs that are
bit-identical to the rounded reference (fp32 summation-order noise only flips ~0.1 % of roundings)."""
import math
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr import _native as N  # noqa: E402

lib = N.load()
dev = "cuda"
g = torch.Generator().manual_seed(1)


def rnd(shape, s=1.0):
    return (torch.randn(shape, generator=g) * s).to(dev)


def report(name, out, ref):
    out = out.float()
    rel = float((out - ref).norm() / ref.norm())
    same = float((out == ref.bfloat16().float()).float().mean())
    ulp = float(((out - ref).abs() / ref.abs().clamp_min(1e-3)).median())
    print(f"{name:28s} rel={rel:.2e} bit-identical-to-rounded-ref={same:.4f} median|rel diff|={ulp:.2e}", flush=True)


def sync():
    torch.cuda.synchronize()


# fe layer 0
B, L = 2, 16000
wave = rnd((B, L))
w = rnd((512, 1, 10), 0.4)
bias, ga, be = rnd((512,), 0.3), 1 + rnd((512,), 0.1), rnd((512,), 0.1)
T0 = (L - 10) // 5 + 1
out = torch.zeros((B, T0, 512), dtype=torch.bfloat16, device=dev)
N.check(lib.oasr_fe_layer0(N.ptr(wave), B, L, N.ptr(w[:, 0, :].t().contiguous()), N.ptr(bias), N.ptr(ga), N.ptr(be),
                           N.ptr(out), N.stream_ptr()))
sync()
ref = F.gelu(F.layer_norm(F.conv1d(wave[:, None].double(), w.double(), bias.double(), stride=5).transpose(1, 2),
                          (512,), ga.double(), be.double(), 1e-5)).float()
report("fe_layer0", out, ref)

# conv layer k=3
L_in = 1001
L_pad = (L_in + 3) & ~1
x = rnd((B, L_pad, 512)).bfloat16()
wc = rnd((512, 512, 3), 1 / math.sqrt(1536)).bfloat16().float()
w_tap = wc.permute(0, 2, 1).reshape(512, 1536).contiguous().bfloat16()
L_out = (L_in - 3) // 2 + 1
out = torch.zeros((B, L_out, 512), dtype=torch.bfloat16, device=dev)
N.check(lib.oasr_conv_ln_gelu(N.ptr(x), B, L_in, L_pad, 3, N.ptr(w_tap), N.ptr(bias), N.ptr(ga), N.ptr(be), N.ptr(out),
                              N.stream_ptr()))
sync()
y = F.conv1d(x[:, :L_in].double().transpose(1, 2), wc.double(), bias.double(), stride=2).transpose(1, 2)
pre = y.float()
ref = F.gelu(F.layer_norm(y, (512,), ga.double(), be.double(), 1e-5)).float()
report("conv_ln_gelu k=3", out, ref)

# plain GEMM epilogues
M, Nn, K = 1000, 1280, 1280
A = rnd((M, K)).bfloat16()
W = rnd((Nn, K), 1 / math.sqrt(K)).bfloat16()
bv = rnd((Nn,))
acc = (A.double() @ W.double().t() + bv.double())
for epi, name, fn in ((N.EPI_BF16, "gemm bf16", lambda t: t), (N.EPI_BF16_GELU, "gemm gelu bf16", F.gelu)):
    out = torch.zeros((M, Nn), dtype=torch.bfloat16, device=dev)
    N.check(lib.oasr_gemm(N.ptr(A), N.ptr(W), N.ptr(bv), M, Nn, K, epi, N.ptr(out), Nn, None, None, None, None,
                          N.stream_ptr()))
    sync()
    report(name, out, fn(acc).float())
out = torch.zeros((M, Nn), dtype=torch.float32, device=dev)
N.check(lib.oasr_gemm(N.ptr(A), N.ptr(W), N.ptr(bv), M, Nn, K, N.EPI_F32, N.ptr(out), Nn, None, None, None, None,
                      N.stream_ptr()))
sync()
print(f"{'gemm f32':28s} rel={float((out - acc.float()).norm() / acc.float().norm()):.2e} "
      f"max abs={float((out - acc.float()).abs().max()):.2e}")
# LN epilogue GEMM
W5 = rnd((512, 1536), 1 / math.sqrt(1536)).bfloat16()
A5 = rnd((M, 1536)).bfloat16()
out = torch.zeros((M, 512), dtype=torch.bfloat16, device=dev)
N.check(lib.oasr_gemm(N.ptr(A5), N.ptr(W5), N.ptr(bias), M, 512, 1536, N.EPI_LN_GELU_BF16, N.ptr(out), 512, None,
                      N.ptr(ga), N.ptr(be), None, N.stream_ptr()))
sync()
ref = F.gelu(F.layer_norm(A5.double() @ W5.double().t() + bias.double(), (512,), ga.double(), be.double(), 1e-5)).float()
report("gemm ln+gelu", out, ref)

# layernorm
xx = rnd((M, 1280), 2.0) + 0.5
g2, b2 = 1 + rnd((1280,), 0.1), rnd((1280,), 0.1)
ob = torch.zeros((M, 1280), dtype=torch.bfloat16, device=dev)
N.check(lib.oasr_layernorm(N.ptr(xx), 0, M, 1280, N.ptr(g2), N.ptr(b2), N.ptr(ob), None, N.stream_ptr()))
sync()
report("layernorm", ob, F.layer_norm(xx.double(), (1280,), g2.double(), b2.double(), 1e-5).float())

# attention
for H, hd in ((4, 64), (4, 80)):
    T, Bq = 300, 2
    d = H * hd
    qkv = rnd((Bq * T, 3 * d)).bfloat16()
    nf = torch.full((Bq,), T, dtype=torch.int32, device=dev)
    out = torch.zeros((Bq * T, d), dtype=torch.bfloat16, device=dev)
    N.check(lib.oasr_attention(N.ptr(qkv), N.ptr(out), N.ptr(nf), Bq, T, H, hd, hd ** -0.5, N.stream_ptr()))
    sync()
    q, k, v = qkv.double().view(Bq, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    sl = (q @ k.transpose(-1, -2)) * (hd ** -0.5) * 1.4426950408889634
    m = torch.ceil(sl.amax(-1, keepdim=True))
    p = torch.exp2(sl - m)
    ref = ((p.float().bfloat16().double() @ v) / p.sum(-1, keepdim=True)).permute(0, 2, 1, 3).reshape(Bq * T, d).float()
    report(f"attention hd={hd}", out, ref)
