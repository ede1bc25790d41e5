"""Per-kernel parity on a B200, through the C-ABI stage entry points (include/oasr.h).

Each CUDA kernel is compared with the same op written in plain PyTorch fp32 on operands rounded to bf16
where the kernel consumes bf16 (tolerances are written next to each assert).  Integer outputs are bit-exact.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from omnilingual_asr import _native as N
from oracle import ctc_oracle as O
from tests._util import bf16_round, gemm, lib, rel_err, sync, unpack_keys

pytestmark = pytest.mark.gpu


def _rand(shape, gen, scale=1.0, device="cuda"):
    return (torch.randn(shape, generator=gen) * scale).to(device)


@pytest.fixture()
def gen():
    return torch.Generator().manual_seed(1234)


# ------------------------------------------------------------------------------------------- GEMM family
@pytest.mark.parametrize("M,Nn,K", [(128, 256, 64), (300, 320, 192), (5120, 1280, 1280), (1499, 3840, 1280),
                                    (777, 5120, 1280)])
def test_gemm_bf16_bias(device, gen, M, Nn, K):
    A = _rand((M, K), gen).bfloat16()
    W = _rand((Nn, K), gen, 1 / math.sqrt(K)).bfloat16()
    bias = _rand((Nn,), gen)
    out = torch.zeros((M, Nn), dtype=torch.bfloat16, device=device)
    gemm(A, W, bias, N.EPI_BF16, out)
    ref = A.float() @ W.float().t() + bias
    # bf16 output rounding: 2^-9 relative per element; accumulation order differences are ~1e-6
    assert rel_err(out, ref) < 4e-3
    assert torch.allclose(out.float(), ref, rtol=1e-2, atol=1e-2)


def test_gemm_gelu(device, gen):
    M, Nn, K = 1000, 640, 320
    A = _rand((M, K), gen).bfloat16()
    W = _rand((Nn, K), gen, 1 / math.sqrt(K)).bfloat16()
    bias = _rand((Nn,), gen)
    out = torch.zeros((M, Nn), dtype=torch.bfloat16, device=device)
    gemm(A, W, bias, N.EPI_BF16_GELU, out)
    ref = F.gelu(A.float() @ W.float().t() + bias)
    assert torch.allclose(out.float(), ref, rtol=1e-2, atol=1e-2)
    assert rel_err(out, ref) < 4e-3


def test_gemm_f32_and_residual(device, gen):
    M, Nn, K = 900, 1280, 5120
    A = _rand((M, K), gen).bfloat16()
    W = _rand((Nn, K), gen, 1 / math.sqrt(K)).bfloat16()
    bias = _rand((Nn,), gen)
    ref = A.float() @ W.float().t() + bias
    out = torch.zeros((M, Nn), dtype=torch.float32, device=device)
    gemm(A, W, bias, N.EPI_F32, out)
    # fp32 accumulate of bf16 products: only summation order differs
    assert torch.allclose(out, ref, rtol=1e-4, atol=2e-4)
    x = _rand((M, Nn), gen)
    x0 = x.clone()
    gemm(A, W, bias, N.EPI_F32_RESID, x, resid=x)  # in place, as the engine uses it
    assert torch.allclose(x, x0 + ref, rtol=1e-4, atol=2e-4)


def test_gemm_argmax_never_stores_logits(device, gen):
    M, V, K = 1499, 9812, 1280
    A = _rand((M, K), gen).bfloat16()
    W = _rand((V, K), gen, 1 / math.sqrt(K)).bfloat16()
    bias = _rand((V,), gen, 0.1)
    keys = torch.zeros((M,), dtype=torch.int64, device=device)
    gemm(A, W, bias, N.EPI_ARGMAX, None, keys=keys, ldo=0)
    got = unpack_keys(keys)
    logits = (A.float() @ W.float().t() + bias).cpu()
    ref = logits.argmax(-1).numpy()
    agree = (got == ref)
    # where the index differs the two logits must be an fp32-summation-order tie
    bad = np.nonzero(~agree)[0]
    for r in bad:
        assert abs(float(logits[r, got[r]] - logits[r, ref[r]])) < 1e-4
    assert agree.mean() > 0.999


def test_gemm_argmax_lowest_index_on_exact_ties(device):
    # identical weight rows -> exactly equal logits: torch.argmax semantics = lowest index
    M, V, K = 200, 520, 64
    A = torch.ones((M, K), device=device).bfloat16()
    W = torch.zeros((V, K), device=device)
    W[7] = 1.0
    W[300] = 1.0
    W[519] = 1.0
    keys = torch.zeros((M,), dtype=torch.int64, device=device)
    gemm(A, W.bfloat16(), torch.zeros(V, device=device), N.EPI_ARGMAX, None, keys=keys, ldo=0)
    assert (unpack_keys(keys) == 7).all()


def test_gemm_ln_gelu(device, gen):
    M, K = 700, 1536
    A = _rand((M, K), gen).bfloat16()
    W = _rand((512, K), gen, 1 / math.sqrt(K)).bfloat16()
    bias, g, b = _rand((512,), gen, 0.3), 1 + _rand((512,), gen, 0.1), _rand((512,), gen, 0.1)
    out = torch.zeros((M, 512), dtype=torch.bfloat16, device=device)
    gemm(A, W, bias, N.EPI_LN_GELU_BF16, out, ln_g=g, ln_b=b)
    ref = F.gelu(F.layer_norm(A.float() @ W.float().t() + bias, (512,), g, b, 1e-5))
    assert torch.allclose(out.float(), ref, rtol=1e-2, atol=1e-2)
    assert rel_err(out, ref) < 4e-3


# ------------------------------------------------------------------------------------------- FE conv layers
@pytest.mark.parametrize("k,L_in,B", [(3, 1001, 2), (2, 600, 3), (3, 4799, 1)])
def test_conv_ln_gelu_implicit_gemm(device, gen, k, L_in, B):
    L_pad = (L_in + 3) & ~1
    x = _rand((B, L_pad, 512), gen).bfloat16()
    w = _rand((512, 512, k), gen, 1 / math.sqrt(512 * k))
    wq = bf16_round(w)
    w_tap = wq.permute(0, 2, 1).reshape(512, k * 512).contiguous().bfloat16()   # [n][j*512 + c]
    bias, g, b = _rand((512,), gen, 0.3), 1 + _rand((512,), gen, 0.1), _rand((512,), gen, 0.1)
    L_out = (L_in - k) // 2 + 1
    out = torch.zeros((B, L_out, 512), dtype=torch.bfloat16, device=device)
    N.check(lib().oasr_conv_ln_gelu(N.ptr(x), B, L_in, L_pad, k, N.ptr(w_tap), N.ptr(bias), N.ptr(g), N.ptr(b),
                                    N.ptr(out), N.stream_ptr()), "conv")
    sync()
    y = F.conv1d(x[:, :L_in].float().transpose(1, 2), wq, bias, stride=2).transpose(1, 2)
    ref = F.gelu(F.layer_norm(y, (512,), g, b, 1e-5))
    assert ref.shape == out.shape
    assert torch.allclose(out.float(), ref, rtol=1e-2, atol=1e-2)
    assert rel_err(out, ref) < 4e-3


def test_fe_layer0(device, gen):
    B, L = 3, 16000
    wave = _rand((B, L), gen)
    w = _rand((512, 1, 10), gen, 0.4)
    bias, g, b = _rand((512,), gen, 0.3), 1 + _rand((512,), gen, 0.1), _rand((512,), gen, 0.1)
    T0 = (L - 10) // 5 + 1
    out = torch.zeros((B, T0, 512), dtype=torch.bfloat16, device=device)
    w_t = w[:, 0, :].t().contiguous()   # [10][512]
    N.check(lib().oasr_fe_layer0(N.ptr(wave), B, L, N.ptr(w_t), N.ptr(bias), N.ptr(g), N.ptr(b), N.ptr(out),
                                 N.stream_ptr()), "fe0")
    sync()
    y = F.conv1d(wave[:, None], w, bias, stride=5).transpose(1, 2)
    ref = F.gelu(F.layer_norm(y, (512,), g, b, 1e-5))
    assert torch.allclose(out.float(), ref, rtol=1e-2, atol=1e-2)   # bf16 output rounding
    assert rel_err(out, ref) < 4e-3


def test_wave_norm(device, gen):
    B, L = 4, 48000
    wave = _rand((B, L), gen, 3.0) + 0.7
    ns = [48000, 31337, 400, 0]
    out = torch.full((B, L), 7.0, device=device)
    nsd = torch.tensor(ns, dtype=torch.int32, device=device)
    N.check(lib().oasr_wave_norm(N.ptr(wave), N.ptr(out), N.ptr(nsd), B, L, N.stream_ptr()), "wave_norm")
    sync()
    ref = O.wave_layer_norm(wave.cpu(), ns)
    assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-5)   # fp32 both sides
    assert (out[1, 31337:] == 0).all() and (out[3] == 0).all()


@pytest.mark.parametrize("D,bf16_in", [(512, True), (1280, False), (2048, False), (256, False), (320, False)])
def test_layernorm(device, gen, D, bf16_in):
    rows = 1001
    x = _rand((rows, D), gen, 2.0) + 0.5
    if bf16_in:
        x = x.bfloat16()
    g, b = 1 + _rand((D,), gen, 0.1), _rand((D,), gen, 0.1)
    ob = torch.zeros((rows, D), dtype=torch.bfloat16, device=device)
    of = torch.zeros((rows, D), dtype=torch.float32, device=device)
    N.check(lib().oasr_layernorm(N.ptr(x), int(bf16_in), rows, D, N.ptr(g), N.ptr(b), N.ptr(ob), N.ptr(of),
                                 N.stream_ptr()), "layernorm")
    sync()
    ref = F.layer_norm(x.float(), (D,), g, b, 1e-5)
    assert torch.allclose(of, ref, rtol=1e-5, atol=2e-5)            # fp32
    assert torch.allclose(ob.float(), ref, rtol=1e-2, atol=1e-2)    # bf16 rounding


# ------------------------------------------------------------------------------------------- pos-conv
@pytest.mark.parametrize("d,groups,k,T,B", [(256, 16, 128, 300, 2), (320, 4, 128, 200, 2), (1280, 16, 128, 150, 1),
                                            (1024, 16, 128, 130, 1), (2048, 16, 128, 129, 1)])
def test_posconv(device, gen, d, groups, k, T, B):
    cg = d // groups
    x = _rand((B * T, d), gen)
    w = _rand((d, cg, k), gen, 1 / math.sqrt(cg * k))
    bias = _rand((d,), gen, 0.2)
    wq = bf16_round(w)
    k_pad = ((cg + 63) // 64) * 64
    wt = torch.zeros((d, k, k_pad), device=device)
    wt[:, :, :cg] = wq.permute(0, 2, 1)
    wt = wt.reshape(d, k * k_pad).contiguous().bfloat16()
    scratch = torch.zeros((B, T + k, d), dtype=torch.bfloat16, device=device)
    x_in = x.clone()
    N.check(lib().oasr_posconv(N.ptr(x), B, T, d, groups, k, N.ptr(wt), N.ptr(bias), N.ptr(scratch), N.stream_ptr()),
            "posconv")
    sync()
    xin = bf16_round(x_in).view(B, T, d).transpose(1, 2)
    y = F.conv1d(xin, wq, bias, padding=k // 2, groups=groups)[..., :-1]
    ref = x_in.view(B, T, d) + F.gelu(y).transpose(1, 2)
    assert torch.allclose(x.view(B, T, d), ref, rtol=1e-3, atol=1e-3)   # fp32 accumulate, order differs


# ------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("H,hd,T,B,ragged", [(4, 64, 300, 2, False), (4, 80, 300, 2, True), (2, 128, 257, 2, True),
                                             (16, 80, 1499, 1, False), (3, 16, 100, 1, False),
                                             (2, 96, 400, 2, True), (2, 112, 333, 2, True), (2, 128, 1499, 1, False),
                                             (2, 32, 129, 1, False), (2, 48, 515, 2, True),
                                             # >= 148 work items: the three-tile persistent shape (fewer: one tile per CTA)
                                             (16, 80, 500, 5, True), (16, 64, 400, 10, True)])
def test_attention(device, gen, H, hd, T, B, ragged):
    d = H * hd
    qkv = _rand((B * T, 3 * d), gen).bfloat16()
    nf = [T] * B
    if ragged:
        nf[-1] = T // 2 + 3
    nfd = torch.tensor(nf, dtype=torch.int32, device=device)
    out = torch.zeros((B * T, d), dtype=torch.bfloat16, device=device)
    scale = hd ** -0.5
    N.check(lib().oasr_attention(N.ptr(qkv), N.ptr(out), N.ptr(nfd), B, T, H, hd, scale, N.stream_ptr()), "attention")
    sync()
    q, k, v = qkv.float().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) * scale
    mask = torch.arange(T, device=device)[None, :] >= nfd[:, None]
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    sl = s * 1.4426950408889634
    m = torch.ceil(sl.amax(-1, keepdim=True))                # integer reference in the log2 domain
    p = torch.exp2(sl - m)
    ref = (bf16_round(p) @ v) / p.sum(-1, keepdim=True)      # the engine's numerics contract
    ref = ref.permute(0, 2, 1, 3).reshape(B * T, d)
    o = out.float().view(B, T, d)
    r = ref.view(B, T, d)
    for b in range(B):   # padded query rows are don't-care
        assert torch.allclose(o[b, :nf[b]], r[b, :nf[b]], rtol=1e-2, atol=1e-2)
        assert rel_err(o[b, :nf[b]], r[b, :nf[b]]) < 5e-3
    # and against the textbook softmax (P not rounded): bf16-level agreement
    ref2 = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B, T, d)
    assert rel_err(o[0], ref2[0]) < 1e-2


@pytest.mark.parametrize("hd,top,H,T,B", [(64, 6.0, 2, 700, 1), (64, 30.0, 2, 700, 1), (80, 30.0, 2, 700, 1),
                                          (128, 24.0, 2, 700, 1),
                                          # >= 148 work items: the three-tile shape with its relay (a block that is
                                          # computed twice must hand over only once)
                                          (80, 30.0, 16, 500, 5)])
def test_attention_growing_scores_take_the_rescale_path(device, gen, hd, top, H, T, B):
    """Keys whose scores grow along the sequence push the running reference up several times (top = 30: by far more
    than the 2^80 margin, within a block and across blocks)."""
    d = H * hd
    qkv = _rand((B * T, 3 * d), gen, 0.5)
    ramp = torch.linspace(0.0, top, T, device=device).repeat(B)
    qkv[:, :d] = 1.0 + 0.1 * qkv[:, :d]                       # q ~ all ones
    qkv[:, d:2 * d] = ramp[:, None] + 0.1 * qkv[:, d:2 * d]   # k grows with the position
    qkv = qkv.bfloat16()
    out = torch.zeros((B * T, d), dtype=torch.bfloat16, device=device)
    nfd = torch.tensor([T] * B, dtype=torch.int32, device=device)
    scale = hd ** -0.5
    N.check(lib().oasr_attention(N.ptr(qkv), N.ptr(out), N.ptr(nfd), B, T, H, hd, scale, N.stream_ptr()), "attention")
    sync()
    q, k, v = qkv.float().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) * scale
    assert float(s.amax() - s[..., :128].amax()) * 1.4427 > 20      # the reference really has to move
    sl = s * 1.4426950408889634
    m = torch.ceil(sl.amax(-1, keepdim=True))
    p = torch.exp2(sl - m)
    ref = ((bf16_round(p) @ v) / p.sum(-1, keepdim=True)).permute(0, 2, 1, 3).reshape(B * T, d)
    assert torch.allclose(out.float(), ref, rtol=1e-2, atol=1e-2)
    assert rel_err(out, ref) < 5e-3


def test_attention_fully_padded_window_is_zero(device, gen):
    H, hd, T, B = 2, 64, 140, 2
    qkv = _rand((B * T, 3 * H * hd), gen).bfloat16()
    nfd = torch.tensor([T, 0], dtype=torch.int32, device=device)
    out = torch.full((B * T, H * hd), 5.0, dtype=torch.bfloat16, device=device)
    N.check(lib().oasr_attention(N.ptr(qkv), N.ptr(out), N.ptr(nfd), B, T, H, hd, 0.125, N.stream_ptr()), "attention")
    sync()
    assert (out.view(B, T, -1)[1] == 0).all()


# ------------------------------------------------------------------------------------------- decode
def test_ctc_collapse_kats_and_random(device):
    blank = 0
    cases = [([0, 0, 5, 5, 0, 5, 7, 7, 7, 0], [5, 5, 7]), ([0] * 10, []), ([1, 2, 3, 4, 5, 6, 7, 8, 9, 10], list(range(1, 11)))]
    T = 10
    ids = torch.tensor([c[0] for c in cases], dtype=torch.int32, device=device)
    nf = torch.tensor([T] * len(cases), dtype=torch.int32, device=device)
    oi = torch.zeros_like(ids)
    of = torch.zeros_like(ids)
    ol = torch.zeros(len(cases), dtype=torch.int32, device=device)
    N.check(lib().oasr_ctc_collapse(N.ptr(ids), N.ptr(nf), len(cases), T, blank, N.ptr(oi), N.ptr(of), N.ptr(ol),
                                    N.stream_ptr()), "collapse")
    sync()
    for i, (_, want) in enumerate(cases):
        n = int(ol[i])
        assert oi[i, :n].tolist() == want
    # random ragged batch at the full window size, bit-exact against the oracle
    rng = np.random.default_rng(0)
    B, T = 33, 1499
    ids = rng.integers(0, 4, size=(B, T)).astype(np.int32)
    nfr = rng.integers(0, T + 1, size=B).astype(np.int32)
    nfr[0], nfr[1] = T, 0
    want_ids, want_pos, want_len = O.collapse_batch(ids, nfr, blank)
    idd = torch.from_numpy(ids).to(device)
    oi = torch.zeros_like(idd)
    of = torch.zeros_like(idd)
    ol = torch.zeros(B, dtype=torch.int32, device=device)
    N.check(lib().oasr_ctc_collapse(N.ptr(idd), N.ptr(torch.from_numpy(nfr).to(device)), B, T, blank, N.ptr(oi),
                                    N.ptr(of), N.ptr(ol), N.stream_ptr()), "collapse")
    sync()
    assert (ol.cpu().numpy() == want_len).all()
    assert (oi.cpu().numpy() == want_ids).all()
    assert (of.cpu().numpy() == want_pos).all()


def _attention_variant(tmp_path, env, B, T, H, hd):
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    out = tmp_path / ("att_" + "_".join(f"{k}{v}" for k, v in env.items()) + f"_{hd}.npy")
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, str(root / "scripts" / "attention_variant.py"), str(out), str(B), str(T), str(H), str(hd)],
                       env=e, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


def test_attention_relay_switch(device, tmp_path):
    """OASR_ATT_RELAY (exponential phases of the tiles kept apart by named barriers: 0 free-running, 1, 2 default):
    scheduling only - the outputs are bit-identical."""
    # 5 windows x 16 heads x 2 query triples = 160 work items: at least one per SM, so the three-tile shape (the one with
    # a relay) runs, not the one-tile shape of small batches
    ref = _attention_variant(tmp_path, {"OASR_ATT_RELAY": "0"}, 5, 500, 16, 80)
    for r in ("1", "2"):
        assert (_attention_variant(tmp_path, {"OASR_ATT_RELAY": r}, 5, 500, 16, 80) == ref).all()


def test_attention_env_switch_v4(device, tmp_path):
    """OASR_ATTN=4 forces the two-tile kernel (the default above head_dim 80) at head_dim 80: same result as v7."""
    a = _attention_variant(tmp_path, {"OASR_ATTN": "7"}, 2, 500, 4, 80)
    b = _attention_variant(tmp_path, {"OASR_ATTN": "4"}, 2, 500, 4, 80)
    assert float((a == b).mean()) > 0.995 and float(np.abs(a - b).max()) < 2e-2 * float(np.abs(a).max())


# ------------------------------------------------------------------------------------------- exactness
def test_outputs_bit_identical_to_rounded_reference(device, gen):
    """Each bf16-producing kernel must equal 'fp64 reference rounded once to bf16' on >= 99.5 % of its outputs
    (the remainder are fp32 summation-order flips of a rounding boundary); fp32 outputs agree to ~1e-6."""
    def same(out, ref64):
        return float((out.float() == ref64.float().bfloat16().float()).float().mean())

    # FE layer 0 (fp32 CUDA-core conv + LN + GELU)
    B, L = 2, 16000
    wave = _rand((B, L), gen)
    w = _rand((512, 1, 10), gen, 0.4)
    bias, g, b = _rand((512,), gen, 0.3), 1 + _rand((512,), gen, 0.1), _rand((512,), gen, 0.1)
    T0 = (L - 10) // 5 + 1
    out = torch.zeros((B, T0, 512), dtype=torch.bfloat16, device=device)
    N.check(lib().oasr_fe_layer0(N.ptr(wave), B, L, N.ptr(w[:, 0, :].t().contiguous()), N.ptr(bias), N.ptr(g), N.ptr(b),
                                 N.ptr(out), N.stream_ptr()), "fe0")
    sync()
    ref = F.gelu(F.layer_norm(F.conv1d(wave[:, None].double(), w.double(), bias.double(), stride=5).transpose(1, 2),
                              (512,), g.double(), b.double(), 1e-5))
    assert same(out, ref) > 0.995

    # tcgen05 GEMM with fused epilogues
    M, Nn, K = 1000, 1280, 1280
    A = _rand((M, K), gen).bfloat16()
    W = _rand((Nn, K), gen, 1 / math.sqrt(K)).bfloat16()
    bv = _rand((Nn,), gen)
    acc = A.double() @ W.double().t() + bv.double()
    for epi, fn in ((N.EPI_BF16, lambda t: t), (N.EPI_BF16_GELU, F.gelu)):
        o = torch.zeros((M, Nn), dtype=torch.bfloat16, device=device)
        gemm(A, W, bv, epi, o)
        assert same(o, fn(acc)) > 0.995
    o = torch.zeros((M, Nn), dtype=torch.float32, device=device)
    gemm(A, W, bv, N.EPI_F32, o)
    assert rel_err(o, acc.float()) < 5e-6

    # implicit-GEMM conv with the LayerNorm+GELU epilogue
    A5 = _rand((M, 1536), gen).bfloat16()
    W5 = _rand((512, 1536), gen, 1 / math.sqrt(1536)).bfloat16()
    o = torch.zeros((M, 512), dtype=torch.bfloat16, device=device)
    gemm(A5, W5, bias, N.EPI_LN_GELU_BF16, o, ln_g=g, ln_b=b)
    ref = F.gelu(F.layer_norm(A5.double() @ W5.double().t() + bias.double(), (512,), g.double(), b.double(), 1e-5))
    assert same(o, ref) > 0.995

    # attention (single pass, integer log2 reference)
    for H, hd in ((4, 64), (4, 80), (2, 128)):
        T, Bq = 300, 2
        d = H * hd
        qkv = _rand((Bq * T, 3 * d), gen).bfloat16()
        nf = torch.full((Bq,), T, dtype=torch.int32, device=device)
        o = torch.zeros((Bq * T, d), dtype=torch.bfloat16, device=device)
        N.check(lib().oasr_attention(N.ptr(qkv), N.ptr(o), N.ptr(nf), Bq, T, H, hd, hd ** -0.5, N.stream_ptr()), "attn")
        sync()
        q, k, v = qkv.double().view(Bq, T, 3, H, hd).permute(2, 0, 3, 1, 4)
        sl = (q @ k.transpose(-1, -2)) * (hd ** -0.5) * 1.4426950408889634
        p = torch.exp2(sl - torch.ceil(sl.amax(-1, keepdim=True)))
        ref = ((p.float().bfloat16().double() @ v) / p.sum(-1, keepdim=True)).permute(0, 2, 1, 3).reshape(Bq * T, d)
        assert same(o, ref) > 0.995


# ------------------------------------------------------------------------------------------- device-side front end
@pytest.mark.parametrize("sr_in,channels,pcm16", [(22050, 1, False), (44100, 2, True), (48000, 1, True), (8000, 1, False),
                                                  (11025, 2, False), (16000, 2, True)])
def test_resample_matches_torchaudio(device, gen, sr_in, channels, pcm16):
    """oasr_resample (channel mean + polyphase windowed sinc) against torchaudio.functional.resample's defaults on the
    CPU, the filter the host path applies (audio.to_mono_16k); fp32 accumulation order differs: 5e-5 absolute."""
    import torchaudio.functional as AF
    n = 50_000 + 37
    x = torch.randn(n, channels, generator=torch.Generator().manual_seed(5)) * 0.3
    if pcm16:
        xi = (x * 32768.0).clamp(-32768, 32767).to(torch.int16)
        xf = xi.float() / 32768.0
        src = xi
    else:
        xf = x
        src = x
    want = AF.resample(xf.mean(dim=1), sr_in, 16000)
    n_out = int(lib().oasr_resample_length(n, sr_in, 16000))
    assert n_out == want.numel()
    d_in = src.contiguous().to(device)
    out = torch.empty(n_out, dtype=torch.float32, device=device)
    N.check(lib().oasr_resample(N.ptr(d_in), 1 if pcm16 else 0, n, channels, sr_in, 16000, N.ptr(out), n_out,
                                N.stream_ptr()), "oasr_resample")
    sync()
    assert torch.allclose(out.cpu(), want, atol=5e-5, rtol=0)   # fp32 taps and accumulation on both sides
