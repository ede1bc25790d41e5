"""CPU oracle for the omniASR CTC inference path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED BY THE REFERENCE: /root/reference (Nathan-Roll1/omnilingual-asr)
ships no local CTC code, no tests and no golden vectors for this path
(SURVEY.md F1/F3, section 8c).  The arithmetic lives upstream in
facebookresearch/omnilingual-asr (models/inference/pipeline.py ::
ASRInferencePipeline) and facebookresearch/fairseq2 (fairseq2.models.wav2vec2,
fairseq2.models.wav2vec2.asr), neither of which is installable here; the last
pin visible in the reference is CI's `fairseq2[arrow]` nightly for pt2.5.1/cpu
(.github/workflows/lint_and_test.yaml:34-40).  This file restates that published
algorithm in plain PyTorch fp32 on the CPU.  It is pinned two ways instead:
  * parameter totals equal the four published omniASR-CTC counts exactly
    (tests/test_oracle.py::test_param_counts), which fixes the architecture;
  * outputs equal an independent implementation of the same graph,
    transformers' Wav2Vec2ForCTC (modeling_wav2vec2.py:275-299, 326-379,
    422-435, 612-655, 730-803, 1697-1710) to fp32 round-off, both live
    (tests/test_oracle.py) and through fixtures in tests/golden/ written by
    tests/golden/make_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product path
(omnilingual-asr_b200/) never does; it fails loudly when the CUDA library is
missing.

Stage map:
  a8  wave_layer_norm      F.layer_norm(x, x.shape) per window, eps 1e-5
  a9  FE layer 0           Conv1d(1,512,k10,s5,bias) -> LN(512) -> GELU(erf)
  a10 FE layers 1-4        Conv1d(512,512,k3,s2,bias) -> LN -> GELU
  a11 FE layers 5-6        Conv1d(512,512,k2,s2,bias) -> LN -> GELU
  a12 feature projection   LN(512) -> Linear(512,d)
  a13 positional conv      zero pads; weight-normed grouped Conv1d(d,d,k128,
                           pad64,groups16); drop last frame; GELU; x + .
  a14 encoder              pre-LN blocks, bias everywhere, scale hd^-0.5,
                           bidirectional, key-padding mask, final LN
  a15 CTC head             Linear(d,V) -> argmax (lowest index on ties)
  a16 greedy collapse      drop repeats, then drop blank (0)

`emulate_bf16=True` is the "fp32-accum check" numerics: every tensor-core
operand (weights and the activation feeding each contraction, and the softmax
probabilities feeding P.V) is rounded to bf16 exactly where the CUDA engine
rounds it, everything else (accumulation, LayerNorm, softmax, GELU, residual
stream) stays fp32.  See DESIGN.md "Numerics contract".
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SAMPLE_RATE = 16_000
LN_EPS = 1e-5
BLANK_ID = 0
LOG2E = 1.4426950408889634
# (out_channels, kernel, stride) of the 7 feature-extractor layers (a9-a11).
FE_LAYERS: Tuple[Tuple[int, int, int], ...] = (
    (512, 10, 5), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 2, 2), (512, 2, 2))


@dataclass(frozen=True)
class CtcModelConfig:
    """Architecture of one omniASR CTC model."""
    name: str
    d_model: int
    n_layers: int
    n_heads: int
    d_ffn: int
    vocab: int = 9812
    fe_dim: int = 512
    pos_kernel: int = 128
    pos_groups: int = 16
    fe_layers: Tuple[Tuple[int, int, int], ...] = FE_LAYERS

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_heads


PRESETS: Dict[str, CtcModelConfig] = {
    "omniASR_CTC_300M": CtcModelConfig("omniASR_CTC_300M", 1024, 24, 16, 4096),
    "omniASR_CTC_1B": CtcModelConfig("omniASR_CTC_1B", 1280, 48, 16, 5120),
    "omniASR_CTC_3B": CtcModelConfig("omniASR_CTC_3B", 2048, 60, 16, 8192),
    "omniASR_CTC_7B": CtcModelConfig("omniASR_CTC_7B", 2048, 128, 16, 8192),
    # Small shapes for fast parity tests; same graph, legal tensor-core shapes.
    "tiny": CtcModelConfig("tiny", 256, 2, 4, 512, vocab=300),
    "tiny80": CtcModelConfig("tiny80", 320, 2, 4, 640, vocab=500, pos_groups=4),   # head_dim 80, group width 80 (1B-like)
    # 3B / 7B widths (d 2048, head_dim 128, FFN 8192, group width 128) with two layers: the shapes of configs 3 and 4
    "wide2l": CtcModelConfig("wide2l", 2048, 2, 16, 8192, vocab=9812),
}

PUBLISHED_PARAM_COUNTS = {
    "omniASR_CTC_300M": 325_494_996,
    "omniASR_CTC_1B": 975_065_300,
    "omniASR_CTC_3B": 3_080_423_636,
    "omniASR_CTC_7B": 6_504_786_132,
}


# ----------------------------------------------------------------------------
# lengths and chunking
# ----------------------------------------------------------------------------
def feature_length(n_samples: int, cfg: Optional[CtcModelConfig] = None) -> int:
    """Frames produced by the conv feature extractor: chain of floor((L-k)/s)+1.

    fairseq2 Wav2Vec2FeatureExtractor length rule; same as
    transformers modeling_wav2vec2.py:1005-1025.
    """
    layers = cfg.fe_layers if cfg is not None else FE_LAYERS
    n = int(n_samples)
    for _, k, s in layers:
        n = (n - k) // s + 1 if n >= k else 0
    return max(n, 0)


def split_into_windows(n_samples: int, window: int) -> List[Tuple[int, int]]:
    """Fixed non-overlapping windows (start, length); the last one may be short.

    Restates split_audio_into_chunks (gemini_pipeline.py:243-310): start = i*chunk,
    no overlap, last chunk short, at least one chunk.
    """
    if n_samples <= 0 or window <= 0:
        return [(0, max(n_samples, 0))]
    out = []
    start = 0
    while start < n_samples:
        out.append((start, min(window, n_samples - start)))
        start += window
    return out


def wave_layer_norm(wave: torch.Tensor, n_samples: Sequence[int]) -> torch.Tensor:
    """a8: per-window zero-mean/unit-variance over the valid samples, pads stay 0.

    upstream ASRInferencePipeline: `layer_norm(wav, wav.shape)` on each window.
    """
    out = torch.zeros_like(wave, dtype=torch.float32)
    for b, n in enumerate(n_samples):
        n = int(n)
        if n > 0:
            out[b, :n] = F.layer_norm(wave[b, :n].float(), (n,), eps=LN_EPS)
    return out


# ----------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------
def weight_shapes(cfg: CtcModelConfig) -> Dict[str, Tuple[int, ...]]:
    """Name -> shape of every parameter, in load order."""
    d, f = cfg.d_model, cfg.d_ffn
    shapes: Dict[str, Tuple[int, ...]] = {}
    c_in = 1
    for i, (c, k, _) in enumerate(cfg.fe_layers):
        shapes[f"fe.{i}.conv.weight"] = (c, c_in, k)
        shapes[f"fe.{i}.conv.bias"] = (c,)
        shapes[f"fe.{i}.ln.weight"] = (c,)
        shapes[f"fe.{i}.ln.bias"] = (c,)
        c_in = c
    shapes["proj.ln.weight"] = (cfg.fe_dim,)
    shapes["proj.ln.bias"] = (cfg.fe_dim,)
    shapes["proj.linear.weight"] = (d, cfg.fe_dim)
    shapes["proj.linear.bias"] = (d,)
    shapes["pos.weight_g"] = (1, 1, cfg.pos_kernel)
    shapes["pos.weight_v"] = (d, d // cfg.pos_groups, cfg.pos_kernel)
    shapes["pos.bias"] = (d,)
    for l in range(cfg.n_layers):
        p = f"enc.{l}."
        shapes[p + "attn_ln.weight"] = (d,)
        shapes[p + "attn_ln.bias"] = (d,)
        for n in ("q", "k", "v", "o"):
            shapes[p + f"{n}.weight"] = (d, d)
            shapes[p + f"{n}.bias"] = (d,)
        shapes[p + "ffn_ln.weight"] = (d,)
        shapes[p + "ffn_ln.bias"] = (d,)
        shapes[p + "ffn1.weight"] = (f, d)
        shapes[p + "ffn1.bias"] = (f,)
        shapes[p + "ffn2.weight"] = (d, f)
        shapes[p + "ffn2.bias"] = (d,)
    shapes["final_ln.weight"] = (d,)
    shapes["final_ln.bias"] = (d,)
    shapes["ctc.weight"] = (cfg.vocab, d)
    shapes["ctc.bias"] = (cfg.vocab,)
    return shapes


def param_count(cfg: CtcModelConfig) -> int:
    return sum(int(np.prod(s)) for s in weight_shapes(cfg).values())


def init_weights(cfg: CtcModelConfig, seed: int = 0, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """Deterministic random init (documented law; there are no checkpoints offline).

    Law: weights ~ N(0, sigma^2) with sigma = gain/sqrt(fan_in) (gain sqrt(2) before
    a GELU, else 1); LN weight ~ 1 + 0.1 N(0,1), LN bias and linear/conv bias ~ 0.1 N(0,1)
    (non-trivial so that bias/affine bugs show up); pos-conv g ~ |N(1, 0.1)| * ||v||-scale.
    One generator, fixed parameter order => identical on every machine.
    """
    gen = torch.Generator(device="cpu").manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in weight_shapes(cfg).items():
        if name.endswith("ln.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=gen)
        elif name.endswith("bias"):
            t = 0.1 * torch.randn(shape, generator=gen)
        elif name == "pos.weight_g":
            t = (1.0 + 0.1 * torch.randn(shape, generator=gen)).abs()
        elif name == "pos.weight_v":
            fan_in = shape[1] * shape[2]
            t = torch.randn(shape, generator=gen) * math.sqrt(2.0 / fan_in)
        else:
            fan_in = int(np.prod(shape[1:]))
            pre_gelu = name.startswith("fe.") or name.endswith("ffn1.weight")
            t = torch.randn(shape, generator=gen) * math.sqrt((2.0 if pre_gelu else 1.0) / fan_in)
        out[name] = t.float().to(device)
    # make g comparable to ||v|| per tap so the folded weight keeps the variance law
    v = out["pos.weight_v"]
    out["pos.weight_g"] = out["pos.weight_g"] * v.norm(dim=(0, 1), keepdim=True)
    return out


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """w = g * v / ||v||, norm over dims (0,1) per tap (weight_norm(dim=2))."""
    return g * v / v.norm(dim=(0, 1), keepdim=True)


# ----------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------
def _q(x: torch.Tensor, on: bool) -> torch.Tensor:
    """Round to bf16 and back when emulating tensor-core operand precision."""
    return x.to(torch.bfloat16).to(torch.float32) if on else x


@dataclass
class OracleOutput:
    frame_ids: torch.Tensor              # [B, Tmax] int64, padded frames = blank
    n_frames: List[int]                  # valid frames per window
    hidden: torch.Tensor                 # [B, Tmax, d] final-LN output (fp32)
    logits: Optional[torch.Tensor] = None  # [B, Tmax, V] when requested
    taps: Dict[str, torch.Tensor] = field(default_factory=dict)


def feature_extractor(w: Dict[str, torch.Tensor], wave: torch.Tensor, cfg: CtcModelConfig,
                      emulate_bf16: bool = False, upto: Optional[int] = None) -> torch.Tensor:
    """a9-a11: [B, L] -> [B, T, 512] channels-last.  LN over channels per frame, fp32."""
    x = wave.float().unsqueeze(1)                       # [B, 1, L]
    for i, (c, k, s) in enumerate(cfg.fe_layers):
        wt = w[f"fe.{i}.conv.weight"]
        if i > 0:                                       # layer 0 runs in fp32 on CUDA cores
            wt = _q(wt, emulate_bf16)
            x = _q(x, emulate_bf16)
        x = F.conv1d(x, wt, w[f"fe.{i}.conv.bias"], stride=s)
        x = x.transpose(1, 2)
        x = F.layer_norm(x, (c,), w[f"fe.{i}.ln.weight"], w[f"fe.{i}.ln.bias"], LN_EPS)
        x = F.gelu(x)                                   # erf GELU
        x = x.transpose(1, 2)
        if upto is not None and i == upto:
            break
    return x.transpose(1, 2).contiguous()


def _row_parallel_linear(x: torch.Tensor, a: torch.Tensor, wt: torch.Tensor, bias: torch.Tensor, tp_world: int) -> torch.Tensor:
    """Tensor-parallel row-parallel linear as the engine computes it (engine.cu, tp_fused.cu): rank r multiplies its
    slice of the input columns (rank 0 adds the bias), rounds its PARTIAL SUM to bf16 - the partial sums cross NVLink
    in bf16 - and the partials are added to the fp32 residual row x in rank order.  Returns the new x."""
    K = a.shape[-1]
    step = K // tp_world
    out = x
    for r in range(tp_world):
        part = F.linear(a[..., r * step:(r + 1) * step], wt[:, r * step:(r + 1) * step], bias if r == 0 else None)
        out = out + _q(part, True)
    return out


def forward(w: Dict[str, torch.Tensor], wave: torch.Tensor, n_samples: Sequence[int],
            cfg: CtcModelConfig, *, emulate_bf16: bool = False, return_logits: bool = False,
            taps: bool = False, tp_world: int = 1) -> OracleOutput:
    """Full path a9-a15 on already-normalised, zero-padded windows `wave` [B, L].

    tp_world > 1 (with emulate_bf16): the rounding points of the tensor-parallel engine (BASELINE config 4) - the
    out-proj and FFN2 partial sums of each rank are rounded to bf16 before they are added (_row_parallel_linear)."""
    q = emulate_bf16
    if tp_world > 1 and not q:
        raise ValueError("tp_world only changes the bf16-operand emulation")
    B = wave.shape[0]
    d, H, hd = cfg.d_model, cfg.n_heads, cfg.head_dim
    n_frames = [feature_length(int(n), cfg) for n in n_samples]
    tp: Dict[str, torch.Tensor] = {}

    feats = _q(feature_extractor(w, wave, cfg, q), q)   # [B, T, 512]; the engine keeps FE activations in bf16
    T = feats.shape[1]
    if taps:
        tp["fe"] = feats.clone()
    valid = torch.arange(T)[None, :] < torch.tensor(n_frames)[:, None]   # [B, T]

    # a12 feature projection
    x = F.layer_norm(feats, (cfg.fe_dim,), w["proj.ln.weight"], w["proj.ln.bias"], LN_EPS)
    x = F.linear(_q(x, q), _q(w["proj.linear.weight"], q), w["proj.linear.bias"])
    x = x * valid[..., None]                            # zero padded frames
    if taps:
        tp["proj"] = x.clone()

    # a13 positional conv (+ residual)
    wpos = fold_weight_norm(w["pos.weight_g"], w["pos.weight_v"])
    y = F.conv1d(_q(x, q).transpose(1, 2), _q(wpos, q), w["pos.bias"],
                 padding=cfg.pos_kernel // 2, groups=cfg.pos_groups)
    y = y[..., :-1] if cfg.pos_kernel % 2 == 0 else y   # SamePad: drop the last frame
    x = x + F.gelu(y).transpose(1, 2)
    if taps:
        tp["pos"] = x.clone()

    # a14 encoder
    key_bias = torch.zeros(B, 1, 1, T)
    key_bias.masked_fill_(~valid[:, None, None, :], float("-inf"))
    scale = hd ** -0.5
    for l in range(cfg.n_layers):
        p = f"enc.{l}."
        h = _q(F.layer_norm(x, (d,), w[p + "attn_ln.weight"], w[p + "attn_ln.bias"], LN_EPS), q)
        qh = _q(F.linear(h, _q(w[p + "q.weight"], q), w[p + "q.bias"]), q)
        kh = _q(F.linear(h, _q(w[p + "k.weight"], q), w[p + "k.bias"]), q)
        vh = _q(F.linear(h, _q(w[p + "v.weight"], q), w[p + "v.bias"]), q)
        qh = qh.view(B, T, H, hd).transpose(1, 2)
        kh = kh.view(B, T, H, hd).transpose(1, 2)
        vh = vh.view(B, T, H, hd).transpose(1, 2)
        s = torch.matmul(qh, kh.transpose(-1, -2)) * scale + key_bias
        if q:
            # engine (attention_v2.cu): P = bf16(2^(s*c - m_ref)) feeds P.V, with c = scale*log2(e) and an
            # INTEGER reference m_ref = ceil(rowmax * c) (any integer gives the same bf16 mantissas); the row
            # sum is taken over the unrounded fp32 P.
            sl = s * LOG2E
            m = torch.ceil(sl.amax(dim=-1, keepdim=True))
            m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
            pr = torch.exp2(sl - m)
            den = pr.sum(dim=-1, keepdim=True)
            a = torch.matmul(_q(pr, True), vh) / den
        else:
            a = torch.matmul(torch.softmax(s, dim=-1), vh)
        a = torch.nan_to_num(a)                         # fully padded windows
        a = _q(a.transpose(1, 2).reshape(B, T, d), q)
        if tp_world > 1:
            x = _row_parallel_linear(x, a, _q(w[p + "o.weight"], q), w[p + "o.bias"], tp_world)
        else:
            x = x + F.linear(a, _q(w[p + "o.weight"], q), w[p + "o.bias"])
        h = _q(F.layer_norm(x, (d,), w[p + "ffn_ln.weight"], w[p + "ffn_ln.bias"], LN_EPS), q)
        h = _q(F.gelu(F.linear(h, _q(w[p + "ffn1.weight"], q), w[p + "ffn1.bias"])), q)
        if tp_world > 1:
            x = _row_parallel_linear(x, h, _q(w[p + "ffn2.weight"], q), w[p + "ffn2.bias"], tp_world)
        else:
            x = x + F.linear(h, _q(w[p + "ffn2.weight"], q), w[p + "ffn2.bias"])
        if taps:
            tp[f"enc.{l}"] = x.clone()
    hidden = F.layer_norm(x, (d,), w["final_ln.weight"], w["final_ln.bias"], LN_EPS)

    # a15 CTC head + argmax
    logits = F.linear(_q(hidden, q), _q(w["ctc.weight"], q), w["ctc.bias"])
    ids = torch.argmax(logits, dim=-1)
    ids = ids * valid                                   # padded frames -> blank (0)
    return OracleOutput(ids, n_frames, hidden, logits if return_logits else None, tp)


def top2_margin(logits: torch.Tensor) -> torch.Tensor:
    """Per-frame gap between the best and second-best logit (for agreement reports)."""
    t2 = torch.topk(logits, 2, dim=-1).values
    return t2[..., 0] - t2[..., 1]


# ----------------------------------------------------------------------------
# decode
# ----------------------------------------------------------------------------
def greedy_collapse(frame_ids: Sequence[int], n_frames: Optional[int] = None,
                    blank: int = BLANK_ID) -> Tuple[List[int], List[int]]:
    """a16: keep ids[t] iff t==0 or ids[t]!=ids[t-1]; then drop blank.

    upstream: mask = seq[1:] != seq[:-1] then tokenizer decode with
    skip_special_tokens (which removes blank).  Returns (ids, first-frame index of each).
    """
    ids = [int(v) for v in (frame_ids[:n_frames] if n_frames is not None else frame_ids)]
    out, pos = [], []
    prev = None
    for t, v in enumerate(ids):
        if prev is None or v != prev:
            if v != blank:
                out.append(v)
                pos.append(t)
        prev = v
    return out, pos


def collapse_batch(frame_ids: np.ndarray, n_frames: Sequence[int], blank: int = BLANK_ID):
    """Vectorised a16 over a [B, T] array -> (ids [B,T] padded with -1, frame idx, lens)."""
    ids = np.asarray(frame_ids)
    B, T = ids.shape
    out_ids = np.full((B, T), -1, dtype=np.int32)
    out_pos = np.full((B, T), -1, dtype=np.int32)
    lens = np.zeros(B, dtype=np.int32)
    for b in range(B):
        n = int(n_frames[b])
        row = ids[b, :n]
        if n == 0:
            continue
        keep = np.ones(n, dtype=bool)
        keep[1:] = row[1:] != row[:-1]
        keep &= row != blank
        idx = np.nonzero(keep)[0]
        lens[b] = len(idx)
        out_ids[b, :len(idx)] = row[idx]
        out_pos[b, :len(idx)] = idx
    return out_ids, out_pos, lens


# ----------------------------------------------------------------------------
# cross-check plumbing: same weights inside transformers' Wav2Vec2ForCTC
# ----------------------------------------------------------------------------
def to_hf_state_dict(w: Dict[str, torch.Tensor], cfg: CtcModelConfig) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    for i in range(len(cfg.fe_layers)):
        b = f"wav2vec2.feature_extractor.conv_layers.{i}."
        sd[b + "conv.weight"] = w[f"fe.{i}.conv.weight"]
        sd[b + "conv.bias"] = w[f"fe.{i}.conv.bias"]
        sd[b + "layer_norm.weight"] = w[f"fe.{i}.ln.weight"]
        sd[b + "layer_norm.bias"] = w[f"fe.{i}.ln.bias"]
    sd["wav2vec2.feature_projection.layer_norm.weight"] = w["proj.ln.weight"]
    sd["wav2vec2.feature_projection.layer_norm.bias"] = w["proj.ln.bias"]
    sd["wav2vec2.feature_projection.projection.weight"] = w["proj.linear.weight"]
    sd["wav2vec2.feature_projection.projection.bias"] = w["proj.linear.bias"]
    pc = "wav2vec2.encoder.pos_conv_embed.conv."
    sd[pc + "parametrizations.weight.original0"] = w["pos.weight_g"]
    sd[pc + "parametrizations.weight.original1"] = w["pos.weight_v"]
    sd[pc + "bias"] = w["pos.bias"]
    for l in range(cfg.n_layers):
        p, b = f"enc.{l}.", f"wav2vec2.encoder.layers.{l}."
        sd[b + "layer_norm.weight"] = w[p + "attn_ln.weight"]
        sd[b + "layer_norm.bias"] = w[p + "attn_ln.bias"]
        for n, hn in (("q", "q_proj"), ("k", "k_proj"), ("v", "v_proj"), ("o", "out_proj")):
            sd[b + f"attention.{hn}.weight"] = w[p + f"{n}.weight"]
            sd[b + f"attention.{hn}.bias"] = w[p + f"{n}.bias"]
        sd[b + "final_layer_norm.weight"] = w[p + "ffn_ln.weight"]
        sd[b + "final_layer_norm.bias"] = w[p + "ffn_ln.bias"]
        sd[b + "feed_forward.intermediate_dense.weight"] = w[p + "ffn1.weight"]
        sd[b + "feed_forward.intermediate_dense.bias"] = w[p + "ffn1.bias"]
        sd[b + "feed_forward.output_dense.weight"] = w[p + "ffn2.weight"]
        sd[b + "feed_forward.output_dense.bias"] = w[p + "ffn2.bias"]
    sd["wav2vec2.encoder.layer_norm.weight"] = w["final_ln.weight"]
    sd["wav2vec2.encoder.layer_norm.bias"] = w["final_ln.bias"]
    sd["lm_head.weight"] = w["ctc.weight"]
    sd["lm_head.bias"] = w["ctc.bias"]
    return sd


def hf_config(cfg: CtcModelConfig):
    from transformers import Wav2Vec2Config
    return Wav2Vec2Config(
        vocab_size=cfg.vocab, hidden_size=cfg.d_model, num_hidden_layers=cfg.n_layers,
        num_attention_heads=cfg.n_heads, intermediate_size=cfg.d_ffn,
        feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True,
        conv_dim=tuple(c for c, _, _ in cfg.fe_layers),
        conv_stride=tuple(s for _, _, s in cfg.fe_layers),
        conv_kernel=tuple(k for _, k, _ in cfg.fe_layers),
        num_conv_pos_embeddings=cfg.pos_kernel, num_conv_pos_embedding_groups=cfg.pos_groups,
        mask_time_prob=0.0, mask_feature_prob=0.0, hidden_dropout=0.0, activation_dropout=0.0,
        attention_dropout=0.0, feat_proj_dropout=0.0, final_dropout=0.0, layerdrop=0.0,
        hidden_act="gelu", layer_norm_eps=LN_EPS, attn_implementation="eager")
