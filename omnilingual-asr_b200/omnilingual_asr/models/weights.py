"""Weight sources for the CTC engine: a state dict, a torch checkpoint on disk, or a seeded random init.

There are no omniASR checkpoints offline, so benchmarks and smoke tests use weights='random'.  The random
law is the one parity tests use (oracle.init_weights): N(0, gain/fan_in) matrices, LayerNorm weights around
1, small non-zero biases, so that every bias/affine path is exercised.
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Any, Dict, Mapping

import numpy as np
import torch

from omnilingual_asr.models.config import CtcModelConfig


def random_weights(cfg: CtcModelConfig, seed: int = 0, device: Any = "cpu") -> Dict[str, torch.Tensor]:
    """Seeded random init generated directly on `device` (fast for the 1B-7B models on a GPU)."""
    device = torch.device(device)
    gen = torch.Generator(device=device).manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}

    def randn(shape):
        return torch.randn(shape, generator=gen, device=device, dtype=torch.float32)

    for name, shape in cfg.weight_shapes().items():
        if name.endswith("ln.weight"):
            t = 1.0 + 0.1 * randn(shape)
        elif name.endswith("bias"):
            t = 0.1 * randn(shape)
        elif name == "pos.weight_g":
            t = (1.0 + 0.1 * randn(shape)).abs()
        elif name == "pos.weight_v":
            t = randn(shape) * math.sqrt(2.0 / (shape[1] * shape[2]))
        else:
            fan_in = int(np.prod(shape[1:]))
            pre_gelu = name.startswith("fe.") or name.endswith("ffn1.weight")
            t = randn(shape) * math.sqrt((2.0 if pre_gelu else 1.0) / fan_in)
        out[name] = t
    v = out["pos.weight_v"]
    out["pos.weight_g"] = out["pos.weight_g"] * v.norm(dim=(0, 1), keepdim=True)
    return out


_HEAD_KEYS = ("num_attention_heads", "num_encoder_attn_heads", "encoder_attention_heads", "n_heads", "num_heads")


def check_head_count(cfg: CtcModelConfig, meta: Any) -> None:
    """The head count changes no tensor shape, so a checkpoint loads cleanly under a wrong one and then transcribes
    wrongly.  The 16 heads of the 2048-wide cards (3B, 7B) come from SURVEY.md's recollection of upstream's arch
    registry, not from a file available here: when a checkpoint carries its own head count (a config / metadata
    mapping with one of the usual keys, nested or not) it must agree with `cfg`; when a REAL checkpoint of a 2048-wide
    model carries none, the caller is told what is being assumed."""
    found = None

    def walk(m: Any, depth: int) -> None:
        nonlocal found
        if found is not None or depth > 3 or not isinstance(m, Mapping):
            return
        for k in _HEAD_KEYS:
            if k in m and isinstance(m[k], (int, np.integer)):
                found = int(m[k])
                return
        for k, v in m.items():
            if isinstance(v, Mapping) and not (isinstance(k, str) and k in ("model", "state_dict")):
                walk(v, depth + 1)

    walk(meta, 0)
    if found is not None:
        if found != cfg.n_heads:
            raise ValueError(f"checkpoint says {found} attention heads, model card {cfg.name} is configured with "
                             f"{cfg.n_heads}: pass a CtcModelConfig with n_heads={found}")
    elif cfg.d_model >= 2048:
        import warnings
        warnings.warn(f"{cfg.name}: the checkpoint carries no head count; assuming {cfg.n_heads} heads "
                      f"(head_dim {cfg.d_model // cfg.n_heads}), which is upstream's value as recalled in SURVEY.md, not "
                      "verified offline - pass a CtcModelConfig with the right n_heads if it differs", stacklevel=3)


def resolve_weights(cfg: CtcModelConfig, weights: Any, seed: int, device: Any) -> Mapping[str, torch.Tensor]:
    if isinstance(weights, str) and weights == "random":
        return random_weights(cfg, seed, device)
    if isinstance(weights, (str, Path)):
        p = Path(weights)
        if not p.exists():
            raise ValueError(f"checkpoint not found: {p}")
        sd = torch.load(str(p), map_location="cpu", weights_only=True)
        check_head_count(cfg, sd)
        if isinstance(sd, Mapping) and "model" in sd and isinstance(sd["model"], Mapping):
            sd = sd["model"]
        return convert_state_dict(sd, cfg)
    if isinstance(weights, Mapping):
        return convert_state_dict(weights, cfg)
    raise ValueError("weights must be 'random', a checkpoint path or a mapping name -> tensor")


# ---------------------------------------------------------------------------------------------------------
# Foreign checkpoint layouts.  The engine's own names are CtcModelConfig.weight_shapes().
#  * Hugging Face `Wav2Vec2ForCTC` (feat_extract_norm="layer", do_stable_layer_norm=True): checked against the
#    live transformers module in tests/test_weights.py.
#  * fairseq2 `Wav2Vec2AsrModel`, the upstream home of omniASR_CTC_*: the module tree as published upstream
#    (encoder_frontend.feature_extractor / post_extract_layer_norm / model_dim_proj / pos_encoder, encoder.layers.N.
#    self_attn{,_layer_norm} / ffn{,_layer_norm}, encoder.layer_norm, final_proj).  fairseq2 is not installable here,
#    so this table is pinned only by its own round-trip test; unknown keys raise instead of being dropped.
# Weight-normed positional conv: both the old (`weight_g` / `weight_v`) and the parametrized
# (`parametrizations.weight.original0` / `original1`) spellings are accepted.
# ---------------------------------------------------------------------------------------------------------
def _hf_key_map(cfg: CtcModelConfig) -> Dict[str, str]:
    m: Dict[str, str] = {}
    for i in range(len(cfg.fe_layers)):
        b = f"wav2vec2.feature_extractor.conv_layers.{i}."
        m[b + "conv.weight"] = f"fe.{i}.conv.weight"
        m[b + "conv.bias"] = f"fe.{i}.conv.bias"
        m[b + "layer_norm.weight"] = f"fe.{i}.ln.weight"
        m[b + "layer_norm.bias"] = f"fe.{i}.ln.bias"
    fp = "wav2vec2.feature_projection."
    m[fp + "layer_norm.weight"] = "proj.ln.weight"
    m[fp + "layer_norm.bias"] = "proj.ln.bias"
    m[fp + "projection.weight"] = "proj.linear.weight"
    m[fp + "projection.bias"] = "proj.linear.bias"
    pc = "wav2vec2.encoder.pos_conv_embed.conv."
    m[pc + "parametrizations.weight.original0"] = "pos.weight_g"
    m[pc + "parametrizations.weight.original1"] = "pos.weight_v"
    m[pc + "weight_g"] = "pos.weight_g"
    m[pc + "weight_v"] = "pos.weight_v"
    m[pc + "bias"] = "pos.bias"
    for l in range(cfg.n_layers):
        b, p = f"wav2vec2.encoder.layers.{l}.", f"enc.{l}."
        m[b + "layer_norm.weight"] = p + "attn_ln.weight"
        m[b + "layer_norm.bias"] = p + "attn_ln.bias"
        for hn, n in (("q_proj", "q"), ("k_proj", "k"), ("v_proj", "v"), ("out_proj", "o")):
            m[b + f"attention.{hn}.weight"] = p + f"{n}.weight"
            m[b + f"attention.{hn}.bias"] = p + f"{n}.bias"
        m[b + "final_layer_norm.weight"] = p + "ffn_ln.weight"
        m[b + "final_layer_norm.bias"] = p + "ffn_ln.bias"
        m[b + "feed_forward.intermediate_dense.weight"] = p + "ffn1.weight"
        m[b + "feed_forward.intermediate_dense.bias"] = p + "ffn1.bias"
        m[b + "feed_forward.output_dense.weight"] = p + "ffn2.weight"
        m[b + "feed_forward.output_dense.bias"] = p + "ffn2.bias"
    m["wav2vec2.encoder.layer_norm.weight"] = "final_ln.weight"
    m["wav2vec2.encoder.layer_norm.bias"] = "final_ln.bias"
    m["lm_head.weight"] = "ctc.weight"
    m["lm_head.bias"] = "ctc.bias"
    return m


def _fairseq2_key_map(cfg: CtcModelConfig) -> Dict[str, str]:
    m: Dict[str, str] = {}
    for i in range(len(cfg.fe_layers)):
        b = f"encoder_frontend.feature_extractor.layers.{i}."
        m[b + "conv.weight"] = f"fe.{i}.conv.weight"
        m[b + "conv.bias"] = f"fe.{i}.conv.bias"
        m[b + "layer_norm.weight"] = f"fe.{i}.ln.weight"
        m[b + "layer_norm.bias"] = f"fe.{i}.ln.bias"
    ef = "encoder_frontend."
    m[ef + "post_extract_layer_norm.weight"] = "proj.ln.weight"
    m[ef + "post_extract_layer_norm.bias"] = "proj.ln.bias"
    m[ef + "model_dim_proj.weight"] = "proj.linear.weight"
    m[ef + "model_dim_proj.bias"] = "proj.linear.bias"
    pc = ef + "pos_encoder.conv."
    m[pc + "parametrizations.weight.original0"] = "pos.weight_g"
    m[pc + "parametrizations.weight.original1"] = "pos.weight_v"
    m[pc + "weight_g"] = "pos.weight_g"
    m[pc + "weight_v"] = "pos.weight_v"
    m[pc + "bias"] = "pos.bias"
    for l in range(cfg.n_layers):
        b, p = f"encoder.layers.{l}.", f"enc.{l}."
        m[b + "self_attn_layer_norm.weight"] = p + "attn_ln.weight"
        m[b + "self_attn_layer_norm.bias"] = p + "attn_ln.bias"
        for fn, n in (("q_proj", "q"), ("k_proj", "k"), ("v_proj", "v"), ("output_proj", "o")):
            m[b + f"self_attn.{fn}.weight"] = p + f"{n}.weight"
            m[b + f"self_attn.{fn}.bias"] = p + f"{n}.bias"
        m[b + "ffn_layer_norm.weight"] = p + "ffn_ln.weight"
        m[b + "ffn_layer_norm.bias"] = p + "ffn_ln.bias"
        m[b + "ffn.inner_proj.weight"] = p + "ffn1.weight"
        m[b + "ffn.inner_proj.bias"] = p + "ffn1.bias"
        m[b + "ffn.output_proj.weight"] = p + "ffn2.weight"
        m[b + "ffn.output_proj.bias"] = p + "ffn2.bias"
    m["encoder.layer_norm.weight"] = "final_ln.weight"
    m["encoder.layer_norm.bias"] = "final_ln.bias"
    m["final_proj.weight"] = "ctc.weight"
    m["final_proj.bias"] = "ctc.bias"
    return m


# parameters of the training-time modules that inference does not use
_IGNORED_SUFFIXES = ("masked_spec_embed", "masker.temporal_mask_embed", "num_batches_tracked")


def detect_layout(sd: Mapping[str, Any]) -> str:
    """'native' | 'hf' | 'fairseq2', from the key prefixes."""
    keys = list(sd.keys())
    if any(k.startswith("wav2vec2.") or k.startswith("lm_head.") for k in keys):
        return "hf"
    if any(k.startswith("encoder_frontend.") or k.startswith("final_proj.") for k in keys):
        return "fairseq2"
    return "native"


def convert_state_dict(sd: Mapping[str, Any], cfg: CtcModelConfig) -> Mapping[str, Any]:
    """Renames a Hugging Face Wav2Vec2ForCTC or fairseq2 Wav2Vec2AsrModel state dict to the engine's names
    (native dicts pass through).  Raises ValueError on keys it does not know, on missing parameters and on shape
    mismatches: a silently half-loaded model would still transcribe - wrongly."""
    layout = detect_layout(sd)
    if layout == "native":
        return sd
    table = _hf_key_map(cfg) if layout == "hf" else _fairseq2_key_map(cfg)
    out: Dict[str, Any] = {}
    unknown = []
    for k, v in sd.items():
        if k in table:
            out[table[k]] = v
        elif not k.endswith(_IGNORED_SUFFIXES):
            unknown.append(k)
    if unknown:
        raise ValueError(f"{layout} checkpoint has parameters this architecture does not: {unknown[:8]}"
                         + (" ..." if len(unknown) > 8 else ""))
    shapes = cfg.weight_shapes()
    missing = [n for n in shapes if n not in out]
    if missing:
        raise ValueError(f"{layout} checkpoint lacks: {missing[:8]}" + (" ..." if len(missing) > 8 else ""))
    for n, shp in shapes.items():
        if tuple(out[n].shape) != tuple(shp):
            raise ValueError(f"{n}: checkpoint shape {tuple(out[n].shape)} != model shape {tuple(shp)}")
    return out
