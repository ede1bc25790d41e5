"""Architecture presets of the omniASR CTC family.

The four published models share one graph (wav2vec2 layer-norm conv feature extractor with conv bias,
pre-LN Transformer encoder, weight-normed grouped positional conv k=128/g=16, CTC head of 9812) and differ
only in width/depth.  The numbers are pinned by the published parameter totals:
325,494,996 / 975,065,300 / 3,080,423,636 / 6,504,786,132.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

SAMPLE_RATE = 16_000
FE_LAYERS: Tuple[Tuple[int, int, int], ...] = (
    (512, 10, 5), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 2, 2), (512, 2, 2))


@dataclass(frozen=True)
class CtcModelConfig:
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
    blank_id: int = 0

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_heads

    def feature_length(self, n_samples: int) -> int:
        """Frames out of the conv feature extractor: chain of floor((L-k)/s)+1."""
        n = int(n_samples)
        for _, k, s in self.fe_layers:
            n = (n - k) // s + 1 if n >= k else 0
        return max(n, 0)

    def weight_shapes(self) -> Dict[str, Tuple[int, ...]]:
        """Name -> shape of every parameter the engine expects (oasr_load_weight names)."""
        d, f = self.d_model, self.d_ffn
        shapes: Dict[str, Tuple[int, ...]] = {}
        c_in = 1
        for i, (c, k, _) in enumerate(self.fe_layers):
            shapes[f"fe.{i}.conv.weight"] = (c, c_in, k)
            shapes[f"fe.{i}.conv.bias"] = (c,)
            shapes[f"fe.{i}.ln.weight"] = (c,)
            shapes[f"fe.{i}.ln.bias"] = (c,)
            c_in = c
        shapes["proj.ln.weight"] = (self.fe_dim,)
        shapes["proj.ln.bias"] = (self.fe_dim,)
        shapes["proj.linear.weight"] = (d, self.fe_dim)
        shapes["proj.linear.bias"] = (d,)
        shapes["pos.weight_g"] = (1, 1, self.pos_kernel)
        shapes["pos.weight_v"] = (d, d // self.pos_groups, self.pos_kernel)
        shapes["pos.bias"] = (d,)
        for l in range(self.n_layers):
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
        shapes["ctc.weight"] = (self.vocab, d)
        shapes["ctc.bias"] = (self.vocab,)
        return shapes


MODEL_CARDS: Dict[str, CtcModelConfig] = {
    "omniASR_CTC_300M": CtcModelConfig("omniASR_CTC_300M", 1024, 24, 16, 4096),
    "omniASR_CTC_1B": CtcModelConfig("omniASR_CTC_1B", 1280, 48, 16, 5120),
    "omniASR_CTC_3B": CtcModelConfig("omniASR_CTC_3B", 2048, 60, 16, 8192),
    "omniASR_CTC_7B": CtcModelConfig("omniASR_CTC_7B", 2048, 128, 16, 8192),
}


def get_model_config(card: str | CtcModelConfig) -> CtcModelConfig:
    if isinstance(card, CtcModelConfig):
        return card
    try:
        return MODEL_CARDS[card]
    except KeyError:
        raise ValueError(f"unknown CTC model card {card!r}; known: {sorted(MODEL_CARDS)}") from None
