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


def resolve_weights(cfg: CtcModelConfig, weights: Any, seed: int, device: Any) -> Mapping[str, torch.Tensor]:
    if isinstance(weights, str) and weights == "random":
        return random_weights(cfg, seed, device)
    if isinstance(weights, (str, Path)):
        p = Path(weights)
        if not p.exists():
            raise ValueError(f"checkpoint not found: {p}")
        sd = torch.load(str(p), map_location="cpu", weights_only=True)
        if isinstance(sd, Mapping) and "model" in sd and isinstance(sd["model"], Mapping):
            sd = sd["model"]
        return sd
    if isinstance(weights, Mapping):
        return weights
    raise ValueError("weights must be 'random', a checkpoint path or a mapping name -> tensor")
