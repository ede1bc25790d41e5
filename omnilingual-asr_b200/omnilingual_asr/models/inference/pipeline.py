"""`ASRInferencePipeline`: the upstream-shaped entry point the reference still documents
(CONTRIBUTING.md:21, `from omnilingual_asr.models.inference.pipeline import ASRInferencePipeline`).

Upstream shape (facebookresearch/omnilingual-asr): ASRInferencePipeline(model_card, device, dtype)
.transcribe(inp: list[path | tensor | {"waveform","sample_rate"}], *, lang=None, batch_size=...) -> list[str].
Inputs longer than 40 s are rejected, as upstream does (MAX_ALLOWED_AUDIO_SEC).
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

from omnilingual_asr.models.config import SAMPLE_RATE
from omnilingual_asr.models.inference.ctc_pipeline import (MAX_ALLOWED_AUDIO_SEC, CTCASRPipeline, _resolve_audio)


class ASRInferencePipeline:
    def __init__(self, model_card: str = "omniASR_CTC_1B", device: Any = None, dtype: Any = None, *,
                 weights: Any = None, vocabulary: Any = None, engine: Any = None) -> None:
        # dtype is accepted for signature compatibility; the engine computes in bf16 with fp32 accumulation
        self._ctc = CTCASRPipeline(model_card, weights=weights, vocabulary=vocabulary, device=device, engine=engine,
                                   window_seconds=MAX_ALLOWED_AUDIO_SEC, distributed=False)

    def transcribe(self, inp: Sequence[Any], *, lang: Optional[Sequence[Optional[str]]] = None,
                   batch_size: int = 32) -> List[str]:
        if isinstance(inp, (str, bytes)) or not isinstance(inp, Sequence):
            raise ValueError("inp must be a list of paths, waveforms or {'waveform','sample_rate'} dicts")
        if batch_size <= 0:
            raise ValueError("batch_size must be positive")
        waves = [_resolve_audio(a, None) for a in inp]
        for w in waves:
            if len(w) > MAX_ALLOWED_AUDIO_SEC * SAMPLE_RATE:
                raise ValueError(f"audio longer than {MAX_ALLOWED_AUDIO_SEC} s is not supported by this entry point")
        # clips sorted by length so that a batch pads little; the pool packs them (batch_size bounds a batch)
        order = sorted(range(len(waves)), key=lambda i: len(waves[i]))
        texts: List[str] = [""] * len(waves)
        for b0 in range(0, len(order), batch_size):
            idx = order[b0:b0 + batch_size]
            toks = self._ctc.pool.run_clips([waves[i] for i in idx])
            for i, t in zip(idx, toks):
                texts[i] = self._ctc.vocab.decode(t.token_ids)
        return texts
