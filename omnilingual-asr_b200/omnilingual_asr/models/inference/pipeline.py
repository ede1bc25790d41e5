"""`ASRInferencePipeline`: the upstream-shaped entry point the reference still documents
(CONTRIBUTING.md:21, `from omnilingual_asr.models.inference.pipeline import ASRInferencePipeline`).

Upstream shape (facebookresearch/omnilingual-asr): ASRInferencePipeline(model_card, device, dtype)
.transcribe(inp: list[path | tensor | {"waveform","sample_rate"}], *, lang=None, batch_size=...) -> list[str].
Inputs longer than 40 s are rejected, as upstream does (MAX_ALLOWED_AUDIO_SEC).
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

import numpy as np

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
        texts: List[str] = [""] * len(waves)
        order = sorted(range(len(waves)), key=lambda i: len(waves[i]))   # bucket by length, pad with zeros
        for b0 in range(0, len(order), batch_size):
            idx = order[b0:b0 + batch_size]
            L = max(max(len(waves[i]) for i in idx), 1)
            batch = np.zeros((len(idx), L), dtype=np.float32)
            for r, i in enumerate(idx):
                batch[r, :len(waves[i])] = waves[i]
            with self._ctc._lock:
                res = self._ctc.engine.transcribe_host(batch, [len(waves[i]) for i in idx])
            for r, i in enumerate(idx):
                texts[i] = self._ctc.vocab.decode(res.token_ids[r])
        return texts
