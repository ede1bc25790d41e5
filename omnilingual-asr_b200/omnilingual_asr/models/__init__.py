"""Omnilingual ASR models."""

from omnilingual_asr.models import inference as inference

__all__ = [
    "inference",
]
