"""Omnilingual ASR - local omniASR CTC transcription on NVIDIA B200 (sm_100a).

Same package layout and public names as the reference (src/omnilingual_asr/__init__.py), with the CTC
pipeline where the reference exports its Gemini client.
"""

__version__ = "0.2.0"

from omnilingual_asr.diarization import CTCTranscriptionPipeline, DiarizedTranscriptSegment, WordTimestamp
from omnilingual_asr.models.inference import (
    CTCASRPipeline,
    CTCTranscriptionResult,
    CTCTranscriptSegment,
)

__all__ = [
    "__version__",
    "CTCASRPipeline",
    "CTCTranscriptionResult",
    "CTCTranscriptSegment",
    "CTCTranscriptionPipeline",
    "DiarizedTranscriptSegment",
    "WordTimestamp",
]
