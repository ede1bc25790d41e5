"""CTC speech transcription pipeline on B200 (replaces the reference's Gemini engine layer)."""

from omnilingual_asr.models.inference.ctc_pipeline import (
    CTCASRPipeline,
    CTCTranscriptionResult,
    CTCTranscriptSegment,
    WordTimestamp,
)

__all__ = [
    "CTCASRPipeline",
    "CTCTranscriptionResult",
    "CTCTranscriptSegment",
    "WordTimestamp",
]
