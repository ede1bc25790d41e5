"""CTC transcription pipeline behind the reference's diarized-segment API."""

from omnilingual_asr.diarization.pipeline import (
    CTCTranscriptionPipeline,
    DiarizedTranscriptSegment,
    WordTimestamp,
)

__all__ = [
    "DiarizedTranscriptSegment",
    "CTCTranscriptionPipeline",
    "WordTimestamp",
]
