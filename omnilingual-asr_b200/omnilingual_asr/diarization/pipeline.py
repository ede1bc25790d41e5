"""Drop-in boundary: `pipeline.transcribe(audio) -> List[DiarizedTranscriptSegment]` for the CTC models.

Same surface as the reference's GeminiDiarizedTranscriptionPipeline
(src/omnilingual_asr/diarization/pipeline.py:39-126): keyword-only constructor, read-only `summary` and
`detected_languages`, and `transcribe(audio_path, *, word_timestamps, progress_callback, language,
speaker_count, **kwargs)`; the record types are field-for-field the reference's (:15-36).
Differences, all deliberate:
  * `word_timestamps=True` is honoured (CTC frames give them for free; the reference ignores it, :78),
  * `summary` / `detected_languages` are kept per calling thread, fixing the reference's race when one
    pipeline object serves several threads (:104-106 with workflows/wav2elan_web/app.py:38-54, 384-389),
  * audio may also be an in-memory waveform (ndarray / tensor / {"waveform", "sample_rate"}).
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Any, Callable, List, Optional

from omnilingual_asr.models.inference.ctc_pipeline import CTCASRPipeline


@dataclass(frozen=True)
class WordTimestamp:
    """Word-level timestamp information."""
    word: str
    start: float
    end: float


@dataclass(frozen=True)
class DiarizedTranscriptSegment:
    """A transcribed segment with speaker and timing information."""
    start: float
    end: float
    speaker: str
    text: str
    words: list[WordTimestamp] | None = None
    language: str | None = None
    language_code: str | None = None
    languages: list[dict] | None = None
    emotion: str | None = None
    translation: str | None = None


class CTCTranscriptionPipeline:
    """omniASR CTC transcription on B200 behind the reference's diarized-pipeline API.

    CTC models do not diarize: every segment carries the reference's default speaker label "Speaker 1"
    (gemini_pipeline.py:435); language/emotion/translation stay None (language_code echoes the hint).
    """

    def __init__(self, *, model_card: str = "omniASR_CTC_1B", weights: Any = None, vocabulary: Any = None,
                 device: Any = None, engine: Any = None, **engine_kwargs: Any) -> None:
        self.ctc = CTCASRPipeline(model_card, weights=weights, vocabulary=vocabulary, device=device, engine=engine,
                                  **engine_kwargs)
        self._local = threading.local()

    @property
    def summary(self) -> Optional[str]:
        """Summary of the calling thread's last transcription."""
        return getattr(self._local, "summary", None)

    @property
    def detected_languages(self) -> Optional[List[dict]]:
        """Languages (the hint, if any) of the calling thread's last transcription."""
        return getattr(self._local, "detected_languages", None)

    def transcribe(
        self,
        audio_path: Any,
        *,
        word_timestamps: bool = False,
        progress_callback: Optional[Callable[[str, int], None]] = None,
        language: Optional[str] = None,
        speaker_count: Optional[str] = None,
        **kwargs: Any,  # accepted and ignored, like the reference (:82)
    ) -> List[DiarizedTranscriptSegment]:
        result = self.ctc.transcribe_with_retry(
            audio_path,
            progress_callback=progress_callback,
            language=language,
            speaker_count=speaker_count,
            sample_rate=kwargs.get("sample_rate"),
            word_timestamps=word_timestamps,
        )
        self._local.summary = result.summary
        self._local.detected_languages = result.detected_languages
        segments: List[DiarizedTranscriptSegment] = []
        for seg in result.segments:
            words = None
            if seg.words is not None:
                words = [WordTimestamp(w.word, w.start, w.end) for w in seg.words]
            segments.append(
                DiarizedTranscriptSegment(
                    start=seg.start, end=seg.end, speaker=seg.speaker, text=seg.text, words=words,
                    language=seg.language, language_code=seg.language_code, languages=seg.languages,
                    emotion=seg.emotion, translation=seg.translation))
        return segments
