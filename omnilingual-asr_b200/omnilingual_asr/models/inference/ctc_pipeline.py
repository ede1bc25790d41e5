"""CTC engine-level pipeline: the local replacement of the reference's GeminiASRPipeline.

Mirrors src/omnilingual_asr/models/inference/gemini_pipeline.py of the reference:
  * dataclasses CTCTranscriptSegment / CTCTranscriptionResult  <- GeminiTranscriptSegment / -Result (:47-70)
  * transcribe / transcribe_chunked / transcribe_with_retry      <- :474-539, :577-682, :684-741
  * window law (fixed, non-overlapping, start = i * window)      <- split_audio_into_chunks :243-310
  * timestamp rebase by the window start                         <- _transcribe_chunk :555-569
  * merge = windows in start order, segments concatenated        <- :646-654
  * progress steps ("uploading",0) ("transcribing",1) ("processing",2) ("done",3)   <- :486-487
  * errors: ValueError for configuration, RuntimeError("Failed to transcribe after N attempts: ...")
    after retries (:329-334, :739-741)
The per-chunk HTTPS call (:512-530) is replaced by liboasr (sm_100a): windows go through an EnginePool
(engine_pool.py) - one engine and one worker thread per GPU behind a queue shared by all callers, two batches in
flight per engine (oasr_transcribe_host_async / oasr_wait).  One pipeline object in one process drives every GPU it
was given (`devices="all"`), which is how the reference's web app holds it (workflows/wav2elan_web/app.py:38-54).
Unlike the reference a failed window is never dropped silently (:635-641): a CUDA failure raises.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Callable, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from omnilingual_asr.models.config import SAMPLE_RATE, CtcModelConfig, get_model_config
from omnilingual_asr.models.inference.audio import (get_audio_duration, load_audio_16k, ownership_bounds, read_wav,
                                                    shard_range, split_into_overlapping_windows, split_into_windows,
                                                    to_mono_16k)
from omnilingual_asr.models.inference.engine_pool import EnginePool, WindowTokens
from omnilingual_asr.models.inference.tokenizer import CtcVocabulary

# Window constants (the reference's CHUNK_DURATION_SECONDS / MIN_DURATION_FOR_CHUNKING / MAX_PARALLEL_CHUNKS,
# gemini_pipeline.py:217-219, re-valued for the CTC models: upstream caps one input at 40 s).
CHUNK_DURATION_SECONDS = 30.0
MIN_DURATION_FOR_CHUNKING = 30.0
MAX_ALLOWED_AUDIO_SEC = 40.0
MAX_PARALLEL_CHUNKS = 32          # windows per device step (batch), not threads
DEFAULT_SPEAKER = "Speaker 1"     # gemini_pipeline.py:435


@dataclass(frozen=True)
class WordTimestamp:
    """Word-level timestamp information (gemini_pipeline.py:39-45)."""
    word: str
    start: float
    end: float


@dataclass
class CTCTranscriptSegment:
    """One transcribed segment; field-for-field GeminiTranscriptSegment (gemini_pipeline.py:47-61)."""
    start: float
    end: float
    speaker: str
    text: str
    language: Optional[str] = None
    language_code: Optional[str] = None
    languages: Optional[List[dict]] = None
    emotion: Optional[str] = None
    translation: Optional[str] = None
    words: Optional[List[WordTimestamp]] = None


@dataclass
class CTCTranscriptionResult:
    """Complete result; mirrors GeminiTranscriptionResult (gemini_pipeline.py:64-70)."""
    summary: Optional[str] = None
    segments: List[CTCTranscriptSegment] = field(default_factory=list)
    detected_languages: Optional[List[dict]] = None


AudioInput = Any  # path | np.ndarray | torch.Tensor | {"waveform": ..., "sample_rate": ...}


def _resolve_audio(audio: AudioInput, sample_rate: Optional[int], engine: Any = None):
    """Anything the boundary accepts -> mono samples at 16 kHz as a host array (float32, or PCM16 left as it is); audio
    that needs a channel mix or a rate change goes through the device-side front end (oasr_resample) when the engine
    has one."""
    if isinstance(audio, (str, Path)):
        p = Path(audio)
        if engine is not None and hasattr(engine, "resample_to_model_rate") and p.exists() \
                and p.suffix.lower() in (".wav", ".wave"):
            x, sr = read_wav(p)
            return _resolve_audio(x, sr, engine)
        return load_audio_16k(audio)
    if isinstance(audio, Mapping):
        return _resolve_audio(audio["waveform"], int(audio.get("sample_rate", sample_rate or SAMPLE_RATE)), engine)
    if hasattr(audio, "detach"):  # torch.Tensor
        audio = audio.detach().cpu().numpy()
    x = np.asarray(audio)
    sr = int(sample_rate or SAMPLE_RATE)
    if x.dtype == np.int16 and x.ndim == 1 and sr == SAMPLE_RATE:
        return x   # mono PCM16 at 16 kHz stays as it is: the device converts it (OASR_FLAG_INPUT_I16)
    if x.ndim == 2 and x.shape[0] < x.shape[1] and x.shape[0] <= 8:
        x = x.T  # [channels, n] -> [n, channels]
    if x.ndim == 2 and x.shape[1] == 1:
        x = x[:, 0]
    needs_front_end = sr != SAMPLE_RATE or x.ndim == 2
    if needs_front_end and engine is not None and hasattr(engine, "resample_to_model_rate"):
        # channel mix + polyphase resampling on the device (oasr_resample); the 16 kHz mono result comes back to the host
        # because its windows may be served by any GPU of the pool
        return engine.resample_to_model_rate(x, sr).cpu().numpy()
    if x.dtype == np.int16:
        x = x.astype(np.float32) / 32768.0
    return to_mono_16k(x, sr)


def build_segments(win: WindowTokens, vocab: CtcVocabulary, *, word_timestamps: bool,
                   split_gap_sec: Optional[float], language: Optional[str]) -> List[CTCTranscriptSegment]:
    """Tokens of one window -> segments with absolute times (rebase = gemini_pipeline.py:555-569).

    Frame t of a window starts at window_start + t * (window_duration / n_frames); a segment runs from its
    first token's frame to the end of its last token's frame.  With split_gap_sec the window is cut where no
    token is emitted for at least that long (step *after* the path).
    """
    if win.n_frames <= 0 or len(win.token_ids) == 0:
        return []
    offset = win.start_sample / SAMPLE_RATE
    frame_dur = (win.n_samples / SAMPLE_RATE) / win.n_frames
    ids, frames = win.token_ids, win.token_frames
    cuts = [0]
    if split_gap_sec is not None and split_gap_sec > 0:
        gap_frames = split_gap_sec / frame_dur
        for i in range(1, len(frames)):
            if frames[i] - frames[i - 1] >= gap_frames:
                cuts.append(i)
    cuts.append(len(ids))
    out: List[CTCTranscriptSegment] = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        text = vocab.decode(ids[lo:hi])
        if not text:
            continue
        words = None
        if word_timestamps:
            words = [WordTimestamp(w, offset + f0 * frame_dur, offset + (f1 + 1) * frame_dur)
                     for (w, f0, f1) in vocab.words_with_frames(ids[lo:hi], frames[lo:hi])]
        out.append(CTCTranscriptSegment(
            start=offset + float(frames[lo]) * frame_dur,
            end=offset + (float(frames[hi - 1]) + 1.0) * frame_dur,
            speaker=DEFAULT_SPEAKER, text=text, language_code=language, words=words))
    return out


def _token_centres(w: WindowTokens) -> np.ndarray:
    """Absolute sample position of the centre of every token's frame."""
    if w.n_frames <= 0 or len(w.token_ids) == 0:
        return np.zeros((0,), dtype=np.float64)
    return w.start_sample + (w.token_frames.astype(np.float64) + 0.5) * (w.n_samples / w.n_frames)


def stitch_bounds(tokens: Sequence[WindowTokens], windows: Sequence[Tuple[int, int]]) -> List[Tuple[float, float]]:
    """Ownership bounds with every cut moved, inside the overlap of the two windows it separates, to the middle of the
    longest stretch in which NEITHER window emits a token - a silence both transcriptions agree on - so that the cut
    does not run through a token one window places just before it and the other just after.  Overlaps without any
    token, and windows without tokens, keep the geometric middle (audio.ownership_bounds)."""
    bounds = [list(b) for b in ownership_bounds(list(windows))]
    by_index = {w.index: w for w in tokens}
    for i in range(len(windows) - 1):
        # the overlap of windows i and i + 1; cuts never cross: the search starts at the previous cut at the earliest
        a = max(float(windows[i + 1][0]), float(bounds[i][0]))
        b = float(windows[i][0] + windows[i][1])
        if b <= a or i not in by_index or i + 1 not in by_index:
            continue
        t = np.concatenate([_token_centres(by_index[i]), _token_centres(by_index[i + 1])])
        t = np.sort(t[(t >= a) & (t < b)])
        if t.size == 0:
            continue
        edges = np.concatenate([[a], t, [b]])
        k = int(np.argmax(np.diff(edges)))
        cut = 0.5 * (edges[k] + edges[k + 1])
        bounds[i][1] = cut
        bounds[i + 1][0] = cut
    return [(lo, hi) for lo, hi in bounds]


def trim_to_ownership(tokens: Sequence[WindowTokens], windows: Sequence[Tuple[int, int]]) -> List[WindowTokens]:
    """Overlapping windows: keep, per window, the tokens whose frame centre lies in the span the window owns
    (stitch_bounds: neighbours meet in the overlap, at a silence common to both).  A sample position belongs to
    exactly one window, so nothing is emitted twice by construction."""
    bounds = stitch_bounds(tokens, windows)
    out: List[WindowTokens] = []
    for w in tokens:
        lo, hi = bounds[w.index]
        if w.n_frames <= 0 or len(w.token_ids) == 0:
            out.append(w)
            continue
        centre = _token_centres(w)
        keep = (centre >= lo) & (centre < hi)
        out.append(WindowTokens(w.index, w.start_sample, w.n_samples, w.n_frames, w.token_ids[keep], w.token_frames[keep]))
    return out


def merge_window_results(per_window: Sequence[Tuple[WindowTokens, List[CTCTranscriptSegment]]],
                         language: Optional[str]) -> CTCTranscriptionResult:
    """Sort by window start and concatenate (gemini_pipeline.py:646-678)."""
    ordered = sorted(per_window, key=lambda r: r[0].start_sample)
    segments: List[CTCTranscriptSegment] = []
    for _, segs in ordered:
        segments.extend(segs)
    n_tok = sum(len(w.token_ids) for w, _ in ordered)
    dur = sum(w.n_samples for w, _ in ordered) / SAMPLE_RATE
    summary = f"{len(segments)} segment(s), {n_tok} token(s), {dur:.2f} s of audio in {len(ordered)} window(s)"
    langs = [{"name": language, "code": language}] if language else None
    return CTCTranscriptionResult(summary=summary, segments=segments, detected_languages=langs)


class CTCASRPipeline:
    """Local CTC pipeline with the reference engine's surface (GeminiASRPipeline, gemini_pipeline.py:313-741)."""

    def __init__(self, model_card: str | CtcModelConfig = "omniASR_CTC_1B", *,
                 weights: Any = None, vocabulary: Optional[CtcVocabulary | Sequence[str]] = None,
                 device: Any = None, devices: Any = None, engine: Any = None, engines: Optional[Sequence[Any]] = None,
                 window_seconds: float = CHUNK_DURATION_SECONDS,
                 batch_windows: int = MAX_PARALLEL_CHUNKS, split_gap_sec: Optional[float] = None,
                 overlap_seconds: float = 0.0, seed: int = 0, distributed: bool = True) -> None:
        self.cfg = get_model_config(model_card)
        if not (0 < window_seconds <= MAX_ALLOWED_AUDIO_SEC):
            raise ValueError(f"window_seconds must be in (0, {MAX_ALLOWED_AUDIO_SEC}]")
        if batch_windows <= 0:
            raise ValueError("batch_windows must be positive")
        self.window_samples = int(round(window_seconds * SAMPLE_RATE))
        if not (0 <= overlap_seconds <= window_seconds / 2):
            # above half a window the overlaps of (i, i+1) and (i+1, i+2) intersect and two cuts could cross
            raise ValueError("overlap_seconds must be in [0, window_seconds / 2]")
        # 0 = the reference's window law (back-to-back windows).  > 0: consecutive windows share this much audio and
        # each keeps only the tokens of the part it owns (overlap-and-stitch)
        self.overlap_samples = int(round(overlap_seconds * SAMPLE_RATE))
        self.batch_windows = int(batch_windows)
        self.split_gap_sec = split_gap_sec
        self.distributed = distributed
        if vocabulary is None:
            vocabulary = CtcVocabulary.synthetic(self.cfg.vocab)
        elif not isinstance(vocabulary, CtcVocabulary):
            vocabulary = CtcVocabulary(list(vocabulary))
        if len(vocabulary) != self.cfg.vocab:
            raise ValueError(f"vocabulary has {len(vocabulary)} entries, model expects {self.cfg.vocab}")
        self.vocab = vocabulary
        self._owns_engines = False
        if engine is not None or engines:
            self.engines = list(engines) if engines else [engine]
        else:
            if weights is None:
                raise ValueError(
                    "no weights given: pass weights=<state dict | path to a torch checkpoint> or weights='random' "
                    "(deterministic random init; there are no omniASR checkpoints offline)")
            self.engines = self._build_engines(device, devices, weights, seed)
            self._owns_engines = True
        self.engine = self.engines[0]     # the audio front end (oasr_resample) and single-engine callers use this one
        # concurrent transcribe calls (app.py shares one pipeline over 4 threads) meet in the pool's queue: their
        # windows are packed into common batches instead of waiting for each other on a lock
        self.pool = EnginePool(self.engines, self.batch_windows)

    def _build_engines(self, device: Any, devices: Any, weights: Any, seed: int) -> List[Any]:
        """One CtcEngine per GPU: `devices` = "all" | a list of devices / indices | None (= `device`, or the current one).
        Replicas are loaded in parallel, each on its own device."""
        import torch
        from concurrent.futures import ThreadPoolExecutor
        from omnilingual_asr.models.inference.ctc_engine import CtcEngine  # needs liboasr.so + CUDA
        from omnilingual_asr.models.weights import resolve_weights
        if devices is None:
            devs = [device]
        elif isinstance(devices, str):
            if devices != "all":
                raise ValueError('devices must be "all", a list of devices or None')
            if not torch.cuda.is_available():
                raise RuntimeError("omniASR CTC engine needs a CUDA device (B200, sm_100a); none is visible")
            devs = [torch.device("cuda", i) for i in range(torch.cuda.device_count())]
        else:
            devs = [torch.device("cuda", d) if isinstance(d, int) else torch.device(d) for d in devices]
            if not devs:
                raise ValueError("devices is empty")

        def make(dev):
            eng = CtcEngine(self.cfg, device=dev)
            eng.load_state_dict(resolve_weights(self.cfg, weights, seed, eng.device))
            return eng

        if len(devs) == 1:
            return [make(devs[0])]
        with ThreadPoolExecutor(max_workers=len(devs)) as ex:
            return list(ex.map(make, devs))

    def close(self) -> None:
        """Stop the worker threads; engines the pipeline created itself are destroyed."""
        self.pool.close()
        if self._owns_engines:
            for e in self.engines:
                e.close()

    # ------------------------------------------------------------------ device step
    def _rank_world(self) -> Tuple[int, int]:
        """(data-parallel rank, data-parallel world) of this process.  The ranks of one tensor-parallel group must
        feed IDENTICAL batches into their shared reductions, so they form one data-parallel replica: with a
        tensor-parallel engine of degree W the replica index is rank // W of world // W."""
        if not self.distributed:
            return 0, 1
        try:
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized()):
                return 0, 1
            rank, world = dist.get_rank(), dist.get_world_size()
        except Exception:
            return 0, 1
        tp = int(getattr(self.engine, "tp_world", 1) or 1)
        if getattr(self.engine, "tp_emulated", False):
            tp = 1
        if tp > 1:
            if world % tp != 0:
                raise ValueError(f"world size {world} is not a multiple of the engine's tensor-parallel degree {tp}")
            return rank // tp, world // tp
        return rank, world

    def _run_windows(self, wave: np.ndarray, windows: List[Tuple[int, int]], lo: int, hi: int, post=None):
        """Windows [lo, hi) through the engine pool; host samples in, token ids out."""
        return self.pool.run_windows(wave, windows, lo, hi, post)

    def _transcribe_wave(self, wave: np.ndarray, *, progress_callback, language, word_timestamps,
                         chunked: bool) -> CTCTranscriptionResult:
        def _report(step: str, idx: int) -> None:
            if progress_callback:
                progress_callback(step, idx)

        if not chunked and len(wave) > int(MAX_ALLOWED_AUDIO_SEC * SAMPLE_RATE):
            raise ValueError(f"audio longer than {MAX_ALLOWED_AUDIO_SEC} s needs chunking "
                             "(use transcribe_chunked / transcribe_with_retry)")
        if chunked and self.overlap_samples > 0:
            windows = split_into_overlapping_windows(len(wave), self.window_samples, self.overlap_samples)
        else:
            windows = split_into_windows(len(wave), self.window_samples if chunked else max(len(wave), 1))
        _report("transcribing", 1)
        rank, world = self._rank_world()
        lo, hi = shard_range(len(windows), rank, world)
        def shape(w: WindowTokens) -> List[CTCTranscriptSegment]:
            return build_segments(w, self.vocab, word_timestamps=word_timestamps, split_gap_sec=self.split_gap_sec,
                                  language=language)

        stitch = chunked and self.overlap_samples > 0
        if world == 1 and not stitch:
            # one process, back-to-back windows: a window's text is shaped by the pool's worker when its batch comes
            # back, under the next batch's device step, not after the last one (1.2 M tokens for a 9.5 h recording)
            shaped = self._run_windows(wave, windows, lo, hi, post=shape)
            _report("processing", 2)
            result = merge_window_results(shaped, language)
            _report("done", 3)
            return result
        failure: Optional[BaseException] = None
        mine: List[WindowTokens] = []
        try:
            mine = self._run_windows(wave, windows, lo, hi)
        except Exception as e:  # noqa: BLE001 - a failing rank must still take part in the gather below
            if world <= 1:
                raise
            failure = e
        if world > 1:
            import torch.distributed as dist
            n_proc = dist.get_world_size()
            gathered: List[Any] = [None] * n_proc
            # host-side gather of token ids only; a rank that failed sends a marker instead of leaving the others blocked
            dist.all_gather_object(gathered, ("error", repr(failure)) if failure is not None else ("ok", mine))
            bad = [(r, g[1]) for r, g in enumerate(gathered) if g[0] == "error"]
            if failure is not None:
                raise failure
            if bad:
                raise RuntimeError(f"window shard failed on rank {bad[0][0]}: {bad[0][1]}")
            tp = n_proc // world               # the ranks of a tensor-parallel group hold the same windows: keep one copy
            mine = [w for r in range(0, n_proc, tp) for w in gathered[r][1]]
        _report("processing", 2)
        if stitch:
            mine = trim_to_ownership(mine, windows)
        shaped = [(w, shape(w)) for w in mine]
        result = merge_window_results(shaped, language)
        _report("done", 3)
        return result

    # ------------------------------------------------------------------ reference surface
    def transcribe(self, audio_path: AudioInput, *, progress_callback: Optional[Callable[[str, int], None]] = None,
                   language: Optional[str] = None, speaker_count: Optional[str] = None,
                   sample_rate: Optional[int] = None, word_timestamps: bool = False) -> CTCTranscriptionResult:
        """One window (<= 40 s), no chunking: GeminiASRPipeline.transcribe (gemini_pipeline.py:474-539)."""
        if progress_callback:
            progress_callback("uploading", 0)
        wave = _resolve_audio(audio_path, sample_rate, self.engine)
        return self._transcribe_wave(wave, progress_callback=progress_callback, language=language,
                                     word_timestamps=word_timestamps, chunked=False)

    def transcribe_chunked(self, audio_path: AudioInput, *,
                           progress_callback: Optional[Callable[[str, int], None]] = None,
                           language: Optional[str] = None, speaker_count: Optional[str] = None,
                           sample_rate: Optional[int] = None, word_timestamps: bool = False) -> CTCTranscriptionResult:
        """Long audio: fixed windows, batched over the device(s), merged in order (gemini_pipeline.py:577-682)."""
        if progress_callback:
            progress_callback("uploading", 0)
        wave = _resolve_audio(audio_path, sample_rate, self.engine)
        return self._transcribe_wave(wave, progress_callback=progress_callback, language=language,
                                     word_timestamps=word_timestamps, chunked=True)

    def transcribe_with_retry(self, audio_path: AudioInput, *, max_retries: int = 3,
                              progress_callback: Optional[Callable[[str, int], None]] = None,
                              language: Optional[str] = None, speaker_count: Optional[str] = None,
                              sample_rate: Optional[int] = None, word_timestamps: bool = False) -> CTCTranscriptionResult:
        """Chunk iff the audio is longer than one window, retry runtime failures with 2^n backoff
        (gemini_pipeline.py:684-741).  Configuration/input errors (ValueError, FileNotFoundError) are not
        retried: a second attempt cannot change them."""
        if progress_callback:
            progress_callback("uploading", 0)
        wave = _resolve_audio(audio_path, sample_rate, self.engine)
        duration = len(wave) / SAMPLE_RATE
        use_chunking = duration > min(MIN_DURATION_FOR_CHUNKING, self.window_samples / SAMPLE_RATE)
        last_error: Optional[BaseException] = None
        for attempt in range(max_retries):
            try:
                return self._transcribe_wave(wave, progress_callback=progress_callback, language=language,
                                             word_timestamps=word_timestamps, chunked=use_chunking)
            except (ValueError, FileNotFoundError):
                raise
            except Exception as e:  # noqa: BLE001 - mirrors the reference's catch-all
                last_error = e
                if attempt < max_retries - 1:
                    time.sleep(2 ** attempt)
        raise RuntimeError(f"Failed to transcribe after {max_retries} attempts: {last_error}")


__all__ = [
    "CTCASRPipeline", "CTCTranscriptionResult", "CTCTranscriptSegment", "WordTimestamp", "WindowTokens",
    "build_segments", "merge_window_results", "get_audio_duration", "split_into_windows",
    "CHUNK_DURATION_SECONDS", "MIN_DURATION_FOR_CHUNKING", "MAX_PARALLEL_CHUNKS", "MAX_ALLOWED_AUDIO_SEC",
]
