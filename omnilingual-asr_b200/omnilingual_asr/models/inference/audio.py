"""Host-side audio front: decode PCM, mix to mono, resample to 16 kHz, cut fixed windows.

Replaces the ffprobe/ffmpeg subprocesses of the reference (gemini_pipeline.py:222-240 get_audio_duration,
:243-310 split_audio_into_chunks) with in-memory slicing of one 16 kHz tensor: same window law
(start = i * chunk_duration, no overlap, last window short, at least one window).
"""
from __future__ import annotations

import wave as _wave
from pathlib import Path
from typing import List, Tuple

import numpy as np

SAMPLE_RATE = 16_000


def read_wav(path: str | Path) -> Tuple[np.ndarray, int]:
    """PCM WAV (8/16/24/32-bit integer) -> (float32 [n_samples, n_channels] in [-1, 1), sample rate)."""
    with _wave.open(str(path), "rb") as wf:
        sr, nch, width, n = wf.getframerate(), wf.getnchannels(), wf.getsampwidth(), wf.getnframes()
        raw = wf.readframes(n)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = v.astype(np.float32) / 8388608.0
    else:
        raise ValueError(f"unsupported WAV sample width: {width} bytes")
    return x.reshape(-1, nch), sr


def to_mono_16k(x: np.ndarray, sr: int) -> np.ndarray:
    """[n] or [n, ch] at `sr` Hz -> float32 [m] at 16 kHz (channel mean, band-limited sinc resampling)."""
    x = np.asarray(x, dtype=np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1)
    elif x.ndim != 1:
        raise ValueError("audio must be [n] or [n, channels]")
    if sr == SAMPLE_RATE:
        return np.ascontiguousarray(x)
    if sr <= 0:
        raise ValueError("sample rate must be positive")
    import torch
    import torchaudio.functional as AF
    y = AF.resample(torch.from_numpy(np.ascontiguousarray(x)), int(sr), SAMPLE_RATE)
    return y.numpy().astype(np.float32, copy=False)


def load_audio_16k(path: str | Path) -> np.ndarray:
    p = Path(path)
    if not p.exists():
        raise FileNotFoundError(str(p))
    if p.suffix.lower() not in (".wav", ".wave"):
        raise ValueError(f"only PCM WAV files can be decoded in-process (got {p.suffix}); "
                         "pass a waveform array and its sample rate instead")
    x, sr = read_wav(p)
    return to_mono_16k(x, sr)


def get_audio_duration(audio_path: str | Path) -> float:
    """Duration in seconds; 0.0 when it cannot be determined (reference fallback, gemini_pipeline.py:238-240)."""
    try:
        with _wave.open(str(audio_path), "rb") as wf:
            return wf.getnframes() / float(wf.getframerate())
    except Exception:
        return 0.0


def split_into_windows(n_samples: int, window_samples: int) -> List[Tuple[int, int]]:
    """(start, length) of every window; restates split_audio_into_chunks (gemini_pipeline.py:243-310)."""
    if n_samples <= 0 or window_samples <= 0:
        return [(0, max(int(n_samples), 0))]
    out = []
    start = 0
    while start < n_samples:
        out.append((start, min(window_samples, n_samples - start)))
        start += window_samples
    return out


def split_into_overlapping_windows(n_samples: int, window_samples: int, overlap_samples: int) -> List[Tuple[int, int]]:
    """Windows of `window_samples` whose starts are `window - overlap` apart (the last one short).  overlap = 0 is the
    reference law above.  Used with `ownership_bounds`: a window transcribes its whole span but keeps only the tokens
    of the part it owns, so no word is cut at a window edge."""
    if overlap_samples <= 0:
        return split_into_windows(n_samples, window_samples)
    if 2 * overlap_samples > window_samples:
        # above half a window the overlaps of (i, i+1) and (i+1, i+2) intersect: a sample could be owned twice
        raise ValueError("overlap must be at most half the window")
    if n_samples <= 0:
        return [(0, max(int(n_samples), 0))]
    hop = window_samples - overlap_samples
    out = []
    start = 0
    while True:
        out.append((start, min(window_samples, n_samples - start)))
        if start + window_samples >= n_samples:
            return out
        start += hop


def ownership_bounds(windows: List[Tuple[int, int]]) -> List[Tuple[float, float]]:
    """[lo, hi) in samples that each window answers for: neighbours meet in the middle of their overlap; the first
    window owns from 0, the last to the end.  Without overlap this is every window's own span."""
    out = []
    for i, (s, n) in enumerate(windows):
        lo = 0.0 if i == 0 else (s + (windows[i - 1][0] + windows[i - 1][1])) / 2.0
        hi = float("inf") if i + 1 == len(windows) else (windows[i + 1][0] + (s + n)) / 2.0
        out.append((lo, hi))
    return out


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition [lo, hi) of window indices for one rank (merge = concat in rank order)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (rank * n_items) // world, ((rank + 1) * n_items) // world
