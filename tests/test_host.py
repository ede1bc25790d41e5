"""Host-side logic behind the reference's pipeline API, on the CPU with an oracle-backed engine double."""
import threading
import wave as wavelib

import numpy as np
import pytest

import omnilingual_asr
from omnilingual_asr import CTCASRPipeline, CTCTranscriptionPipeline, DiarizedTranscriptSegment, WordTimestamp
from omnilingual_asr.models.inference import ctc_pipeline as P
from omnilingual_asr.models.inference.audio import (get_audio_duration, load_audio_16k, read_wav, shard_range,
                                                    split_into_windows, to_mono_16k)
from omnilingual_asr.models.inference.pipeline import ASRInferencePipeline
from omnilingual_asr.models.inference.tokenizer import WORD_BOUNDARY, CtcVocabulary
from oracle import ctc_oracle as O
from tests._fake_engine import FlakyEngine, OracleEngine


def noise(seconds, seed=0, sr=16000):
    return np.random.default_rng(seed).standard_normal(int(seconds * sr)).astype(np.float32)


@pytest.fixture(scope="module")
def engine():
    return OracleEngine("tiny")


def test_public_names_and_version():
    assert omnilingual_asr.__version__ == "0.2.0"
    for n in ("CTCASRPipeline", "CTCTranscriptionResult", "CTCTranscriptSegment", "CTCTranscriptionPipeline",
              "DiarizedTranscriptSegment", "WordTimestamp"):
        assert hasattr(omnilingual_asr, n)
    seg = DiarizedTranscriptSegment(0.0, 1.0, "Speaker 1", "hi")
    assert (seg.words, seg.language, seg.language_code, seg.languages, seg.emotion, seg.translation) == (None,) * 6
    with pytest.raises(Exception):
        seg.start = 2.0  # frozen, like the reference's record


def test_constructor_errors_are_value_errors():
    with pytest.raises(ValueError, match="no weights"):
        CTCTranscriptionPipeline()
    with pytest.raises(ValueError, match="unknown CTC model card"):
        CTCASRPipeline("omniASR_CTC_9B", weights="random")
    with pytest.raises(ValueError):
        CTCASRPipeline("omniASR_CTC_1B", engine=object(), window_seconds=45)
    with pytest.raises(ValueError, match="vocabulary"):
        CTCASRPipeline(OracleEngine("tiny").cfg, engine=object(), vocabulary=["a", "b"])


def test_window_law_matches_reference_chunker():
    assert split_into_windows(0, 480000) == [(0, 0)]
    w = split_into_windows(3600 * 16000, 480000)
    assert len(w) == 120 and w[7] == (7 * 480000, 480000)
    w = split_into_windows(34200 * 16000, 480000)
    assert len(w) == 1140
    w = split_into_windows(281233, 480000)
    assert w == [(0, 281233)]
    w = split_into_windows(1_000_000, 480000)
    assert w == [(0, 480000), (480000, 480000), (960000, 40000)]
    assert w == O.split_into_windows(1_000_000, 480000)


def test_shard_ranges_cover_everything_in_order():
    for n in (0, 1, 7, 120, 1140):
        for world in (1, 2, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1


def test_wav_reader_and_resampler(tmp_path):
    sr = 22050
    t = np.arange(sr) / sr
    x = (0.5 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    st = np.stack([x, -x * 0.5], axis=1)
    p = tmp_path / "a.wav"
    with wavelib.open(str(p), "wb") as wf:
        wf.setnchannels(2)
        wf.setsampwidth(2)
        wf.setframerate(sr)
        wf.writeframes((st * 32767).astype("<i2").tobytes())
    y, r = read_wav(p)
    assert r == sr and y.shape == (sr, 2)
    assert abs(get_audio_duration(p) - 1.0) < 1e-9
    assert get_audio_duration(tmp_path / "missing.wav") == 0.0
    m = load_audio_16k(p)
    assert m.dtype == np.float32 and abs(len(m) - 16000) <= 1
    assert np.allclose(to_mono_16k(x, 16000), x)
    with pytest.raises(ValueError):
        load_audio_16k(tmp_path / "a.mp3") if (tmp_path / "a.mp3").write_bytes(b"x") else None
    with pytest.raises(FileNotFoundError):
        load_audio_16k(tmp_path / "nope.wav")


def test_tokenizer_decode_and_words():
    v = CtcVocabulary.synthetic(300)
    ids = [4, 5, 6, 4, 7, 0, 8]          # "▁ab▁cd" with a special in between
    assert v.decode(ids) == "ab cd"
    words = v.words_with_frames(ids, [0, 3, 5, 9, 11, 12, 20])
    assert words == [("ab", 3, 5), ("cd", 11, 20)]
    assert v.decode([0, 1, 2, 3]) == ""
    assert len(v) == 300 and v.pieces[4] == WORD_BOUNDARY


def test_single_window_segments_match_oracle(engine):
    pipe = CTCASRPipeline(engine.cfg, engine=engine)
    x = noise(3.0, 1)
    res = pipe.transcribe(x, word_timestamps=True)
    wave = O.wave_layer_norm(__import__("torch").from_numpy(x)[None], [len(x)])
    out = O.forward(engine.w, wave, [len(x)], engine.ocfg)
    ids, pos = O.greedy_collapse(out.frame_ids[0], out.n_frames[0])
    assert len(res.segments) == 1
    seg = res.segments[0]
    assert seg.text == pipe.vocab.decode(ids)
    fd = 3.0 / out.n_frames[0]
    assert seg.start == pytest.approx(pos[0] * fd) and seg.end == pytest.approx((pos[-1] + 1) * fd)
    assert seg.speaker == "Speaker 1" and seg.words is not None
    assert all(isinstance(w, P.WordTimestamp) and seg.start <= w.start <= w.end <= seg.end + 1e-9 for w in seg.words)


def test_long_audio_is_chunked_rebased_and_merged_in_order(engine):
    pipe = CTCASRPipeline(engine.cfg, engine=engine, window_seconds=2.0, batch_windows=2)
    x = noise(5.5, 2)
    steps = []
    res = pipe.transcribe_with_retry(x, progress_callback=lambda s, i: steps.append((s, i)))
    assert steps == [("uploading", 0), ("transcribing", 1), ("processing", 2), ("done", 3)]
    # 3 windows (2 s, 2 s, 1.5 s) in 2 device steps
    assert [c[0][0] for c in engine.calls[-2:]] == [2, 1]
    starts = [s.start for s in res.segments]
    assert starts == sorted(starts) and len(res.segments) == 3
    for i, seg in enumerate(res.segments):
        assert 2.0 * i <= seg.start < seg.end <= min(2.0 * (i + 1), 5.5) + 1e-6
    # each window equals the single-window transcription of the same samples, rebased by its start offset
    for i, seg in enumerate(res.segments):
        one = pipe.transcribe(x[i * 32000:(i + 1) * 32000]).segments[0]
        assert one.text == seg.text
        assert seg.start == pytest.approx(one.start + 2.0 * i)


def test_too_long_for_unchunked_entry_point(engine):
    pipe = CTCASRPipeline(engine.cfg, engine=engine)
    with pytest.raises(ValueError, match="needs chunking"):
        pipe.transcribe(np.zeros(41 * 16000, dtype=np.float32))


def test_retry_then_success_and_exhaustion(monkeypatch):
    monkeypatch.setattr(P.time, "sleep", lambda s: None)
    eng = FlakyEngine(2, name="tiny")
    pipe = CTCASRPipeline(eng.cfg, engine=eng)
    assert pipe.transcribe_with_retry(noise(1.0)).segments
    eng = FlakyEngine(5, name="tiny")
    pipe = CTCASRPipeline(eng.cfg, engine=eng)
    with pytest.raises(RuntimeError, match="Failed to transcribe after 3 attempts: injected device failure"):
        pipe.transcribe_with_retry(noise(1.0))


def test_boundary_pipeline_surface_and_thread_local_summary(engine, tmp_path):
    pipe = CTCTranscriptionPipeline(model_card=engine.cfg, engine=engine)
    assert pipe.summary is None and pipe.detected_languages is None
    p = tmp_path / "x.wav"
    with wavelib.open(str(p), "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(16000)
        wf.writeframes((np.clip(noise(1.5, 3) * 0.2, -1, 1) * 32767).astype("<i2").tobytes())
    segs = pipe.transcribe(str(p), word_timestamps=True, language="en", speaker_count="2", unknown_kwarg=1)
    assert segs and all(isinstance(s, DiarizedTranscriptSegment) for s in segs)
    assert all(isinstance(w, WordTimestamp) for w in segs[0].words)
    assert segs[0].language_code == "en" and pipe.detected_languages == [{"name": "en", "code": "en"}]
    assert "segment" in pipe.summary
    seen = {}

    def worker(name, secs):
        pipe.transcribe(noise(secs, 4))
        seen[name] = pipe.summary

    ts = [threading.Thread(target=worker, args=(i, 0.5 + 0.5 * i)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(f"{0.5 + 0.5 * i:.2f} s" in seen[i] for i in range(4))   # no cross-thread overwrite


def test_asr_inference_pipeline_shim(engine):
    shim = ASRInferencePipeline(model_card=engine.cfg, engine=engine)
    a, b = noise(1.0, 5), noise(2.0, 6)
    texts = shim.transcribe([a, {"waveform": b, "sample_rate": 16000}], batch_size=2)
    pipe = CTCASRPipeline(engine.cfg, engine=engine)
    # batched with zero padding == alone (the engine masks padded frames)
    assert texts[1] == pipe.transcribe(b).segments[0].text
    assert isinstance(texts[0], str)
    with pytest.raises(ValueError):
        shim.transcribe("not-a-list")
    with pytest.raises(ValueError):
        shim.transcribe([np.zeros(41 * 16000, dtype=np.float32)])


def test_split_on_silence_shapes_segments():
    v = CtcVocabulary.synthetic(300)
    win = P.WindowTokens(0, 0, 480000, 1499, np.array([5, 6, 7, 8], np.int32), np.array([10, 12, 400, 405], np.int32))
    one = P.build_segments(win, v, word_timestamps=False, split_gap_sec=None, language=None)
    two = P.build_segments(win, v, word_timestamps=False, split_gap_sec=1.0, language=None)
    assert len(one) == 1 and len(two) == 2 and two[0].text == "ab" and two[1].text == "cd"
    assert P.build_segments(P.WindowTokens(0, 0, 100, 0, np.array([], np.int32), np.array([], np.int32)), v,
                            word_timestamps=True, split_gap_sec=None, language=None) == []


def test_pcm16_recording_stays_int16(engine):
    """Mono PCM16 at 16 kHz is handed to the engine as it is (the device converts it: OASR_FLAG_INPUT_I16), two windows
    per batch; the result equals the float32 route."""
    from omnilingual_asr.models.inference.ctc_pipeline import CTCASRPipeline
    rng = np.random.default_rng(7)
    win = 8000
    pcm = (rng.standard_normal(3 * win + 1234) * 3000).astype(np.int16)
    pipe = CTCASRPipeline(engine.cfg, engine=engine, window_seconds=win / 16000, batch_windows=2, distributed=False)
    a = pipe.transcribe_chunked(pcm, sample_rate=16000)
    assert engine.last_dtype == np.int16
    assert [c[0] for c in engine.calls[-2:]] == [(2, win), (2, win)]      # [w0, w1], [w2, ragged tail padded to the window]
    b = pipe.transcribe_chunked(pcm.astype(np.float32) / 32768.0, sample_rate=16000)
    assert engine.last_dtype == np.float32
    assert [(s.start, s.end, s.text) for s in a.segments] == [(s.start, s.end, s.text) for s in b.segments]


def test_overlapping_window_law_and_ownership():
    from omnilingual_asr.models.inference.audio import ownership_bounds, split_into_overlapping_windows
    assert split_into_overlapping_windows(100, 40, 0) == [(0, 40), (40, 40), (80, 20)]          # the reference law
    w = split_into_overlapping_windows(100, 40, 10)
    assert w == [(0, 40), (30, 40), (60, 40)]                                                     # hop 30, ends at 100
    assert split_into_overlapping_windows(101, 40, 10)[-1] == (90, 11)
    assert split_into_overlapping_windows(40, 40, 10) == [(0, 40)]
    b = ownership_bounds(w)
    assert b[0] == (0.0, 35.0) and b[1] == (35.0, 65.0) and b[2][0] == 65.0 and b[2][1] == float("inf")
    with pytest.raises(ValueError):
        split_into_overlapping_windows(100, 40, 40)


def test_overlap_and_stitch_keeps_each_frame_once(engine):
    """With overlap every window transcribes its whole span, but a token survives only in the window that owns its
    frame: the kept tokens are exactly the per-window oracle tokens filtered by the stitch bounds."""
    from omnilingual_asr.models.inference.audio import split_into_overlapping_windows
    import torch
    wave = noise(2.5, seed=21)
    engine.calls.clear()
    pipe = CTCASRPipeline(engine.cfg, engine=engine, window_seconds=1.0, overlap_seconds=0.25, distributed=False)
    res = pipe.transcribe_chunked(wave, sample_rate=16000, word_timestamps=False)
    windows = split_into_overlapping_windows(40000, 16000, 4000)
    assert engine.calls[0][1][0] == 16000 and len(windows) == 3 and windows[1] == (12000, 16000)
    toks = []
    for i, (s0, n) in enumerate(windows):
        x = torch.from_numpy(wave[s0:s0 + n]).float()[None]
        out = O.forward(engine.w, O.wave_layer_norm(x, [n]), [n], engine.ocfg)
        ids, frames = O.greedy_collapse(out.frame_ids[0], out.n_frames[0])
        toks.append(P.WindowTokens(i, s0, n, out.n_frames[0], np.array(ids, np.int32), np.array(frames, np.int32)))
    bounds = P.stitch_bounds(toks, windows)
    assert bounds[0][0] == 0.0 and bounds[-1][1] == float("inf")
    for i in range(len(windows) - 1):                      # cuts are shared and lie inside the overlaps
        assert bounds[i][1] == bounds[i + 1][0]
        assert windows[i + 1][0] <= bounds[i][1] <= windows[i][0] + windows[i][1]
    expect = []
    for w, (lo, hi) in zip(toks, bounds):
        c = P._token_centres(w)
        expect += [int(t) for t, k in zip(w.token_ids, (c >= lo) & (c < hi)) if k]
    got = "".join(seg.text for seg in res.segments)
    want = pipe.vocab.decode(np.array(expect, dtype=np.int32))
    assert got.replace(" ", "") == want.replace(" ", "")
    starts = [seg.start for seg in res.segments]
    assert starts == sorted(starts)
    with pytest.raises(ValueError):
        CTCASRPipeline(engine.cfg, engine=engine, window_seconds=1.0, overlap_seconds=1.0)


def test_stitch_cut_moves_to_the_common_silence():
    """Two windows of 100 samples overlapping on [60, 100): both see a token near 70 (window 0 at 69, window 1 at 71 -
    the geometric middle, 80, is fine here, but a token pair straddling 80 would be doubled or lost) and silence on
    [72, 95).  The cut goes to the middle of that silence and each token is kept exactly once."""
    W = P.WindowTokens
    windows = [(0, 100), (60, 100)]
    # 10 frames of 10 samples each; frame f of window w is centred at start + 10 f + 5
    w0 = W(0, 0, 100, 10, np.array([5, 6, 7], np.int32), np.array([2, 6, 7], np.int32))     # centres 25, 65, 75
    w1 = W(1, 60, 100, 10, np.array([6, 8, 9], np.int32), np.array([0, 1, 5], np.int32))    # centres 65, 75, 115
    lo_hi = P.stitch_bounds([w0, w1], windows)
    cut = lo_hi[0][1]
    assert cut == lo_hi[1][0] == 0.5 * (75 + 100)          # longest token-free stretch inside [60, 100) is [75, 100)
    out = P.trim_to_ownership([w0, w1], windows)
    assert list(out[0].token_ids) == [5, 6, 7] and list(out[1].token_ids) == [9]
    # without tokens in the overlap the geometric middle stays
    e0 = W(0, 0, 100, 10, np.array([5], np.int32), np.array([1], np.int32))
    e1 = W(1, 60, 100, 10, np.array([9], np.int32), np.array([8], np.int32))
    assert P.stitch_bounds([e0, e1], windows)[0][1] == 80.0


def test_overlap_above_half_a_window_is_rejected_and_half_is_safe(engine):
    """Above half a window the overlaps of (i, i+1) and (i+1, i+2) intersect and two cuts could cross:
    rejected at construction and by the window law.  At exactly half a window every sample still has one owner: the
    cuts are non-decreasing and no token is kept twice."""
    from omnilingual_asr.models.inference.audio import split_into_overlapping_windows
    with pytest.raises(ValueError):
        CTCASRPipeline(engine.cfg, engine=engine, window_seconds=1.0, overlap_seconds=0.75)
    with pytest.raises(ValueError):
        split_into_overlapping_windows(100, 40, 30)
    W = P.WindowTokens
    windows = split_into_overlapping_windows(100, 40, 20)      # starts 0, 20, 40, 60: overlaps [20,40) [40,60) [60,80)
    assert windows == [(0, 40), (20, 40), (40, 40), (60, 40)]
    rng = np.random.default_rng(5)
    toks = []
    for i, (s0, n) in enumerate(windows):                      # 40 frames of one sample; random tokens everywhere
        frames = np.sort(rng.choice(40, size=12, replace=False)).astype(np.int32)
        toks.append(W(i, s0, n, 40, rng.integers(1, 9, size=12).astype(np.int32), frames))
    bounds = P.stitch_bounds(toks, windows)
    cuts = [b[1] for b in bounds[:-1]]
    assert cuts == sorted(cuts) and all(bounds[i][1] == bounds[i + 1][0] for i in range(3))
    kept = P.trim_to_ownership(toks, windows)
    centres = np.concatenate([P._token_centres(w) for w in kept])
    owners = np.concatenate([np.full(len(w.token_ids), w.index) for w in kept])
    for c, o in zip(centres, owners):                          # every kept token lies in its owner's span only
        lo, hi = bounds[o]
        assert lo <= c < hi
        assert sum(1 for (l2, h2) in bounds if l2 <= c < h2) == 1


# ------------------------------------------------------------------------------------------- engine pool
def _texts(res):
    return [(s.start, s.end, s.text) for s in res.segments]


def test_pool_packs_windows_of_concurrent_callers_into_one_batch():
    """SURVEY 3b / VERDICT r1 item 4: concurrent transcribe calls on ONE pipeline (workflows/wav2elan_web/app.py:38-54,
    384-389) share batches instead of queueing on a lock.  The first caller's batch is held inside the engine until two
    more callers have queued their windows: those must then leave in ONE batch, and every caller gets exactly what a
    lone call returns."""
    import threading
    import time
    from tests._fake_engine import OracleEngine

    class GatedEngine(OracleEngine):
        def __init__(self):
            super().__init__("tiny")
            self.gate = threading.Event()
            self.first = True

        def submit_host(self, *a, **kw):
            if self.first:
                self.first = False
                assert self.gate.wait(60)
            return super().submit_host(*a, **kw)

    eng = GatedEngine()
    pipe = CTCASRPipeline(eng.cfg, engine=eng, window_seconds=0.5, batch_windows=4, distributed=False)
    clips = [noise(1.0, seed=30 + i) for i in range(3)]          # two windows each
    lone = OracleEngine("tiny")
    lone_pipe = CTCASRPipeline(lone.cfg, engine=lone, window_seconds=0.5, batch_windows=4, distributed=False)
    want = [_texts(lone_pipe.transcribe_chunked(c, sample_rate=16000)) for c in clips]
    got = [None] * 3

    def work(i):
        got[i] = _texts(pipe.transcribe_chunked(clips[i], sample_rate=16000))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(3)]
    ts[0].start()
    while eng.first:                       # caller 0's batch is inside the engine
        time.sleep(0.01)
    ts[1].start()
    ts[2].start()
    deadline = time.time() + 60
    while len(pipe.pool._queue) < 4 and time.time() < deadline:
        time.sleep(0.01)
    eng.gate.set()
    [t.join(120) for t in ts]
    assert got == want
    assert [c[0][0] for c in eng.calls] == [2, 4]               # caller 0 alone, then callers 1 and 2 together
    assert pipe.pool.stats["mixed_batches"] == 1 and pipe.pool.stats["windows"] == 6
    pipe.close()
    lone_pipe.close()


def test_pool_spreads_one_recording_over_several_engines():
    """One pipeline object, several engines (one per GPU in production: devices="all"): the windows of ONE recording
    are served by all of them from a single process - no torchrun, no process group - and the transcript equals the
    single-engine one."""
    from tests._fake_engine import OracleEngine
    engs = [OracleEngine("tiny") for _ in range(3)]
    pipe = CTCASRPipeline(engs[0].cfg, engines=engs, window_seconds=0.5, batch_windows=2, distributed=False)
    x = noise(6.2, seed=44)                                       # 13 windows
    res = pipe.transcribe_chunked(x, sample_rate=16000)
    single = OracleEngine("tiny")
    sp = CTCASRPipeline(single.cfg, engine=single, window_seconds=0.5, batch_windows=2, distributed=False)
    assert _texts(res) == _texts(sp.transcribe_chunked(x, sample_rate=16000))
    per = [sum(c[0][0] for c in e.calls) for e in engs]
    assert sum(per) == 13 and all(n > 0 for n in per)            # every engine took part
    assert pipe.pool.stats["per_engine"] == [len(e.calls) for e in engs]
    pipe.close()
    sp.close()


def test_pool_failure_reaches_the_caller_and_the_pool_lives_on():
    from tests._fake_engine import FlakyEngine
    eng = FlakyEngine(1, name="tiny")
    pipe = CTCASRPipeline(eng.cfg, engine=eng, window_seconds=0.5, batch_windows=4, distributed=False)
    x = noise(1.0, seed=5)
    with pytest.raises(RuntimeError, match="injected device failure"):
        pipe.transcribe_chunked(x, sample_rate=16000)
    assert len(pipe.transcribe_chunked(x, sample_rate=16000).segments) > 0      # next call is served
    pipe.close()
    with pytest.raises(RuntimeError):
        pipe.transcribe_chunked(x, sample_rate=16000)                            # a closed pool refuses work


def test_empty_and_sub_frame_audio_give_no_segments(engine):
    """Edge cases of the boundary: an empty recording is one empty window (the reference's chunker always yields at
    least one chunk, gemini_pipeline.py:243-310), and audio shorter than the feature extractor's receptive field (400
    samples) has no frame at all: both return no segments instead of failing."""
    pipe = CTCASRPipeline(engine.cfg, engine=engine, window_seconds=1.0, batch_windows=4, distributed=False)
    for x in (np.zeros((0,), np.float32), noise(399 / 16000.0, seed=1)[:399], np.zeros((100,), np.int16)):
        res = pipe.transcribe_chunked(x, sample_rate=16000)
        assert res.segments == []
    # a recording whose LAST window is below one frame: the full windows are transcribed, the stub adds nothing
    x = noise(2.0, seed=2)
    x = np.concatenate([x, np.zeros((120,), np.float32)])
    a = pipe.transcribe_chunked(x, sample_rate=16000)
    b = pipe.transcribe_chunked(x[:32000], sample_rate=16000)
    assert _texts(a) == _texts(b) and len(a.segments) > 0
    pipe.close()


def test_pool_run_clips_keeps_input_order_and_mixes_sample_types():
    """ASRInferencePipeline's path: independent clips of different lengths (and PCM16 beside float32) come back in
    input order, each equal to its own single-clip result."""
    from omnilingual_asr.models.inference.engine_pool import EnginePool
    from tests._fake_engine import OracleEngine
    eng = OracleEngine("tiny")
    pool = EnginePool([eng], batch_windows=3)
    clips = [noise(0.3 + 0.11 * i, seed=60 + i) for i in range(5)]
    clips[2] = (clips[2] * 20000).astype(np.int16)
    got = pool.run_clips(clips)
    assert [t.index for t in got] == list(range(5))
    for c, t in zip(clips, got):
        alone = pool.run_clips([c])[0]
        assert t.token_ids.tolist() == alone.token_ids.tolist() and t.n_samples == len(c)
    assert pool.run_clips([]) == []
    pool.close()
