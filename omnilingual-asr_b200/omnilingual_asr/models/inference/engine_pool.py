"""Engine pool: every visible GPU driven from ONE process, one submission queue shared by all callers.

Why it exists (reference behaviour this replaces):
  * the reference's caller holds ONE pipeline object per process (workflows/wav2elan_web/app.py:38-54) and calls it from
    up to four executor threads (app.py:195-206, 384-389).  A drop-in `pipeline.transcribe(path)` therefore has to reach
    all GPUs of the box without torchrun, and concurrent calls have to share the device instead of queueing on a lock;
  * the reference keeps MAX_PARALLEL_CHUNKS = 4 window requests of a recording in flight
    (gemini_pipeline.py:217-219, 623-641).  Here the unit in flight is a BATCH of windows per GPU: two per engine
    (oasr_transcribe_host_async / oasr_wait), so that batch k + 1 crosses PCIe and waits in the stream while batch k
    computes, and the host shapes the tokens of batch k - 1 meanwhile.

Shape: `run_windows` turns one caller's windows into jobs on a shared deque and blocks on the caller's own event.  One
worker thread per engine takes up to `batch_windows` jobs - from WHICHEVER callers queued them (cross-caller batching) -
copies their samples into one of its two pinned staging buffers, submits the batch and, once two are in flight (or the
queue is empty), waits for the older one and hands its tokens back.  Windows are independent (the reference fans them
out over threads for the same reason), so any packing is legal; results are re-ordered per caller by window index.
"""
from __future__ import annotations

import collections
import threading
from dataclasses import dataclass
from typing import Any, Deque, Dict, List, Optional, Sequence, Tuple

import numpy as np

from omnilingual_asr.models.config import SAMPLE_RATE  # noqa: F401  (documented unit of start / n)


@dataclass
class WindowTokens:
    """Decoded tokens of one window, before text shaping (what travels between ranks)."""
    index: int
    start_sample: int
    n_samples: int
    n_frames: int
    token_ids: np.ndarray
    token_frames: np.ndarray


class _Request:
    """The windows of one transcribe call."""

    def __init__(self, wave: Any, windows: Sequence[Tuple[int, int]], indices: Sequence[int], dtype, post=None) -> None:
        self.wave = wave            # one recording (1-D array) or, for a list of clips, one array per window
        self.dtype = dtype
        self.post = post            # optional per-window host work (text shaping) done by the worker at delivery,
        self.extras: Dict[int, Any] = {}   # i.e. under the device step of the next batch instead of after the last one
        self.windows = windows
        self.results: Dict[int, WindowTokens] = {}
        self.remaining = len(indices)
        self.error: Optional[BaseException] = None
        self.done = threading.Event()
        if self.remaining == 0:
            self.done.set()

    def samples(self, i: int) -> np.ndarray:
        s0, n = self.windows[i]
        src = self.wave if isinstance(self.wave, np.ndarray) else self.wave[i]
        return src[s0:s0 + n]


class _Staging:
    """One of a worker's two batch slots: pinned input samples and pinned outputs."""

    def __init__(self) -> None:
        self.wave: Dict[Any, np.ndarray] = {}     # dtype -> flat pinned buffer
        self.out_ids: Optional[np.ndarray] = None
        self.out_frames: Optional[np.ndarray] = None
        self.out_lens: Optional[np.ndarray] = None
        self._keep: List[Any] = []                # the torch tensors that own the pinned memory


def _pinned_array(n: int, dtype, engine: Any, keep: List[Any]) -> np.ndarray:
    """n elements of page-locked host memory for a CUDA engine (a DMA needs it; pageable memory would make every
    'asynchronous' copy synchronous); plain memory for the CPU test doubles."""
    dev = getattr(engine, "device", None)
    if dev is not None and getattr(dev, "type", "cpu") == "cuda":
        import torch
        t = torch.empty((n,), dtype={np.dtype(np.int16): torch.int16, np.dtype(np.float32): torch.float32,
                                     np.dtype(np.int32): torch.int32}[np.dtype(dtype)], pin_memory=True)
        keep.append(t)
        return t.numpy()
    return np.empty((n,), dtype=dtype)


class EnginePool:
    """One worker thread per engine (= per GPU) behind a single queue of window jobs."""

    SLOTS = 2      # batches in flight per engine (liboasr: ASYNC_SLOTS)
    MIN_TAKE = 8   # windows: smallest batch an engine adds to work it already has in flight

    def __init__(self, engines: Sequence[Any], batch_windows: int = 32) -> None:
        if not engines:
            raise ValueError("engine pool needs at least one engine")
        if batch_windows <= 0:
            raise ValueError("batch_windows must be positive")
        self.engines = list(engines)
        self.batch_windows = int(batch_windows)
        self._cv = threading.Condition()
        self._queue: Deque[Tuple[_Request, int]] = collections.deque()
        self._stop = False
        self._inflight = [0] * len(self.engines)      # windows each engine has in flight (guarded by _cv)
        # what the tests and the stress script read: batches submitted, windows in them, batches that held windows of
        # more than one caller, batches per engine
        self.stats = {"batches": 0, "windows": 0, "mixed_batches": 0, "per_engine": [0] * len(self.engines)}
        self._threads = [threading.Thread(target=self._worker, args=(k,), name=f"oasr-engine-{k}", daemon=True)
                         for k in range(len(self.engines))]
        for t in self._threads:
            t.start()

    # ------------------------------------------------------------------ caller side
    def run_windows(self, wave: np.ndarray, windows: Sequence[Tuple[int, int]], lo: int, hi: int, post=None):
        """Windows [lo, hi) of `wave` (1-D float32 or PCM16 host array) -> their tokens, in window order.  Blocks the
        calling thread only; other callers' windows share the batches.  With `post` (WindowTokens -> anything) the
        result is a list of (tokens, post(tokens)) and post runs in the worker thread when a batch comes back."""
        if self._stop:
            raise RuntimeError("engine pool is closed")
        if wave.ndim != 1:
            raise ValueError("the pool takes one mono recording per call")
        if wave.dtype != np.int16 and wave.dtype != np.float32:
            wave = np.asarray(wave, dtype=np.float32)
        idx = list(range(lo, hi))
        req = _Request(wave, windows, idx, wave.dtype, post)
        toks = self._run(req, idx)
        return toks if post is None else [(t, req.extras[t.index]) for t in toks]

    def run_clips(self, clips: Sequence[np.ndarray]) -> List[WindowTokens]:
        """A list of independent clips (each one window long at most) -> their tokens, in input order; the clips are
        packed into batches by length like any other windows."""
        arrs = [np.asarray(c, dtype=np.float32) if np.asarray(c).dtype != np.int16 else np.asarray(c) for c in clips]
        if any(a.ndim != 1 for a in arrs):
            raise ValueError("clips must be mono")
        if len({a.dtype for a in arrs}) > 1:
            arrs = [a.astype(np.float32) / 32768.0 if a.dtype == np.int16 else a for a in arrs]
        idx = list(range(len(arrs)))
        dtype = arrs[0].dtype if arrs else np.dtype(np.float32)
        return self._run(_Request(arrs, [(0, len(a)) for a in arrs], idx, dtype), idx)

    def _run(self, req: _Request, idx: List[int]) -> List[WindowTokens]:
        if self._stop:
            raise RuntimeError("engine pool is closed")
        with self._cv:
            self._queue.extend((req, i) for i in idx)
            self._cv.notify_all()
        req.done.wait()
        if req.error is not None:
            raise req.error
        return [req.results[i] for i in idx]

    def close(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        for t in self._threads:
            if t is not threading.current_thread():
                t.join(timeout=30)

    # ------------------------------------------------------------------ worker side
    def _take(self, k: int, block: bool) -> List[Tuple[_Request, int]]:
        """Up to batch_windows jobs of one sample type, whoever queued them, for engine k.  With several engines the
        aim is that all of them finish together: an engine may hold at most its fair share of ALL outstanding windows
        (queued + in flight anywhere), so 120 windows on 8 idle GPUs leave as 15 each, a long recording keeps two full
        batches in flight per GPU while there is plenty, and at its tail nobody sits on two batches while the others
        have run dry.  A share below MIN_TAKE is not worth a device step of its own (~4 ms fixed cost): the engine
        finishes what it has first and asks again."""
        with self._cv:
            while True:
                while self._queue and self._queue[0][0].error is not None:   # its caller has been told already
                    self._queue.popleft()
                if self._queue or self._stop or not block:
                    break
                self._cv.wait()
            if not self._queue:
                return []
            outstanding = len(self._queue) + sum(self._inflight)
            fair = -(-outstanding // len(self.engines))                     # ceil(outstanding / engines)
            want = fair - self._inflight[k]
            if self._inflight[k] > 0 and want < min(self.MIN_TAKE, self.batch_windows):
                return []                                                   # has its share in flight already
            limit = max(1, min(self.batch_windows, want))
            dtype = self._queue[0][0].dtype
            taken: List[Tuple[_Request, int]] = []
            skipped: List[Tuple[_Request, int]] = []
            while self._queue and len(taken) < limit:
                req, i = self._queue.popleft()
                if req.error is not None:
                    continue
                # Any lengths share a batch (rows are padded to the longest; the marginal cost of one more padded row in
                # a batch, ~3.7 ms at 1B, is below the cost of a batch of its own); PCM16 and float32 samples do not.
                if req.dtype == dtype:
                    taken.append((req, i))
                else:
                    skipped.append((req, i))
                    if len(skipped) >= 4 * self.batch_windows:
                        break
            self._queue.extendleft(reversed(skipped))
            self._inflight[k] += len(taken)   # counted from the take on: the next engine to ask sees them as assigned
            return taken

    def _deliver(self, engine: Any, jobs: List[Tuple[_Request, int]], st: _Staging, T: int) -> None:
        lens = st.out_lens
        ids = st.out_ids[: len(jobs) * T].reshape(len(jobs), T)
        frames = st.out_frames[: len(jobs) * T].reshape(len(jobs), T)
        finished: List[_Request] = []
        for r, (req, i) in enumerate(jobs):
            s0, n = req.windows[i]
            k = int(lens[r])
            tok = WindowTokens(i, s0, n, int(engine.cfg.feature_length(int(n))), ids[r, :k].copy(), frames[r, :k].copy())
            extra = req.post(tok) if req.post is not None else None
            with self._cv:
                req.results[i] = tok
                req.extras[i] = extra
                req.remaining -= 1
                if req.remaining == 0:
                    finished.append(req)
        for req in finished:
            req.done.set()

    def _deliver_empty(self, engine: Any, jobs: List[Tuple[_Request, int]]) -> None:
        finished: List[_Request] = []
        for req, i in jobs:
            s0, n = req.windows[i]
            tok = WindowTokens(i, s0, n, 0, np.zeros((0,), np.int32), np.zeros((0,), np.int32))
            extra = req.post(tok) if req.post is not None else None
            with self._cv:
                req.results[i] = tok
                req.extras[i] = extra
                req.remaining -= 1
                if req.remaining == 0:
                    finished.append(req)
        for req in finished:
            req.done.set()

    def _fail(self, jobs: List[Tuple[_Request, int]], err: BaseException) -> None:
        for req in {id(r): r for r, _ in jobs}.values():
            with self._cv:
                first = req.error is None
                req.error = err
            if first:
                req.done.set()

    def _worker(self, k: int) -> None:
        engine = self.engines[k]
        dev = getattr(engine, "device", None)
        if dev is not None and getattr(dev, "type", "cpu") == "cuda":
            import torch
            torch.cuda.set_device(dev)
        slots = [_Staging() for _ in range(self.SLOTS)]
        inflight: Deque[Tuple[int, List[Tuple[_Request, int]], int, int]] = collections.deque()   # ticket, jobs, slot, T
        n_sub = 0
        while True:
            jobs = self._take(k, block=not inflight)
            if not jobs and not inflight:
                if self._stop:
                    return
                continue
            if jobs:
                slot = n_sub % self.SLOTS
                st = slots[slot]
                try:
                    L = max(req.windows[i][1] for req, i in jobs)
                    B = len(jobs)
                    dtype = jobs[0][0].dtype
                    if int(engine.cfg.feature_length(int(L))) == 0:
                        # nothing in this batch reaches one frame (empty recording, a stub shorter than the feature
                        # extractor's 400-sample receptive field): no tokens, and no device step to find that out
                        self._deliver_empty(engine, jobs)
                        with self._cv:
                            self._inflight[k] -= len(jobs)
                            self._cv.notify_all()
                        continue
                    T = max(int(engine.cfg.feature_length(int(L))), 1)
                    buf = st.wave.get(dtype)
                    if buf is None or buf.size < B * L:
                        buf = _pinned_array(max(B, self.batch_windows) * L, dtype, engine, st._keep)
                        st.wave[dtype] = buf
                    if st.out_ids is None or st.out_ids.size < B * T:
                        cap = max(B, self.batch_windows) * T
                        st.out_ids = _pinned_array(cap, np.int32, engine, st._keep)
                        st.out_frames = _pinned_array(cap, np.int32, engine, st._keep)
                        st.out_lens = _pinned_array(max(B, self.batch_windows), np.int32, engine, st._keep)
                    batch = buf[: B * L].reshape(B, L)
                    ns = []
                    for r, (req, i) in enumerate(jobs):
                        s0, n = req.windows[i]
                        batch[r, :n] = req.samples(i)           # samples past n are never read as signal (a8 zero-fills them)
                        ns.append(int(n))
                    ticket = engine.submit_host(batch, ns, st.out_ids[: B * T].reshape(B, T),
                                                st.out_frames[: B * T].reshape(B, T), st.out_lens)
                    inflight.append((ticket, jobs, slot, T))
                    n_sub += 1
                    with self._cv:
                        self.stats["batches"] += 1
                        self.stats["windows"] += B
                        self.stats["per_engine"][k] += 1
                        if len({id(req) for req, _ in jobs}) > 1:
                            self.stats["mixed_batches"] += 1
                except BaseException as e:  # noqa: BLE001 - the callers get the failure, the worker lives on
                    with self._cv:
                        self._inflight[k] -= len(jobs)
                    self._fail(jobs, e)
            if inflight and (len(inflight) >= self.SLOTS or not jobs):
                ticket, done_jobs, slot, T = inflight.popleft()
                try:
                    engine.wait(ticket)
                    self._deliver(engine, done_jobs, slots[slot], T)
                except BaseException as e:  # noqa: BLE001
                    self._fail(done_jobs, e)
                with self._cv:
                    self._inflight[k] -= len(done_jobs)
                    self._cv.notify_all()      # shares are re-evaluated when work completes
