"""Python owner of one liboasr engine handle (one per GPU).

This is the slot the reference fills with an HTTPS call (gemini_pipeline.py:512-530): a batch of
fixed-length windows goes in, token ids and their frame positions come out.  PyTorch is used only for
device memory, streams and pinned buffers; every arithmetic step runs inside liboasr's sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import List, Mapping, Optional, Sequence

import numpy as np
import torch

from omnilingual_asr import _native as N
from omnilingual_asr.models.config import CtcModelConfig, get_model_config


@dataclass
class CtcBatchResult:
    """Greedy CTC output of one batch of windows."""
    token_ids: List[np.ndarray]      # per window: collapsed ids (int32)
    token_frames: List[np.ndarray]   # per window: first frame index of each id
    n_frames: List[int]              # valid frames per window
    frame_ids: Optional[np.ndarray] = None   # [B, Tmax] per-frame arg-max (padded frames = blank)
    hidden: Optional[torch.Tensor] = None    # [B, Tmax, d] final-LayerNorm output (fp32, device)


def _as_config_struct(cfg: CtcModelConfig) -> N.OasrConfig:
    c = N.OasrConfig()
    c.d_model, c.n_layers, c.n_heads, c.d_ffn = cfg.d_model, cfg.n_layers, cfg.n_heads, cfg.d_ffn
    c.vocab, c.fe_dim, c.pos_kernel, c.pos_groups = cfg.vocab, cfg.fe_dim, cfg.pos_kernel, cfg.pos_groups
    if len(cfg.fe_layers) > 8:
        raise ValueError("at most 8 feature-extractor layers")
    c.n_fe_layers = len(cfg.fe_layers)
    for i, (ch, k, s) in enumerate(cfg.fe_layers):
        if ch != cfg.fe_dim:
            raise ValueError("all feature-extractor layers must have fe_dim channels")
        c.fe_kernel[i], c.fe_stride[i] = k, s
    c.blank_id = cfg.blank_id
    return c


class CtcEngine:
    """One CUDA engine: weights resident in HBM, forward = a8..a16 on the calling thread's stream."""

    def __init__(self, model: str | CtcModelConfig, device: Optional[torch.device | str | int] = None, *,
                 tp_rank: int = 0, tp_world: int = 1, tp_id: Optional[bytes] = None, tp_emulate: int = 0):
        """tp_world > 1: this engine is rank `tp_rank` of a tensor-parallel group (one process per GPU; `tp_id` is the
        128-byte id from `CtcEngine.tp_unique_id()` of rank 0, broadcast by the caller).  tp_emulate = W computes all
        W shards on this one GPU and sums them locally (parity check of the slicing, no communicator)."""
        self.cfg = get_model_config(model)
        self._lib = N.load()  # raises when liboasr.so is missing: no CPU fallback
        if not torch.cuda.is_available():
            raise RuntimeError("omniASR CTC engine needs a CUDA device (B200, sm_100a); none is visible")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("CtcEngine device must be a CUDA device")
        self._cstruct = _as_config_struct(self.cfg)
        self._handle = C.c_void_p(0)
        self._lock = threading.Lock()
        self._finalized = False
        with torch.cuda.device(self.device):
            N.check(self._lib.oasr_create(C.byref(self._cstruct), C.byref(self._handle)), "oasr_create")
            if tp_emulate and tp_emulate > 1:
                N.check(self._lib.oasr_tp_emulate(self._handle, int(tp_emulate)), "oasr_tp_emulate")
            elif tp_world > 1:
                if tp_id is None or len(tp_id) != 128:
                    raise ValueError("tensor parallelism needs the 128-byte id of rank 0 (CtcEngine.tp_unique_id())")
                buf = C.create_string_buffer(bytes(tp_id), 128)
                N.check(self._lib.oasr_tp_init(self._handle, int(tp_rank), int(tp_world), C.cast(buf, C.c_void_p)),
                        "oasr_tp_init")
        self.tp_world = int(tp_emulate) if tp_emulate and tp_emulate > 1 else int(tp_world)

    def tp_enable_peer_memory(self, max_batch: int, max_samples: int, group=None) -> None:
        """Switch the tensor-parallel group to the fused peer-memory kernel (all-reduce + residual + LayerNorm over
        NVLink in one launch).  Collective over `group` (torch.distributed); call before the first forward."""
        import torch.distributed as dist
        buf = C.create_string_buffer(64)
        with self._lock, torch.cuda.device(self.device):
            N.check(self._lib.oasr_tp_ipc_export(self._handle, int(max_batch), int(max_samples), C.cast(buf, C.c_void_p)),
                    "oasr_tp_ipc_export")
        mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(self.device)
        world = dist.get_world_size(group)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gathered)
        hb = C.create_string_buffer(blob, len(blob))
        with self._lock, torch.cuda.device(self.device):
            N.check(self._lib.oasr_tp_ipc_import(self._handle, C.cast(hb, C.c_void_p)), "oasr_tp_ipc_import")

    @staticmethod
    def tp_unique_id() -> bytes:
        """128-byte NCCL id for a tensor-parallel group; rank 0 creates it, the caller broadcasts it."""
        buf = C.create_string_buffer(128)
        N.check(N.load().oasr_tp_unique_id(C.cast(buf, C.c_void_p)), "oasr_tp_unique_id")
        return buf.raw

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, weights: Mapping[str, torch.Tensor | np.ndarray], finalize: bool = True) -> None:
        """Copy every tensor named in cfg.weight_shapes() into the engine (host or device sources)."""
        shapes = self.cfg.weight_shapes()
        missing = [k for k in shapes if k not in weights]
        if missing:
            raise ValueError(f"missing weights: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        with self._lock, torch.cuda.device(self.device):
            for name, shape in shapes.items():
                t = weights[name]
                if isinstance(t, np.ndarray):
                    t = torch.from_numpy(t)
                if tuple(t.shape) != tuple(shape):
                    raise ValueError(f"{name}: expected shape {shape}, got {tuple(t.shape)}")
                if t.dtype == torch.bfloat16:
                    dt = N.DTYPE_BF16
                else:
                    t = t.to(torch.float32)
                    dt = N.DTYPE_F32
                t = t.contiguous()
                shp = (C.c_int64 * len(shape))(*shape)
                N.check(self._lib.oasr_load_weight(self._handle, name.encode(), N.ptr(t), dt, shp, len(shape)),
                        f"oasr_load_weight({name})")
            if finalize:
                N.check(self._lib.oasr_finalize_weights(self._handle), "oasr_finalize_weights")
                self._finalized = True

    def finalize(self) -> None:
        with self._lock, torch.cuda.device(self.device):
            N.check(self._lib.oasr_finalize_weights(self._handle), "oasr_finalize_weights")
            self._finalized = True

    # ------------------------------------------------------------------ forward
    def feature_length(self, n_samples: int) -> int:
        return int(self._lib.oasr_feature_length(C.byref(self._cstruct), int(n_samples)))

    def forward(self, wave: torch.Tensor, n_samples: Sequence[int], *, normalised: bool = False,
                return_hidden: bool = False, return_frame_ids: bool = True) -> CtcBatchResult:
        """wave: [B, L] fp32 on this engine's device, zero padded past n_samples[b]."""
        if wave.dim() != 2 or wave.dtype != torch.float32 or wave.device != self.device:
            raise ValueError("wave must be a [B, L] float32 tensor on the engine's device")
        B, L = wave.shape
        if len(n_samples) != B:
            raise ValueError("n_samples must have one entry per window")
        if wave.stride(1) != 1:
            wave = wave.contiguous()
        T = self.feature_length(L)
        ns = (C.c_int32 * B)(*[int(v) for v in n_samples])
        with self._lock, torch.cuda.device(self.device):
            frame_ids = torch.empty((B, max(T, 1)), dtype=torch.int32, device=self.device)
            out_ids = torch.empty_like(frame_ids)
            out_frames = torch.empty_like(frame_ids)
            out_lens = torch.zeros((B,), dtype=torch.int32, device=self.device)
            hidden = (torch.empty((B, T, self.cfg.d_model), dtype=torch.float32, device=self.device)
                      if return_hidden else None)
            flags = N.FLAG_INPUT_NORMALISED if normalised else 0
            N.check(self._lib.oasr_forward_ctc(
                self._handle, N.ptr(wave), wave.stride(0), C.cast(ns, C.c_void_p), B, L, flags,
                N.ptr(frame_ids), N.ptr(hidden), N.ptr(out_ids), N.ptr(out_frames), N.ptr(out_lens),
                N.stream_ptr()), "oasr_forward_ctc")
            lens = out_lens.cpu().numpy()          # synchronises the stream
            ids_h = out_ids.cpu().numpy()
            frames_h = out_frames.cpu().numpy()
            fids = frame_ids[:, :T].cpu().numpy() if return_frame_ids else None
        return CtcBatchResult(
            token_ids=[ids_h[b, :lens[b]].copy() for b in range(B)],
            token_frames=[frames_h[b, :lens[b]].copy() for b in range(B)],
            n_frames=[self.cfg.feature_length(int(v)) for v in n_samples],
            frame_ids=fids, hidden=hidden)

    def transcribe_host(self, wave: np.ndarray | torch.Tensor, n_samples: Sequence[int], *,
                        normalised: bool = False, return_frame_ids: bool = False) -> CtcBatchResult:
        """Same from host memory through oasr_transcribe_host (H2D + forward + D2H inside the call)."""
        flags = N.FLAG_INPUT_NORMALISED if normalised else 0
        if isinstance(wave, torch.Tensor):
            if wave.device.type != "cpu" or wave.dtype not in (torch.float32, torch.int16) or wave.dim() != 2 \
                    or wave.stride(1) != 1:
                raise ValueError("wave must be a [B, L] float32 (or PCM16 int16) CPU tensor with unit inner stride")
            B, L = wave.shape
            stride = wave.stride(0)
            if wave.dtype == torch.int16:
                flags |= N.FLAG_INPUT_I16
        else:
            wave = np.asarray(wave)
            if wave.ndim != 2:
                raise ValueError("wave must be [B, L]")
            if wave.dtype != np.int16:
                wave = np.asarray(wave, dtype=np.float32)
            if wave.strides[1] != wave.itemsize or wave.strides[0] % wave.itemsize:
                wave = np.ascontiguousarray(wave)
            B, L = wave.shape
            stride = wave.strides[0] // wave.itemsize      # rows may be a strided view of one long recording
            if wave.dtype == np.int16:                      # PCM16: converted on the device (half the H2D bytes)
                flags |= N.FLAG_INPUT_I16
        if (flags & N.FLAG_INPUT_I16) and normalised:
            raise ValueError("PCM16 input cannot be flagged as normalised")
        T = self.feature_length(L)
        ns = (C.c_int32 * B)(*[int(v) for v in n_samples])
        out_ids = np.empty((B, max(T, 1)), dtype=np.int32)
        out_frames = np.empty_like(out_ids)
        out_lens = np.zeros((B,), dtype=np.int32)
        fids = np.empty_like(out_ids) if return_frame_ids else None
        with self._lock, torch.cuda.device(self.device):
            N.check(self._lib.oasr_transcribe_host(
                self._handle, N.ptr(wave), stride, C.cast(ns, C.c_void_p), B, L,
                flags, N.ptr(out_ids), N.ptr(out_frames), N.ptr(out_lens),
                N.ptr(fids), N.stream_ptr()), "oasr_transcribe_host")
        return CtcBatchResult(
            token_ids=[out_ids[b, :out_lens[b]].copy() for b in range(B)],
            token_frames=[out_frames[b, :out_lens[b]].copy() for b in range(B)],
            n_frames=[self.cfg.feature_length(int(v)) for v in n_samples],
            frame_ids=fids[:, :T] if fids is not None else None)

    # ------------------------------------------------------------------ asynchronous host loop (engine_pool.py)
    def submit_host(self, wave: np.ndarray, n_samples: Sequence[int], out_ids: np.ndarray, out_frames: np.ndarray,
                    out_lens: np.ndarray, *, stream=0) -> int:
        """oasr_transcribe_host_async: enqueue H2D + forward + D2H of a [B, L] batch held in PINNED host memory and
        return a ticket at once (at most two outstanding per engine).  `wave` (float32 or PCM16) and the three output
        arrays (int32, pinned, [B, >= T] / [B]) belong to the engine until `wait(ticket)` returns.  stream = 0: the
        engine's own non-blocking stream."""
        if wave.ndim != 2 or wave.dtype not in (np.float32, np.int16) or wave.strides[1] != wave.itemsize:
            raise ValueError("wave must be a [B, L] float32 or int16 array with unit inner stride")
        B, L = wave.shape
        if len(n_samples) != B:
            raise ValueError("n_samples must have one entry per window")
        T = max(self.feature_length(L), 1)
        for a, shape in ((out_ids, (B, T)), (out_frames, (B, T))):
            if a.dtype != np.int32 or a.ndim != 2 or a.shape[0] < B or a.shape[1] != T or not a.flags.c_contiguous:
                raise ValueError(f"output arrays must be C-contiguous int32 [>= {B}, {T}]")
        if out_lens.dtype != np.int32 or out_lens.size < B:
            raise ValueError("out_lens must be int32 [>= B]")
        flags = N.FLAG_INPUT_I16 if wave.dtype == np.int16 else 0
        ns = (C.c_int32 * B)(*[int(v) for v in n_samples])
        ticket = C.c_int64(-1)
        with self._lock, torch.cuda.device(self.device):
            N.check(self._lib.oasr_transcribe_host_async(
                self._handle, N.ptr(wave), wave.strides[0] // wave.itemsize, C.cast(ns, C.c_void_p), B, L, flags,
                N.ptr(out_ids), N.ptr(out_frames), N.ptr(out_lens), C.c_void_p(0),
                C.c_void_p(int(getattr(stream, "cuda_stream", stream) or 0)), C.byref(ticket)),
                "oasr_transcribe_host_async")
        return int(ticket.value)

    def wait(self, ticket: int) -> None:
        """oasr_wait: block (GIL released) until the ticket's outputs are in host memory.  Not under the engine lock:
        it only waits for an event, so another thread may submit meanwhile."""
        with torch.cuda.device(self.device):
            N.check(self._lib.oasr_wait(self._handle, int(ticket)), "oasr_wait")

    # ------------------------------------------------------------------ device-side audio front end
    def resample_to_model_rate(self, samples: np.ndarray | torch.Tensor, sample_rate: int) -> torch.Tensor:
        """[n] or [n, channels] fp32 / PCM16 samples at `sample_rate` -> mono fp32 DEVICE tensor at 16 kHz
        (oasr_resample: channel mean + the windowed-sinc filter of torchaudio.functional.resample's defaults)."""
        from omnilingual_asr.models.config import SAMPLE_RATE
        if sample_rate <= 0:
            raise ValueError("sample rate must be positive")
        t = samples if isinstance(samples, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(samples))
        if t.dim() == 1:
            t = t[:, None]
        if t.dim() != 2 or t.shape[1] > 8:
            raise ValueError("audio must be [n] or [n, channels] with at most 8 channels")
        if t.dtype != torch.int16:
            t = t.to(torch.float32)
        with self._lock, torch.cuda.device(self.device):
            t = t.contiguous().to(self.device, non_blocking=True)
            n_in, ch = int(t.shape[0]), int(t.shape[1])
            n_out = int(self._lib.oasr_resample_length(n_in, int(sample_rate), SAMPLE_RATE))
            out = torch.empty((n_out,), dtype=torch.float32, device=self.device)
            N.check(self._lib.oasr_resample(N.ptr(t), 1 if t.dtype == torch.int16 else 0, n_in, ch, int(sample_rate),
                                            SAMPLE_RATE, N.ptr(out), n_out, N.stream_ptr()), "oasr_resample")
        return out

    # ------------------------------------------------------------------ debug / accounting
    def debug_forward(self, wave: torch.Tensor, n_samples: Sequence[int], stop_stage: int, *,
                      normalised: bool = False) -> None:
        B, L = wave.shape
        ns = (C.c_int32 * B)(*[int(v) for v in n_samples])
        with self._lock, torch.cuda.device(self.device):
            N.check(self._lib.oasr_debug_forward(
                self._handle, N.ptr(wave), wave.stride(0), C.cast(ns, C.c_void_p), B, L,
                N.FLAG_INPUT_NORMALISED if normalised else 0, stop_stage, N.stream_ptr()), "oasr_debug_forward")
            torch.cuda.synchronize(self.device)

    def debug_buffer(self, name: str) -> torch.Tensor:
        """Copy of an internal device buffer (see oasr_debug_buffer) as a torch tensor."""
        p = C.c_void_p(0)
        shape = (C.c_int64 * 4)()
        dt = C.c_int32(0)
        N.check(self._lib.oasr_debug_buffer(self._handle, name.encode(), C.byref(p), shape, C.byref(dt)),
                "oasr_debug_buffer")
        dims = [int(v) for v in shape if v > 0]
        dtype = torch.bfloat16 if dt.value == N.DTYPE_BF16 else torch.float32
        out = torch.empty(dims, dtype=dtype, device=self.device)
        nbytes = out.numel() * out.element_size()
        with torch.cuda.device(self.device):
            N.check(self._lib.oasr_debug_copy(self._handle, name.encode(), N.ptr(out), nbytes), "oasr_debug_copy")
        return out

    def profile(self, on: bool) -> None:
        N.check(self._lib.oasr_profile_enable(self._handle, 1 if on else 0), "oasr_profile_enable")

    def profile_read(self) -> dict:
        """{stage: (milliseconds, launches)} accumulated since the last read (device time, CUDA events)."""
        n = len(N.PROF_CATEGORIES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        with self._lock, torch.cuda.device(self.device):
            N.check(self._lib.oasr_profile_read(self._handle, ms, cnt, n), "oasr_profile_read")
        return {N.PROF_CATEGORIES[i]: (float(ms[i]), int(cnt[i])) for i in range(n) if N.PROF_CATEGORIES[i] != "end"}

    @property
    def launch_count(self) -> int:
        return int(self._lib.oasr_launch_count(self._handle))

    def close(self) -> None:
        if getattr(self, "_handle", None) and self._handle.value:
            with torch.cuda.device(self.device):
                self._lib.oasr_destroy(self._handle)
            self._handle = C.c_void_p(0)

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass
