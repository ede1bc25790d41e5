"""Test double for CtcEngine: same `transcribe_host` contract, computed by the CPU oracle.

Lives under tests/ only, so host logic (chunking, batching, merging, sharding) can be exercised without a GPU.
The product path never imports this."""
import numpy as np
import torch

from omnilingual_asr.models.config import CtcModelConfig
from omnilingual_asr.models.inference.ctc_engine import CtcBatchResult
from oracle import ctc_oracle as O


class OracleEngine:
    def __init__(self, name="tiny", seed=0):
        self.ocfg = O.PRESETS[name]
        self.cfg = CtcModelConfig(self.ocfg.name, self.ocfg.d_model, self.ocfg.n_layers, self.ocfg.n_heads,
                                  self.ocfg.d_ffn, vocab=self.ocfg.vocab, pos_groups=self.ocfg.pos_groups)
        self.w = O.init_weights(self.ocfg, seed)
        self.device = torch.device("cpu")
        self.calls = []

    def transcribe_host(self, wave, n_samples, *, normalised=False, return_frame_ids=False):
        arr = np.asarray(wave)
        self.calls.append((tuple(arr.shape), list(n_samples)))
        self.last_dtype = arr.dtype
        wave = torch.as_tensor(arr.astype(np.float32) / 32768.0 if arr.dtype == np.int16 else arr, dtype=torch.float32)
        if not normalised:
            wave = O.wave_layer_norm(wave, n_samples)
        out = O.forward(self.w, wave, n_samples, self.ocfg)
        ids, frames = [], []
        for b, nf in enumerate(out.n_frames):
            i, p = O.greedy_collapse(out.frame_ids[b], nf)
            ids.append(np.array(i, dtype=np.int32))
            frames.append(np.array(p, dtype=np.int32))
        return CtcBatchResult(ids, frames, out.n_frames, out.frame_ids.numpy() if return_frame_ids else None)


    # the asynchronous pair the engine pool drives (CtcEngine.submit_host / wait): computed at submit time here
    def submit_host(self, wave, n_samples, out_ids, out_frames, out_lens, *, stream=0):
        res = self.transcribe_host(wave, n_samples)
        for b, (i, f) in enumerate(zip(res.token_ids, res.token_frames)):
            out_lens[b] = len(i)
            out_ids[b, :len(i)] = i
            out_frames[b, :len(i)] = f
        self.tickets = getattr(self, "tickets", 0) + 1
        return self.tickets - 1

    def wait(self, ticket):
        return None


class FlakyEngine(OracleEngine):
    """Fails the first `fail` calls with a RuntimeError (retry path)."""
    def __init__(self, fail, **kw):
        super().__init__(**kw)
        self.fail = fail

    def transcribe_host(self, *a, **kw):
        if self.fail > 0:
            self.fail -= 1
            raise RuntimeError("injected device failure")
        return super().transcribe_host(*a, **kw)
