"""CUDA-event timing of the device-side audio front end: oasr_resample (channel mean + polyphase filter) and the PCM16
window normalisation.  Algorithmic bytes = input samples + output samples."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr import _native as N  # noqa: E402

lib = N.load()
for sr, ch, seconds in ((48000, 2, 1800), (44100, 2, 1800), (22050, 1, 3600)):
    n = sr * seconds
    x = (torch.randn(n, ch, device="cuda") * 3000).to(torch.int16)
    n_out = int(lib.oasr_resample_length(n, sr, 16000))
    out = torch.empty(n_out, dtype=torch.float32, device="cuda")
    run = lambda: N.check(lib.oasr_resample(N.ptr(x), 1, n, ch, sr, 16000, N.ptr(out), n_out, N.stream_ptr()))
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = (n * ch * 2 + n_out * 4) / 1e9
    print(f"resample {sr} Hz x{ch} PCM16, {seconds} s of audio: {ms:.3f} ms, {gb / ms * 1e3:.0f} GB/s, "
          f"{seconds / (ms / 1e3):.3g} audio-s/s")
