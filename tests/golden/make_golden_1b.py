"""Golden vector for the headline model (BASELINE configs[1] shape, one window): omniASR_CTC_1B, weights seed 0, one
30 s synthetic window (bench.synthetic_windows(1, 1234)), through the CPU oracle in both modes.
    python tests/golden/make_golden_1b.py        (about two minutes on 8 cores; writes oracle_1b_window.npz)"""
import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
import bench  # noqa: E402
from oracle import ctc_oracle as O  # noqa: E402

torch.set_num_threads(8)
cfg = O.PRESETS["omniASR_CTC_1B"]
w = O.init_weights(cfg, seed=0)
wave = bench.synthetic_windows(1, 1234)
ns = [wave.shape[1]]
wn = O.wave_layer_norm(wave, ns)
with torch.no_grad():
    emu = O.forward(w, wn, ns, cfg, emulate_bf16=True, return_logits=True)
    f32 = O.forward(w, wn, ns, cfg, emulate_bf16=False, return_logits=True)
ids_emu = emu.frame_ids[0].numpy().astype(np.int32)
ids_f32 = f32.frame_ids[0].numpy().astype(np.int32)
rows = np.arange(0, emu.n_frames[0], 16)
np.savez_compressed(Path(__file__).resolve().parent / "oracle_1b_window.npz",
                    ids_emu=ids_emu, ids_f32=ids_f32,
                    margin_emu=O.top2_margin(emu.logits[0]).numpy().astype(np.float32),
                    margin_f32=O.top2_margin(f32.logits[0]).numpy().astype(np.float32),
                    hidden_rows=rows.astype(np.int32),
                    hidden_emu=emu.hidden[0, rows].numpy().astype(np.float32),
                    hidden_f32=f32.hidden[0, rows].numpy().astype(np.float32),
                    wave_sha256=hashlib.sha256(wave.numpy().tobytes()).hexdigest(),
                    sha256=hashlib.sha256(ids_emu.tobytes()).hexdigest())
print("1B window: frames", emu.n_frames, "emu/f32 id agreement", float((ids_emu == ids_f32).mean()),
      "median margin", float(np.median(O.top2_margin(emu.logits[0]).numpy())))
