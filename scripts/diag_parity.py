"""Per-stage relative error of the engine against the oracle's emulated-operand taps (GPU diagnostic)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr.models.config import CtcModelConfig  # noqa: E402
from omnilingual_asr.models.inference.ctc_engine import CtcEngine  # noqa: E402
from oracle import ctc_oracle as O  # noqa: E402
from tests.golden.make_golden import golden_inputs  # noqa: E402


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


for name in sys.argv[1:] or ["tiny", "tiny80"]:
    ocfg = O.PRESETS[name]
    w = O.init_weights(ocfg, 0)
    cfg = CtcModelConfig(ocfg.name, ocfg.d_model, ocfg.n_layers, ocfg.n_heads, ocfg.d_ffn, vocab=ocfg.vocab,
                         pos_groups=ocfg.pos_groups)
    dev = torch.device("cuda", 0)
    eng = CtcEngine(cfg, device=dev)
    eng.load_state_dict(w)
    wave, ns = golden_inputs()
    B = len(ns)
    T = O.feature_length(wave.shape[1], ocfg)
    for mode in (True, False):
        ref = O.forward(w, wave, ns, ocfg, emulate_bf16=mode, taps=True, return_logits=True)
        nf = ref.n_frames
        wd = wave.to(dev)
        eng.debug_forward(wd, ns, 1, normalised=True)
        fe = eng.debug_buffer("fe")[:, :T].float().cpu()
        line = [f"fe={rel(fe, ref.taps['fe']):.2e}"]
        for stage, tap in [(2, "proj"), (3, "pos")] + [(4 + l, f"enc.{l}") for l in range(ocfg.n_layers)]:
            eng.debug_forward(wd, ns, stage, normalised=True)
            x = eng.debug_buffer("x").view(B, T, -1).cpu()
            e = max(rel(x[b, :nf[b]], ref.taps[tap][b, :nf[b]]) for b in range(B))
            line.append(f"{tap}={e:.2e}")
        res = eng.forward(wd, ns, normalised=True, return_hidden=True)
        hid = res.hidden.cpu()
        e = max(rel(hid[b, :nf[b]], ref.hidden[b, :nf[b]]) for b in range(B))
        line.append(f"hidden={e:.2e}")
        agree = np.mean(np.concatenate([res.frame_ids[b, :nf[b]] == ref.frame_ids[b, :nf[b]].numpy() for b in range(B)]))
        print(name, "emu" if mode else "f32", " ".join(line), f"ids={agree:.4f}", flush=True)
    eng.close()
