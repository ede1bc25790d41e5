"""Runs the tiny80 engine three times on the golden windows (the third forward of a small batch is a CUDA-graph replay
unless OASR_GRAPH_MAX_B=0) and saves the frame ids of the last run.    python scripts/engine_variant.py out.npy"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
import torch  # noqa: E402
from omnilingual_asr.models.config import CtcModelConfig  # noqa: E402
from omnilingual_asr.models.inference.ctc_engine import CtcEngine  # noqa: E402
from oracle import ctc_oracle as O  # noqa: E402  (weights + inputs only: this script is a test helper)
from tests.golden.make_golden import golden_inputs  # noqa: E402

o = O.PRESETS["tiny80"]
eng = CtcEngine(CtcModelConfig(o.name, o.d_model, o.n_layers, o.n_heads, o.d_ffn, vocab=o.vocab, pos_groups=o.pos_groups),
                device=torch.device("cuda", 0))
eng.load_state_dict(O.init_weights(o, seed=0))
wave, ns = golden_inputs()
x = wave.cuda()
for _ in range(3):
    res = eng.forward(x, ns, normalised=True)
np.save(sys.argv[1], np.concatenate([res.frame_ids.ravel(), np.array([eng.launch_count])]))
eng.close()
