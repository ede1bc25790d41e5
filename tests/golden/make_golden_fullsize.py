"""Full-size pins of the oracle and golden vectors for the full-depth GPU tests.

    python tests/golden/make_golden_fullsize.py [300m] [1b] [3b]      (build container; ~8 min and ~40 GB of RAM for all)

The reference holds no code, test or vector for this path, so the strongest pin this box allows is an
INDEPENDENT implementation of the same graph at the REAL sizes: transformers' Wav2Vec2ForCTC
(modeling_wav2vec2.py:612-655 encoder layer, 730-803 stable-layer-norm encoder, 275-299 positional conv, 326-379 conv
feature layers, 1697-1710 CTC head) run in fp32 on the CPU with the oracle's weights.  For each model the script
asserts oracle(fp32) == HF(fp32) to fp32 round-off at full width, full depth and T = 1499, and stores

  oracle_1b_batch.npz   omniASR_CTC_1B (48 layers), the BENCH batch bench.synthetic_windows(32, 1234): windows 0 and 13
                        in full and window 31 cut to RAGGED_SAMPLES samples - oracle ids / top-2 margins in both modes
                        (bf16-operand emulation, fp32), hidden rows, and the HF ids / hidden rows / max |dlogit| of
                        window 0.  Read by tests/test_gpu_engine.py::test_config2_1b_full_depth_batch32.
  oracle_3b_window.npz  omniASR_CTC_3B (d 2048, 60 layers, head_dim 128), window 0 of the same batch: the same fields.
                        Read by tests/test_gpu_engine.py::test_config3_3b_full_depth_window.

  oracle_7b_window.npz  omniASR_CTC_7B (128 layers), the same window, oracle only (`7b` on the command line: 26 GB of fp32
                        weights, no room for a second model).  Read by test_config4_7b_full_depth_window.

  oracle_300m_gettysburg_emu.npz
                        BASELINE configs[0] (omniASR_CTC_300M on the committed 16 kHz gettysburg fixture): ids and margins
                        in BOTH oracle modes plus the HF ids (oracle_300m_gettysburg.npz holds the fp32 ids only).

tests/test_oracle.py::test_fullsize_hf_pins checks the recorded oracle-vs-HF differences and re-runs the fp32 oracle on
one full 1B window against the stored HF vectors on the CPU.
"""
import hashlib
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
import bench  # noqa: E402
from oracle import ctc_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent
RAGGED_SAMPLES = 300_001          # window 31 of the bench batch is cut to this length (937 frames)
BENCH_SEED = 1234
ROW_STEP = 16                     # every 16th hidden row is stored


def bench_windows(indices):
    full = bench.synthetic_windows(32, BENCH_SEED)
    return full[list(indices)].clone(), hashlib.sha256(full.numpy().tobytes()).hexdigest()


def hf_forward(cfg, w, wave_norm, ns):
    """transformers' Wav2Vec2ForCTC, fp32, eager attention, built on the meta device and given the oracle's tensors
    (no second copy of the weights)."""
    from transformers import Wav2Vec2ForCTC
    with torch.device("meta"):
        m = Wav2Vec2ForCTC(O.hf_config(cfg))
    res = m.load_state_dict(O.to_hf_state_dict(w, cfg), strict=False, assign=True)
    assert all("masked_spec_embed" in k for k in res.missing_keys) and not res.unexpected_keys, res
    m.eval()
    am = torch.zeros(wave_norm.shape, dtype=torch.long)
    for b, n in enumerate(ns):
        am[b, :n] = 1
    with torch.no_grad():
        out = m(wave_norm, attention_mask=am, output_hidden_states=True)
    return out.logits, out.hidden_states[-1]


def pack(prefix, o, b, rows):
    nf = o.n_frames[b]
    return {
        f"{prefix}_ids": o.frame_ids[b, :nf].numpy().astype(np.int32),
        f"{prefix}_margin": O.top2_margin(o.logits[b, :nf]).numpy().astype(np.float32),
        f"{prefix}_hidden": o.hidden[b, rows].numpy().astype(np.float32),
    }


def run(model: str, indices, ns, out_name: str, hf_window: int = 0, wave=None, with_hf: bool = True):
    t0 = time.time()
    cfg = O.PRESETS[model]
    w = O.init_weights(cfg, seed=0)
    if wave is None:
        wave, batch_sha = bench_windows(indices)
    else:
        batch_sha = hashlib.sha256(wave.numpy().tobytes()).hexdigest()
    for b, n in enumerate(ns):
        wave[b, n:] = 0
    wn = O.wave_layer_norm(wave, ns)
    fields = {"windows": np.array(indices, dtype=np.int32), "n_samples": np.array(ns, dtype=np.int32),
              "bench_batch_sha256": batch_sha, "hidden_row_step": ROW_STEP}
    with torch.no_grad():
        f32 = O.forward(w, wn, ns, cfg, emulate_bf16=False, return_logits=True)
        print(f"{model}: fp32 oracle done ({time.time() - t0:.0f} s)", flush=True)
        emu = O.forward(w, wn, ns, cfg, emulate_bf16=True, return_logits=True)
        print(f"{model}: bf16-operand oracle done ({time.time() - t0:.0f} s)", flush=True)
    fields["n_frames"] = np.array(f32.n_frames, dtype=np.int32)
    for b in range(len(indices)):
        rows = np.arange(0, f32.n_frames[b], ROW_STEP)
        for k, v in {**pack(f"w{b}_f32", f32, b, rows), **pack(f"w{b}_emu", emu, b, rows)}.items():
            fields[k] = v
    if not with_hf:     # 7B: the oracle alone (a second 26 GB model does not fit beside it in this container)
        fields["hf_window"] = -1
        np.savez_compressed(OUT / out_name, **fields)
        print(f"{model}: oracle vectors written (no HF run), emu/f32 id agreement w0 "
              f"{float((fields['w0_emu_ids'] == fields['w0_f32_ids']).mean()):.4f}; {(OUT / out_name).stat().st_size} bytes, "
              f"{time.time() - t0:.0f} s", flush=True)
        return
    # ---- the independent implementation at full size (one window: its cost is the oracle's)
    b = hf_window
    nf = f32.n_frames[b]
    logits_hf, hidden_hf = hf_forward(cfg, w, wn[b:b + 1, :ns[b]].contiguous(), [ns[b]])
    print(f"{model}: transformers Wav2Vec2ForCTC done ({time.time() - t0:.0f} s)", flush=True)
    dl = float((f32.logits[b, :nf] - logits_hf[0, :nf]).abs().max())
    dh = float((f32.hidden[b, :nf] - hidden_hf[0, :nf]).abs().max())
    ids_hf = logits_hf[0, :nf].argmax(-1).numpy().astype(np.int32)
    same = float((ids_hf == fields[f"w{b}_f32_ids"]).mean())
    rows = np.arange(0, nf, ROW_STEP)
    fields.update(hf_window=b, hf_ids=ids_hf, hf_hidden=hidden_hf[0, rows].numpy().astype(np.float32),
                  hf_margin=O.top2_margin(logits_hf[0, :nf]).numpy().astype(np.float32),
                  hf_max_abs_dlogit=dl, hf_max_abs_dhidden=dh, hf_id_agreement=same,
                  logit_scale=float(f32.logits[b, :nf].abs().max()))
    # fp32 round-off over 48-60 layers: two fp32 implementations with different summation orders
    assert dl < 5e-3 and dh < 5e-3, (dl, dh)
    flips = np.nonzero(ids_hf != fields[f"w{b}_f32_ids"])[0]
    assert all(fields[f"w{b}_f32_margin"][t] < 2 * dl + 1e-6 for t in flips), "HF and oracle differ on a clear frame"
    np.savez_compressed(OUT / out_name, **fields)
    print(f"{model}: oracle(fp32) vs HF(fp32): max |dlogit| {dl:.2e} (logits up to {fields['logit_scale']:.1f}), max |dhidden| {dh:.2e}, "
          f"ids equal on {same:.4f} of {nf} frames ({len(flips)} flips, all at margins < 2 |dlogit|); "
          f"emu/f32 id agreement w0 {float((fields['w0_emu_ids'] == fields['w0_f32_ids']).mean()):.4f}; "
          f"{(OUT / out_name).stat().st_size} bytes, {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = [a.lower() for a in sys.argv[1:]] or ["300m", "1b", "3b"]
    L = int(bench.WINDOW_SEC * bench.SR)
    if "1b" in which:
        run("omniASR_CTC_1B", (0, 13, 31), [L, L, RAGGED_SAMPLES], "oracle_1b_batch.npz")
    if "300m" in which:
        pcm = np.load(OUT / "gettysburg_16k_i16.npz")["pcm"]
        g = torch.from_numpy(pcm.astype(np.float32) / 32768.0)[None]
        run("omniASR_CTC_300M", (0,), [g.shape[1]], "oracle_300m_gettysburg_emu.npz", wave=g)
    if "3b" in which:
        run("omniASR_CTC_3B", (0,), [L], "oracle_3b_window.npz")
    if "7b" in which:      # not in the default list: 26 GB of fp32 weights, ~6 min
        run("omniASR_CTC_7B", (0,), [L], "oracle_7b_window.npz", with_hf=False)
