"""Writes the committed golden fixtures.  Run in the build container (needs /root/reference and transformers):

    python tests/golden/make_golden.py

Fixtures:
  gettysburg_16k_i16.npz   the reference's bundled gettysburg.wav (sha256 7630daff...0d46, PCM16 mono 22 050 Hz,
                           BASELINE.json configs[0] input) resampled on the host to 16 kHz (281 233 samples)
                           and stored as int16, so that GPU-box tests do not need /root/reference.
  hf_tiny.npz / hf_tiny80.npz
                           outputs of transformers' Wav2Vec2ForCTC (an implementation of the same graph that
                           is independent of oracle/ctc_oracle.py) for the seeded tiny configs: logits,
                           final hidden states and per-frame arg-max on a seeded ragged batch.
  oracle_300m_gettysburg.npz
                           the oracle's own fp32 frame ids / collapsed ids on gettysburg (300M, seed 0), and
                           their sha256, to detect drift of the oracle itself.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from oracle import ctc_oracle as O  # noqa: E402
from omnilingual_asr.models.inference.audio import read_wav, to_mono_16k  # noqa: E402

OUT = Path(__file__).resolve().parent
REF_WAV = Path("/root/reference/gettysburg.wav")


def golden_inputs(seed=1234, B=3, L=16000):
    g = torch.Generator().manual_seed(seed)
    wave = torch.randn(B, L, generator=g)
    t = torch.arange(L) / 16000.0
    wave += 0.5 * torch.sin(2 * np.pi * 220.0 * t)[None] + 0.25 * torch.sin(2 * np.pi * 1333.0 * t)[None]
    ns = [L, L - 3217, 9000][:B]
    for b, n in enumerate(ns):
        wave[b, n:] = 0
    return O.wave_layer_norm(wave, ns), ns


def main():
    torch.set_num_threads(8)
    # --- gettysburg
    raw = REF_WAV.read_bytes()
    sha = hashlib.sha256(raw).hexdigest()
    assert sha == "7630daffb2f28f2724d81f1ff2039eb69a5fa360db3919721a77032a58db0d46", sha
    x, sr = read_wav(REF_WAV)
    assert sr == 22050 and x.shape == (387574, 1)
    y = to_mono_16k(x, sr)
    assert len(y) == 281233, len(y)
    i16 = np.clip(np.round(y * 32768.0), -32768, 32767).astype(np.int16)
    np.savez_compressed(OUT / "gettysburg_16k_i16.npz", pcm=i16, source_sha256=sha, source_rate=sr)
    print("gettysburg:", i16.shape, (OUT / "gettysburg_16k_i16.npz").stat().st_size, "bytes")

    # --- HF cross-check vectors for the tiny configs
    from transformers import Wav2Vec2ForCTC
    for name in ("tiny", "tiny80"):
        cfg = O.PRESETS[name]
        w = O.init_weights(cfg, seed=0)
        wave, ns = golden_inputs()
        m = Wav2Vec2ForCTC(O.hf_config(cfg)).eval()
        res = m.load_state_dict(O.to_hf_state_dict(w, cfg), strict=False)
        assert not res.missing_keys or all("masked_spec_embed" in k for k in res.missing_keys), res
        am = torch.zeros(wave.shape, dtype=torch.long)
        for b, n in enumerate(ns):
            am[b, :n] = 1
        with torch.no_grad():
            out = m(wave, attention_mask=am, output_hidden_states=True)
        logits = out.logits.numpy()
        hidden = out.hidden_states[-1].numpy()
        nf = [O.feature_length(n, cfg) for n in ns]
        ids = logits.argmax(-1).astype(np.int32)
        np.savez_compressed(OUT / f"hf_{name}.npz", n_samples=np.array(ns), n_frames=np.array(nf), ids=ids,
                            logits=logits.astype(np.float32), hidden=hidden.astype(np.float32))
        o = O.forward(w, wave, ns, cfg, return_logits=True)
        for b in range(len(ns)):
            err = np.abs(o.logits[b, :nf[b]].numpy() - logits[b, :nf[b]]).max()
            assert err < 1e-4, err
        print(name, "HF vectors written; oracle max |dlogit| ok")

    # --- oracle drift check on config #1
    cfg = O.PRESETS["omniASR_CTC_300M"]
    w = O.init_weights(cfg, seed=0)
    wave = torch.from_numpy(i16.astype(np.float32) / 32768.0)[None]
    wn = O.wave_layer_norm(wave, [wave.shape[1]])
    o = O.forward(w, wn, [wave.shape[1]], cfg, return_logits=True)
    ids = o.frame_ids[0].numpy().astype(np.int32)
    col, pos = O.greedy_collapse(ids, o.n_frames[0])
    margin = O.top2_margin(o.logits[0]).numpy().astype(np.float32)
    np.savez_compressed(OUT / "oracle_300m_gettysburg.npz", frame_ids=ids, collapsed=np.array(col, dtype=np.int32),
                        positions=np.array(pos, dtype=np.int32), margin=margin,
                        sha256=hashlib.sha256(ids.tobytes()).hexdigest())
    print("300M gettysburg: frames", o.n_frames, "tokens", len(col), "median margin", float(np.median(margin)))


if __name__ == "__main__":
    main()
