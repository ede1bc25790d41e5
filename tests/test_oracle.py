"""The oracle against its pins: published parameter counts, length/collapse/chunker known answers, and an
independent implementation of the same graph (transformers' Wav2Vec2ForCTC), live and through the committed
fixtures of tests/golden/make_golden.py.  CPU only."""
import hashlib
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ctc_oracle as O
from tests.golden.make_golden import golden_inputs

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_param_counts_equal_published():
    for name, want in O.PUBLISHED_PARAM_COUNTS.items():
        assert O.param_count(O.PRESETS[name]) == want


def test_feature_extractor_breakdown():
    cfg = O.PRESETS["omniASR_CTC_1B"]
    fe = sum(int(np.prod(s)) for n, s in O.weight_shapes(cfg).items() if n.startswith("fe."))
    assert fe == 4_210_176


@pytest.mark.parametrize("n,frames", [(480_000, 1499), (640_000, 1999), (281_233, 878), (400, 1), (399, 0), (0, 0)])
def test_frame_count_kats(n, frames):
    assert O.feature_length(n) == frames


def test_collapse_kats():
    assert O.greedy_collapse([0, 0, 5, 5, 0, 5, 7, 7, 7, 0])[0] == [5, 5, 7]
    assert O.greedy_collapse([0] * 9)[0] == []
    assert O.greedy_collapse([3, 1, 4, 1, 5, 9, 2, 6])[0] == [3, 1, 4, 1, 5, 9, 2, 6]
    ids, pos = O.greedy_collapse([4, 4, 0, 4, 2, 2], n_frames=5)
    assert ids == [4, 4, 2] and pos == [0, 3, 4]
    a, p, l = O.collapse_batch(np.array([[0, 0, 5, 5, 0, 5, 7, 7, 7, 0]]), [10])
    assert a[0, :l[0]].tolist() == [5, 5, 7] and p[0, :l[0]].tolist() == [2, 5, 6]


@pytest.mark.parametrize("seconds,window,count", [(3600, 30, 120), (34200, 30, 1140), (17.577, 30, 1), (3600, 300, 12)])
def test_chunker_kats(seconds, window, count):
    w = O.split_into_windows(int(round(seconds * 16000)), window * 16000)
    assert len(w) == count
    assert [s for s, _ in w] == [i * window * 16000 for i in range(count)]
    assert sum(n for _, n in w) == int(round(seconds * 16000))


def test_wave_layer_norm_matches_definition():
    x = torch.randn(2, 1000) * 3 + 1
    y = O.wave_layer_norm(x, [1000, 600])
    assert abs(float(y[0].mean())) < 1e-5 and abs(float(y[0].var(unbiased=False)) - 1) < 1e-3
    assert (y[1, 600:] == 0).all() and abs(float(y[1, :600].mean())) < 1e-5


@pytest.mark.parametrize("name", ["tiny", "tiny80"])
def test_oracle_equals_committed_hf_vectors(name):
    g = np.load(GOLDEN / f"hf_{name}.npz")
    cfg = O.PRESETS[name]
    w = O.init_weights(cfg, seed=0)
    wave, ns = golden_inputs()
    out = O.forward(w, wave, ns, cfg, return_logits=True)
    assert out.n_frames == list(g["n_frames"])
    for b, nf in enumerate(out.n_frames):
        assert np.abs(out.logits[b, :nf].numpy() - g["logits"][b, :nf]).max() < 1e-4     # fp32 round-off
        assert np.abs(out.hidden[b, :nf].numpy() - g["hidden"][b, :nf]).max() < 1e-4
        assert (out.frame_ids[b, :nf].numpy() == g["ids"][b, :nf]).all()


def test_oracle_equals_hf_live():
    transformers = pytest.importorskip("transformers")
    cfg = O.PRESETS["tiny"]
    w = O.init_weights(cfg, seed=3)
    torch.manual_seed(5)
    wave = torch.randn(2, 8000)
    ns = [8000, 5000]
    wave[1, 5000:] = 0
    wave = O.wave_layer_norm(wave, ns)
    m = transformers.Wav2Vec2ForCTC(O.hf_config(cfg)).eval()
    m.load_state_dict(O.to_hf_state_dict(w, cfg), strict=False)
    am = torch.zeros(2, 8000, dtype=torch.long)
    am[0] = 1
    am[1, :5000] = 1
    with torch.no_grad():
        ref = m(wave, attention_mask=am).logits
    out = O.forward(w, wave, ns, cfg, return_logits=True)
    for b, nf in enumerate(out.n_frames):
        assert (out.logits[b, :nf] - ref[b, :nf]).abs().max() < 1e-4


def test_emulated_operand_mode_is_close_to_fp32():
    cfg = O.PRESETS["tiny80"]
    w = O.init_weights(cfg, seed=0)
    wave, ns = golden_inputs()
    a = O.forward(w, wave, ns, cfg)
    b = O.forward(w, wave, ns, cfg, emulate_bf16=True)
    assert float((a.hidden - b.hidden).norm() / a.hidden.norm()) < 1e-2
    assert float((a.frame_ids == b.frame_ids).float().mean()) > 0.95


def test_gettysburg_fixture_and_oracle_drift():
    g = np.load(GOLDEN / "gettysburg_16k_i16.npz")
    assert len(g["pcm"]) == 281_233
    assert str(g["source_sha256"]) == "7630daffb2f28f2724d81f1ff2039eb69a5fa360db3919721a77032a58db0d46"
    gold = np.load(GOLDEN / "oracle_300m_gettysburg.npz")
    assert hashlib.sha256(gold["frame_ids"].tobytes()).hexdigest() == str(gold["sha256"])
    ids, pos = O.greedy_collapse(gold["frame_ids"], 878)
    assert ids == gold["collapsed"].tolist() and pos == gold["positions"].tolist()


@pytest.mark.slow
def test_config1_oracle_reproduces_golden_ids():
    """BASELINE configs[0] on the CPU: 300M random-init on gettysburg (about 10 s of CPU)."""
    g = np.load(GOLDEN / "gettysburg_16k_i16.npz")
    gold = np.load(GOLDEN / "oracle_300m_gettysburg.npz")
    cfg = O.PRESETS["omniASR_CTC_300M"]
    w = O.init_weights(cfg, seed=0)
    wave = torch.from_numpy(g["pcm"].astype(np.float32) / 32768.0)[None]
    out = O.forward(w, O.wave_layer_norm(wave, [wave.shape[1]]), [wave.shape[1]], cfg)
    ids = out.frame_ids[0].numpy().astype(np.int32)
    # summation order depends on the host's core count/BLAS; near-ties may flip, everything else is exact
    agree = (ids == gold["frame_ids"])
    assert agree[gold["margin"] > 1e-3].all()
    assert agree.mean() > 0.995


@pytest.mark.parametrize("fixture,bound", [("oracle_1b_batch.npz", 1e-4), ("oracle_3b_window.npz", 2e-4),
                                           ("oracle_300m_gettysburg_emu.npz", 1e-4)])
def test_fullsize_hf_pins_recorded(fixture, bound):
    """What tests/golden/make_golden_fullsize.py measured when it ran transformers' Wav2Vec2ForCTC (fp32, the oracle's
    weights) at the REAL model sizes: oracle(fp32) and HF agree to fp32 round-off in logits and hidden states, and on
    every frame id outside exact near-ties."""
    g = np.load(GOLDEN / fixture)
    assert float(g["hf_max_abs_dlogit"]) < bound and float(g["hf_max_abs_dhidden"]) < bound
    b = int(g["hf_window"])
    flips = np.nonzero(g["hf_ids"] != g[f"w{b}_f32_ids"])[0]
    assert len(flips) <= 2 and all(g[f"w{b}_f32_margin"][t] < 2 * bound for t in flips)
    assert np.abs(g["hf_hidden"] - g[f"w{b}_f32_hidden"]).max() < bound
    # the bf16-operand mode stays within the north_star bars of the fp32 mode at full size
    assert float((g[f"w{b}_emu_ids"] == g[f"w{b}_f32_ids"]).mean()) >= 0.95


@pytest.mark.slow
def test_fullsize_1b_window_oracle_reproduces_hf_vectors():
    """Live: today's fp32 oracle on one full 30 s window of omniASR_CTC_1B (48 layers, T = 1499; ~30 s of CPU) against
    the stored transformers vectors - the full-size pin does not depend on the oracle that wrote the fixture."""
    import bench
    from tests.golden.make_golden_fullsize import BENCH_SEED, ROW_STEP
    g = np.load(GOLDEN / "oracle_1b_batch.npz")
    cfg = O.PRESETS["omniASR_CTC_1B"]
    w = O.init_weights(cfg, seed=0)
    wave = bench.synthetic_windows(32, BENCH_SEED)
    assert hashlib.sha256(wave.numpy().tobytes()).hexdigest() == str(g["bench_batch_sha256"])
    wave = wave[:1].contiguous()
    ns = [wave.shape[1]]
    with torch.no_grad():
        out = O.forward(w, O.wave_layer_norm(wave, ns), ns, cfg, return_logits=True)
    ids = out.frame_ids[0].numpy()
    rows = np.arange(0, 1499, ROW_STEP)
    assert np.abs(out.hidden[0, rows].numpy() - g["hf_hidden"]).max() < 2e-4
    agree = ids == g["hf_ids"]
    assert agree[g["hf_margin"] > 1e-3].all() and agree.mean() > 0.995
