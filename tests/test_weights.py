"""Checkpoint layouts the engine accepts: Hugging Face Wav2Vec2ForCTC and fairseq2 Wav2Vec2AsrModel key
names are renamed to CtcModelConfig.weight_shapes().  The HF table is pinned by the live transformers module (its
own state dict, its own forward); the fairseq2 table only by a round trip, since fairseq2 cannot be installed here."""
import pytest
import torch

from omnilingual_asr.models import weights as W
from omnilingual_asr.models.config import CtcModelConfig
from oracle import ctc_oracle as O


def _cfg(name="tiny"):
    o = O.PRESETS[name]
    return o, CtcModelConfig(o.name, o.d_model, o.n_layers, o.n_heads, o.d_ffn, vocab=o.vocab, pos_groups=o.pos_groups)


def test_native_dict_passes_through():
    o, cfg = _cfg()
    w = O.init_weights(o, seed=0)
    assert W.detect_layout(w) == "native"
    assert W.convert_state_dict(w, cfg) is w


def test_hf_module_state_dict_converts_and_reproduces_its_logits():
    transformers = pytest.importorskip("transformers")
    o, cfg = _cfg()
    torch.manual_seed(11)
    m = transformers.Wav2Vec2ForCTC(O.hf_config(o)).eval()
    with torch.no_grad():      # HF initialises LayerNorm to (1, 0) and most biases to 0: perturb every parameter
        for prm in m.parameters():
            prm.add_(0.05 * torch.randn_like(prm))
    sd = m.state_dict()
    assert W.detect_layout(sd) == "hf"
    w = W.convert_state_dict(sd, cfg)
    assert set(w) == set(cfg.weight_shapes())
    wave = torch.randn(2, 8000)
    ns = [8000, 5000]
    wave[1, 5000:] = 0
    wave = O.wave_layer_norm(wave, ns)
    am = torch.zeros(2, 8000, dtype=torch.long)
    am[0] = 1
    am[1, :5000] = 1
    with torch.no_grad():
        ref = m(wave, attention_mask=am).logits
    out = O.forward({k: v.detach().clone() for k, v in w.items()}, wave, ns, o, return_logits=True)
    for b, nf in enumerate(out.n_frames):
        assert (out.logits[b, :nf] - ref[b, :nf]).abs().max() < 1e-4


def test_old_weight_norm_spelling_is_accepted():
    o, cfg = _cfg()
    sd = dict(O.to_hf_state_dict(O.init_weights(o, seed=1), o))
    pc = "wav2vec2.encoder.pos_conv_embed.conv."
    sd[pc + "weight_g"] = sd.pop(pc + "parametrizations.weight.original0")
    sd[pc + "weight_v"] = sd.pop(pc + "parametrizations.weight.original1")
    w = W.convert_state_dict(sd, cfg)
    assert torch.equal(w["pos.weight_g"], sd[pc + "weight_g"])


def test_fairseq2_names_round_trip():
    o, cfg = _cfg("tiny80")
    native = O.init_weights(o, seed=2)
    inv = {v: k for k, v in W._fairseq2_key_map(cfg).items() if "weight_g" not in k and "weight_v" not in k}
    sd = {inv[n]: t for n, t in native.items()}
    sd["encoder_frontend.masker.temporal_mask_embed"] = torch.zeros(o.d_model)   # training-only, ignored
    assert W.detect_layout(sd) == "fairseq2"
    w = W.convert_state_dict(sd, cfg)
    assert set(w) == set(native)
    for n in native:
        assert torch.equal(w[n], native[n])


def test_foreign_checkpoints_fail_loudly():
    o, cfg = _cfg()
    sd = dict(O.to_hf_state_dict(O.init_weights(o, seed=1), o))
    extra = dict(sd)
    extra["wav2vec2.adapter.layers.0.conv.weight"] = torch.zeros(1)
    with pytest.raises(ValueError, match="does not"):
        W.convert_state_dict(extra, cfg)
    short = dict(sd)
    short.pop("lm_head.bias")
    with pytest.raises(ValueError, match="lacks"):
        W.convert_state_dict(short, cfg)
    bad = dict(sd)
    bad["lm_head.weight"] = torch.zeros(3, 3)
    with pytest.raises(ValueError, match="shape"):
        W.convert_state_dict(bad, cfg)


def test_checkpoint_file_in_hf_layout_resolves(tmp_path):
    o, cfg = _cfg()
    native = O.init_weights(o, seed=4)
    path = tmp_path / "ckpt.pt"
    torch.save({"model": O.to_hf_state_dict(native, o)}, str(path))
    w = W.resolve_weights(cfg, str(path), seed=0, device="cpu")
    for n in native:
        assert torch.equal(w[n], native[n])


def test_checkpoint_head_count_is_checked_when_present():
    """the head count changes no tensor shape; a checkpoint that states its own must agree with the card,
    and a 2048-wide checkpoint without one triggers a warning naming the assumption."""
    import warnings
    from omnilingual_asr.models.config import get_model_config
    from omnilingual_asr.models.weights import check_head_count
    c3 = get_model_config("omniASR_CTC_3B")
    check_head_count(c3, {"config": {"num_attention_heads": 16}, "model": {}})
    with pytest.raises(ValueError, match="32 attention heads"):
        check_head_count(c3, {"model_config": {"encoder": {"num_encoder_attn_heads": 32}}})
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        check_head_count(c3, {"model": {}})
        check_head_count(get_model_config("omniASR_CTC_1B"), {"model": {}})     # 1280-wide: nothing to warn about
    assert len(w) == 1 and "assuming 16 heads" in str(w[0].message)
