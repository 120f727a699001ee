"""CPU: oracle self-consistency, GGML format round trip, synthetic data."""
import os

import numpy as np
import pytest

from oracle import logmel, whisper_ref
from spittle_b200 import ggml_format, synth


def test_mel_filterbank_matches_transformers():
    tf = pytest.importorskip("transformers.audio_utils")
    for n_mel in (80, 128):
        ref = tf.mel_filter_bank(num_frequency_bins=201, num_mel_filters=n_mel, min_frequency=0.0,
                                 max_frequency=8000.0, sampling_rate=16000, norm="slaney", mel_scale="slaney").T
        ours = synth.mel_filterbank(n_mel)
        assert ours.shape == (n_mel, 201)
        np.testing.assert_allclose(ours, ref, rtol=1e-5, atol=1e-8)


def test_logmel_geometry_and_floor():
    x = synth.make_clip(1, seconds=2.0)
    filt = synth.mel_filterbank(80)
    mel, n_len_org = logmel.logmel_f64(x, filt)
    assert mel.shape == (80, (32000 + 480000) // 160)
    assert n_len_org == 1 + (32000 + 200 - 400) // 160
    # frames past the audio are the constant clamp floor
    assert np.all(mel[:, 300:] == mel[0, -1])
    assert abs(float(mel.max()) - (float(mel.min()) + 2.0)) < 1e-5 or mel.min() == mel[0, -1]


def test_logmel_f32_faithful_close_to_f64():
    filt = synth.mel_filterbank(80)
    for i in range(3):
        x = synth.make_clip(i, seconds=3.0)
        a, _ = logmel.logmel_f32_faithful(x, filt)
        b, _ = logmel.logmel_f64(x, filt)
        assert np.abs(a - b).max() / np.abs(b).max() <= 1e-4


def test_logmel_interior_matches_hf_extractor():
    """Independent cross-check: agrees with HF on interior frames; the
    tail differs by design (whisper.cpp zero-pads, HF reflect-pads)."""
    fe_mod = pytest.importorskip("transformers")
    fe = fe_mod.WhisperFeatureExtractor(feature_size=80)
    x = synth.make_clip(3, seconds=30.0)
    hf = fe(x, sampling_rate=16000, return_tensors="np")["input_features"][0]   # [80, 3000]
    ours, _ = logmel.logmel_f64(x, synth.mel_filterbank(80))
    assert np.abs(ours[:, :2980] - hf[:, :2980]).max() < 2e-4


def test_ggml_roundtrip(tmp_path):
    m = synth.make_synthetic_model("nano", seed=7)
    p = str(tmp_path / "ggml-nano.bin")
    ggml_format.write_ggml(p, m)
    r = ggml_format.read_ggml(p)
    assert r.hparams == m.hparams
    assert r.vocab == m.vocab
    np.testing.assert_array_equal(r.mel_filters, m.mel_filters)
    assert list(r.tensors) == list(m.tensors)
    for k in m.tensors:
        assert r.tensors[k].dtype == m.tensors[k].dtype
        np.testing.assert_array_equal(r.tensors[k], m.tensors[k])
    # converter convention (App. D item 7)
    assert r.tensors["encoder.conv1.bias"].shape == (128, 1) and r.tensors["encoder.conv1.bias"].dtype == np.float32
    assert r.tensors["decoder.token_embedding.weight"].dtype == np.float16
    with open(p, "rb") as f:
        assert f.read(4) == bytes.fromhex("6c6d6767")


def test_special_tokens():
    s = ggml_format.SpecialTokens.from_n_vocab(51865)
    assert (s.eot, s.sot, s.translate, s.transcribe, s.solm, s.prev, s.nosp, s.not_, s.beg) == \
        (50257, 50258, 50358, 50359, 50360, 50361, 50362, 50363, 50364)
    s = ggml_format.SpecialTokens.from_n_vocab(51866)
    assert (s.eot, s.sot, s.translate, s.transcribe, s.solm, s.prev, s.nosp, s.not_, s.beg) == \
        (50257, 50258, 50359, 50360, 50361, 50362, 50363, 50364, 50365)
    assert s.num_languages == 100


def test_logits_filter_rules():
    m = synth.make_synthetic_model("nano", seed=3)
    o = whisper_ref.WhisperOracle(m)
    sp = o.sp
    rng = np.random.default_rng(0)
    logits = rng.normal(0, 1, m.hparams.n_vocab).astype(np.float32)
    cfg = whisper_ref.DecodeConfig()
    lg, lp, pr = o.process_logits(logits, [], False, 3000, cfg)
    assert lg[sp.eot] == -np.inf and lg[sp.blank] == -np.inf and lg[sp.not_] == -np.inf
    assert np.all(lg[sp.beg + 51:] == -np.inf)          # max_initial_ts = 1.0 s -> +50
    assert np.isfinite(lg[sp.beg + 50]) or np.all(lg[:sp.beg] == -np.inf)
    assert np.all(lg[sp.lang_first:sp.lang_first + sp.num_languages] == -np.inf)
    # after one timestamp: text tokens < eot are suppressed, next must be ts or eot
    lg, _, _ = o.process_logits(logits, [5, sp.beg + 10], True, 20, cfg)
    assert np.all(lg[:sp.eot] == -np.inf) and np.all(lg[sp.beg:sp.beg + 10] == -np.inf)
    # after a timestamp pair: timestamps are suppressed
    lg, _, _ = o.process_logits(logits, [sp.beg + 10, sp.beg + 10], True, 20, cfg)
    assert np.all(lg[sp.beg:] == -np.inf)
    # ties: lowest index wins
    assert o.sample_best(np.array([0.1, 0.4, 0.4, 0.1], np.float32)) == 1


def test_oracle_window_runs_and_is_audio_dependent():
    m = synth.make_synthetic_model("nano", seed=42)
    o = whisper_ref.WhisperOracle(m)
    outs = []
    for i in (1, 2):
        x = synth.make_clip(i, seconds=30.0)
        mel, n_len_org = logmel.logmel_f64(x, m.mel_filters)
        enc = o.encode(logmel.mel_window(mel, 0))
        assert enc.shape == (1500, 128) and np.isfinite(enc).all()
        cfg = whisper_ref.DecodeConfig(n_max_override=12)
        w = o.decode_window(enc, 0, n_len_org, cfg)
        assert 1 <= len(w.tokens) <= 12
        outs.append(tuple(w.tokens))
    assert outs[0] != outs[1]
