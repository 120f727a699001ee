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


def test_ggml_block_quantisation_round_trip(tmp_path):
    """q4_0 / q4_1 / q5_0 / q5_1 / q8_0 tensors (the reference catalog ships Medium as q4_1 and
    Large-v3 as q5_0).  Known answers for the block layouts + write/read round trip of a quantised model file."""
    from spittle_b200 import ggml_format as g, synth
    # hand-built blocks: d = 0.5 (f16 0x3800), m = -1.0 (f16 0xBC00)
    qs = bytes(((j & 0xF) | (((15 - j) & 0xF) << 4)) for j in range(16))
    x = g.dequantize_blocks(b"\x00\x38" + qs, g.GGML_TYPE_Q4_0, 32)
    assert np.array_equal(x[:16], (np.arange(16) - 8) * 0.5) and np.array_equal(x[16:], (15 - np.arange(16) - 8) * 0.5)
    x = g.dequantize_blocks(b"\x00\x38\x00\xbc" + qs, g.GGML_TYPE_Q4_1, 32)
    assert np.array_equal(x[:16], np.arange(16) * 0.5 - 1.0)
    qh = (0x0000FFFF).to_bytes(4, "little")           # fifth bit set for the low half only
    x = g.dequantize_blocks(b"\x00\x38" + qh + qs, g.GGML_TYPE_Q5_0, 32)
    assert np.array_equal(x[:16], (np.arange(16) + 16 - 16) * 0.5) and np.array_equal(x[16:], (15 - np.arange(16) - 16) * 0.5)
    x = g.dequantize_blocks(b"\x00\x38" + bytes(np.arange(-16, 16, dtype=np.int8).view(np.uint8)), g.GGML_TYPE_Q8_0, 32)
    assert np.array_equal(x, np.arange(-16, 16) * 0.5)
    # quantise -> dequantise stays within half a step of the block scale
    rng = np.random.default_rng(1)
    w = rng.normal(0, 0.1, (8, 64)).astype(np.float32)
    for t, steps in ((g.GGML_TYPE_Q4_0, 8), (g.GGML_TYPE_Q4_1, 15), (g.GGML_TYPE_Q5_0, 16), (g.GGML_TYPE_Q5_1, 31), (g.GGML_TYPE_Q8_0, 127)):
        y = g.dequantize_blocks(g.quantize_blocks(w, t), t, w.size).reshape(w.shape)
        span = np.abs(w.reshape(-1, 32)).max(axis=1) * (2 if t in (g.GGML_TYPE_Q4_1, g.GGML_TYPE_Q5_1) else 1)
        # the symmetric types anchor the scale on the largest-magnitude value: the other side clips by up to one step
        k = 1.01 if t in (g.GGML_TYPE_Q4_0, g.GGML_TYPE_Q5_0) else 0.51
        assert (np.abs(y - w).reshape(-1, 32).max(axis=1) <= k * span / steps + 1e-3).all(), t
    # model file: matrices quantised, vectors / conv kernels / positional embeddings untouched
    path = synth.ensure_model_file("nano", str(tmp_path), quant_type=g.GGML_TYPE_Q5_0)
    ref = g.read_ggml(synth.ensure_model_file("nano", str(tmp_path)))
    q = g.read_ggml(path)
    assert q.hparams.ftype == 8
    name = "decoder.blocks.0.mlp.0.weight"
    assert q.tensors[name].dtype == np.float32 and q.tensors[name].shape == ref.tensors[name].shape
    err = np.abs(q.tensors[name] - ref.tensors[name].astype(np.float32)).max()
    assert 0 < err < 0.1 * np.abs(ref.tensors[name].astype(np.float32)).max()
    assert np.array_equal(q.tensors["encoder.conv1.weight"], ref.tensors["encoder.conv1.weight"])


def test_ggml_q5_k_layout_known_answer_and_round_trip(tmp_path):
    """q5_K (the catalog's breeze-asr-q5_k.bin).  A hand-built super-block pins the bit layout
    (get_scale_min_k4 packing, nibble / fifth-bit placement); quantise -> dequantise stays within a step."""
    from spittle_b200 import ggml_format as g, synth
    d, dmin = b"\x00\x38", b"\x00\x34"                  # 0.5, 0.25
    sc = [1, 2, 3, 4, 17, 34, 51, 63]                   # 6-bit scales of the 8 sub-blocks (j >= 4 need the high-bit fields)
    mn = [0, 1, 2, 3, 16, 33, 50, 63]
    scales = bytearray(12)
    for j in range(4):
        scales[j] = sc[j]
        scales[j + 4] = mn[j]
    for j in range(4, 8):
        scales[j + 4] = (sc[j] & 0xF) | ((mn[j] & 0xF) << 4)
        scales[j - 4] |= (sc[j] >> 4) << 6
        scales[j] |= (mn[j] >> 4) << 6
    q = np.zeros((8, 32), np.int32)
    for j in range(8):
        q[j] = (np.arange(32) + 3 * j) % 32
    qs, qh = bytearray(128), bytearray(32)
    for grp in range(4):
        for l in range(32):
            lo, hi = int(q[2 * grp, l]), int(q[2 * grp + 1, l])
            qs[32 * grp + l] = (lo & 0xF) | ((hi & 0xF) << 4)
            qh[l] |= ((lo >> 4) & 1) << (2 * grp)
            qh[l] |= ((hi >> 4) & 1) << (2 * grp + 1)
    x = g.dequantize_blocks(d + dmin + bytes(scales) + bytes(qh) + bytes(qs), g.GGML_TYPE_Q5_K, 256).reshape(8, 32)
    for j in range(8):
        assert np.array_equal(x[j], (0.5 * sc[j]) * q[j] - 0.25 * mn[j]), j
    rng = np.random.default_rng(2)
    w = rng.normal(0, 0.1, (4, 512)).astype(np.float32)
    y = g.dequantize_blocks(g.quantize_blocks(w, g.GGML_TYPE_Q5_K), g.GGML_TYPE_Q5_K, w.size).reshape(w.shape)
    span = (w.reshape(-1, 32).max(axis=1) - np.minimum(w.reshape(-1, 32).min(axis=1), 0))
    assert (np.abs(y - w).reshape(-1, 32).max(axis=1) <= 1.6 * span / 31 + 2e-3).all()
    path = synth.ensure_model_file("micro", str(tmp_path), quant_type=g.GGML_TYPE_Q5_K)
    m = g.read_ggml(path)
    assert m.hparams.ftype == 13 and m.tensors["decoder.blocks.0.mlp.0.weight"].dtype == np.float32
