"""CPU: front-end oracles (rubato restatement, Silero graph, SmoothedVad) and the golden fixtures."""
import os

import numpy as np
import pytest

from oracle import logmel, resample, silero, vad_gate, whisper_ref
from spittle_b200 import silero_weights, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ONNX = "/root/reference/src-tauri/resources/models/silero_vad_v4.onnx"


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "oracle_golden.npz"))


@pytest.fixture(scope="module")
def sw():
    return silero_weights.load_npz(os.path.join(GOLD, "silero_v4_16k.npz"))


def test_rubato_geometry_and_forms_agree():
    assert resample.geometry(48000, 16000) == (1026, 342)
    assert resample.geometry(44100, 16000) == (1323, 480)
    h = resample.make_filter(1026, 342)
    assert abs(h.sum() - 1.0) < 1e-12 and h.argmax() == 513
    x = synth.make_clip(3, seconds=2.0, sr=48000)
    a = resample.resample_block_fft(x)
    b = resample.resample_direct_fir(x)
    assert a.shape == b.shape == ((96000 + 1023) // 1024 * 1024 // 1026 * 342,)
    assert np.abs(a - b).max() < 1e-8                    # SURVEY App. B validation
    # group delay 513 input samples = 171 output samples, unity pass-band gain
    imp = np.zeros(48000); imp[3000] = 1.0
    y = resample.resample_block_fft(imp)
    assert y.argmax() == 1171 and abs(y.sum() * 3 - 1.0) < 1e-9
    fr = resample.frame_resampler(x)
    assert fr.shape == (-(-a.shape[0] // 480), 480) and np.all(fr.reshape(-1)[a.shape[0]:] == 0)


def test_silero_fixture_matches_reference_onnx(sw):
    if not os.path.exists(ONNX):
        pytest.skip("reference tree not mounted (GPU box)")
    w = silero_weights.silero_v4_16k_from_onnx(ONNX)
    for k, _ in silero_weights.BLOB_LAYOUT:
        np.testing.assert_array_equal(w[k], sw[k])
    assert silero_weights.to_blob(w).shape == (155908,)


def test_silero_behaves_as_vad(sw):
    o = silero.SileroOracle(sw)
    p_tone = o.score(synth.make_clip(0, seconds=2.0, kind="tone"))
    o.reset()
    p_vowel = o.score(synth.make_clip(4, seconds=2.0, kind="vowel"))
    assert p_tone.max() < 0.3 and (p_vowel > 0.3).mean() > 0.5


def test_smoothed_vad_semantics():
    # onset needs 2 consecutive voiced frames; the emission at onset is the buffered prefill + current
    plan = vad_gate.smoothed_vad_plan(np.array([0, 1, 0, 1, 1, 1, 0, 0, 0, 0], bool), prefill=3, hangover=2, onset=2)
    assert plan == [(0, 0), (1, 0), (2, 0), (3, 0), (1, 4), (5, 1), (6, 1), (7, 1), (8, 0), (9, 0)]
    # a re-onset shortly after a segment re-emits frames that were already emitted (reference behaviour)
    plan = vad_gate.smoothed_vad_plan(np.array([1, 1, 0, 1, 1], bool), prefill=15, hangover=0, onset=2)
    assert plan == [(0, 0), (0, 2), (2, 0), (3, 0), (0, 5)]
    x = np.arange(5 * 480, dtype=np.float32).reshape(5, 480)
    out = vad_gate.gate_audio(x, np.array([1, 1, 0, 1, 1], np.float32), threshold=0.3, prefill=15, hangover=0, onset=2)
    assert out.shape[0] == 7 * 480
    assert vad_gate.stop_recording_pad(np.ones(100, np.float32)).shape == (20000,)
    assert vad_gate.stop_recording_pad(np.ones(16000, np.float32)).shape == (16000,)
    assert vad_gate.stop_recording_pad(np.zeros(0, np.float32)).shape == (0,)


def test_oracle_reproduces_golden_fixtures(gold, sw):
    """Drift detection: the committed fixtures were produced by this oracle at fixed seeds."""
    filt = synth.mel_filterbank(80)
    for i in (1, 3, 4):
        mel, n_len_org = logmel.logmel_f64(synth.make_clip(i, seconds=5.0), filt)
        np.testing.assert_allclose(mel[::8, :500:25], gold[f"mel_clip{i}_grid"], atol=2e-6)
        assert n_len_org == int(gold[f"mel_clip{i}_sum"][1])
    y = resample.resample_block_fft(synth.make_clip(3, seconds=1.0, sr=48000))[:2000]
    np.testing.assert_allclose(y, gold["resample_clip3_head"], atol=1e-12)
    so = silero.SileroOracle(sw)
    np.testing.assert_allclose(so.score(synth.make_clip(4, seconds=3.0, kind="vowel")), gold["silero_vowel_probs"], atol=1e-9)
    plan = vad_gate.smoothed_vad_plan(gold["gate_pattern"], 15, 15, 2)
    np.testing.assert_array_equal(np.array(plan, np.int32), gold["gate_plan"])


def test_oracle_tokens_match_golden(gold):
    model = synth.make_synthetic_model("nano", 42)
    oracle = whisper_ref.WhisperOracle(model, act_f16=True)
    x = synth.make_clip(1, seconds=30.0)
    mel, n_len_org = logmel.logmel_f64(x, model.mel_filters)
    enc = oracle.encode(logmel.mel_window(mel, 0))
    np.testing.assert_allclose(enc[::100, ::16], gold["nano_clip1_enc_grid"], atol=2e-3)
    w = oracle.decode_window(enc, 0, n_len_org, whisper_ref.DecodeConfig(n_max_override=16))
    assert w.tokens == list(gold["nano_clip1_tokens"])


def test_downmix_oracle_known_answers():
    """cpal to_sample conventions + channel mean (recorder.rs:182-201)."""
    from oracle import vad_gate
    x = np.array([32767, -32768, 0, 16384], np.int16)
    assert np.array_equal(vad_gate.downmix_mono(x, 1), np.array([32767 / 32768, -1.0, 0.0, 0.5], np.float32))
    assert np.array_equal(vad_gate.downmix_mono(x, 2), np.array([(32767 / 32768 - 1.0) / 2, 0.25], np.float32))
    u = np.array([0, 32768, 65535], np.uint16)
    assert np.array_equal(vad_gate.downmix_mono(u, 1), np.array([-1.0, 0.0, 32767 / 32768], np.float32))
    f = np.array([0.1, 0.2, 0.3, 0.4, 0.5, 0.6], np.float32)
    assert np.array_equal(vad_gate.downmix_mono(f, 3), ((f[0::3] + f[1::3]).astype(np.float32) + f[2::3]).astype(np.float32) / np.float32(3))
