"""Oracle of the capture-side formats against known answers derived from the reference source."""
import numpy as np

from oracle import capture_formats as cf


def test_pcm_to_i16_rust_cast_semantics():
    # audio_toolkit/audio/utils.rs:19: (sample * i16::MAX as f32) as i16 -- truncation toward zero, saturation, NaN -> 0
    x = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 1.5, -1.5, 3.0517578e-05, -3.0517578e-05, 0.99999, np.nan, np.inf, -np.inf], np.float32)
    want = np.array([0, 32767, -32767, 16383, -16383, 32767, -32768, 0, 0, 32766, 0, 32767, -32768], np.int16)
    assert (cf.pcm_f32_to_i16(x) == want).all()


def test_bucket_ranges_match_the_reference_formula():
    # visualizer.rs:38-66 at 48 kHz, window 512, 16 buckets, 400-4000 Hz: quadratic spacing, >= 1 bin per bucket
    r = cf.bucket_ranges(48000)
    assert r[0] == (4, 5) and r[-1][1] == 42 and len(r) == 16
    assert all(e > s for s, e in r) and all(r[i][0] <= r[i + 1][0] for i in range(15))
    # 16 kHz: Nyquist 8 kHz does not clip the 4 kHz upper edge; bins are three times wider in Hz
    r16 = cf.bucket_ranges(16000)
    assert r16[0][0] == 12 and r16[-1][1] == 128


def test_visualiser_tone_lands_in_the_right_bucket_and_silence_is_zero():
    sr, n = 48000, 1024
    t = np.arange(n) / sr
    v = cf.AudioVisualiser(sr)
    assert v.feed(np.zeros(100, np.float32)) is None               # fewer than 512 buffered samples: nothing emitted
    lv = v.feed(np.zeros(n, np.float32))
    assert lv.shape == (16,) and (lv == 0).all()                  # zero power -> -80 dB -> clamped to 0
    tone = (0.5 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    lv = cf.AudioVisualiser(sr).feed(tone)
    ranges = cf.bucket_ranges(sr)
    k = int(1000.0 * 512 / sr)
    hot = [i for i, (s, e) in enumerate(ranges) if s <= k < e][0]
    assert int(np.argmax(lv)) == hot and lv[hot] > 0.8 and (lv >= 0).all() and (lv <= 1).all()
    # only the first 512 samples of a chunk are analysed: changing the tail changes nothing
    tone2 = tone.copy(); tone2[512:] = 0
    assert np.array_equal(cf.AudioVisualiser(sr).feed(tone2), lv)


def test_oracle_reproduces_capture_format_fixtures():
    """tests/golden/capture_formats_golden.npz (tests/golden/make_golden.py) pins the oracle against drift."""
    import os
    from spittle_b200 import synth
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "capture_formats_golden.npz"))
    x = synth.make_clip(41, seconds=0.3, sr=48000)[: 12 * 1024]
    assert np.allclose(cf.visualiser_levels(x, 1024, 48000), gold["vis_clip41_48k_chunk1024"], atol=1e-6)
    x16 = synth.make_clip(42, seconds=0.5)[: 12 * 512]
    assert np.allclose(cf.visualiser_levels(x16, 512, 16000), gold["vis_clip42_16k_chunk512"], atol=1e-6)
    assert np.array_equal(cf.pcm_f32_to_i16(x16[:256] * 4.0), gold["i16_clip42_head"])
    assert np.abs(gold["i16_clip42_head"]).max() == 32767 or np.abs(gold["i16_clip42_head"].astype(np.int32)).max() == 32768
