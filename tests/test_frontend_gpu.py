"""GPU parity: resampler, Silero VAD, SmoothedVad gate vs the oracles, through the C ABI."""
import os

import numpy as np
import pytest

from oracle import resample, silero, vad_gate
from spittle_b200 import audio_toolkit, capi, silero_weights, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SILERO = os.path.join(GOLD, "silero_v4_16k.npz")

# max abs error vs the f64 rubato restatement.  SURVEY 8(d) proposed 1e-5; measured on B200 in round 2: 1.7e-6 for the shipped
# polyphase f16 (3-pass) kernel (3.1e-6 for the 3xTF32 Toeplitz form it replaced), 1.2e-6 for the dense block operator of the
# non-integer ratios -- the stated tolerance is 3x the largest.
RESAMPLE_TOL = 5e-6
# max abs error on the speech probability vs the f64 oracle on identical input; decisions (prob > 0.3) must be identical
# outside it.  The shipped path evaluates the STFT as f64 FFTs plus the exact residual of the stored basis (frontend.cu):
# measured 1.9e-6 (round 1's f32 convolution: ~1.1e-4, tolerance 3e-4, because log(1 + 2^20 |X|) magnifies the round-off
# of a 256-term f32 dot product on quiet bins).  Stated tolerance: 1e-5.  The direct-convolution kernel kept for a basis
# that is not a windowed DFT still has the f32 behaviour and keeps the old bound.
SILERO_TOL = 1e-5
SILERO_DIRECT_TOL = 3e-4


def test_resampler_matches_rubato_oracle(cuda_dev):
    import torch
    rs = audio_toolkit.FrameResampler(48000, 16000)
    for ids, secs in [([0, 1, 2, 3, 4], 3.0), ([5], 30.0), ([3, 4], 0.04), ([1], 1025 / 48000.0)]:
        x = np.stack([synth.make_clip(i, seconds=secs, sr=48000) for i in ids])
        got = rs.process(x).cpu().numpy()
        for s in range(len(ids)):
            ref = resample.frame_resampler(x[s])
            assert got[s].shape == ref.shape
            err = np.abs(got[s] - ref).max() if ref.size else 0.0
            assert err <= RESAMPLE_TOL, (ids[s], secs, err)
    # golden vector (first 2000 samples of clip 3)
    gold = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    got = rs.process(synth.make_clip(3, seconds=1.0, sr=48000)[None]).cpu().numpy().reshape(-1)[:2000]
    assert np.abs(got - gold["resample_clip3_head"]).max() <= RESAMPLE_TOL
    # pass-through when the device already delivers 16 kHz (resampler.rs:38-41)
    x16 = synth.make_clip(2, seconds=1.0)
    fr = audio_toolkit.FrameResampler(16000, 16000).process(x16[None]).cpu().numpy().reshape(-1)
    assert np.array_equal(fr[:16000], x16) and np.all(fr[16000:] == 0) and fr.shape[0] == 34 * 480


def test_resampler_rational_ratios(cuda_dev):
    """44.1 / 22.05 / 11.025 / 8 kHz capture devices: rubato's FftFixedIn block operator (1323 -> 480 samples at 44.1 kHz) as a dense
    split-precision GEMM, against the f64 restatement of the block FFT algorithm (reference: resampler.rs:16-27 accepts any rate)."""
    for fs in (44100, 22050, 11025, 8000, 24000, 12000):     # 24 kHz: fft_size_out = 684 is not a multiple of 8 (padded GEMM)
        x = np.stack([synth.make_clip(50 + i, seconds=2.0, sr=fs, kind=k) for i, k in enumerate(["vowel", "noise", "mix"])])
        got = audio_toolkit.FrameResampler(fs, 16000).process(x).cpu().numpy()
        for s_ in range(3):
            ref = resample.frame_resampler(x[s_], fs, 16000)
            assert got[s_].shape == ref.shape, (fs, got[s_].shape, ref.shape)
            err = np.abs(got[s_] - ref).max()
            print(f"resample {fs} -> 16000: max err {err:.2e} (max |y| {np.abs(ref).max():.2f})")
            assert err <= RESAMPLE_TOL, (fs, err)
    # 30 s at 44.1 kHz, several streams: the chunked workspace path and the zero tail of the last frame
    x = np.stack([synth.make_clip(60 + i, seconds=30.0, sr=44100) for i in range(3)])
    got = audio_toolkit.FrameResampler(44100, 16000).process(x).cpu().numpy()
    for s_ in range(3):
        ref = resample.frame_resampler(x[s_], 44100, 16000)
        assert got[s_].shape == ref.shape and np.abs(got[s_] - ref).max() <= RESAMPLE_TOL
    # a ratio whose block operator would not fit (47 999 Hz: 47 999 x 16 000 samples per block) fails loudly
    with pytest.raises(capi.SbError):
        audio_toolkit.FrameResampler(47999, 16000)


def test_resampler_other_integer_ratios(cuda_dev):
    """The polyphase f16 kernel at decimation 2, 4 and 6 (32 / 64 / 96 kHz capture devices) against the f64 rubato restatement."""
    for fs in (32000, 64000, 96000):
        x = np.stack([synth.make_clip(30 + i, seconds=1.5, sr=fs, kind=k) for i, k in enumerate(["vowel", "noise"])])
        got = audio_toolkit.FrameResampler(fs, 16000).process(x).cpu().numpy()
        for s_ in range(2):
            ref = resample.frame_resampler(x[s_], fs, 16000)
            assert got[s_].shape == ref.shape
            err = np.abs(got[s_] - ref).max()
            print(f"resample {fs} -> 16000: max err {err:.2e}")
            assert err <= RESAMPLE_TOL, (fs, err)


def test_resampler_properties_full_size(cuda_dev):
    """Size-independent properties at the C5 stream length (30 s @ 48 kHz): linearity, unity DC gain,
    group delay of 171 output samples."""
    rs = audio_toolkit.FrameResampler(48000, 16000)
    n = 1440000
    a = synth.make_clip(3, seconds=30.0, sr=48000)
    b = synth.make_clip(4, seconds=30.0, sr=48000)
    imp = np.zeros(n, np.float32); imp[30000] = 1.0
    dc = np.full(n, 0.25, np.float32)
    y = rs.process(np.stack([a, b, 2 * a - 3 * b, imp, dc])).cpu().numpy().reshape(5, -1)
    assert np.abs(y[2] - (2 * y[0] - 3 * y[1])).max() < 5e-6
    assert y[3].argmax() == (30000 + 513) // 3 and abs(y[3].sum() * 3 - 1.0) < 1e-5
    assert np.abs(y[4][400:470000] - 0.25).max() < 1e-6


def test_silero_probs_and_decisions_match_oracle(cuda_dev):
    import torch
    w = silero_weights.load_npz(SILERO)
    vad = audio_toolkit.SileroVad(SILERO, 0.3)
    kinds = ["vowel", "noise", "mix", "tone", "vowel", "mix"]
    clips = np.stack([synth.make_clip(10 + i, seconds=4.5, kind=k) for i, k in enumerate(kinds)])
    n_frames = clips.shape[1] // 480
    frames = torch.from_numpy(clips[:, : n_frames * 480]).cuda().view(len(kinds), n_frames, 480)
    probs = vad.score(frames).cpu().numpy()
    borderline = 0
    for s in range(len(kinds)):
        o = silero.SileroOracle(w)
        ref = o.score(clips[s])
        err = np.abs(probs[s] - ref).max()
        print(f"silero {kinds[s]}: max |dp| {err:.2e}")
        assert err <= SILERO_TOL, (kinds[s], err)
        near = np.abs(ref - 0.3) < SILERO_TOL
        borderline += int(near.sum())
        assert np.array_equal((probs[s] > 0.3)[~near], (ref > 0.3)[~near])
    print("frames within tolerance of the 0.3 threshold:", borderline)
    # state carry: two half-length calls == one call (vad-rs keeps h/c between compute() calls)
    vad.reset()
    h = n_frames // 2
    p2 = np.concatenate([vad.score(frames[:, :h].contiguous()).cpu().numpy(), vad.score(frames[:, h:].contiguous()).cpu().numpy()], axis=1)
    assert np.abs(p2 - probs).max() < 1e-6
    # golden vector
    gold = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    vad.reset()
    xg = synth.make_clip(4, seconds=3.0, kind="vowel")
    pg = vad.score(torch.from_numpy(xg[: 100 * 480]).cuda().view(1, 100, 480)).cpu().numpy()[0]
    assert np.abs(pg - gold["silero_vowel_probs"]).max() <= SILERO_TOL


def test_silero_direct_convolution_path(cuda_dev, monkeypatch):
    """sb_vad_create picks the FFT form of the STFT when the basis is a windowed DFT (the shipped model) and the
    direct-convolution kernel otherwise: both against the oracle, (a) forced through SB_SILERO_DIRECT, (b) on a basis
    that is NOT a DFT (every row scaled by a different gain -- |X| changes, so a wrongly taken FFT path would show)."""
    import torch
    w = silero_weights.load_npz(SILERO)
    clips = np.stack([synth.make_clip(40 + i, seconds=2.4, kind=k) for i, k in enumerate(["vowel", "mix", "tone"])])
    n_frames = clips.shape[1] // 480
    frames = torch.from_numpy(clips[:, : n_frames * 480]).cuda().view(3, n_frames, 480)

    monkeypatch.setenv("SB_SILERO_DIRECT", "1")
    sv = audio_toolkit.SileroVad(SILERO, 0.3)
    direct = sv.score(frames).cpu().numpy()
    monkeypatch.delenv("SB_SILERO_DIRECT")
    sv2 = audio_toolkit.SileroVad(SILERO, 0.3)
    fft = sv2.score(frames).cpu().numpy()
    for s_ in range(3):
        ref = silero.SileroOracle(w).score(clips[s_])
        assert np.abs(direct[s_] - ref).max() <= SILERO_DIRECT_TOL
        assert np.abs(fft[s_] - ref).max() <= SILERO_TOL
    print("FFT vs direct STFT path, max |dp|:", float(np.abs(fft - direct).max()))
    # (b) a non-DFT basis
    w2 = dict(w)
    gains = (1.0 + 0.25 * np.cos(np.arange(258))).astype(np.float32)
    w2["stft_basis"] = (w["stft_basis"] * gains[:, None]).astype(np.float32)
    sv3 = audio_toolkit.SileroVad(w2, 0.3)
    got = sv3.score(frames).cpu().numpy()
    for s_ in range(3):
        ref = silero.SileroOracle(w2).score(clips[s_])
        assert np.abs(got[s_] - ref).max() <= SILERO_DIRECT_TOL


def test_frontend_is_deterministic(cuda_dev):
    """Same input, same bits: the tensor-core resampler (48 k polyphase, 44.1 k dense), the Silero front (f64 FFT exchange, partial
    tiles summed in a fixed order) and the tensor-core LSTM are run three times each -- a shared-memory race or an order-dependent
    reduction would show as differing bits (compute-sanitizer is not available on the GPU pool)."""
    import torch
    x48 = np.stack([synth.make_clip(70 + i, seconds=3.0, sr=48000, kind=k) for i, k in enumerate(["vowel", "noise", "mix", "tone", "chirp"])])
    x44 = np.stack([synth.make_clip(75 + i, seconds=3.0, sr=44100, kind=k) for i, k in enumerate(["vowel", "noise", "mix"])])
    rs48, rs44 = audio_toolkit.FrameResampler(48000), audio_toolkit.FrameResampler(44100)
    f0 = rs48.process(x48)
    g0 = rs44.process(x44)
    sv = audio_toolkit.SileroVad(SILERO, 0.3)
    p0 = sv.score(f0)
    for _ in range(2):
        assert torch.equal(rs48.process(x48), f0) and torch.equal(rs44.process(x44), g0)
        sv.reset()
        assert torch.equal(sv.score(f0), p0)


@pytest.mark.parametrize("n_streams,n_frames", [(1, 1), (1, 5), (9, 3), (3, 33)])
def test_silero_ragged_shapes(cuda_dev, n_streams, n_frames):
    """Frame counts that are not multiples of the kernels' tiles (4 frames per CTA in the front, 32 in the conv blocks) and stream
    counts that are not multiples of the LSTM's 8 streams per CTA."""
    import torch
    w = silero_weights.load_npz(SILERO)
    clips = np.stack([synth.make_clip(80 + i, seconds=(n_frames * 480 + 7) / 16000.0, kind=["vowel", "noise", "mix"][i % 3])
                      for i in range(n_streams)])
    frames = torch.from_numpy(clips[:, : n_frames * 480].copy()).cuda().view(n_streams, n_frames, 480)
    probs = audio_toolkit.SileroVad(SILERO, 0.3).score(frames).cpu().numpy()
    assert probs.shape == (n_streams, n_frames)
    for s_ in range(n_streams):
        ref = silero.SileroOracle(w).score(clips[s_, : n_frames * 480])
        assert np.abs(probs[s_] - ref).max() <= SILERO_TOL, (s_, np.abs(probs[s_] - ref).max())


def test_silero_full_size_properties(cuda_dev):
    """Size-independent properties at the C5 stream length (30 s = 1000 frames): a stream's probabilities do not depend on which
    other streams share the batch (bit-identical alone and inside a batch of 11), and two calls over the halves of the stream
    equal one call over the whole (the LSTM state is carried like vad-rs does between compute() calls)."""
    import torch
    kinds = ["vowel", "noise", "mix", "tone", "chirp"]
    clips = np.stack([synth.make_clip(90 + i, seconds=30.0, kind=kinds[i % 5]) for i in range(11)])
    frames = torch.from_numpy(clips).cuda().view(11, 1000, 480)
    sv = audio_toolkit.SileroVad(SILERO, 0.3)
    whole = sv.score(frames)
    assert whole.shape == (11, 1000) and bool(torch.isfinite(whole).all()) and float(whole.min()) >= 0.0 and float(whole.max()) <= 1.0
    for s_ in (0, 4, 10):
        sv.reset()
        alone = sv.score(frames[s_: s_ + 1].contiguous())
        assert torch.equal(alone[0], whole[s_]), s_
    sv.reset()
    a = sv.score(frames[:, :517].contiguous())
    b = sv.score(frames[:, 517:].contiguous())
    assert float((torch.cat([a, b], dim=1) - whole).abs().max()) < 1e-6
    # one stream against the f64 oracle over the full 30 s
    w = silero_weights.load_npz(SILERO)
    ref = silero.SileroOracle(w).score(clips[2])
    assert np.abs(whole[2].cpu().numpy() - ref).max() <= SILERO_TOL


def test_gate_is_bit_exact(cuda_dev):
    import torch
    vad = audio_toolkit.SileroVad(SILERO, 0.3)
    sm = audio_toolkit.SmoothedVad(vad, 15, 15, 2)
    rng = np.random.default_rng(5)
    n_streams, n_frames = 33, 257
    frames = rng.uniform(-1, 1, (n_streams, n_frames, 480)).astype(np.float32)
    probs = rng.uniform(0, 1, (n_streams, n_frames)).astype(np.float32)
    # bursty voiced patterns so onset / hangover / prefill re-emission all occur
    for s in range(n_streams):
        run = rng.integers(1, 12)
        t = 0
        while t < n_frames:
            v = rng.random() < 0.4
            probs[s, t:t + run] = 0.9 if v else 0.05
            t += run
            run = rng.integers(1, 25)
    probs[0, :] = 0.0
    probs[1, :] = 1.0
    probs[2, :] = 0.3                      # exactly at the threshold: prob > 0.3 is false
    got = sm.gate(torch.from_numpy(frames).cuda(), torch.from_numpy(probs).cuda())
    for s in range(n_streams):
        ref = vad_gate.gate_audio(frames[s], probs[s], 0.3, 15, 15, 2)
        assert got[s].shape == ref.shape and np.array_equal(got[s], ref), s
    assert got[0].size == 0 and got[2].size == 0 and got[1].size == (n_frames - 1 + 2) * 480 - 480


def test_capture_chain_matches_oracle(cuda_dev):
    """48 kHz streams -> FrameResampler -> SileroVad -> SmoothedVad -> kept 16 kHz samples (run_consumer).

    Silero's log(1 + 2^20 |X|) front magnifies f32-level differences of its *input* on quiet bins (a
    1e-6 change of a sample moves a probability by up to ~3e-3), so stage parity is asserted stage by
    stage on identical inputs, and the end-to-end kept audio must be identical unless a frame's
    probability lies within CHAIN_TOL of the 0.3 threshold."""
    import torch
    CHAIN_TOL = 5e-3
    w = silero_weights.load_npz(SILERO)
    sv = audio_toolkit.SileroVad(SILERO, 0.3)
    vad = audio_toolkit.SmoothedVad(sv, 15, 15, 2)
    kinds = ["vowel", "mix", "tone", "mix"]
    x48 = np.stack([synth.make_clip(20 + i, seconds=6.0, sr=48000, kind=k) for i, k in enumerate(kinds)])
    frames = audio_toolkit.FrameResampler(48000).process(x48)
    frames_h = frames.cpu().numpy()
    probs = sv.score(frames).cpu().numpy()
    got = vad.gate(frames, torch.from_numpy(probs).cuda())
    sv.reset()
    got2 = audio_toolkit.run_consumer(x48, 48000, vad)
    n_same = 0
    for s in range(len(kinds)):
        fr = resample.frame_resampler(x48[s])
        assert np.abs(frames_h[s] - fr).max() <= RESAMPLE_TOL                       # stage 1
        o = silero.SileroOracle(w)
        p_same_in = np.array([o.compute(f) for f in frames_h[s]])
        assert np.abs(p_same_in - probs[s]).max() <= SILERO_TOL                      # stage 2, identical input
        ref_same = vad_gate.gate_audio(frames_h[s], probs[s], 0.3, 15, 15, 2)
        assert np.array_equal(got[s], ref_same) and np.array_equal(got2[s], ref_same)   # stage 3, bit exact
        o = silero.SileroOracle(w)
        p64 = np.array([o.compute(f) for f in fr])                                   # all-f64 chain
        ref = vad_gate.gate_audio(fr.astype(np.float32), p64, 0.3, 15, 15, 2)
        if got[s].shape == ref.shape:
            assert ref.size == 0 or np.abs(got[s] - ref).max() <= RESAMPLE_TOL
            n_same += 1
        else:
            flips = np.nonzero((p64 > 0.3) != (probs[s] > 0.3))[0]
            assert len(flips) > 0 and np.all(np.abs(p64[flips] - 0.3) < CHAIN_TOL), (s, flips, p64[flips])
    assert n_same >= len(kinds) - 1
    assert got[2].size == 0                       # a pure tone never opens the gate
    assert audio_toolkit.stop_recording_pad(np.ones(10, np.float32)).shape == (20000,)


@pytest.mark.parametrize("fmt,dtype", [(capi.SB_SAMPLE_F32, np.float32), (capi.SB_SAMPLE_I16, np.int16), (capi.SB_SAMPLE_U16, np.uint16)])
@pytest.mark.parametrize("channels", [1, 2, 6])
def test_downmix_mono_bit_exact(cuda_dev, fmt, dtype, channels):
    """row a14: cpal callback down-mix (recorder.rs:182-201), bit-exact vs the oracle; ragged frame counts."""
    import torch
    from oracle import vad_gate
    rng = np.random.default_rng(7 + channels)
    n_streams, n_frames = 3, 4801
    if dtype == np.float32:
        host = rng.uniform(-1, 1, (n_streams, n_frames * channels + 5)).astype(np.float32)
    else:
        info = np.iinfo(dtype)
        host = rng.integers(info.min, info.max + 1, (n_streams, n_frames * channels + 5)).astype(dtype)
    tdt = {np.float32: torch.float32, np.int16: torch.int16, np.uint16: torch.uint16}[dtype]
    dev_in = torch.from_numpy(host).to("cuda")
    assert dev_in.dtype == tdt
    dev_out = torch.zeros((n_streams, n_frames + 3), dtype=torch.float32, device="cuda")
    capi.downmix_mono_dev(dev_in.data_ptr(), fmt, channels, host.shape[1], n_frames, n_streams, dev_out.data_ptr(), n_frames + 3)
    torch.cuda.synchronize()
    got = dev_out.cpu().numpy()
    for s in range(n_streams):
        ref = vad_gate.downmix_mono(host[s, : n_frames * channels], channels)
        assert np.array_equal(got[s, :n_frames], ref)
        assert np.all(got[s, n_frames:] == 0)


def test_blueprint_named_host_entries_match_the_batched_forms(cuda_dev):
    """sb_resample_48k_16k / sb_silero_v4 / sb_vad_gate (SURVEY 8(b) names; host pointers, one stream): same results as
    the oracle / the batched device forms they wrap."""
    x48 = synth.make_clip(6, seconds=2.0, sr=48000, kind="vowel")
    y = capi.resample_48k_16k(x48)
    ref = resample.resample_block_fft(x48)
    n = min(y.shape[0], ref.shape[0])
    assert y.shape[0] % 480 == 0 and y.shape[0] >= n > 30000
    assert np.abs(y[:n] - ref[:n]).max() <= RESAMPLE_TOL
    # Silero on the resampled audio, state carried across two calls
    w = silero_weights.load_npz(SILERO)
    vad = capi.Vad(silero_weights.to_blob(w))
    h, c = np.zeros((2, 64), np.float32), np.zeros((2, 64), np.float32)
    half = (y.shape[0] // 480 // 2) * 480
    probs = np.concatenate([capi.silero_v4(vad, y[:half], h, c), capi.silero_v4(vad, y[half:], h, c)])
    o = silero.SileroOracle(w)
    assert np.abs(probs - o.score(y)).max() <= SILERO_TOL
    assert np.abs(h).max() > 0                                   # the state came back
    # gate: bit-exact against the oracle plan
    kept = capi.vad_gate(probs, y, 0.3, 15, 15, 2)
    want = vad_gate.gate_audio(y.reshape(-1, 480), probs, 0.3, 15, 15, 2)
    assert kept.shape == want.shape and np.array_equal(kept, want)
    vad.close()
