"""GPU parity of the capture-side formats through the C ABI against the oracle."""
import wave

import numpy as np
import pytest

from oracle import capture_formats as cf
from spittle_b200 import audio_toolkit, capi, synth

pytestmark = pytest.mark.gpu


def test_pcm_to_i16_bit_exact(cuda_dev):
    import torch
    rng = np.random.default_rng(3)
    edge = np.array([0.0, 1.0, -1.0, 1.5, -1.5, 0.99999, 3.0517578e-05, -3.0517578e-05, np.nan, np.inf, -np.inf], np.float32)
    for n in (0, 1, 7, 8, 9, 4099, 480000):
        x = rng.uniform(-1.2, 1.2, n).astype(np.float32)
        x[: min(n, edge.size)] = edge[: min(n, edge.size)]
        xd = torch.from_numpy(x).to(cuda_dev)
        out = torch.full((max(n, 1),), 12345, dtype=torch.int16, device=cuda_dev)
        if n:
            capi.pcm_f32_to_i16_dev(xd.data_ptr(), out.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert (out.cpu().numpy()[:n] == cf.pcm_f32_to_i16(x)).all(), n


def test_pcm_to_i16_full_size_properties(cuda_dev):
    """10 minutes of audio: odd symmetry (trunc toward zero) and monotonicity, without the oracle."""
    import torch
    n = 16000 * 600
    x = torch.empty(n, device=cuda_dev).uniform_(-1.0, 1.0)
    a = torch.empty(n, dtype=torch.int16, device=cuda_dev)
    b = torch.empty(n, dtype=torch.int16, device=cuda_dev)
    st = torch.cuda.current_stream().cuda_stream
    capi.pcm_f32_to_i16_dev(x.data_ptr(), a.data_ptr(), n, st)
    xm = (-x).contiguous()
    capi.pcm_f32_to_i16_dev(xm.data_ptr(), b.data_ptr(), n, st)
    torch.cuda.synchronize()
    assert torch.equal(a, -b)
    idx = torch.argsort(x)
    assert (a[idx][1:] >= a[idx][:-1]).all()
    assert (a.float() - x * 32767.0).abs().max().item() < 1.0 + 1e-3


def test_save_wav_file_round_trip(cuda_dev, tmp_path):
    x = synth.make_clip(2, seconds=2.0)
    p = str(tmp_path / "clip.wav")
    audio_toolkit.save_wav_file(p, x)
    with wave.open(p, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 16000, x.shape[0])
        data = np.frombuffer(w.readframes(x.shape[0]), np.int16)
    assert (data == cf.pcm_f32_to_i16(x)).all()


@pytest.mark.parametrize("sr,chunk", [(48000, 1024), (16000, 512), (44100, 2048)])
def test_visualiser_levels_match_oracle(cuda_dev, sr, chunk):
    n_streams, n_chunks = 5, 12
    streams = np.stack([synth.make_clip(40 + i, seconds=n_chunks * chunk / sr + 0.01, sr=sr)[: n_chunks * chunk]
                        for i in range(n_streams)])
    streams[3] = 0.0                                            # silence: every level 0
    streams[4, : chunk] *= 1e-4                                 # near the -55 dB floor: the steep end of the x^0.7 curve
    got = audio_toolkit.AudioVisualiser(sr).levels(streams, chunk).cpu().numpy()
    want = np.stack([cf.visualiser_levels(streams[s], chunk, sr) for s in range(n_streams)])
    assert got.shape == want.shape == (n_streams, n_chunks, 16)
    assert np.abs(got - want).max() <= 2e-4, np.abs(got - want).max()
    assert (got[3] == 0).all() and got.min() >= 0.0 and got.max() <= 1.0


def test_host_pointer_forms_match_oracle(cuda_dev):
    """sb_pcm_f32_to_i16 / sb_visualiser_levels: what a CUDA-free host (the reference's Rust side) calls."""
    x = synth.make_clip(43, seconds=0.6, sr=48000)
    assert (capi.pcm_f32_to_i16(x * 4.0) == cf.pcm_f32_to_i16(x * 4.0)).all()
    assert capi.pcm_f32_to_i16(np.zeros(0, np.float32)).shape == (0,)
    got = capi.visualiser_levels(x, 1024, 48000)
    want = cf.visualiser_levels(x, 1024, 48000)
    assert got.shape == want.shape == (x.shape[0] // 1024, 16)
    assert np.abs(got - want).max() <= 2e-4
    assert capi.visualiser_levels(x[:700], 1024, 48000).shape == (0, 16)      # no whole chunk: nothing emitted
