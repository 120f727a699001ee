"""Oracle: rubato 0.16.2 ``FftFixedIn<f32>`` as wrapped by the reference's FrameResampler.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: rubato is not vendored
(src-tauri/Cargo.toml:58, Cargo.lock:5384); its published algorithm is restated from
SURVEY.md Appendix B and anchored on the reference call sites
  audio_toolkit/audio/resampler.rs:24     FftFixedIn::<f32>::new(in_hz, out_hz, 1024, 1, 1)
  audio_toolkit/audio/resampler.rs:37-64  push(): re-chunk to 1024, process(), emit 480-sample frames
  audio_toolkit/audio/resampler.rs:66-84  finish(): zero-pad the last chunk and the last frame
Float64 throughout ("truth").  Two forms:
  * resample_block_fft   -- rubato's overlap-add block algorithm (what the reference runs)
  * resample_direct_fir  -- y[m] = sum_k h[k] x[D*m - k], the polyphase form the CUDA kernel uses
They agree to ~1e-9 (stop-band leakage of the 1026-tap filter, SURVEY App. B validation).
"""
from __future__ import annotations

from math import gcd

import numpy as np

CHUNK = 1024
FRAME = 480


def geometry(fs_in: int, fs_out: int, chunk: int = CHUNK):
    g = gcd(fs_in, fs_out)
    fft_chunks = -(-chunk // (fs_in // g))
    return fft_chunks * (fs_in // g), fft_chunks * (fs_out // g)     # fft_size_in, fft_size_out


def make_filter(n_in: int, n_out: int) -> np.ndarray:
    """make_sincs(npoints=n_in, factor=1, cutoff, BlackmanHarris2), unit sum."""
    if n_in > n_out:
        cutoff = np.float32(0.4) ** np.float32(16.0 / n_out) * n_out / n_in
    else:
        cutoff = np.float32(0.4) ** np.float32(16.0 / n_in)
    cutoff = float(np.float32(cutoff))
    x = np.arange(n_in, dtype=np.float64)
    w = (0.35875 - 0.48829 * np.cos(2 * np.pi * x / n_in) + 0.14128 * np.cos(4 * np.pi * x / n_in)
         - 0.01168 * np.cos(6 * np.pi * x / n_in)) ** 2
    t = (x - n_in // 2) * cutoff
    s = np.where(t == 0, 1.0, np.sin(np.pi * t) / np.where(t == 0, 1.0, np.pi * t))
    h = w * s
    return h / h.sum()


def fed_stream(x: np.ndarray, chunk: int = CHUNK) -> np.ndarray:
    """What push()+finish() hand to rubato: the input zero-padded to a whole number of chunks."""
    n = x.shape[0]
    total = -(-n // chunk) * chunk
    out = np.zeros(total, np.float64)
    out[:n] = x
    return out


def resample_block_fft(x: np.ndarray, fs_in: int = 48000, fs_out: int = 16000) -> np.ndarray:
    n_in, n_out = geometry(fs_in, fs_out)
    h = make_filter(n_in, n_out)
    filt = np.fft.rfft(np.concatenate([h / (2 * n_in), np.zeros(n_in)]))       # n_in + 1 bins
    fed = fed_stream(x)
    n_blocks = fed.shape[0] // n_in
    overlap = np.zeros(n_out)
    out = np.zeros(n_blocks * n_out)
    new_len = n_in + 1 if n_in < n_out else n_out
    for b in range(n_blocks):
        X = np.fft.rfft(np.concatenate([fed[b * n_in:(b + 1) * n_in], np.zeros(n_in)]))
        Y = np.zeros(n_out + 1, np.complex128)
        Y[:new_len] = X[:new_len] * filt[:new_len]
        y = np.fft.irfft(Y, 2 * n_out) * (2 * n_out)          # realfft inverse is unnormalised
        out[b * n_out:(b + 1) * n_out] = y[:n_out] + overlap
        overlap = y[n_out:]
    return out


def resample_direct_fir(x: np.ndarray, fs_in: int = 48000, fs_out: int = 16000) -> np.ndarray:
    n_in, n_out = geometry(fs_in, fs_out)
    assert n_in % n_out == 0, "direct form implemented for integer decimation only"
    D = n_in // n_out
    h = make_filter(n_in, n_out)
    fed = fed_stream(x)
    n_blocks = fed.shape[0] // n_in
    full = np.convolve(fed[: n_blocks * n_in], h)
    return full[0: n_blocks * n_in: D][: n_blocks * n_out].copy()


def frame_resampler(x: np.ndarray, fs_in: int = 48000, fs_out: int = 16000, form: str = "block") -> np.ndarray:
    """FrameResampler.push(all) + finish(): [n_frames, 480] with the last frame zero-padded."""
    if fs_in == fs_out:
        y = np.asarray(x, np.float64)
    else:
        y = resample_block_fft(x, fs_in, fs_out) if form == "block" else resample_direct_fir(x, fs_in, fs_out)
    n_frames = -(-y.shape[0] // FRAME)
    out = np.zeros(n_frames * FRAME)
    out[: y.shape[0]] = y
    return out.reshape(n_frames, FRAME)
