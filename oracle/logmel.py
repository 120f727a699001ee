"""Oracle: whisper.cpp ``log_mel_spectrogram``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the reference has no
log-mel fixture; whisper.cpp is not vendored.  Call path restated:
  reference  managers/transcription.rs:501-503  whisper_engine.transcribe_samples(audio, ..)
   -> transcribe-rs 0.2.3 -> whisper-rs 0.13.2 ``WhisperState::full`` -> whisper.cpp
      ``whisper_pcm_to_mel_with_state`` -> ``log_mel_spectrogram``.

Two variants:
  * ``logmel_f64``  -- the mathematical definition in float64 (numpy rfft).  "Truth".
  * ``logmel_f32_faithful`` -- the same *operation order* as whisper.cpp: f32 Hann window from
    ``cosf``, f32 recursive radix-2 FFT 400->200->100->50->25 with an O(N^2) f32 DFT at N=25
    driven by a 400-entry sin/cos table, f32 power spectrum, mel accumulated in double.
    It quantifies the reference's own round-off so the GPU tolerance can be stated fairly.

Both return the mel-major [n_mel, n_len] float32 array whisper.cpp builds (n_len = 6000 for
a 30 s clip), plus n_len_org (= seek_end).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
CHUNK_S = 30
N_BINS = 201


def _pad(samples: np.ndarray, dtype) -> np.ndarray:
    """[reflect(samples[1..200]) | samples | zeros(480000 + 200)]  (App. C.1 step 2)."""
    n = samples.shape[0]
    stage1 = SAMPLE_RATE * CHUNK_S
    stage2 = N_FFT // 2
    out = np.zeros(n + stage1 + 2 * stage2, dtype=dtype)
    out[stage2:stage2 + n] = samples
    # std::reverse_copy(samples + 1, samples + 1 + 200, padded.begin())
    m = min(stage2, max(n - 1, 0))
    refl = samples[1:1 + stage2][::-1]
    out[stage2 - refl.shape[0]:stage2] = refl  # if n < 201 whisper.cpp would read OOB; we require n >= 201
    return out


def n_len_of(n_samples: int) -> tuple[int, int]:
    padded = n_samples + SAMPLE_RATE * CHUNK_S + N_FFT
    n_len = (padded - N_FFT) // HOP
    n_len_org = 1 + (n_samples + N_FFT // 2 - N_FFT) // HOP
    return n_len, n_len_org


def _finish(logspec: np.ndarray) -> np.ndarray:
    """global max-8 clamp, (x+4)/4 in double as whisper.cpp does, stored to f32."""
    mmax = float(logspec.max()) - 8.0
    out = np.maximum(logspec.astype(np.float64), mmax)
    return ((out + 4.0) / 4.0).astype(np.float32)


def logmel_f64(samples: np.ndarray, filters: np.ndarray, raw: bool = False):
    samples = np.asarray(samples, dtype=np.float32)
    n = samples.shape[0]
    assert n >= 201, "whisper.cpp reflect pad needs > 200 samples"
    n_len, n_len_org = n_len_of(n)
    padded = _pad(samples.astype(np.float64), np.float64)
    hann = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT))
    n_eff = n + N_FFT // 2
    n_frames = min(n_eff // HOP + 1, n_len)
    idx = np.arange(n_frames)[:, None] * HOP + np.arange(N_FFT)[None, :]
    frames = padded[np.minimum(idx, padded.shape[0] - 1)]
    frames = np.where(idx < n_eff, frames, 0.0) * hann[None, :]
    spec = np.fft.rfft(frames, axis=1)
    power = spec.real ** 2 + spec.imag ** 2                     # [frames, 201]
    mel = power @ filters.astype(np.float64).T                  # [frames, n_mel]
    logspec = np.full((filters.shape[0], n_len), -10.0, dtype=np.float64)
    logspec[:, :n_frames] = np.log10(np.maximum(mel, 1e-10)).T
    # whisper.cpp stores the log10 value into a float array before the clamp pass
    logspec32 = logspec.astype(np.float32)
    if raw:
        return logspec32, n_len_org
    return _finish(logspec32), n_len_org


# ---- f32-faithful variant --------------------------------------------------------------
_SIN = np.sin(2.0 * np.pi * np.arange(N_FFT) / N_FFT).astype(np.float32)   # sinf(theta)
_COS = np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT).astype(np.float32)   # cosf(theta)


def _dft32(x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """whisper.cpp dft(): re += in[n]*cos, im -= in[n]*sin, sequential in n, all f32."""
    F, N = x.shape
    step = N_FFT // N
    re = np.zeros((F, N), np.float32)
    im = np.zeros((F, N), np.float32)
    k = np.arange(N)
    for n in range(N):
        idx = (k * n * step) % N_FFT
        xn = x[:, n:n + 1]
        re = (re + xn * _COS[idx][None, :]).astype(np.float32)
        im = (im - xn * _SIN[idx][None, :]).astype(np.float32)
    return re, im


def _fft32(x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """whisper.cpp fft(): recursive radix-2, odd N falls back to dft().  x: [F, N] f32."""
    F, N = x.shape
    if N == 1:
        return x.copy(), np.zeros_like(x)
    half = N // 2
    if N - half * 2 == 1:
        return _dft32(x)
    er, ei = _fft32(np.ascontiguousarray(x[:, 0::2]))
    orr, oi = _fft32(np.ascontiguousarray(x[:, 1::2]))
    step = N_FFT // N
    idx = np.arange(half) * step
    re = _COS[idx][None, :]
    im = (-_SIN[idx])[None, :]
    f = np.float32
    # out[k] = even + (re*re_odd - im*im_odd), evaluated left to right in f32 like the C code
    t_re = ((er + re * orr).astype(f) - (im * oi).astype(f)).astype(f)
    t_im = ((ei + re * oi).astype(f) + (im * orr).astype(f)).astype(f)
    b_re = ((er - re * orr).astype(f) + (im * oi).astype(f)).astype(f)
    b_im = ((ei - re * oi).astype(f) - (im * orr).astype(f)).astype(f)
    return np.concatenate([t_re, b_re], axis=1), np.concatenate([t_im, b_im], axis=1)


def logmel_f32_faithful(samples: np.ndarray, filters: np.ndarray, raw: bool = False):
    samples = np.asarray(samples, dtype=np.float32)
    n = samples.shape[0]
    assert n >= 201
    n_len, n_len_org = n_len_of(n)
    padded = _pad(samples, np.float32)
    # fill_hann_window: 0.5*(1 - cosf(2*pi*i/400)) in f32
    hann = (0.5 * (1.0 - np.cos((2.0 * np.pi * np.arange(N_FFT)) / N_FFT))).astype(np.float32)
    n_eff = n + N_FFT // 2
    n_frames = min(n_eff // HOP + 1, n_len)
    logspec = np.full((filters.shape[0], n_len), np.float32(np.log10(1e-10)), dtype=np.float32)
    filt64 = filters.astype(np.float32)
    B = 512
    for s in range(0, n_frames, B):
        e = min(s + B, n_frames)
        idx = np.arange(s, e)[:, None] * HOP + np.arange(N_FFT)[None, :]
        fr = padded[np.minimum(idx, padded.shape[0] - 1)]
        fr = np.where(idx < n_eff, (hann[None, :] * fr).astype(np.float32), np.float32(0))
        re, im = _fft32(fr.astype(np.float32))
        power = (re[:, :N_BINS] * re[:, :N_BINS] + im[:, :N_BINS] * im[:, :N_BINS]).astype(np.float32)
        # f32 products summed in double (the 4-way unrolled f32 partial sums of the C code
        # differ from this by < 1 ulp of f32 on non-negative terms)
        prod = (power[:, None, :] * filt64[None, :, :]).astype(np.float32)
        mel = prod.astype(np.float64).sum(axis=2)
        logspec[:, s:e] = np.log10(np.maximum(mel, 1e-10)).T.astype(np.float32)
    if raw:
        return logspec, n_len_org
    return _finish(logspec), n_len_org


def mel_window(mel: np.ndarray, seek: int, n_ctx: int = 1500) -> np.ndarray:
    """Encoder input for a window starting at frame ``seek``: [n_mel, 2*n_ctx], zero-filled past
    n_len (whisper_encode_internal's copy loop, App. C.1 step 7)."""
    n_mel, n_len = mel.shape
    out = np.zeros((n_mel, 2 * n_ctx), np.float32)
    i0 = min(seek, n_len)
    i1 = min(seek + 2 * n_ctx, n_len)
    out[:, :i1 - i0] = mel[:, i0:i1]
    return out
