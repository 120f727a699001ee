"""Oracle: the two data formats on the capture side of the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Unlike the rest of oracle/, both functions restate code that IS in
the reference tree (first-hand, not from memory); only the FFT itself (rustfft) and the WAV container (hound) are
external crates, and a DFT / a 44-byte PCM header are standard.

  pcm_f32_to_i16        audio_toolkit/audio/utils.rs:17-20   ``(sample * i16::MAX as f32) as i16`` -- Rust's float->int
                        ``as`` cast truncates toward zero, saturates at the i16 range and maps NaN to 0
  AudioVisualiser       audio_toolkit/audio/visualizer.rs:20-149, constructed at audio/recorder.rs:276-282 with
                        (device rate, 512, 16, 400 Hz, 4000 Hz); ``feed`` is called once per captured chunk
                        (recorder.rs:323) and analyses the FIRST 512 samples buffered, then clears the buffer.
                        The adaptive noise floor (visualizer.rs:124-129) is state that never reaches the returned levels
                        and is not restated.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

DB_MIN, DB_MAX, GAIN, CURVE_POWER = F32(-55.0), F32(-8.0), F32(1.3), F32(0.7)     # visualizer.rs:4-7


def pcm_f32_to_i16(samples: np.ndarray) -> np.ndarray:
    x = samples.astype(F32) * F32(32767.0)
    y = np.trunc(x.astype(np.float64))
    y = np.where(np.isnan(y), 0.0, np.clip(y, -32768.0, 32767.0))
    return y.astype(np.int16)


def bucket_ranges(sample_rate: int, window_size: int = 512, buckets: int = 16, freq_min: float = 400.0,
                  freq_max: float = 4000.0):
    """visualizer.rs:38-66, every operation in f32 like the reference."""
    nyq = F32(sample_rate) / F32(2.0)
    fmin, fmax = min(F32(freq_min), nyq), min(F32(freq_max), nyq)
    out = []
    for b in range(buckets):
        log_start = F32(F32(b) / F32(buckets)) ** 2
        log_end = F32(F32(b + 1) / F32(buckets)) ** 2
        start_hz = F32(fmin + F32(fmax - fmin) * F32(log_start))
        end_hz = F32(fmin + F32(fmax - fmin) * F32(log_end))
        start_bin = int(F32(start_hz * F32(window_size)) / F32(sample_rate))
        end_bin = int(F32(end_hz * F32(window_size)) / F32(sample_rate))
        if end_bin <= start_bin:
            end_bin = start_bin + 1
        out.append((min(start_bin, window_size // 2), min(end_bin, window_size // 2)))
    return out


class AudioVisualiser:
    def __init__(self, sample_rate: int, window_size: int = 512, buckets: int = 16, freq_min: float = 400.0,
                 freq_max: float = 4000.0):
        self.n, self.buckets = window_size, buckets
        i = np.arange(window_size, dtype=F32)
        self.window = (F32(0.5) * (F32(1.0) - np.cos(F32(2.0) * F32(np.pi) * i / F32(window_size), dtype=F32))).astype(F32)
        self.ranges = bucket_ranges(sample_rate, window_size, buckets, freq_min, freq_max)
        self.buffer = np.zeros(0, F32)

    def feed(self, samples: np.ndarray):
        self.buffer = np.concatenate([self.buffer, samples.astype(F32)])
        if self.buffer.shape[0] < self.n:
            return None
        w = self.buffer[: self.n]
        mean = F32(w.sum(dtype=F32) / F32(self.n))
        x = ((w - mean) * self.window).astype(F32)
        spec = np.fft.fft(x.astype(np.float64))                    # rustfft f32: differs from this by ~1e-7 relative
        mag2 = (np.abs(spec) ** 2)
        out = np.zeros(self.buckets, F32)
        for bi, (s, e) in enumerate(self.ranges):
            if s >= e or e > self.n // 2:
                continue
            avg_power = F32(mag2[s:e].sum() / (e - s))
            db = F32(20.0) * F32(np.log10(np.sqrt(avg_power, dtype=F32) / F32(self.n))) if avg_power > 1e-12 else F32(-80.0)
            normalized = min(max(F32((db - DB_MIN) / (DB_MAX - DB_MIN)), F32(0.0)), F32(1.0))
            out[bi] = min(max(F32(np.power(F32(normalized * GAIN), CURVE_POWER)), F32(0.0)), F32(1.0))
        for i in range(1, self.buckets - 1):                       # in place: bucket i-1 is already smoothed
            out[i] = F32(out[i] * F32(0.7) + out[i - 1] * F32(0.15) + out[i + 1] * F32(0.15))
        self.buffer = np.zeros(0, F32)
        return out


def visualiser_levels(pcm: np.ndarray, chunk_len: int, sample_rate: int) -> np.ndarray:
    """Levels of every chunk of one stream fed through a fresh AudioVisualiser: [n_chunks, 16] (chunk_len >= 512)."""
    v = AudioVisualiser(sample_rate)
    n_chunks = pcm.shape[0] // chunk_len
    return np.stack([v.feed(pcm[c * chunk_len:(c + 1) * chunk_len]) for c in range(n_chunks)])
