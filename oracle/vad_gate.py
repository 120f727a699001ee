"""Oracle: SmoothedVad onset/hangover/prefill gate + the capture consumer's concatenation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates reference code that IS in the tree:
  audio_toolkit/vad/smoothed.rs:41-96        SmoothedVad::push_frame
  audio_toolkit/audio/recorder.rs:284-314    handle_frame: Speech(buf) => out.extend(buf)
  managers/audio.rs:132-134                  threshold 0.3, prefill 15, hangover 15, onset 2
  managers/audio.rs:466-475                  short-clip rule: 0 < n < 16000 => resize to 20000
Integer/index work: parity with the CUDA path is bit-exact.
"""
from __future__ import annotations

from collections import deque
from typing import List, Tuple

import numpy as np

FRAME = 480


def smoothed_vad_plan(is_voice: np.ndarray, prefill: int = 15, hangover: int = 15, onset: int = 2) -> List[Tuple[int, int]]:
    """For each frame t: (first_source_frame, n_frames) emitted (n_frames = 0 for Noise)."""
    buf = deque()
    in_speech = False
    onset_counter = 0
    hangover_counter = 0
    plan = []
    for t, v in enumerate(is_voice):
        buf.append(t)
        while len(buf) > prefill + 1:
            buf.popleft()
        v = bool(v)
        if not in_speech and v:
            onset_counter += 1
            if onset_counter >= onset:
                in_speech = True
                hangover_counter = hangover
                onset_counter = 0
                plan.append((buf[0], len(buf)))          # prefill + current, whatever is buffered
            else:
                plan.append((t, 0))
        elif in_speech and v:
            hangover_counter = hangover
            plan.append((t, 1))
        elif in_speech and not v:
            if hangover_counter > 0:
                hangover_counter -= 1
                plan.append((t, 1))
            else:
                in_speech = False
                plan.append((t, 0))
        else:
            onset_counter = 0
            plan.append((t, 0))
    return plan


def gate_audio(frames: np.ndarray, probs: np.ndarray, threshold: float = 0.3, prefill: int = 15,
               hangover: int = 15, onset: int = 2) -> np.ndarray:
    """frames [n_frames, 480] -> concatenated kept samples (recorder.rs handle_frame)."""
    plan = smoothed_vad_plan(probs > threshold, prefill, hangover, onset)
    out = [frames[s:s + n].reshape(-1) for s, n in plan if n > 0]
    return np.concatenate(out) if out else np.zeros(0, np.float32)


def stop_recording_pad(samples: np.ndarray, rate: int = 16000) -> np.ndarray:
    n = samples.shape[0]
    if 0 < n < rate:
        out = np.zeros(rate * 5 // 4, samples.dtype)
        out[:n] = samples
        return out
    return samples


def downmix_mono(interleaved: np.ndarray, channels: int) -> np.ndarray:
    """cpal input callback of AudioRecorder::build_stream (audio/recorder.rs:182-201): to_sample::<f32>()
    per sample (i16: x / 32768, u16: (x - 32768) / 32768, f32: identity), then per frame the f32 sum over the
    channels in order, divided by the channel count.  Bit-exact f32 arithmetic."""
    x = np.asarray(interleaved)
    if x.dtype == np.int16:
        f = x.astype(np.float32) / np.float32(32768.0)
    elif x.dtype == np.uint16:
        f = (x.astype(np.int32) - 32768).astype(np.float32) / np.float32(32768.0)
    else:
        f = x.astype(np.float32)
    if channels == 1:
        return f.copy()
    fr = f[: (f.shape[0] // channels) * channels].reshape(-1, channels)
    acc = np.zeros(fr.shape[0], np.float32)
    for c in range(channels):
        acc = (acc + fr[:, c]).astype(np.float32)
    return (acc / np.float32(channels)).astype(np.float32)
