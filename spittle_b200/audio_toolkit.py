"""Host-side mirror of the reference's capture toolkit, batched over independent streams on the GPU.

Same names and argument meaning as the reference (src-tauri/src/audio_toolkit):

  FrameResampler(in_hz, out_hz, frame_ms)     audio/resampler.rs:16-98   (rubato FftFixedIn)
  SileroVad(model_path, threshold)            vad/silero.rs:19-51        (vad-rs + onnxruntime)
  SmoothedVad(inner, prefill, hangover, onset)  vad/smoothed.rs:20-96
  run_consumer(...)                           audio/recorder.rs:255-373  (resample -> VAD -> append)
  stop_recording_pad(samples)                 managers/audio.rs:466-475
  save_wav_file(path, samples)                audio/utils.rs:7-26        (hound: 16 kHz mono i16)
  AudioVisualiser(rate, 512, 16, 400, 4000)   audio/visualizer.rs:20-149 (mic-level buckets per captured chunk)

The reference is a streaming, single-microphone loop; here a whole recording per stream is pushed
at once (push(all) + finish()), many streams per call.  torch is used only for device memory.
Every numeric step runs in libspittle_b200.so; there is no CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import capi, silero_weights

WHISPER_SAMPLE_RATE = 16000          # audio_toolkit/constants.rs:1
SILERO_FRAME_SAMPLES = 480           # vad/silero.rs:9-11


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("spittle_b200.audio_toolkit needs a CUDA device (no CPU fallback)")
    return torch


class FrameResampler:
    def __init__(self, in_hz: int, out_hz: int = WHISPER_SAMPLE_RATE, frame_ms: int = 30):
        self.frame_samples = int(round(out_hz * frame_ms / 1000.0))
        if self.frame_samples != SILERO_FRAME_SAMPLES:
            raise capi.SbError(-7, "only 30 ms frames at 16 kHz are implemented")
        self._r = capi.Resampler(in_hz, out_hz)

    def process(self, streams):
        """streams: [n_streams, n_in] f32 (numpy or CUDA tensor) -> CUDA tensor [n_streams, n_frames, 480]."""
        torch = _torch()
        x = streams if hasattr(streams, "is_cuda") else torch.from_numpy(np.ascontiguousarray(streams, np.float32))
        x = x.cuda().contiguous()
        n_streams, n_in = x.shape
        _, _, n_frames = self._r.geometry(n_in)
        out = torch.empty((n_streams, n_frames * SILERO_FRAME_SAMPLES), dtype=torch.float32, device=x.device)
        if n_frames:
            self._r.run_dev(x.data_ptr(), x.stride(0), n_in, n_streams, out.data_ptr(), out.stride(0),
                            torch.cuda.current_stream().cuda_stream)
        return out.view(n_streams, n_frames, SILERO_FRAME_SAMPLES)


class SileroVad:
    """Probability scoring + `prob > threshold`; LSTM state persists across calls until reset()
    (the reference never resets the inner Silero state: vad/mod.rs:25)."""

    def __init__(self, model, threshold: float):
        if not (0.0 <= threshold <= 1.0):
            raise ValueError("threshold must be between 0.0 and 1.0")     # vad/silero.rs:20-22
        if isinstance(model, str):
            w = silero_weights.load_npz(model) if model.endswith(".npz") else silero_weights.silero_v4_16k_from_onnx(model)
        else:
            w = model
        self._v = capi.Vad(silero_weights.to_blob(w))
        self.threshold = float(threshold)
        self._state = None

    def reset(self):
        self._state = None

    def score(self, frames):
        """frames: CUDA tensor [n_streams, n_frames, 480] -> probabilities [n_streams, n_frames]."""
        torch = _torch()
        f = frames.contiguous()
        n_streams, n_frames, fs = f.shape
        assert fs == SILERO_FRAME_SAMPLES, f"expected {SILERO_FRAME_SAMPLES} samples, got {fs}"
        if self._state is None or self._state[0].shape[1] != n_streams:
            self._state = (torch.zeros((2, n_streams, 64), dtype=torch.float32, device=f.device),
                           torch.zeros((2, n_streams, 64), dtype=torch.float32, device=f.device))
        probs = torch.empty((n_streams, n_frames), dtype=torch.float32, device=f.device)
        if n_frames == 0:
            return probs
        ws = torch.empty(capi.Vad.workspace_bytes(n_streams, n_frames), dtype=torch.uint8, device=f.device)
        self._v.score_dev(f.data_ptr(), n_frames * fs, n_streams, n_frames, self._state[0].data_ptr(),
                          self._state[1].data_ptr(), probs.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return probs

    def is_voice(self, frames):
        return self.score(frames) > self.threshold


class SmoothedVad:
    def __init__(self, inner: SileroVad, prefill_frames: int, hangover_frames: int, onset_frames: int):
        self.inner = inner
        self.prefill, self.hangover, self.onset = prefill_frames, hangover_frames, onset_frames

    def reset(self):
        pass   # SmoothedVad::reset does not touch the inner VAD (vad/smoothed.rs:98-104)

    def gate(self, frames, probs=None) -> List[np.ndarray]:
        """Kept samples per stream (what run_consumer accumulates in processed_samples)."""
        torch = _torch()
        f = frames.contiguous()
        n_streams, n_frames, fs = f.shape
        if n_frames == 0:
            return [np.zeros(0, np.float32) for _ in range(n_streams)]
        if probs is None:
            probs = self.inner.score(f)
        # every onset may re-emit the prefill ring: worst case (prefill + 1) frames per (onset) voiced frames
        max_frames = n_frames * (self.prefill + 1 + self.onset) // max(1, self.onset) + self.prefill + 1
        out = torch.empty((n_streams, max_frames * fs), dtype=torch.float32, device=f.device)
        counts = torch.zeros(n_streams, dtype=torch.int32, device=f.device)
        ws = torch.empty(capi.vad_gate_workspace_bytes(n_streams, n_frames), dtype=torch.uint8, device=f.device)
        capi.vad_gate_dev(probs.data_ptr(), f.data_ptr(), n_frames * fs, n_streams, n_frames, self.inner.threshold,
                          self.prefill, self.hangover, self.onset, out.data_ptr(), out.stride(0), counts.data_ptr(),
                          ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        cnt = counts.cpu().numpy()
        host = out.cpu().numpy()
        return [host[s, : int(cnt[s]) * fs].copy() for s in range(n_streams)]


def stop_recording_pad(samples: np.ndarray) -> np.ndarray:
    """0 < n < 16000  =>  resize to 20000 with zeros (managers/audio.rs:466-475)."""
    n = samples.shape[0]
    if 0 < n < WHISPER_SAMPLE_RATE:
        out = np.zeros(WHISPER_SAMPLE_RATE * 5 // 4, np.float32)
        out[:n] = samples
        return out
    return samples


def run_consumer(streams, in_sample_rate: int, vad: Optional[SmoothedVad]) -> List[np.ndarray]:
    """resample -> 30 ms frames -> VAD gate -> kept 16 kHz samples per stream (recorder.rs:255-373,
    Cmd::Start ... Cmd::Stop over one whole recording per stream)."""
    frames = FrameResampler(in_sample_rate).process(streams)
    if vad is None:
        host = frames.cpu().numpy()
        return [host[s].reshape(-1).copy() for s in range(host.shape[0])]
    return vad.gate(frames)


def save_wav_file(file_path: str, samples) -> None:
    """16 kHz mono 16-bit PCM WAV of one recording (audio/utils.rs:7-26; called by managers/history.rs:180-214).
    The f32 -> i16 conversion `(s * 32767) as i16` runs on the GPU; the 44-byte header is written here."""
    import struct
    torch = _torch()
    x = samples if hasattr(samples, "is_cuda") else torch.from_numpy(np.ascontiguousarray(samples, np.float32))
    x = x.cuda().contiguous().view(-1)
    out = torch.empty(x.numel(), dtype=torch.int16, device=x.device)
    if x.numel():
        capi.pcm_f32_to_i16_dev(x.data_ptr(), out.data_ptr(), x.numel(), torch.cuda.current_stream().cuda_stream)
    data = out.cpu().numpy().tobytes()
    with open(file_path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, WHISPER_SAMPLE_RATE, WHISPER_SAMPLE_RATE * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", len(data)))
        f.write(data)


class AudioVisualiser:
    """Batched mirror of AudioVisualiser::feed: levels(streams, chunk_len) returns what the reference's visualiser
    would have emitted for every chunk of every stream ([n_streams, n_chunks, 16]); the reference analyses the first
    512 samples of each fed chunk."""

    def __init__(self, sample_rate: int, window_size: int = 512, buckets: int = 16, freq_min: float = 400.0,
                 freq_max: float = 4000.0):
        if (window_size, buckets, freq_min, freq_max) != (512, 16, 400.0, 4000.0):
            raise capi.SbError(-7, "only the reference's visualiser geometry (512, 16, 400 Hz, 4000 Hz) is implemented")
        self.sample_rate = sample_rate

    def levels(self, streams, chunk_len: int):
        torch = _torch()
        x = streams if hasattr(streams, "is_cuda") else torch.from_numpy(np.ascontiguousarray(streams, np.float32))
        x = x.cuda().contiguous()
        n_streams, n_in = x.shape
        n_chunks = n_in // chunk_len
        out = torch.empty((n_streams, n_chunks, 16), dtype=torch.float32, device=x.device)
        if n_chunks:
            capi.visualiser_levels_dev(x.data_ptr(), x.stride(0), n_streams, n_chunks, chunk_len, self.sample_rate,
                                       out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return out
