"""Oracle: Silero VAD v4, 16 kHz branch + the reference's thresholding.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED by the reference; the graph
and the weights are first-hand from the reference's own artefact
``src-tauri/resources/models/silero_vad_v4.onnx`` (onnxruntime itself is absent here).

Restates what the reference executes per 30 ms frame at
``audio_toolkit/vad/silero.rs:41-50``: ``vad_rs::Vad::compute(frame)`` -> ort Session::run of the
ONNX graph with input[1,480], sr=16000, h,c[2,1,64]; then ``prob > threshold``.
All arithmetic in float64 ("truth"); the LSTM state is carried across frames like vad-rs does.
"""
from __future__ import annotations

import numpy as np

FRAME = 480


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _dw_conv(x, w, b):
    """depthwise k5 pad 2.  x [C,T], w [C,5], b [C]."""
    C, T = x.shape
    xp = np.pad(x, ((0, 0), (2, 2)))
    y = np.zeros((C, T))
    for k in range(5):
        y += w[:, k:k + 1] * xp[:, k:k + T]
    return y + b[:, None]


class SileroOracle:
    def __init__(self, weights):
        self.w = {k: np.asarray(v, np.float64) for k, v in weights.items()}
        self.reset()

    def reset(self):
        self.h = np.zeros((2, 64))
        self.c = np.zeros((2, 64))

    def features(self, frame: np.ndarray) -> np.ndarray:
        """480 samples -> the 64-vector fed to the LSTM."""
        w = self.w
        x = np.asarray(frame, np.float64)
        assert x.shape == (FRAME,)
        xp = np.pad(x, 96, mode="reflect")                                   # 672
        cols = np.stack([xp[64 * t: 64 * t + 256] for t in range(7)], axis=1)    # [256, 7]
        ft = w["stft_basis"] @ cols                                          # [258, 7]
        mag = np.sqrt(ft[:129] ** 2 + ft[129:] ** 2)                         # [129, 7]
        spect = np.log(1.0 + 1048576.0 * mag)
        mean = spect.mean(axis=0)                                            # [7]
        padded = np.concatenate([mean[1:4][::-1], mean, mean[-4:-1][::-1]])  # reflect 3
        mean1 = np.array([np.dot(w["norm_filter"], padded[t:t + 7]) for t in range(7)])
        norm = spect - mean1.mean()
        x1 = np.concatenate([mag, norm], axis=0)                             # [258, 7]
        relu = lambda a: np.maximum(a, 0.0)
        # block 1
        r = relu(_dw_conv(x1, w["b1_dw_w"], w["b1_dw_b"]))
        y = w["b1_pw_w"] @ r + w["b1_pw_b"][:, None] + w["b1_proj_w"] @ x1 + w["b1_proj_b"][:, None]
        y = relu(y)
        y = relu(w["b1_down_w"] @ y[:, ::2] + w["b1_down_b"][:, None])       # T 7 -> 4
        # block 2
        r = relu(_dw_conv(y, w["b2_dw_w"], w["b2_dw_b"]))
        z = w["b2_pw_w"] @ r + w["b2_pw_b"][:, None] + w["b2_proj_w"] @ y + w["b2_proj_b"][:, None]
        z = relu(z)
        z = relu(w["b2_down_w"] @ z[:, ::2] + w["b2_down_b"][:, None])       # T 4 -> 2
        # block 3 (identity residual)
        r = relu(_dw_conv(z, w["b3_dw_w"], w["b3_dw_b"]))
        u = relu(w["b3_pw_w"] @ r + w["b3_pw_b"][:, None] + z)
        u = relu(w["b3_down_w"] @ u[:, ::2] + w["b3_down_b"][:, None])       # T 2 -> 1
        # block 4
        r = relu(_dw_conv(u, w["b4_dw_w"], w["b4_dw_b"]))
        v = w["b4_pw_w"] @ r + w["b4_pw_b"][:, None] + w["b4_proj_w"] @ u + w["b4_proj_b"][:, None]
        v = relu(v)
        v = relu(w["b4_down_w"] @ v + w["b4_down_b"][:, None])               # [64, 1]
        return v[:, 0]

    def _lstm(self, x, layer):
        """ONNX LSTM, gate order i, o, f, c; B = Wb || Rb."""
        w = self.w
        W, R, B = w[f"lstm{layer + 1}_w"], w[f"lstm{layer + 1}_r"], w[f"lstm{layer + 1}_b"]
        g = W @ x + R @ self.h[layer] + B[:256] + B[256:]
        i, o, f, ct = _sigmoid(g[:64]), _sigmoid(g[64:128]), _sigmoid(g[128:192]), np.tanh(g[192:])
        c = f * self.c[layer] + i * ct
        h = o * np.tanh(c)
        self.c[layer], self.h[layer] = c, h
        return h

    def compute(self, frame: np.ndarray) -> float:
        z = self.features(frame)
        h1 = self._lstm(z, 0)
        h2 = self._lstm(h1, 1)
        y = float(np.dot(self.w["dec_w"], np.maximum(h2, 0.0)) + self.w["dec_b"][0])
        return float(_sigmoid(y))

    def score(self, pcm16k: np.ndarray) -> np.ndarray:
        """Whole frames only (the reference's FrameResampler always delivers 480-sample frames)."""
        n = pcm16k.shape[0] // FRAME
        return np.array([self.compute(pcm16k[i * FRAME:(i + 1) * FRAME]) for i in range(n)])
