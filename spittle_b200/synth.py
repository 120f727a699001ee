"""Synthetic inputs for parity tests and benchmarks.

* audio generators: tone / chirp / gated noise burst / mix / vowel-like, 16 kHz or 48 kHz,
  f32 in [-1, 1], seeded ``numpy.random.default_rng(1000 + i)`` for clip ``i``;
* slaney mel filterbank (what whisper.cpp reads from the model file, App. C.1 item 6);
* random-init weight recipes for each named architecture, written to the GGML legacy
  format so the oracle, the CPU baseline and the GPU engine load bit-identical tensors.

There is no network here, so real checkpoints/datasets are not available; every result
produced from this module is labelled ``"data": "synthetic"``.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np

from .ggml_format import ARCHS, GgmlModel, WhisperHParams, write_ggml, read_ggml

SAMPLE_RATE = 16000


# --------------------------------------------------------------------------------------
# audio
# --------------------------------------------------------------------------------------
def tone(n: int, sr: int, f: float = 440.0, a: float = 0.3) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / sr
    return (a * np.sin(2 * np.pi * f * t)).astype(np.float32)


def chirp(n: int, sr: int, f0: float = 100.0, f1: float = 7000.0, a: float = 0.3) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / sr
    dur = n / sr
    phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / dur)
    return (a * np.sin(phase)).astype(np.float32)


def noise_burst(n: int, sr: int, rng: np.random.Generator, sigma: float = 0.1) -> np.ndarray:
    x = rng.normal(0.0, sigma, n)
    gate = ((np.arange(n) // (sr // 2)) % 2 == 0)
    return np.clip(x * gate, -1, 1).astype(np.float32)


def vowel_like(n: int, sr: int, rng: np.random.Generator) -> np.ndarray:
    """Glottal pulse train -> 3 formant resonators -> syllabic AM (scores as speech on Silero,
    SURVEY.md Appendix A validation)."""
    f0 = rng.uniform(90, 220)
    am = rng.uniform(2.0, 4.0)
    t = np.arange(n, dtype=np.float64) / sr
    # pulse train with slight jitter-free period
    phase = (t * f0) % 1.0
    src = (phase < 0.1).astype(np.float64) - 0.1
    y = np.zeros(n)
    formants = [(rng.uniform(600, 800), 80.0), (rng.uniform(1100, 1400), 90.0),
                (rng.uniform(2400, 2800), 120.0)]
    from scipy.signal import lfilter
    for fc, bw in formants:
        r = np.exp(-np.pi * bw / sr)
        th = 2 * np.pi * fc / sr
        b = [1 - r]
        a = [1.0, -2 * r * np.cos(th), r * r]
        y += lfilter(b, a, src)
    env = 0.5 * (1 + np.sin(2 * np.pi * am * t - np.pi / 2))
    y = y * env
    y = 0.3 * y / (np.max(np.abs(y)) + 1e-9)
    return y.astype(np.float32)


def make_clip(i: int, seconds: float = 30.0, sr: int = SAMPLE_RATE, kind: Optional[str] = None) -> np.ndarray:
    """Clip ``i`` of the synthetic corpus. kind=None cycles tone/chirp/noise/mix(/vowel)."""
    rng = np.random.default_rng(1000 + i)
    n = int(round(seconds * sr))
    kinds = ["tone", "chirp", "noise", "mix", "vowel", "mix"]
    k = kind or kinds[i % len(kinds)]
    if k == "tone":
        f = [220.0, 440.0, 1000.0, 3000.0][(i // len(kinds)) % 4]
        x = tone(n, sr, f)
    elif k == "chirp":
        x = chirp(n, sr)
    elif k == "noise":
        x = noise_burst(n, sr, rng)
    elif k == "vowel":
        x = vowel_like(n, sr, rng)
    elif k == "mix":
        w = rng.uniform(0.2, 1.0, 4)
        f = rng.choice([220.0, 440.0, 1000.0, 3000.0])
        x = (w[0] * tone(n, sr, f) + w[1] * chirp(n, sr, rng.uniform(80, 400), rng.uniform(2000, 7000))
             + w[2] * noise_burst(n, sr, rng) + w[3] * vowel_like(n, sr, rng))
        x = x / max(1.0, np.max(np.abs(x)) / 0.9)
        if rng.random() < 0.5:  # leading / trailing silence (exercises VAD onset/hangover)
            lead = int(rng.uniform(1.0, 2.0) * sr)
            trail = int(rng.uniform(1.0, 2.0) * sr)
            x[:lead] = 0
            x[n - trail:] = 0
    else:
        raise ValueError(k)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


# --------------------------------------------------------------------------------------
# mel filterbank (librosa "slaney" scale + slaney norm; what ggml model files carry)
# --------------------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(n_mels: int, n_fft: int = 400, sr: int = SAMPLE_RATE) -> np.ndarray:
    """[n_mels, n_fft//2+1] f32, identical in construction to librosa.filters.mel(norm='slaney')."""
    n_bins = n_fft // 2 + 1
    fftfreqs = np.linspace(0, sr / 2, n_bins)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_bins))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


# --------------------------------------------------------------------------------------
# vocab
# --------------------------------------------------------------------------------------
def synthetic_vocab(n: int = 50257) -> List[bytes]:
    """ASCII pseudo-words, unique per id; id 220 is " " (GPT-2 position of the blank token,
    looked up by text for suppress_blank in whisper.cpp)."""
    out = []
    for i in range(n):
        if i == 220:
            out.append(b" ")
            continue
        s = []
        v = i
        while True:
            s.append(chr(ord("a") + v % 26))
            v //= 26
            if v == 0:
                break
        w = "".join(s)
        out.append(((" " if i % 2 == 0 else "") + w).encode())
    return out


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2))
    st = np.arange(length)[:, None] * inv[None, :]
    return np.concatenate([np.sin(st), np.cos(st)], axis=1).astype(np.float32)


RECIPES: Dict[str, dict] = {
    # HF/OpenAI style init: everything N(0, 0.02^2).  Degenerate under greedy decoding
    #  -- kept as the "baseline recipe".
    "hf": dict(kind="const", std=0.02, emb_std=0.02, ts_emb_scale=1.0, dec_pos_std=0.02),
    # Designed recipe: fan-in scaled projections (activations stay O(1) through depth),
    # sharper cross-attention so the audio actually steers the decoder, large tied token
    # embedding so the softmax is peaked (text tokens win most steps), a slightly up-scaled
    # timestamp block so timestamp pairs appear at a realistic rate, small decoder residual
    # gains so the current token matters, and a random +-1 final-LayerNorm gain that removes the
    # "repeat my own input token" attractor of a tied embedding (see make_synthetic_model).
    "sharp": dict(kind="fanin", gain=1.0, conv_gain=2.0, qk_gain=2.0, cross_qk_gain=4.0,
                  cross_out_gain=0.5, dec_res_gain=0.25, emb_std=0.2, ts_emb_scale=1.05,
                  dec_pos_std=0.5, bias_std=0.02, final_ln_signs=True),
}


def make_synthetic_model(arch: str, seed: int = 42, recipe: str = "sharp",
                         overrides: Optional[dict] = None) -> GgmlModel:
    hp: WhisperHParams = ARCHS[arch]
    rc = dict(RECIPES[recipe])
    if overrides:
        rc.update(overrides)
    rng = np.random.default_rng(seed)
    d, dt = hp.n_audio_state, hp.n_text_state
    T: Dict[str, np.ndarray] = {}

    def lin(out_f, in_f, gain=1.0):
        if rc["kind"] == "const":
            w = rng.standard_normal((out_f, in_f), dtype=np.float32) * rc["std"]
        else:
            w = rng.standard_normal((out_f, in_f), dtype=np.float32) * (gain * rc["gain"] / np.sqrt(in_f))
        return w.astype(np.float16)

    def bias(n):
        s = rc.get("bias_std", 0.0) if rc["kind"] != "const" else 0.0
        return (rng.standard_normal(n, dtype=np.float32) * s).astype(np.float32)

    def ln(prefix, n):
        T[prefix + ".weight"] = np.ones(n, np.float32)
        T[prefix + ".bias"] = np.zeros(n, np.float32)

    def attn(prefix, n, qk, out_gain=1.0):
        T[prefix + ".query.weight"] = lin(n, n, qk)
        T[prefix + ".query.bias"] = bias(n)
        T[prefix + ".key.weight"] = lin(n, n, qk)
        T[prefix + ".value.weight"] = lin(n, n)
        T[prefix + ".value.bias"] = bias(n)
        T[prefix + ".out.weight"] = lin(n, n, out_gain)
        T[prefix + ".out.bias"] = bias(n)

    def mlp(prefix, n, out_gain=1.0):
        T[prefix + ".0.weight"] = lin(4 * n, n)
        T[prefix + ".0.bias"] = bias(4 * n)
        T[prefix + ".2.weight"] = lin(n, 4 * n, out_gain)
        T[prefix + ".2.bias"] = bias(n)

    cg = rc.get("conv_gain", 1.0)
    T["encoder.positional_embedding"] = sinusoids(hp.n_audio_ctx, d)
    if rc["kind"] == "const":
        T["encoder.conv1.weight"] = (rng.standard_normal((d, hp.n_mels, 3), dtype=np.float32) * rc["std"]).astype(np.float16)
        T["encoder.conv2.weight"] = (rng.standard_normal((d, d, 3), dtype=np.float32) * rc["std"]).astype(np.float16)
    else:
        T["encoder.conv1.weight"] = (rng.standard_normal((d, hp.n_mels, 3), dtype=np.float32) * (cg / np.sqrt(3 * hp.n_mels))).astype(np.float16)
        T["encoder.conv2.weight"] = (rng.standard_normal((d, d, 3), dtype=np.float32) * (cg / np.sqrt(3 * d))).astype(np.float16)
    T["encoder.conv1.bias"] = bias(d).reshape(d, 1)
    T["encoder.conv2.bias"] = bias(d).reshape(d, 1)
    qk = rc.get("qk_gain", 1.0)
    for i in range(hp.n_audio_layer):
        p = f"encoder.blocks.{i}"
        ln(p + ".attn_ln", d)
        attn(p + ".attn", d, qk)
        ln(p + ".mlp_ln", d)
        mlp(p + ".mlp", d)
    ln("encoder.ln_post", d)

    T["decoder.positional_embedding"] = (rng.standard_normal((hp.n_text_ctx, dt), dtype=np.float32)
                                         * rc["dec_pos_std"]).astype(np.float32)
    emb = rng.standard_normal((hp.n_vocab, dt), dtype=np.float32) * rc["emb_std"]
    from .ggml_format import SpecialTokens
    sp = SpecialTokens.from_n_vocab(hp.n_vocab)
    emb[sp.beg:] *= rc["ts_emb_scale"]
    T["decoder.token_embedding.weight"] = emb.astype(np.float16)
    cqk = rc.get("cross_qk_gain", 1.0)
    cog = rc.get("cross_out_gain", 1.0)
    drg = rc.get("dec_res_gain", 1.0)
    for i in range(hp.n_text_layer):
        p = f"decoder.blocks.{i}"
        ln(p + ".attn_ln", dt)
        attn(p + ".attn", dt, qk, drg)
        ln(p + ".cross_attn_ln", dt)
        attn(p + ".cross_attn", dt, cqk, cog)
        ln(p + ".mlp_ln", dt)
        mlp(p + ".mlp", dt, drg)
    ln("decoder.ln", dt)
    if rc.get("final_ln_signs", False):
        # random +-1 gain on the final LayerNorm: with a tied embedding and skip connections a
        # random decoder otherwise locks onto repeating its own input token (E_t . E_t term)
        T["decoder.ln.weight"] = rng.choice(np.array([-1.0, 1.0], np.float32), size=dt).astype(np.float32)

    return GgmlModel(hparams=hp, mel_filters=mel_filterbank(hp.n_mels), vocab=synthetic_vocab(50257 if hp.n_vocab >= 51865 else 50256),
                     tensors=T)


def ensure_model_file(arch: str, directory: str, seed: int = 42, recipe: str = "sharp", quant_type: int = None) -> str:
    """Write (once) ``ggml-synth-<arch>-<recipe>-s<seed>[-q<type>].bin`` under ``directory``.
    quant_type: a ggml block type (ggml_format.GGML_TYPE_Q*) for the weight matrices, like whisper.cpp's quantize tool."""
    os.makedirs(directory, exist_ok=True)
    suffix = "" if quant_type is None else f"-q{quant_type}"
    path = os.path.join(directory, f"ggml-synth-{arch}-{recipe}-s{seed}{suffix}.bin")
    if not os.path.exists(path):
        tmp = path + ".tmp%d" % os.getpid()
        write_ggml(tmp, make_synthetic_model(arch, seed, recipe), quant_type=quant_type)
        os.replace(tmp, path)
    return path


def with_non_speech_vocab(model, ids):
    """A copy of `model` whose vocabulary entries at `ids` are whisper.cpp's non-speech strings ('(', ' (', the music notes, ...,
    ' -', " '"): the synthetic vocabulary is plain ASCII words, so suppress_nst would have nothing to suppress.  `ids` are
    usually the tokens a decode emits WITHOUT the rule, so that switching it on has to change the result."""
    import dataclasses
    from oracle.whisper_ref import NON_SPEECH_TOKENS          # test helper only: the product never imports the oracle
    words = []
    for t in NON_SPEECH_TOKENS:
        words += [t.encode("utf-8"), (" " + t).encode("utf-8")]
    words += [b" -", b" '"]
    vocab = list(model.vocab)
    for i, w in zip(ids, words):
        vocab[i] = w
    return dataclasses.replace(model, vocab=vocab)


__all__ = ["make_clip", "mel_filterbank", "make_synthetic_model", "ensure_model_file",
           "synthetic_vocab", "sinusoids", "tone", "chirp", "noise_burst", "vowel_like",
           "read_ggml", "SAMPLE_RATE"]
