"""Regenerates the fixtures under tests/golden/ (run in the build container, where /root/reference
is mounted; the GPU box only reads the committed outputs).

  silero_v4_16k.npz      the 16 kHz-branch tensors of the reference's own
                         src-tauri/resources/models/silero_vad_v4.onnx (weights only; MIT-licensed model)
  capture_formats_golden.npz   outputs of oracle/capture_formats.py (the restatement of the reference's own
                         audio/utils.rs and audio/visualizer.rs) on seeded synthetic inputs
  oracle_golden.npz      outputs of THIS repo's oracle on seeded synthetic inputs -- they pin the oracle
                         against drift; none of them comes from the reference (it has no fixtures on
                         this path)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import capture_formats, logmel, resample, silero, vad_gate, whisper_ref      # noqa: E402
from spittle_b200 import silero_weights, synth                          # noqa: E402

ONNX = "/root/reference/src-tauri/resources/models/silero_vad_v4.onnx"


def capture_formats_fixture():
    c = {}
    x = synth.make_clip(41, seconds=0.3, sr=48000)[: 12 * 1024]
    c["vis_clip41_48k_chunk1024"] = capture_formats.visualiser_levels(x, 1024, 48000)
    x16 = synth.make_clip(42, seconds=0.5)[: 12 * 512]
    c["vis_clip42_16k_chunk512"] = capture_formats.visualiser_levels(x16, 512, 16000)
    c["i16_clip42_head"] = capture_formats.pcm_f32_to_i16(x16[:256] * 4.0)          # x4: exercises the saturation
    np.savez_compressed(os.path.join(HERE, "capture_formats_golden.npz"), **c)
    print("wrote", sorted(c))


def main():
    if "--capture-only" in sys.argv:
        return capture_formats_fixture()
    capture_formats_fixture()
    w = silero_weights.silero_v4_16k_from_onnx(ONNX)
    np.savez_compressed(os.path.join(HERE, "silero_v4_16k.npz"), **w)
    g = {}
    # log-mel: a coarse grid of cells + the clamp floor for three clips
    filt = synth.mel_filterbank(80)
    for i in (1, 3, 4):
        x = synth.make_clip(i, seconds=5.0)
        mel, n_len_org = logmel.logmel_f64(x, filt)
        g[f"mel_clip{i}_grid"] = mel[::8, :500:25].astype(np.float32)
        g[f"mel_clip{i}_sum"] = np.array([mel[:, :501].astype(np.float64).sum(), n_len_org])
    # resampler: first 2000 output samples of a 48 kHz mix
    x48 = synth.make_clip(3, seconds=1.0, sr=48000)
    g["resample_clip3_head"] = resample.resample_block_fft(x48)[:2000].astype(np.float64)
    # Silero: 100 frame probabilities on a vowel-like clip and on gated noise
    so = silero.SileroOracle(w)
    for i, kind in ((4, "vowel"), (2, "noise")):
        so.reset()
        g[f"silero_{kind}_probs"] = so.score(synth.make_clip(i, seconds=3.0, kind=kind))
    # gate plan on a fixed voiced pattern
    pat = np.array([0, 0, 1, 0, 1, 1, 1, 0] + [0] * 20 + [1] * 5 + [0] * 30 + [1, 1, 0, 1, 1], bool)
    g["gate_pattern"] = pat
    g["gate_plan"] = np.array(vad_gate.smoothed_vad_plan(pat, 15, 15, 2), np.int32)
    # greedy tokens of the nano model (16 steps) for two clips
    model = synth.make_synthetic_model("nano", 42)
    oracle = whisper_ref.WhisperOracle(model, act_f16=True)
    for i in (1, 2):
        xc = synth.make_clip(i, seconds=30.0)
        mel, n_len_org = logmel.logmel_f64(xc, model.mel_filters)
        enc = oracle.encode(logmel.mel_window(mel, 0))
        wr = oracle.decode_window(enc, 0, n_len_org, whisper_ref.DecodeConfig(n_max_override=16))
        g[f"nano_clip{i}_tokens"] = np.array(wr.tokens, np.int32)
        g[f"nano_clip{i}_enc_grid"] = enc[::100, ::16].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **g)
    print("wrote", sorted(g))


if __name__ == "__main__":
    main()
