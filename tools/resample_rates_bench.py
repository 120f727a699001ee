"""Resampler throughput at 44.1 kHz (dense block operator through the tcgen05 GEMM) next to 48 kHz (polyphase f16 mma.sync):
1000 synthetic 30 s streams, device time.  Usage: python tools/resample_rates_bench.py"""
import os, sys, json, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spittle_b200 import audio_toolkit, synth
for fs in (44100, 48000):
    base = np.stack([synth.make_clip(i, 30.0, sr=fs, kind=["vowel", "mix", "tone", "noise"][i % 4]) for i in range(8)])
    x = torch.from_numpy(base).cuda().repeat(125, 1).contiguous()
    rs = audio_toolkit.FrameResampler(fs)
    rs.process(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): y = rs.process(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"fs_in": fs, "streams": 1000, "ms": ms, "GBps_alg": 1000 * (fs * 30 * 4 + 1.92e6) / ms / 1e6}))
