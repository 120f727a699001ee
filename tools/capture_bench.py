"""Timing of the capture-side format kernels with CUDA events; prints one JSON line per kernel.
Algorithmic bytes: f32 -> i16: 6 B per sample; visualiser: 512 x 4 B read + 64 B written per chunk."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spittle_b200 import capi
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


n = 1000 * 480000                                   # 1000 recordings of 30 s at 16 kHz: 1.92 GB in, 0.96 GB out (>> L2)
x = torch.empty(n, device=dev).uniform_(-1, 1)
o = torch.empty(n, dtype=torch.int16, device=dev)
ms = timeit(lambda: capi.pcm_f32_to_i16_dev(x.data_ptr(), o.data_ptr(), n, st))
print(json.dumps({"kernel": "k_pcm_f32_to_i16", "samples": n, "ms": ms, "GB/s": 6.0 * n / ms / 1e6, "frac_of_hbm": 6.0 * n / ms / 1e6 / peak}))
del x, o
streams, chunk, n_chunks = 1000, 1024, 1406         # 1000 capture streams of 30 s at 48 kHz in 1024-sample chunks
p = torch.empty((streams, chunk * n_chunks), device=dev).uniform_(-0.3, 0.3)
out = torch.empty((streams, n_chunks, 16), device=dev)
ms = timeit(lambda: capi.visualiser_levels_dev(p.data_ptr(), p.stride(0), streams, n_chunks, chunk, 48000, out.data_ptr(), st))
b = streams * n_chunks * (512 * 4 + 64)
print(json.dumps({"kernel": "k_visualiser_levels", "chunks": streams * n_chunks, "ms": ms, "GB/s": b / ms / 1e6, "frac_of_hbm": b / ms / 1e6 / peak,
                  "note": "reads the first 2 KB of every 4 KB chunk"}))
