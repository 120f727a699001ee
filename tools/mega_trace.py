"""Per-stage timeline of the decoder-step megakernel (CTA 0), SB_MEGA_TRACE=1.
Usage: python tools/mega_trace.py [arch] [clips]"""
import ctypes as C, os, sys, json
os.environ["SB_MEGA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spittle_b200 import capi, synth
arch = sys.argv[1] if len(sys.argv) > 1 else "small"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
path = synth.ensure_model_file(arch, os.environ.get("SB_MODEL_DIR", "/tmp/spittle_b200_models"))
eng = capi.Engine(path, max_batch=n, dtype=capi.SB_DTYPE_F16)
clips = [synth.make_clip(i, 30.0) for i in range(n)]
params = capi.default_params(n_max_tokens=int(os.environ.get("SB_TRACE_STEPS", "40")), max_windows=1)
eng.transcribe_batch(clips, params)
eng.transcribe_batch(clips, params)
L = eng.info.n_text_layer
buf = (C.c_ulonglong * 1024)()
lib = capi.lib()
lib.sb_debug_mega_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.sb_debug_mega_trace(buf, 1024) == 0
t = np.array(buf[: 2 * (11 * L + 1)], dtype=np.int64).reshape(-1, 2)[: 11 * L + 1]   # [stage] = (wait done, work done)
names = ["LN1", "QKV", "SA", "O", "LN2", "CQ", "XA", "CO", "LN3", "FC1", "FC2"]
work = (t[:, 1] - t[:, 0]) / 1e3
gap = np.zeros(len(t)); gap[1:] = (t[1:, 0] - t[:-1, 1]) / 1e3       # own work done -> next stage's dependency resolved
print("total us", (t[-1, 1] - t[0, 0]) / 1e3)
agg = {}
for s in range(11 * L):
    k = names[s % 11]
    a = agg.setdefault(k, [0.0, 0.0, 0])
    a[0] += work[s]; a[1] += gap[s]; a[2] += 1
for k, a in agg.items():
    print(f"{k:4s} n={a[2]:3d} work(cta0) {a[0] / a[2]:7.2f} us  wait-before {a[1] / a[2]:7.2f} us")
print("layer 1 detail (wait-before, work):", [(round(gap[s], 2), round(work[s], 2)) for s in range(11, 22)])
