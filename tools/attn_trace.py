"""Per-CTA phase timeline of k_attn_enc_ts (SB_ATTN_TRACE=1): start, Q staged, pass 1 done, pass 2 done, exit."""
import ctypes as C, os, sys
os.environ["SB_ATTN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spittle_b200 import capi, synth
arch = sys.argv[1] if len(sys.argv) > 1 else "small"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 16
path = synth.ensure_model_file(arch, os.environ.get("SB_MODEL_DIR", "/tmp/spittle_b200_models"))
eng = capi.Engine(path, max_batch=W, dtype=capi.SB_DTYPE_F16)
mel = np.random.default_rng(0).uniform(-1, 1, (W, eng.info.n_mels, 3000)).astype(np.float32)
eng.encode(mel); eng.encode(mel)
buf = (C.c_ulonglong * (4096 * 6))()
lib = capi.lib(); lib.sb_debug_attn_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.sb_debug_attn_trace(buf, 4096 * 6) == 0
t = np.array(buf[:], dtype=np.int64).reshape(4096, 6)
n = min(4096, 12 * eng.info.n_audio_head * W)
t = t[:n]
t0 = t[:, 0].min()
d = np.diff(t[:, :5], axis=1) / 1e3
print("CTAs", n, "kernel span us", (t[:, 4].max() - t0) / 1e3)
for k, name in enumerate(["prologue+Q", "pass 1", "pass 2", "epilogue"]):
    print(f"{name:12s} mean {d[:, k].mean():7.2f} us  p10 {np.percentile(d[:, k], 10):7.2f}  p90 {np.percentile(d[:, k], 90):7.2f}")
print("lifetime mean", (t[:, 4] - t[:, 0]).mean() / 1e3)
# overlap on one SM: for SM of CTA 0, list (start, p1 end, p2 end) of its CTAs
sm = t[0, 5]
rows = t[t[:, 5] == sm]
rows = rows[np.argsort(rows[:, 0])][:8]
for r in rows: print("SM", sm, [round((x - t0) / 1e3, 1) for x in r[:5]])
