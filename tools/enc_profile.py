"""Encoder-only timing: sb_encode on W mel windows (host API), CUDA-event-free wall timing + engine stats.
Usage: python tools/enc_profile.py [arch] [W] [dtype] [reps]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spittle_b200 import capi, synth
arch = sys.argv[1] if len(sys.argv) > 1 else "small"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dtype = capi.SB_DTYPE_F16 if (len(sys.argv) <= 3 or sys.argv[3] == "f16") else capi.SB_DTYPE_BF16
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
path = synth.ensure_model_file(arch, os.environ.get("SB_MODEL_DIR", "/tmp/spittle_b200_models"))
eng = capi.Engine(path, max_batch=W, dtype=dtype)
rng = np.random.default_rng(0)
mel = rng.uniform(-1, 1, (W, eng.info.n_mels, 3000)).astype(np.float32)
eng.set_profile(True)
for r in range(reps):
    eng.stats(reset=True)
    t0 = time.perf_counter()
    out = eng.encode(mel)
    dt = time.perf_counter() - t0
    st = eng.stats(reset=True)
    d, L = eng.info.n_audio_state, eng.info.n_audio_layer
    flops = W * (2 * 3000 * d * eng.info.n_mels * 3 + 2 * 1500 * d * d * 3 + L * (24 * 1500 * d * d + 4 * 1500 * 1500 * d))
    print(json.dumps({"arch": arch, "W": W, "wall_ms": dt * 1e3, "gemm_ms": st["gemm_ms"], "gemm_tflops": st["gemm_flops"] / max(st["gemm_ms"], 1e-9) / 1e9,
                      "attn_ms": st["attn_ms"], "attn_tflops": st["attn_flops"] / max(st["attn_ms"], 1e-9) / 1e9,
                      "enc_flops_T": flops / 1e12,
                      # device time of conv stem + blocks + ln_post + the cross-KV GEMM (which is not in enc_flops_T):
                      "encode_ms": st["encode_ms"],
                      "encoder_tflops_lower_bound": flops / max(st["encode_ms"], 1e-9) / 1e9}), flush=True)
