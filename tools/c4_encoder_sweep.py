"""C4 (BASELINE.json configs[3]): Whisper Large-v3 encoder-dominated throughput sweep, B = 32 ... 512 windows per GPU.

One sb_encode call per batch size (conv stem + 32 blocks + ln_post + the cross-KV GEMM of all 32 decoder layers), device
time from the engine's CUDA events; achieved TFLOP/s = B x 2.2738e12 / t for the encoder proper and with the
cross-KV FLOPs (B x 3.146e11) added, as fractions of the measured sustained / burst bf16 peak.
Usage: python tools/c4_encoder_sweep.py [arch=large-v3] [dtype=f16] [sizes=32,64,128,256,512] > profiles/raw/r2_c4_sweep.jsonl"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from spittle_b200 import capi, synth

arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
dtype = capi.SB_DTYPE_BF16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else capi.SB_DTYPE_F16
sizes = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "32,64,128,256,512").split(",")]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) \
    else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
path = synth.ensure_model_file(arch, os.environ.get("SB_MODEL_DIR", "/tmp/spittle_b200_models"))
rng = np.random.default_rng(0)
for B in sizes:
    eng = capi.Engine(path, max_batch=B, dtype=dtype)
    d, L, Ld, nm = eng.info.n_audio_state, eng.info.n_audio_layer, eng.info.n_text_layer, eng.info.n_mels
    mel = rng.uniform(-1, 1, (B, nm, 3000)).astype(np.float32)
    enc_flops = B * (2 * 3000 * d * nm * 3 + 2 * 1500 * d * d * 3 + L * (24 * 1500 * d * d + 4 * 1500 * 1500 * d))
    ckv_flops = B * Ld * 4 * 1500 * d * d
    eng.set_profile(True)
    best = None
    for rep in range(3):            # rep 0 allocates the workspaces
        eng.stats(reset=True)
        t0 = time.perf_counter()
        eng.encode(mel)
        wall = time.perf_counter() - t0
        st = eng.stats(reset=True)
        if rep and (best is None or st["encode_ms"] < best["encode_ms"]):
            best = dict(st, wall_ms=wall * 1e3)
    ms = best["encode_ms"]
    row = {"workload": f"C4: {arch} encoder, {B} windows (30 s each) in one batch, {'bf16' if dtype == capi.SB_DTYPE_BF16 else 'f16'}",
           "B": B, "encode_ms": ms, "windows_per_s": B / ms * 1e3, "rtfx_encoder_only": B * 30.0 / ms * 1e3,
           "encoder_tflops": enc_flops / ms / 1e9, "with_cross_kv_tflops": (enc_flops + ckv_flops) / ms / 1e9,
           "frac_sustained": (enc_flops + ckv_flops) / ms / 1e9 / peaks["bf16_tflops_sustained"],
           "frac_burst": (enc_flops + ckv_flops) / ms / 1e9 / peaks["bf16_tflops"],
           "gemm_tflops": best["gemm_flops"] / max(best["gemm_ms"], 1e-9) / 1e9, "gemm_ms": best["gemm_ms"],
           "attn_tflops": best["attn_flops"] / max(best["attn_ms"], 1e-9) / 1e9, "attn_ms": best["attn_ms"],
           "other_ms": ms - best["gemm_ms"] - best["attn_ms"]}
    print(json.dumps(row), flush=True)
    eng.close()
    del mel
