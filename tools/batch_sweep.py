"""RTFx of sb_transcribe_batch against the number of clips per call and the number of decode slots (sb_config.max_batch):
with more clips than slots the scheduler refills finished slots with other clips' windows, so the tail windows of some clips
overlap the first windows of others.  Usage: python tools/batch_sweep.py [arch] "clips:slots,clips:slots,..." [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from spittle_b200 import capi, synth

arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3-turbo"
combos = [tuple(int(v) for v in c.split(":")) for c in (sys.argv[2] if len(sys.argv) > 2 else "64:64,128:64,128:128,256:128").split(",")]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
path = synth.ensure_model_file(arch, os.environ.get("SB_MODEL_DIR", "/tmp/spittle_b200_models"))
max_clips = max(c for c, _ in combos)
clips = np.stack([synth.make_clip(i, 30.0) for i in range(max_clips)])
dev = torch.from_numpy(clips).cuda()
n = clips.shape[1]
for n_clips, slots in combos:
    eng = capi.Engine(path, max_batch=slots, dtype=capi.SB_DTYPE_F16)
    ptrs = [dev.data_ptr() + i * n * 4 for i in range(n_clips)]
    sizes = [n] * n_clips
    params = capi.default_params()
    for _ in range(2):
        eng.transcribe_batch_ptrs(ptrs, sizes, params)
    eng.stats(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.transcribe_batch_ptrs(ptrs, sizes, params)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    st = eng.stats(reset=True)
    print(json.dumps({"arch": arch, "clips": n_clips, "slots": slots, "ms_per_call": dt * 1e3, "rtfx": n_clips * 30.0 / dt,
                      "encode_ms": st["encode_ms"] / steps, "decode_ms": st["decode_ms"] / steps, "decoder_steps": st["decoder_steps"] / steps,
                      "encoder_batches": st["rounds"] / steps, "tokens": st["tokens_sampled"] / steps, "prefill_rows": st["prefill_rows"] / steps}), flush=True)
    eng.close()
