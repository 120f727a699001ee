"""One bench-shaped batch for ncu: SB_BENCH_ARCH (default large-v3-turbo), SB_BENCH_CLIPS (default 64) synthetic 30 s clips through
sb_transcribe_batch, one warm-up batch first.  Prints `LAUNCHES warm=<n> batch=<n>` so that an ncu launch list can skip the
warm-up (`-s`) and capture exactly one batch (`-c`).  Usage: python tools/profile_step.py [n_batches=1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from spittle_b200 import capi, synth

arch = os.environ.get("SB_BENCH_ARCH", "large-v3-turbo")
n_clips = int(os.environ.get("SB_BENCH_CLIPS", "64"))
n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 1
path = synth.ensure_model_file(arch, os.environ.get("SB_MODEL_DIR", "/tmp/spittle_b200_models"))
clips = np.stack([synth.make_clip(i, 30.0) for i in range(n_clips)])
dev = torch.from_numpy(clips).cuda()
ptrs = [dev.data_ptr() + i * clips.shape[1] * 4 for i in range(n_clips)]
sizes = [clips.shape[1]] * n_clips
l0 = capi.launch_count()
eng = capi.Engine(path, max_batch=n_clips, dtype=capi.SB_DTYPE_F16)
params = capi.default_params()
eng.transcribe_batch_ptrs(ptrs, sizes, params)
l1 = capi.launch_count()
for _ in range(n_batches):
    res = eng.transcribe_batch_ptrs(ptrs, sizes, params)
l2 = capi.launch_count()
st = eng.stats()
print(f"LAUNCHES warm={l1 - l0} batch={(l2 - l1) // n_batches} windows={st['windows']} decoder_steps={st['decoder_steps']}")
eng.close()
