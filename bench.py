#!/usr/bin/env python
"""bench.py -- headline benchmark of the Whisper transcription hot path on B200.

Metric (BASELINE.json): RTFx = audio-seconds per wall-second.  Workload: BASELINE.json configs[2] "Whisper
Large-v3 Turbo (128-mel, 4-layer decoder) batched clips sharded across 1/2/4/8 B200" -- the model north_star names --
with 128 synthetic 30 s clips per GPU; SB_BENCH_ARCH=small runs configs[1] (Whisper Small, 64 clips), and at N = 1 the
default run carries that configuration as the secondary `config.small_64` entry, plus `config.c4_encoder` (configs[3] in small:
the encoder + cross-KV GEMM on 128 windows) and `config.c5_frontend` (configs[4] in small: the 48 kHz capture front-end over
1000 streams) so that the driver's BENCH line carries all five configurations' figures.
A step = one pass of the whole hot path (log-mel -> encoder -> cross-KV -> greedy decode with
the whisper.cpp seek loop -> text) over one batch of clips through the C ABI
(sb_transcribe_batch).  N > 1: one process per GPU (torchrun), every rank transcribes its own
batch (weak scaling, no collective on the data path -- clips do not interact).

  value  RTFx with the PCM already resident in HBM (device pointers handed to the C ABI)
  e2e    RTFx with pinned HOST buffers: H2D of the PCM and D2H of tokens inside the timed region
  roofline      the kernel with the largest share of the step (the tcgen05 GEMM on Large-v3-Turbo, the decoder-step
                projections on Small): algorithmic bytes or FLOPs per launch / its CUDA-event duration, measured
                live; the other hot kernels follow in roofline_extra
  cpu_baseline  oracle/cpu_ref (C++ restatement of the whisper.cpp CPU path, std::thread, AVX2 / AVX-512) on the host cores,
                all threads and whisper.cpp's default 4, bounded sample
  parity        tokens of the timed GPU run against that CPU restatement on the first clips of the batch
  --impl reference   times the same CPU restatement as the reference arm (the reference itself cannot be
                     built here: no Rust toolchain, crates not vendored -- DESIGN.md)

Timing: W >= 3 warm-up steps; inputs (128 x 1.92 MB PCM plus ~6 GB of activations and 3.9 GB of
cross-KV per step) are far larger than the 126 MB L2, so no explicit flush is needed; device
timing with CUDA events bracketed by barrier + synchronize, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

ARCH = os.environ.get("SB_BENCH_ARCH", "large-v3-turbo")
# clips per GPU: configs[1] fixes 64 for Whisper Small; configs[2] (Turbo) leaves it open -- 128 clips on 128 decode slots is
# where one B200 saturates (profiles/r2_batch_sweep.md: 64 -> 3217, 128 -> 3555 RTFx)
CLIPS_PER_GPU = int(os.environ.get("SB_BENCH_CLIPS", "64" if ARCH == "small" else "128"))
CLIP_SECONDS = 30.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        pw = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(self.samples)}


def make_clips(rank: int, n: int):
    from spittle_b200 import synth
    return [synth.make_clip(rank * n + i, seconds=CLIP_SECONDS) for i in range(n)]


def model_path(arch: str) -> str:
    from spittle_b200 import synth
    d = os.environ.get("SB_MODEL_DIR", "/tmp/spittle_b200_models")
    return synth.ensure_model_file(arch, d)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_ref_run(arch: str, clip_ids, n_threads: int):
    """oracle/cpu_ref (C++ restatement of the whisper.cpp CPU path, std::thread) over the given clips of the bench batch.
    Returns (rtfx, seconds, per-clip results)."""
    from oracle import cpu_ref
    ref = cpu_ref.CpuRef(model_path(arch), n_threads=n_threads)
    from spittle_b200 import synth
    out = []
    t0 = time.perf_counter()
    for i in clip_ids:
        out.append(ref.full(synth.make_clip(i, seconds=CLIP_SECONDS)))
    dt = time.perf_counter() - t0
    isa = ref.isa
    ref.close()
    return (len(clip_ids) * CLIP_SECONDS) / dt, dt, out, isa


def cpu_baseline_block(arch: str, budget_s: float = 20.0):
    """Bounded CPU sample: clips 0, 1, ... of the bench batch with all host threads until ~budget_s of CPU wall time,
    then clip 0 again at 4 threads (whisper.cpp's default n_threads = min(4, hardware_concurrency))."""
    cores = host_threads()
    results, secs, n = [], 0.0, 0
    isa = 0
    while secs < budget_s and n < 8:
        v, dt, out, isa = cpu_ref_run(arch, [n], cores)
        results += out
        secs += dt
        n += 1
    v_all = n * CLIP_SECONDS / secs
    v4, dt4, _, _ = cpu_ref_run(arch, [0], min(4, cores))
    n_tok = sum(len(w["tokens"]) for r in results for w in r["windows"])
    n_win = sum(len(r["windows"]) for r in results)
    block = {"value": v_all, "unit": "x real-time", "cores": cores, "kind": "port",
             "threads4_value": v4,
             "sample": f"clips 0..{n - 1} of the same batch ({n} x 30 s, {n_win} windows, {n_tok} decoded tokens, {secs:.1f} s CPU wall "
                       f"at {cores} threads; clip 0 again at {min(4, cores)} threads = whisper.cpp's default: {dt4:.1f} s); oracle/cpu_ref = C++ "
                       f"restatement of the whisper.cpp CPU path (f16 weights, f16-rounded activations, AVX{isa}), validated against "
                       "the numpy oracle; the reference itself cannot be built here (no cargo/rustc, crates not vendored)"}
    return block, results


def parity_block(gpu_results, cpu_results):
    """GPU tokens of clips 0..n-1 (through sb_transcribe_batch) against the CPU restatement's tokens, window by window."""
    n_win = n_exact = n_tok = 0
    first_div = None
    for ci, (g, c) in enumerate(zip(gpu_results, cpu_results)):
        for wi, cw in enumerate(c["windows"]):
            n_win += 1
            if wi >= len(g.windows):
                first_div = first_div or {"clip": ci, "window": wi, "step": 0, "cpu_margin": None, "note": "GPU produced fewer windows"}
                break
            gw = g.windows[wi]
            got = g.sampled[gw["token_offset"]: gw["token_offset"] + gw["n_tokens"]]
            if got == cw["tokens"]:
                n_exact += 1
                n_tok += len(got)
                continue
            k = next((i for i in range(min(len(got), len(cw["tokens"]))) if got[i] != cw["tokens"][i]), min(len(got), len(cw["tokens"])))
            n_tok += k
            if first_div is None:
                first_div = {"clip": ci, "window": wi, "step": k,
                             "cpu_margin": cw["margins"][k] if k < len(cw["margins"]) else None,
                             "gpu_margin": g.margins[gw["token_offset"] + k] if gw["token_offset"] + k < len(g.margins) else None}
            break          # later windows of this clip depend on the diverged text context
    return {"checker": "oracle/cpu_ref", "clips": len(cpu_results), "windows": n_win, "exact_windows": n_exact,
            "tokens_identical": n_tok, "first_divergence": first_div,
            "note": "a divergence is only legitimate at a top-1/top-2 margin below the logit tolerance (tests/test_parity_sizes_gpu.py)"}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle/cpu_ref, the C++ restatement of whisper.cpp's CPU implementation;
    the reference itself cannot be built here) on all host threads, each step = clip 0 of the bench batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_threads()
    # one probe step sizes the run: the whole --steps / --warmup run must end within a few minutes
    v, dt, out, isa = cpu_ref_run(ARCH, [0], cores)
    budget = 150.0
    warm = 0 if dt > 20 else min(max(args.warmup, 0), 2)
    steps = max(1, min(args.steps, int(max(1.0, (budget - dt * (1 + warm)) / max(dt, 1e-3)))))
    for _ in range(max(0, warm - 1)):
        cpu_ref_run(ARCH, [0], cores)
    vals = []
    for _ in range(steps):
        v, dt, out, isa = cpu_ref_run(ARCH, [0], cores)
        vals.append((v, dt))
    value = CLIP_SECONDS * len(vals) / sum(d for _, d in vals)
    ms = float(np.mean([d for _, d in vals])) * 1e3
    n_tok = sum(len(w["tokens"]) for w in out[0]["windows"])
    line = {
        "impl": "reference", "metric": "RTFx (audio-s per wall-s)", "value": value, "unit": "x real-time",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16 (ggml rounding points) / f32 accumulate", "data": "synthetic",
        "config": workload_config(ARCH, extra={"sample": "each step = clip 0 of that batch (1 x 30 s), CPU only, "
                                                         f"{len(out[0]['windows'])} windows, {n_tok} decoded tokens",
                                               "threads": cores, "isa": f"AVX{isa}"}),
        "cpu_baseline": {"value": value, "unit": "x real-time", "cores": cores, "kind": "port",
                         "sample": f"1 clip x 30 s per step, {cores} std::threads; oracle/cpu_ref = C++ restatement of the whisper.cpp CPU "
                                   "path, validated against the numpy oracle (reference not buildable here: no cargo/rustc, crates not vendored)"},
        "e2e": {"value": value, "unit": "x real-time", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(arch: str, extra=None, n_clips=None):
    which = {"large-v3-turbo": "configs[2]", "small": "configs[1]", "large-v3": "configs[3] model"}.get(arch, "test architecture")
    n_clips = n_clips or CLIPS_PER_GPU
    cfg = {"workload": f"Whisper {arch} greedy decode, batch of {n_clips} synthetic 30 s 16 kHz clips per GPU (BASELINE.json {which}), "
                       "random-init 'sharp' recipe seed 42, language en, timestamps on, text context carried between the windows of a "
                       "clip like whisper_full, no temperature fallback",
           "arch": arch, "clips_per_gpu": n_clips}
    if extra:
        cfg.update(extra)
    return cfg


def measure(arch, args, torch, capi, dist, rank, local_rank, world, steps, warmup, trace=True, n_clips=None):
    """Load `arch`, run warm-up + the two timed regions (device-resident PCM, pinned host PCM) + one traced batch."""
    if rank == 0:
        model_path(arch)                  # rank 0 writes the synthetic model file once; other ranks wait for it
    if dist:
        dist.barrier()
    path = model_path(arch)
    dtype = capi.SB_DTYPE_F16 if args.dtype == "f16" else capi.SB_DTYPE_BF16
    n_clips = n_clips or CLIPS_PER_GPU
    eng = capi.Engine(path, device=local_rank, max_batch=n_clips, dtype=dtype)
    clips = make_clips(rank, n_clips)
    n = clips[0].shape[0]
    host = torch.from_numpy(np.stack(clips)).pin_memory()
    devbuf = host.to(torch.device("cuda", local_rank))
    params = capi.default_params()
    host_ptrs = [host.data_ptr() + i * n * 4 for i in range(n_clips)]
    dev_ptrs = [devbuf.data_ptr() + i * n * 4 for i in range(n_clips)]
    sizes = [n] * n_clips
    eng_stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(ptrs, k, profile):
        eng.set_profile(profile)
        eng.stats(reset=True)
        l0 = capi.launch_count()
        sync_all()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(eng_stream)          # CUDA events on the stream the engine launches on
        res = None
        for _ in range(k):
            res = eng.transcribe_batch_ptrs(ptrs, sizes, params)     # returns host text: device work is complete
        e1.record(eng_stream)
        torch.cuda.synchronize()
        ms_dev = e0.elapsed_time(e1)
        wall = (time.perf_counter() - t0) * 1e3
        ms_rank = max(ms_dev, wall)    # the call is synchronous: wall additionally covers host bookkeeping
        ms = ms_rank
        if dist:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ms_rank, res, eng.stats(reset=True), capi.launch_count() - l0

    for _ in range(warmup):
        eng.transcribe_batch_ptrs(dev_ptrs, sizes, params)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, ms_rank, res, st_dev, launches = timed(dev_ptrs, steps, profile=True)
    ms_e2e, ms_e2e_rank, res2, st_e2e, _ = timed(host_ptrs, steps, profile=False)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    st_dec = None
    if trace:
        # one more (untimed) batch with the device-side launch trace on (sb_engine_set_profile(e, 2)): every decoder-stage
        # launch stamps %globaltimer at its first block's start and its last block's end.  The step keeps its CUDA graph,
        # its lanes and its PDL overlap, so these are the durations inside the real chain.
        eng.set_profile(2)
        eng.stats(reset=True)
        eng.transcribe_batch_ptrs(dev_ptrs, sizes, params)
        st_dec = eng.stats(reset=True)
        eng.set_profile(0)
    n_mels = eng.info.n_mels
    eng.close()
    del devbuf
    torch.cuda.empty_cache()
    return dict(ms_dev=ms_dev, ms_rank=ms_rank, ms_e2e=ms_e2e, ms_e2e_rank=ms_e2e_rank, res=res, st_dev=st_dev, st_e2e=st_e2e,
                st_dec=st_dec, launches=launches, clocks=sampler.summary(), n_mels=n_mels)


def c4_encoder_block(capi, arch, dtype_name, B=128):
    """BASELINE.json configs[3] in small: the encoder (+ cross-KV GEMM) of `arch` on B windows in one sb_encode call, device time
    from the engine's events, as TFLOP/s and as a fraction of the measured sustained bf16 peak.  (Large-v3-Turbo has Large-v3's
    encoder; tools/c4_encoder_sweep.py is the full B = 32 ... 512 sweep on Large-v3 itself.)"""
    try:
        peaks, _src = load_peaks()
        dtype = capi.SB_DTYPE_F16 if dtype_name == "f16" else capi.SB_DTYPE_BF16
        eng = capi.Engine(model_path(arch), max_batch=B, dtype=dtype)
        d, L, Ld, nm = eng.info.n_audio_state, eng.info.n_audio_layer, eng.info.n_text_layer, eng.info.n_mels
        mel = np.random.default_rng(0).uniform(-1, 1, (B, nm, 3000)).astype(np.float32)
        flops = B * (2 * 3000 * d * nm * 3 + 2 * 1500 * d * d * 3 + L * (24 * 1500 * d * d + 4 * 1500 * 1500 * d)) + B * Ld * 4 * 1500 * d * d
        eng.set_profile(True)
        best = None
        for rep in range(3):
            eng.stats(reset=True)
            eng.encode(mel)
            st = eng.stats(reset=True)
            if rep and (best is None or st["encode_ms"] < best):
                best = st["encode_ms"]
        eng.close()
        tf = flops / best / 1e9
        return {"workload": f"C4: {arch} encoder + cross-KV GEMM, {B} windows (30 s each) in one batch, {dtype_name}", "encode_ms": best,
                "tflops": tf, "frac_sustained_bf16_peak": tf / peaks["bf16_tflops_sustained"], "encoder_only_rtfx": B * 30.0 / best * 1e3}
    except Exception as e:          # a secondary entry must never cost the line
        return {"error": repr(e)}


def c5_frontend_block(n_streams=1000):
    """BASELINE.json configs[4] in small: resample -> Silero -> log-mel over `n_streams` synthetic 48 kHz streams
    (tools/frontend_bench.py is the same code at the stated 10 000 streams)."""
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("frontend_bench", os.path.join(ROOT, "tools", "frontend_bench.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.run(n_streams, n_streams)
    except Exception as e:
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default=os.environ.get("SB_BENCH_DTYPE", "f16"), choices=["f16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary entries (Whisper Small configs[1], C4 encoder, C5 front-end)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from spittle_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: spittle_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    warmup = max(3, args.warmup)
    steps = args.steps
    peaks, peak_src = load_peaks()
    audio_s = CLIPS_PER_GPU * CLIP_SECONDS

    m = measure(ARCH, args, torch, capi, dist, rank, local_rank, world, steps, warmup)
    ms_dev, ms_e2e, st_dev, st_e2e, st_dec, launches = m["ms_dev"], m["ms_e2e"], m["st_dev"], m["st_e2e"], m["st_dec"], m["launches"]

    # per-rank breakdown (where does the N > 1 loss come from: device phases or the host side of a rank?)
    mine = [m["ms_rank"] / steps, m["ms_e2e_rank"] / steps, st_dev["mel_ms"] / steps, st_dev["encode_ms"] / steps,
            st_dev["decode_ms"] / steps, st_dev["decoder_steps"] / steps]
    per_rank = [mine]
    if dist:
        t = torch.tensor(mine, device="cuda", dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [[float(v) for v in x.tolist()] for x in allt]

    value = world * audio_s * steps / (ms_dev / 1e3)
    e2e = world * audio_s * steps / (ms_e2e / 1e3)
    gemm_tflops = st_dev["gemm_flops"] / max(st_dev["gemm_ms"], 1e-9) / 1e9
    attn_tflops = st_dev["attn_flops"] / max(st_dev["attn_ms"], 1e-9) / 1e9
    peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    peak_hbm = peaks["hbm_gbs"]
    skinny_gbps = st_dec["skinny_bytes"] / max(st_dec["skinny_ms"], 1e-9) / 1e6
    xattn_gbps = st_dec["xattn_bytes"] / max(st_dec["xattn_ms"], 1e-9) / 1e6
    # log-mel: algorithmic bytes (PCM in + n_mel x 3000 f32 out per 30 s clip) over the engine's mel time
    mel_bytes = CLIPS_PER_GPU * (480000 * 4 + m["n_mels"] * 3000 * 4)
    mel_gbps = mel_bytes * steps / max(st_dev["mel_ms"], 1e-9) / 1e6

    secondary = c4 = c5 = None
    if world == 1 and ARCH != "small" and not args.no_secondary:
        # BASELINE.json configs[1] (Whisper Small, 64 clips, 1 GPU) carried as a secondary entry of the default run
        m2 = measure("small", args, torch, capi, None, rank, local_rank, 1, min(steps, 5), 3, trace=False, n_clips=64)
        k2 = min(steps, 5)
        secondary = {"workload": workload_config("small", n_clips=64)["workload"], "value": 64 * CLIP_SECONDS * k2 / (m2["ms_dev"] / 1e3),
                     "e2e": 64 * CLIP_SECONDS * k2 / (m2["ms_e2e"] / 1e3), "unit": "x real-time", "steps": k2, "ms_per_step": m2["ms_dev"] / k2,
                     "ms_encode": m2["st_dev"]["encode_ms"] / k2, "ms_decode": m2["st_dev"]["decode_ms"] / k2,
                     "decoder_steps_per_step": m2["st_dev"]["decoder_steps"] / k2}
        c4 = c4_encoder_block(capi, ARCH, args.dtype)
        c5 = c5_frontend_block(1000)
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    cpu_base = parity = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base, cpu_results = cpu_baseline_block(ARCH)
        parity = parity_block(m["res"], cpu_results)
    line = {
        "metric": "RTFx (audio-s per wall-s)", "value": value, "unit": "x real-time", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_dev / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(ARCH, extra={
            "l2": "inputs+activations per step >> 126 MB L2; no explicit flush",
            "windows_per_step": st_dev["windows"] / steps, "decoder_steps_per_step": st_dev["decoder_steps"] / steps,
            "tokens_per_step": st_dev["tokens_sampled"] / steps, "encoder_batches_per_step": st_dev["rounds"] / steps,
            "ms_mel": st_dev["mel_ms"] / steps, "ms_encode": st_dev["encode_ms"] / steps,
            "ms_decode": st_dev["decode_ms"] / steps,
            "phase_note": "ms_encode = sum of the encoder batches (refill batches overlap the decode lanes); ms_decode = the rest of the call",
            "small_64": secondary, "c4_encoder": c4, "c5_frontend": c5}),
        "e2e": {"value": e2e, "unit": "x real-time", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": (st_e2e["pcm_bytes"] + st_e2e["h2d_bytes"]) / steps,
                "d2h_bytes_per_step": st_e2e["d2h_bytes"] / steps},
        "gpu_launches": int(launches),
        "per_rank": [dict(zip(("ms_per_step", "ms_per_step_e2e", "ms_mel", "ms_encode", "ms_decode", "decoder_steps"), r)) for r in per_rank],
        "roofline": None,
        "roofline_extra": None,
        "clocks": m["clocks"],
    }
    # ---- rooflines: per-kernel entries, the one with the largest share of the timed step first ----
    skinny_avg_us = 1e3 * st_dec["skinny_ms"] / max(st_dec["skinny_launches"], 1)
    xattn_avg_us = 1e3 * st_dec["xattn_ms"] / max(st_dec["xattn_launches"], 1)
    ln_avg_us = 1e3 * st_dec["dln_ms"] / max(st_dec["dln_launches"], 1)
    self_avg_us = 1e3 * st_dec["dself_ms"] / max(st_dec["dself_launches"], 1)
    ms_step = ms_dev / steps
    # share of the timed step: the decode phase of the step is split between the traced kernel classes in proportion to
    # their summed launch durations in the traced batch (the lanes overlap, so the sum exceeds the wall time; the logits
    # GEMM and the sampler, ~5 % of the decode kernel time, are not traced)
    t_all = max(st_dec["skinny_ms"] + st_dec["xattn_ms"] + st_dec["dln_ms"] + st_dec["dself_ms"], 1e-9)
    dec_share = (st_dev["decode_ms"] / steps) / ms_step
    lane_step_us = 1e3 * st_dec["dstep_ms"] / max(st_dec["dstep_count"], 1)
    entries = [
        {"kernel": "k_skinny_gemm (decoder-step projections, weight streaming, mma.sync)", "bound": "hbm",
         "achieved": skinny_gbps, "peak": peak_hbm, "unit": "GB/s", "frac": skinny_gbps / peak_hbm,
         "alg_bytes_per_launch": st_dec["skinny_bytes"] / max(st_dec["skinny_launches"], 1),
         "avg_launch_us": skinny_avg_us, "launches_per_step": st_dec["skinny_launches"],
         "share_of_step": dec_share * st_dec["skinny_ms"] / t_all,
         "note": "latency-bound launch (1-13 MB of weights) inside the dependent chain of a lane-step "
                 f"(traced lane-step {lane_step_us:.0f} us) that overlaps the other lanes' cross-attention streams; duration = first block "
                 "start -> last block end (device trace); the decode PHASE runs within ~1.5x of its HBM floor (DESIGN.md 4.3); "
                 "ncu (profiles/r1_full_skinny_gemm.md, 768x768 launch): 1.31 MB DRAM read for 1.18 MB of weights",
         "peak_source": f"{peak_src} hbm_gbs", "traffic": 1.31e6 / 1.18e6 * st_dec["skinny_bytes"] / max(st_dec["skinny_launches"], 1)},
        {"kernel": "k_dec_cross_attn (decoder cross-attention over the cached 1500 encoder keys, register streaming, "
                   "mma.sync scores)", "bound": "hbm",
         "achieved": xattn_gbps, "peak": peak_hbm, "unit": "GB/s", "frac": xattn_gbps / peak_hbm,
         "alg_bytes_per_launch": st_dec["xattn_bytes"] / max(st_dec["xattn_launches"], 1),
         "avg_launch_us": xattn_avg_us, "launches_per_step": st_dec["xattn_launches"],
         "share_of_step": dec_share * st_dec["xattn_ms"] / t_all,
         "note": "bytes = K and V of the sequences still decoding (finished ones are skipped); two lanes stream "
                 "concurrently and share HBM; alone at 64 live sequences 49.6 us = 5.95 TB/s "
                 "(profiles/r1_cross_attn_stream.md)", "peak_source": f"{peak_src} hbm_gbs",
         "traffic": 290.4e6 / 294.9e6 * st_dec["xattn_bytes"] / max(st_dec["xattn_launches"], 1)},   # ncu: profiles/r1_full_dec_cross_attn.md
        {"kernel": "k_dec_ln + k_dec_self_attn (decoder LayerNorm / self-attention over the <= 448-token cache)",
         "bound": "hbm", "achieved": None, "peak": peak_hbm, "unit": "GB/s", "frac": None,
         "avg_launch_us": {"ln": ln_avg_us, "self_attn": self_avg_us},
         "launches_per_step": st_dec["dln_launches"] + st_dec["dself_launches"],
         "share_of_step": dec_share * (st_dec["dln_ms"] + st_dec["dself_ms"]) / t_all,
         "note": "latency-bound stages of the chain", "traffic": None},
        {"kernel": "k_gemm_tn (tcgen05 / TMEM / TMA, CTA pairs, 16 epilogue warps: encoder + cross-KV projections)", "bound": "tensor",
         "achieved": gemm_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tflops / peak_tf,
         "alg_flops_per_launch": st_dev["gemm_flops"] / max(st_dev["gemm_launches"], 1),
         "avg_launch_us": 1e3 * st_dev["gemm_ms"] / max(st_dev["gemm_launches"], 1),
         "launches_per_step": st_dev["gemm_launches"] / steps, "share_of_step": st_dev["gemm_ms"] / steps / ms_step,
         "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step); {gemm_tflops / peaks['bf16_tflops']:.3f} of the burst figure",
         "note": "2 M N K FLOP per launch / CUDA-event time of every launch on the engine's stream, summed over the timed steps",
         # DRAM bytes per launch from ncu --set full of one Turbo encoder layer (profiles/r2_full_gemm_tn.md: 1307 MB for 944 GFLOP over
         # the QKV / out / FC1 / FC2 launches), scaled by this run's FLOPs per launch
         "traffic": 1307e6 / 944e9 * st_dev["gemm_flops"] / max(st_dev["gemm_launches"], 1)},
        {"kernel": "k_attn_enc_ts (tcgen05 encoder attention, Q/P as TMEM operands, single exp sweep)", "bound": "tensor",
         "achieved": attn_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": attn_tflops / peak_tf,
         "launches_per_step": st_dev["attn_launches"] / steps, "share_of_step": st_dev["attn_ms"] / steps / ms_step,
         "note": "algorithmic 4 T^2 d FLOP; the kernel is bound by MUFU.EX2 (d_head 64), see profiles/", "traffic": None},
        {"kernel": "k_logmel + k_logmel_norm", "bound": "hbm", "achieved": mel_gbps, "peak": peak_hbm, "unit": "GB/s",
         "frac": mel_gbps / peak_hbm, "share_of_step": st_dev["mel_ms"] / steps / ms_step,
         "note": "fp32 400-point FFT on the CUDA cores: instruction-bound (DESIGN.md 4.2); one ragged launch, device clips read in place, "
                 "normalisation folded into the conv1 im2col", "traffic": None},
    ]
    entries.sort(key=lambda e: -e["share_of_step"])
    if entries[0]["frac"] is None:      # the dominant entry must carry a roofline fraction
        entries[0], entries[1] = entries[1], entries[0]
    line["roofline"] = entries[0]
    line["roofline_extra"] = entries[1:]
    if cpu_base:
        line["cpu_baseline"] = cpu_base
        line["parity"] = parity
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
