"""C5: 48 kHz multi-stream capture front-end only (BASELINE.json configs[4]):
polyphase resample -> Silero VAD (+ SmoothedVad gate) -> log-mel over N synthetic 30 s streams.
Algorithmic bytes per stream-30 s: 5.76 MB in + n_mel*3000*4 B mel + 4 KB probs = 6.724 MB (80 mel);
the materialised 16 kHz intermediate adds 1.92 MB written + 1.92 MB re-read twice (VAD, log-mel).
Usage: python tools/frontend_bench.py [n_streams] [chunk]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from spittle_b200 import audio_toolkit, capi, silero_weights, synth


def run(n_streams: int = 2000, chunk: int = 500, device_index: int = 0) -> dict:
    dev = torch.device("cuda", device_index)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peaks = json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    base = np.stack([synth.make_clip(i, 30.0, sr=48000, kind=["vowel", "mix", "tone", "noise"][i % 4]) for i in range(8)])
    x48 = torch.from_numpy(base).to(dev).repeat((chunk + 7) // 8, 1)[:chunk].contiguous()
    rs = audio_toolkit.FrameResampler(48000)
    sv = audio_toolkit.SileroVad(os.path.join(root, "tests", "golden", "silero_v4_16k.npz"), 0.3)
    plan = capi.MelPlan(synth.mel_filterbank(80))
    st = torch.cuda.current_stream().cuda_stream

    def run_chunk():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record()
        frames = rs.process(x48)                                    # [chunk, n_frames, 480]
        e[1].record()
        sv.reset()
        probs = sv.score(frames)
        e[2].record()
        n_frames = frames.shape[1]
        pcm16 = frames.view(chunk, n_frames * 480)
        n = n_frames * 480
        n_len, n_len_org, n_calc = capi.logmel_geometry(n)
        stride = (n_calc + 31) // 32 * 32
        mel = torch.empty((chunk, 80, stride), dtype=torch.float32, device=dev)
        cmax = torch.empty(chunk, dtype=torch.int32, device=dev)
        e[3].record()
        capi.logmel_batch_dev(plan, pcm16.data_ptr(), chunk, n, mel.data_ptr(), stride, cmax.data_ptr(), 0, st)
        e[4].record()
        torch.cuda.synchronize()
        return dict(resample=e[0].elapsed_time(e[1]), silero=e[1].elapsed_time(e[2]), logmel=e[3].elapsed_time(e[4])), probs

    run_chunk()
    tot = dict(resample=0.0, silero=0.0, logmel=0.0)
    n_chunks = max(1, n_streams // chunk)
    probs = None
    for _ in range(n_chunks):
        t, probs = run_chunk()
        for k in tot:
            tot[k] += t[k]
    n_done = n_chunks * chunk
    alg = n_done * 6.724e6
    ms = sum(tot.values())
    return {"workload": f"C5 front-end, {n_done} synthetic 30 s streams @48 kHz in chunks of {chunk}", "ms_total": ms,
            "ms": tot, "alg_GB": alg / 1e9, "GBps": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / peaks["hbm_gbs"],
            "per_stage_GBps": {"resample (5.76+1.92 MB)": n_done * 7.68e6 / tot["resample"] / 1e6,
                               "silero (1.92 MB + 4 KB)": n_done * 1.924e6 / tot["silero"] / 1e6,
                               "logmel (1.92+0.96 MB)": n_done * 2.88e6 / tot["logmel"] / 1e6},
            "speech_frame_fraction": float((probs > 0.3).float().mean())}


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 2000, int(sys.argv[2]) if len(sys.argv) > 2 else 500)))
