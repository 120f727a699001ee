"""Per-kernel micro-benchmarks (CUDA events, warm-up, L2 flush between iterations).
Usage: python tools/kbench.py [gemm] [logmel]   -- prints one JSON line per case."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from spittle_b200 import capi, synth

dev = torch.device("cuda:0")
PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else \
    {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_buf.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def bench_gemm():
    st = torch.cuda.current_stream().cuda_stream
    for dtype, tdt in ((capi.SB_DTYPE_BF16, torch.bfloat16),):
        for (M, N, K, tag) in [(96000, 2304, 768, "small qkv B64"), (96000, 768, 768, "small out B64"),
                               (96000, 3072, 768, "small mlp1 B64"), (96000, 768, 3072, "small mlp2 B64"),
                               (48000, 3840, 1280, "v3 qkv B32"), (48000, 5120, 1280, "v3 mlp1 B32"),
                               (48000, 1280, 5120, "v3 mlp2 B32"), (8192, 8192, 8192, "square 8192")]:
            A = (torch.randn(M, K, device=dev) * 0.5).to(tdt)
            W = (torch.randn(N, K, device=dev) * 0.05).to(tdt)
            out = torch.empty(M, N, dtype=tdt, device=dev)
            bias = torch.randn(N, device=dev)
            f = lambda: capi.gemm_tn_dev(dtype, A.data_ptr(), K, W.data_ptr(), K, M, N, K, out.data_ptr(), N, False,
                                         bias.data_ptr(), 0, 0, 0, 0, st)
            med, mn = timeit(f, flush=False)
            ref = lambda: torch.addmm(bias.to(tdt), A, W.T)
            rmed, rmn = timeit(ref, flush=False)
            fl = 2.0 * M * N * K
            print(json.dumps({"kernel": "gemm_tcgen05", "case": tag, "M": M, "N": N, "K": K, "ms": med,
                              "tflops": fl / med / 1e9, "frac_burst": fl / med / 1e9 / PEAKS["bf16_tflops"],
                              "cublas_ms": rmed, "cublas_tflops": fl / rmed / 1e9}), flush=True)
            del A, W, out


def bench_gemm_epilogue():
    """the three epilogue kinds of the encoder GEMMs on the Large-v3 / Turbo shapes (64 windows): 16-bit out, f32 out,
    f32 out + f32 residual read (what the O-projection and FC2 do: x += W a + b in the f32 residual stream)"""
    st = torch.cuda.current_stream().cuda_stream
    dtype, tdt = capi.SB_DTYPE_F16, torch.float16
    for (M, N, K, tag) in [(96000, 1280, 1280, "turbo out-proj B64"), (96000, 1280, 5120, "turbo fc2 B64"),
                           (96000, 768, 768, "small out-proj B64"), (96000, 768, 3072, "small fc2 B64")]:
        A = (torch.randn(M, K, device=dev) * 0.5).to(tdt)
        W = (torch.randn(N, K, device=dev) * 0.05).to(tdt)
        o16 = torch.empty(M, N, dtype=tdt, device=dev)
        o32 = torch.zeros(M, N, dtype=torch.float32, device=dev)
        bias = torch.randn(N, device=dev)
        fl = 2.0 * M * N * K
        cases = {
            "out16": lambda: capi.gemm_tn_dev(dtype, A.data_ptr(), K, W.data_ptr(), K, M, N, K, o16.data_ptr(), N, False, bias.data_ptr(), 0, 0, 0, 0, st),
            "out32": lambda: capi.gemm_tn_dev(dtype, A.data_ptr(), K, W.data_ptr(), K, M, N, K, o32.data_ptr(), N, True, bias.data_ptr(), 0, 0, 0, 0, st),
            "out32+res": lambda: capi.gemm_tn_dev(dtype, A.data_ptr(), K, W.data_ptr(), K, M, N, K, o32.data_ptr(), N, True, bias.data_ptr(), 0, o32.data_ptr(), N, 0, st),
        }
        row = {"kernel": "gemm_tcgen05 epilogues", "case": tag, "M": M, "N": N, "K": K}
        for name, f in cases.items():
            med, mn = timeit(f, flush=False)
            row[name + "_ms"] = med
            row[name + "_tflops"] = fl / med / 1e9
        print(json.dumps(row), flush=True)
        del A, W, o16, o32


def bench_attn():
    """encoder attention, exact f32 sweep or packed (env SB_ATTN_PACKED, read once per process), with the error against
    an f32 torch reference on the same 16-bit inputs"""
    st = torch.cuda.current_stream().cuda_stream
    l = capi.lib()
    for dtype, tdt in ((capi.SB_DTYPE_F16, torch.float16), (capi.SB_DTYPE_BF16, torch.bfloat16)):
        for (Wn, H, tag) in [(16, 20, "turbo 16 windows"), (16, 12, "small 16 windows")]:
            d = 64 * H
            T = 1500
            torch.manual_seed(1)
            qkv = (torch.randn(Wn * T, 3 * d, device=dev) * 1.5).to(tdt)
            out = torch.empty(Wn * T, d, dtype=tdt, device=dev)
            f = lambda: capi.check(l.sb_attn_enc_dev(dtype, qkv.data_ptr(), out.data_ptr(), Wn, T, d, H, st))
            med, mn = timeit(f, flush=False)
            # reference on window 0, f32
            q, k, v = [x.float().reshape(T, H, 64).transpose(0, 1) for x in qkv[:T].split(d, dim=1)]
            ref = torch.softmax(q @ k.transpose(1, 2) / 8.0, dim=-1) @ v
            ref = ref.transpose(0, 1).reshape(T, d)
            got = out[:T].float()
            err = (got - ref)
            fl = 4.0 * Wn * H * T * T * 64
            print(json.dumps({"kernel": "k_attn_enc_ts", "packed": os.environ.get("SB_ATTN_PACKED", "0"), "dtype": str(tdt), "case": tag,
                              "ms": med, "tflops": fl / med / 1e9, "rel_rms": (err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item(),
                              "max_abs": err.abs().max().item()}), flush=True)


def bench_logmel():
    st = torch.cuda.current_stream().cuda_stream
    for n_mel in (80, 128):
        plan = capi.MelPlan(synth.mel_filterbank(n_mel))
        for B in (64, 256, 1024):
            n = 480000
            n_len, n_len_org, n_calc = capi.logmel_geometry(n)
            stride = (n_calc + 31) // 32 * 32
            base = torch.from_numpy(np.stack([synth.make_clip(i) for i in range(8)])).to(dev)
            pcm = base.repeat((B + 7) // 8, 1)[:B].contiguous()
            mel = torch.empty((B, n_mel, stride), dtype=torch.float32, device=dev)
            cmax = torch.empty(B, dtype=torch.int32, device=dev)
            f = lambda: capi.logmel_batch_dev(plan, pcm.data_ptr(), B, n, mel.data_ptr(), stride, cmax.data_ptr(), 0, st)
            med, mn = timeit(f, flush=True)
            alg = B * (n * 4 + n_mel * 3000 * 4)
            print(json.dumps({"kernel": "k_logmel(+norm)", "n_mel": n_mel, "clips": B, "ms": med,
                              "GBps": alg / med / 1e6, "frac_hbm": alg / med / 1e6 / PEAKS["hbm_gbs"]}), flush=True)
            del pcm, mel


def bench_skinny():
    st = torch.cuda.current_stream().cuda_stream
    l = capi.lib()
    for (B, N, K, tag) in [(64, 768, 768, "small dxd"), (64, 2304, 768, "small qkv"), (64, 3072, 768, "small fc1"),
                           (64, 768, 3072, "small fc2"), (64, 51865, 768, "small logits"), (64, 1280, 1280, "v3 dxd"),
                           (64, 5120, 1280, "v3 fc1"), (64, 1280, 5120, "v3 fc2"), (64, 51866, 1280, "v3 logits")]:
        X = (torch.randn(B, K, device=dev) * 0.5).half()
        # rotate over several weight copies so the stream comes from HBM, not L2
        n_copies = max(2, int(300e6 // (N * K * 2)) + 1)
        Ws = [(torch.randn(N, K, device=dev) * 0.05).half() for _ in range(min(n_copies, 64))]
        out = torch.empty(B, N, dtype=torch.float32, device=dev)
        bias = torch.randn(N, device=dev)
        it = [0]

        def f():
            W = Ws[it[0] % len(Ws)]
            it[0] += 1
            capi.check(l.sb_skinny_gemm_dev(capi.SB_DTYPE_F16, X.data_ptr(), K, W.data_ptr(), K, B, N, K, bias.data_ptr(), 0,
                                            None, 0, out.data_ptr(), N, None, 0, st))
        med, mn = timeit(f, iters=20, flush=False)
        ref = (X.float() @ Ws[(it[0] - 1) % len(Ws)].float().T + bias)
        err = (out - ref).abs().max().item()
        byt = N * K * 2
        print(json.dumps({"kernel": "k_skinny_gemm", "case": tag, "B": B, "N": N, "K": K, "us": med * 1e3,
                          "GBps": byt / med / 1e6, "frac_hbm": byt / med / 1e6 / PEAKS["hbm_gbs"], "max_err": err}), flush=True)
        del Ws


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "logmel"]
    if "gemm" in what:
        bench_gemm()
    if "logmel" in what:
        bench_logmel()
    if "skinny" in what:
        bench_skinny()
    if "epilogue" in what:
        bench_gemm_epilogue()
    if "attn" in what:
        bench_attn()
