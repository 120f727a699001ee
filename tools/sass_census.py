#!/usr/bin/env python
"""SASS opcode census of libspittle_b200.so: per kernel, how many Blackwell-native instructions the compiled code holds.

  UTCHMMA   tcgen05.mma (f16 / bf16 kinds)        UTMALDG  cp.async.bulk.tensor (TMA load)
  LDTM/STTM tcgen05.ld / tcgen05.st (TMEM)        UTCBAR   tcgen05.commit -> mbarrier
  SYNCS     mbarrier arrive / try_wait            HMMA     legacy mma.sync (decoder-step kernels)
  UTCATOMSWS tcgen05.alloc / dealloc              UCGABAR  cluster barrier (cta_group::2 pairs)
  LDSM      ldmatrix (B operands of the front-end)  DFMA   fp64 FMA (Silero f64 FFT, resampler operator build)

Usage: python tools/sass_census.py > profiles/r2_sass_census.md   (needs cuobjdump; no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spittle_b200", "libspittle_b200.so")
COLS = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "HMMA", "LDSM", "DFMA", "MUFU", "LDGSTS", "total"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur][op] += 1
            counts[cur]["total"] += 1
            if op.startswith("UCGABAR"):
                counts[cur]["UCGABAR"] += 1
    names = demangle(list(counts))
    print("# SASS opcode census of libspittle_b200.so (sm_100a), `python tools/sass_census.py`\n")
    print("Counts of instructions in the compiled kernels (static, per kernel instantiation; f16 instantiations shown, the bf16")
    print("ones are identical up to the operand kind).  Kernels without any listed opcode are omitted.\n")
    print("| kernel | " + " | ".join(COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for k, c in counts.items():
        d = names.get(k, k)
        if "__nv_bfloat16" in d:
            continue
        if not any(c[x] for x in COLS[:-3]):
            continue
        short = re.sub(r"\(.*", "", d).replace("sb::", "").replace("void ", "")
        print(f"| `{short}` | " + " | ".join(str(c[x]) if c[x] else "" for x in COLS) + " |")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("\nWhole library: " + ", ".join(f"{x} {tot[x]}" for x in COLS[:-1]) + f"; {len(counts)} kernels, {tot['total']} instructions.")
    print("\n`wgmma` (sm_90a) does not exist in this binary; the tensor-core work of the encoder (`k_gemm_tn`, `k_attn_enc_ts`) is")
    print("UTCHMMA = `tcgen05.mma` with TMEM accumulators (LDTM / STTM) fed by TMA (UTMALDG); the HMMA counts belong to the")
    print("decoder-step weight-streaming kernels (`k_skinny_gemm*`, `k_dec_cross_attn`: M = 16-row weight tiles against <= 64")
    print("sequences, HBM-bound) and to the capture front-end (`k_resample_poly`: f16 3-pass polyphase Toeplitz GEMM; `k_silero_features_fft`:")
    print("bf16 residual of the STFT basis + 3xTF32 block-1 product next to its f64 FFT (DFMA); `k_silero_lstm`: f16 3-pass recurrent product),")
    print("with LDSM = `ldmatrix` feeding their B operands.")


if __name__ == "__main__":
    sys.exit(main())
