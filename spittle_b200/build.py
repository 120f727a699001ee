"""In-tree build of libspittle_b200.so (nvcc, sm_100a only).

``python -m spittle_b200.build`` or ``spittle_b200.build.build()``.  nvcc cross-compiles
without a GPU; the resulting .so is git-ignored but travels with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libspittle_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return sorted(hs)


def build(verbose: bool = False, force: bool = False, ptxas_info: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr_digest = _digest(headers())
    jobs = []
    objs = []
    for src in sources():
        name = os.path.splitext(os.path.basename(src))[0]
        obj = os.path.join(OBJ, name + ".o")
        stamp = obj + ".sha"
        dg = _digest([src]) + hdr_digest
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dg:
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-x", "cu", "-c", src, "-o", obj]
        jobs.append((cmd, stamp, dg, src))

    def run(job):
        cmd, stamp, dg, src = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose or ptxas_info:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(dg)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    build_host()
    return LIB


def build_host() -> str:
    """C++ host-side mirror of the reference TranscriptionManager + a small CLI over it."""
    root = os.path.dirname(HERE)
    out = os.path.join(root, "host", "sb_transcribe_cli")
    srcs = [os.path.join(root, "host", f) for f in ("transcription_manager.cpp", "text_filters.cpp", "jargon.cpp", "sb_transcribe_cli.cpp")]
    newest = max(os.path.getmtime(p) for p in srcs + [os.path.join(root, "host", "transcription_manager.hpp"),
                                                      os.path.join(root, "host", "text_filters.hpp"),
                                                      os.path.join(root, "host", "jargon.hpp"), LIB])
    if os.path.exists(out) and os.path.getmtime(out) >= newest:
        return out
    cmd = ["g++", "-std=c++17", "-O2", "-I", INCLUDE] + srcs + ["-o", out, "-L", HERE, "-lspittle_b200",
                                                              "-Wl,-rpath,$ORIGIN/../spittle_b200", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build failed:\n%s\n%s" % (r.stdout, r.stderr))
    return out


if __name__ == "__main__":
    p = build(verbose=True, force="--force" in sys.argv, ptxas_info="--ptxas" in sys.argv)
    print(p)
