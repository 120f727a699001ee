"""ctypes wrapper of oracle/cpu_ref/whisper_cpu_ref.cpp: the multi-threaded C++ restatement of the reference's CPU path.

TEST INFRASTRUCTURE ONLY (oracle/__init__.py).  PARITY UNPINNED (see the header of whisper_cpu_ref.cpp).  Used by
``tests/`` (validated against the numpy oracle), by ``__graft_entry__.build()`` (compiled, not used) and by
``bench.py``'s CPU baseline / ``--impl reference`` arm.  Never imported by spittle_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwhisper_cpu_ref.so")
_SRCS = [os.path.join(_HERE, f) for f in ("whisper_cpu_ref.cpp", "kernels.inc")]


def build(force: bool = False) -> str:
    """g++ -O3 with an x86-64-v3 baseline (AVX2 + FMA + F16C); the AVX-512 kernels are selected at run time, so the
    library built in the CPU-only container also runs on the GPU box's host whatever its generation."""
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in _SRCS):
        return LIB_PATH
    cmd = ["g++", "-O3", "-std=c++17", "-mavx2", "-mfma", "-mf16c", "-fPIC", "-shared", "-pthread", "-o", LIB_PATH, _SRCS[0]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("cpu_ref build failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


class _Cfg(C.Structure):
    _fields_ = [("language_id", C.c_int), ("translate", C.c_int), ("no_timestamps", C.c_int), ("suppress_blank", C.c_int),
                ("single_segment", C.c_int), ("max_initial_ts", C.c_float), ("n_max_override", C.c_int),
                ("n_max_text_ctx", C.c_int), ("initial_prompt", C.POINTER(C.c_int32)), ("n_initial_prompt", C.c_int),
                ("suppress_nst", C.c_int)]


class _Window(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("seek", "n_tokens", "result_len", "seek_delta", "failed", "token_offset", "n_prompt")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        vp, i32, fp = C.c_void_p, C.c_int, C.c_void_p
        l.wcr_last_error.restype = C.c_char_p
        l.wcr_load.argtypes = [C.c_char_p, C.POINTER(vp)]
        l.wcr_free.argtypes = [vp]
        l.wcr_free.restype = None
        l.wcr_hparams.argtypes = [vp, vp]
        l.wcr_hparams.restype = None
        l.wcr_logmel_geometry.argtypes = [C.c_size_t, C.POINTER(i32), C.POINTER(i32)]
        l.wcr_logmel.argtypes = [vp, fp, C.c_size_t, i32, fp, C.POINTER(i32), C.POINTER(i32)]
        l.wcr_encode.argtypes = [vp, fp, i32, fp]
        l.wcr_decode_window.argtypes = [vp, fp, i32, i32, C.POINTER(_Cfg), vp, i32, vp, i32, i32, vp, vp, C.POINTER(_Window), vp]
        l.wcr_detect_language.argtypes = [vp, fp, i32]
        l.wcr_full.argtypes = [vp, fp, C.c_size_t, C.POINTER(_Cfg), i32, i32, vp, vp, i32, C.POINTER(_Window), C.POINTER(i32),
                               C.POINTER(i32), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib = l
    return _lib


def _cfg(language_id=0, translate=False, no_timestamps=False, suppress_blank=True, single_segment=False, max_initial_ts=1.0,
         n_max_override=None, n_max_text_ctx=16384, initial_prompt_tokens: Optional[Sequence[int]] = None, suppress_nst=False):
    ip = np.ascontiguousarray(initial_prompt_tokens if initial_prompt_tokens is not None else [], np.int32)
    c = _Cfg(language_id, int(translate), int(no_timestamps), int(suppress_blank), int(single_segment), max_initial_ts,
             n_max_override or 0, n_max_text_ctx, ip.ctypes.data_as(C.POINTER(C.c_int32)) if ip.size else None, int(ip.size),
             int(suppress_nst))
    return c, ip          # keep ip alive


class CpuRef:
    """One loaded GGML model (f32 / f16 tensors)."""

    def __init__(self, model_path: str, n_threads: Optional[int] = None):
        self._h = C.c_void_p()
        if lib().wcr_load(model_path.encode(), C.byref(self._h)) != 0:
            raise RuntimeError("cpu_ref: " + lib().wcr_last_error().decode())
        hp = (C.c_int32 * 11)()
        lib().wcr_hparams(self._h, hp)
        (self.n_vocab, self.n_audio_ctx, self.n_audio_state, self.n_audio_head, self.n_audio_layer, self.n_text_ctx,
         self.n_text_state, self.n_text_head, self.n_text_layer, self.n_mels, self.ftype) = list(hp)
        self.n_threads = n_threads or (os.cpu_count() or 1)
        self.isa = int(lib().wcr_isa())

    def close(self):
        if getattr(self, "_h", None):
            lib().wcr_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def logmel(self, pcm: np.ndarray):
        x = np.ascontiguousarray(pcm, np.float32)
        a, b = C.c_int(), C.c_int()
        lib().wcr_logmel_geometry(x.size, C.byref(a), C.byref(b))
        out = np.empty((self.n_mels, a.value), np.float32)
        if lib().wcr_logmel(self._h, x.ctypes.data, x.size, self.n_threads, out.ctypes.data, C.byref(a), C.byref(b)) != 0:
            raise RuntimeError("cpu_ref: " + lib().wcr_last_error().decode())
        return out, b.value

    def encode(self, mel_win: np.ndarray) -> np.ndarray:
        m = np.ascontiguousarray(mel_win, np.float32)
        assert m.shape == (self.n_mels, 2 * self.n_audio_ctx)
        out = np.empty((self.n_audio_ctx, self.n_audio_state), np.float32)
        lib().wcr_encode(self._h, m.ctypes.data, self.n_threads, out.ctypes.data)
        return out

    def detect_language(self, enc: np.ndarray) -> int:
        e = np.ascontiguousarray(enc, np.float32)
        return int(lib().wcr_detect_language(self._h, e.ctypes.data, self.n_threads))

    def decode_window(self, enc: np.ndarray, seek: int, seek_end: int, prompt_past: Sequence[int] = (), forced=None,
                      trace: bool = False, **cfg_kw):
        """-> dict(tokens, margins, result_len, seek_delta, failed, n_prompt[, logits [n_tokens, n_vocab]])"""
        e = np.ascontiguousarray(enc, np.float32)
        cfg, _keep = _cfg(**cfg_kw)
        n_max = self.n_text_ctx // 2 - 4
        if cfg.n_max_override > 0:
            n_max = min(n_max, cfg.n_max_override)
        toks = np.full(self.n_text_ctx, -1, np.int32)
        marg = np.zeros(self.n_text_ctx, np.float32)
        past = np.ascontiguousarray(list(prompt_past), np.int32)
        f = None if forced is None else np.ascontiguousarray(forced, np.int32)
        logits = np.empty((n_max, self.n_vocab), np.float32) if trace else None
        w = _Window()
        rc = lib().wcr_decode_window(self._h, e.ctypes.data, seek, seek_end, C.byref(cfg), past.ctypes.data if past.size else None,
                                     int(past.size), f.ctypes.data if f is not None else None, 0 if f is None else int(f.size),
                                     self.n_threads, toks.ctypes.data, marg.ctypes.data, C.byref(w),
                                     logits.ctypes.data if logits is not None else None)
        if rc != 0:
            raise RuntimeError("cpu_ref: " + lib().wcr_last_error().decode())
        out = dict(tokens=[int(t) for t in toks[: w.n_tokens]], margins=[float(x) for x in marg[: w.n_tokens]],
                   result_len=w.result_len, seek_delta=w.seek_delta, failed=bool(w.failed), n_prompt=w.n_prompt)
        if trace:
            out["logits"] = logits[: w.n_tokens]
        return out

    def full(self, pcm: np.ndarray, max_windows: int = 64, **cfg_kw):
        """whisper_full -> dict(windows=[{tokens, margins, seek, result_len, seek_delta, failed, n_prompt}], kept, lang,
        t_mel, t_enc, t_dec seconds)"""
        x = np.ascontiguousarray(pcm, np.float32)
        cfg, _keep = _cfg(**cfg_kw)
        cap = max_windows * self.n_text_ctx
        toks = np.empty(cap, np.int32)
        marg = np.empty(cap, np.float32)
        wins = (_Window * max_windows)()
        nw, lang = C.c_int(0), C.c_int(0)
        tm, te, td = C.c_double(0), C.c_double(0), C.c_double(0)
        rc = lib().wcr_full(self._h, x.ctypes.data, x.size, C.byref(cfg), self.n_threads, max_windows, toks.ctypes.data,
                            marg.ctypes.data, cap, wins, C.byref(nw), C.byref(lang), C.byref(tm), C.byref(te), C.byref(td))
        if rc != 0:
            raise RuntimeError("cpu_ref: " + lib().wcr_last_error().decode())
        out_w, kept = [], []
        for i in range(nw.value):
            w = wins[i]
            t = [int(v) for v in toks[w.token_offset: w.token_offset + w.n_tokens]]
            out_w.append(dict(tokens=t, margins=[float(v) for v in marg[w.token_offset: w.token_offset + w.n_tokens]], seek=w.seek,
                              result_len=w.result_len, seek_delta=w.seek_delta, failed=bool(w.failed), n_prompt=w.n_prompt))
            kept.extend(t[: w.result_len])
        return dict(windows=out_w, kept=kept, lang=lang.value, t_mel=tm.value, t_enc=te.value, t_dec=td.value)
