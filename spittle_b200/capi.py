"""ctypes binding of libspittle_b200.so -- the C ABI declared in include/spittle_b200.h.

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspittle_b200.so")

SB_DTYPE_BF16 = 0
SB_DTYPE_F16 = 1


class SbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sb_status {code}: {msg}")
        self.code = code


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m spittle_b200.build` "
                "(there is no CPU fallback in spittle_b200)")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(l: C.CDLL) -> None:
    vp, i32, i64, sz, fp = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.POINTER(C.c_float)
    l.sb_last_error.restype = C.c_char_p
    l.sb_version.restype = C.c_char_p
    l.sb_launch_count.restype = C.c_uint64
    l.sb_melplan_create.argtypes = [vp, i32, C.POINTER(vp)]
    l.sb_melplan_destroy.argtypes = [vp]
    l.sb_logmel_geometry.argtypes = [sz, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    l.sb_logmel.argtypes = [vp, vp, sz, vp, C.POINTER(i32), C.POINTER(i32)]
    l.sb_logmel_batch_dev.argtypes = [vp, vp, i32, sz, vp, i32, vp, vp, vp]
    l.sb_gemm_tn_dev.argtypes = [i32, vp, i64, vp, i64, i32, i32, i32, vp, i64, i32, vp, i32, vp, i64, i32, vp]
    for name in dir(l):
        pass


def check(rc: int) -> None:
    if rc != 0:
        raise SbError(rc, lib().sb_last_error().decode(errors="replace"))


def version() -> str:
    return lib().sb_version().decode()


def launch_count() -> int:
    return int(lib().sb_launch_count())


def logmel_geometry(n_samples: int):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(lib().sb_logmel_geometry(n_samples, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


class MelPlan:
    """Device-resident sparse mel filterbank (sb_melplan)."""

    def __init__(self, filters: np.ndarray):
        f = np.ascontiguousarray(filters, dtype=np.float32)
        assert f.ndim == 2 and f.shape[1] == 201
        self.n_mel = f.shape[0]
        self._h = C.c_void_p()
        check(lib().sb_melplan_create(f.ctypes.data, self.n_mel, C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def logmel(self, pcm: np.ndarray):
        """Host API: one clip -> ([n_mel, n_len] f32, n_len_org) like whisper_pcm_to_mel."""
        x = np.ascontiguousarray(pcm, dtype=np.float32)
        n_len, n_len_org, _ = logmel_geometry(x.shape[0])
        out = np.empty((self.n_mel, n_len), np.float32)
        a, b = C.c_int(), C.c_int()
        check(lib().sb_logmel(self._h, x.ctypes.data, x.shape[0], out.ctypes.data, C.byref(a), C.byref(b)))
        return out, b.value

    def close(self):
        if self._h:
            lib().sb_melplan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def logmel_batch_dev(plan: MelPlan, pcm_ptr: int, n_clips: int, n_samples: int, mel_ptr: int, mel_stride: int,
                     clip_max_ptr: int, floor_ptr: int = 0, stream: int = 0) -> None:
    check(lib().sb_logmel_batch_dev(plan.handle, pcm_ptr, n_clips, n_samples, mel_ptr, mel_stride, clip_max_ptr,
                                    floor_ptr or None, stream or None))


def gemm_tn_dev(dtype: int, a_ptr: int, lda: int, w_ptr: int, ldw: int, M: int, N: int, K: int, out_ptr: int,
                ldo: int, out_f32: bool, bias_ptr: int = 0, act: int = 0, res_ptr: int = 0, ldr: int = 0,
                res_row_mod: int = 0, stream: int = 0) -> None:
    check(lib().sb_gemm_tn_dev(dtype, a_ptr, lda, w_ptr, ldw, M, N, K, out_ptr, ldo, int(out_f32),
                               bias_ptr or None, act, res_ptr or None, ldr, res_row_mod, stream or None))
