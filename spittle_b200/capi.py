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
    _declare_engine(l)
    _declare_frontend(l)


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


# ---------------------------------------------------------------------------------------
# engine
# ---------------------------------------------------------------------------------------
class SbConfig(C.Structure):
    _fields_ = [("model_path", C.c_char_p), ("device", C.c_int), ("max_batch", C.c_int), ("dtype", C.c_int),
                ("use_cuda_graph", C.c_int), ("devices", C.POINTER(C.c_int)), ("n_devices", C.c_int)]


class SbModelInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer", "n_text_ctx", "n_text_state",
        "n_text_head", "n_text_layer", "n_mels", "ftype", "token_eot", "token_sot", "token_beg", "token_blank")]


class SbParams(C.Structure):
    _fields_ = [("language", C.c_char_p), ("translate", C.c_int), ("initial_prompt", C.c_char_p),
                ("no_timestamps", C.c_int), ("suppress_blank", C.c_int), ("single_segment", C.c_int),
                ("max_initial_ts", C.c_float), ("n_max_tokens", C.c_int), ("max_windows", C.c_int),
                ("n_max_text_ctx", C.c_int), ("temperature", C.c_float), ("temperature_inc", C.c_float),
                ("logprob_thold", C.c_float), ("entropy_thold", C.c_float), ("suppress_nst", C.c_int)]


class SbStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "clips", "windows", "rounds", "decoder_steps", "tokens_sampled", "pcm_bytes", "h2d_bytes", "d2h_bytes",
        "mel_ms", "encode_ms", "decode_ms", "gemm_ms", "gemm_flops", "gemm_launches", "attn_ms", "attn_flops",
        "attn_launches", "skinny_ms", "skinny_bytes", "skinny_launches", "xattn_ms", "xattn_bytes", "xattn_launches",
        "dln_ms", "dln_launches", "dself_ms", "dself_launches", "dstep_ms", "dstep_count", "prefill_rows", "fallbacks")]


class SbWindowInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("seek", "n_tokens", "result_len", "seek_delta", "failed", "token_offset",
                                         "n_prompt")] + [("temperature", C.c_float), ("n_attempts", C.c_int32),
                                                         ("avg_logprob", C.c_float)]


class SbSegment(C.Structure):
    _fields_ = [("t0", C.c_int64), ("t1", C.c_int64), ("text", C.POINTER(C.c_char)), ("text_len", C.c_size_t),
                ("token_offset", C.c_int32), ("n_tokens", C.c_int32)]


class SbResult(C.Structure):
    _fields_ = [("text", C.POINTER(C.c_char)), ("text_len", C.c_size_t),
                ("tokens", C.POINTER(C.c_int32)), ("n_tokens", C.c_size_t),
                ("sampled", C.POINTER(C.c_int32)), ("n_sampled", C.c_size_t),
                ("margins", C.POINTER(C.c_float)), ("tids", C.POINTER(C.c_int32)), ("logprobs", C.POINTER(C.c_float)),
                ("windows", C.POINTER(SbWindowInfo)), ("n_windows", C.c_size_t),
                ("segments", C.POINTER(SbSegment)), ("n_segments", C.c_size_t), ("segment_text", C.POINTER(C.c_char)),
                ("ms_mel", C.c_float), ("ms_encode", C.c_float), ("ms_decode", C.c_float),
                ("status", C.c_int), ("lang_id", C.c_int)]


class SbAbiField(C.Structure):
    _fields_ = [("struct_name", C.c_char_p), ("field", C.c_char_p), ("struct_size", C.c_int), ("offset", C.c_int)]


def abi_layout():
    """The library's own sizeof / offsetof table: [(struct, field, sizeof, offset)] (sb_abi_layout)."""
    l = lib()
    n = l.sb_abi_layout(None, 0)
    rows = (SbAbiField * n)()
    l.sb_abi_layout(rows, n)
    return [(r.struct_name.decode(), r.field.decode(), r.struct_size, r.offset) for r in rows]


ABI_STRUCTS = {}     # name -> ctypes mirror; filled below (checked against abi_layout() by tests/test_abi.py)


ABI_STRUCTS.update(sb_config=SbConfig, sb_params=SbParams, sb_window_info=SbWindowInfo, sb_segment=SbSegment,
                   sb_result=SbResult, sb_model_info=SbModelInfo, sb_stats=SbStats)


def _declare_engine(l: C.CDLL) -> None:
    vp, i32 = C.c_void_p, C.c_int
    l.sb_abi_layout.argtypes = [C.POINTER(SbAbiField), i32]
    l.sb_engine_device_count.argtypes = [vp]
    l.sb_params_default.argtypes = [C.POINTER(SbParams)]
    l.sb_params_default.restype = None
    l.sb_engine_create.argtypes = [C.POINTER(SbConfig), C.POINTER(vp)]
    l.sb_engine_destroy.argtypes = [vp]
    l.sb_engine_info.argtypes = [vp, C.POINTER(SbModelInfo)]
    l.sb_token_text.argtypes = [vp, C.c_int32, C.c_char_p, i32]
    l.sb_tokenize.argtypes = [vp, C.c_char_p, vp, i32]
    l.sb_engine_stream.argtypes = [vp]
    l.sb_engine_stream.restype = vp
    l.sb_engine_set_profile.argtypes = [vp, i32]
    l.sb_engine_stats.argtypes = [vp, C.POINTER(SbStats), i32]
    l.sb_transcribe.argtypes = [vp, vp, C.c_size_t, C.POINTER(SbParams), C.POINTER(SbResult)]
    l.sb_transcribe_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(SbParams),
                                      C.POINTER(SbResult)]
    l.sb_result_free.argtypes = [C.POINTER(SbResult)]
    l.sb_result_free.restype = None
    l.sb_encode.argtypes = [vp, vp, i32, vp]
    l.sb_decode_trace.argtypes = [vp, vp, i32, vp, C.POINTER(SbParams), vp, i32, vp, vp, vp]
    l.sb_layernorm_dev.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, vp]
    i64 = C.c_int64
    l.sb_skinny_gemm_dev.argtypes = [i32, vp, i64, vp, i64, i32, i32, i32, vp, i32, vp, i64, vp, i64, vp, i64, vp]
    l.sb_attn_enc_dev.argtypes = [i32, vp, vp, i32, i32, i32, i32, vp]


class ClipResult:
    """Python view of one sb_result (copied out; the C result is freed)."""

    def __init__(self, r: SbResult):
        self.text = C.string_at(r.text, r.text_len) if r.text else b""
        self.tokens = [r.tokens[i] for i in range(r.n_tokens)]
        self.sampled = [r.sampled[i] for i in range(r.n_sampled)]
        self.margins = [r.margins[i] for i in range(r.n_sampled)]
        self.tids = [r.tids[i] for i in range(r.n_sampled)]
        self.logprobs = [r.logprobs[i] for i in range(r.n_sampled)]
        self.windows = [dict(seek=w.seek, n_tokens=w.n_tokens, result_len=w.result_len, seek_delta=w.seek_delta,
                             failed=w.failed, token_offset=w.token_offset, n_prompt=w.n_prompt, temperature=w.temperature,
                             n_attempts=w.n_attempts, avg_logprob=w.avg_logprob)
                        for w in (r.windows[i] for i in range(r.n_windows))]
        self.segments = [dict(t0=g.t0, t1=g.t1, text=C.string_at(g.text, g.text_len), token_offset=g.token_offset,
                              n_tokens=g.n_tokens) for g in (r.segments[i] for i in range(r.n_segments))]
        self.ms_mel, self.ms_encode, self.ms_decode = r.ms_mel, r.ms_encode, r.ms_decode
        self.status = r.status
        self.lang_id = r.lang_id


def default_params(**kw) -> SbParams:
    l = lib()
    p = SbParams()
    l.sb_params_default(C.byref(p))
    for k, v in kw.items():
        if k in ("language", "initial_prompt") and isinstance(v, str):
            v = v.encode()
        setattr(p, k, v)
    return p


class Engine:
    """sb_engine: one loaded model on one CUDA device."""

    def __init__(self, model_path: str, device: int = 0, max_batch: int = 64, dtype: int = SB_DTYPE_BF16,
                 use_cuda_graph: bool = True, devices=None):
        """devices: list of CUDA ordinals -> one model replica per device, clips of a batch are split over them."""
        l = lib()
        devs = (C.c_int * len(devices))(*devices) if devices else None
        cfg = SbConfig(model_path.encode(), device, max_batch, dtype, int(use_cuda_graph),
                       C.cast(devs, C.POINTER(C.c_int)) if devs is not None else None, len(devices) if devices else 0)
        self._h = C.c_void_p()
        check(l.sb_engine_create(C.byref(cfg), C.byref(self._h)))
        self.info = SbModelInfo()
        check(l.sb_engine_info(self._h, C.byref(self.info)))
        self.max_batch = max_batch
        self.dtype = dtype

    def close(self):
        if getattr(self, "_h", None):
            lib().sb_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        """cudaStream_t of the engine (wrap with torch.cuda.ExternalStream to record events on it)."""
        return int(lib().sb_engine_stream(self._h) or 0)

    def set_profile(self, enable) -> None:
        """0 off, 1 encoder GEMM / attention brackets, 2 additionally decoder projections + cross-attention (no graph)."""
        check(lib().sb_engine_set_profile(self._h, int(enable)))

    def stats(self, reset: bool = False) -> dict:
        st = SbStats()
        check(lib().sb_engine_stats(self._h, C.byref(st), int(reset)))
        return {n: getattr(st, n) for n, _ in SbStats._fields_}

    def tokenize(self, text) -> list:
        """whisper.cpp tokenisation of a prompt text with this model's vocabulary."""
        b = text if isinstance(text, bytes) else text.encode("utf-8")
        n = lib().sb_tokenize(self._h, b, None, 0)
        if n < 0:
            check(n)
        out = (C.c_int32 * max(n, 1))()
        lib().sb_tokenize(self._h, b, C.addressof(out), n)
        return [int(out[i]) for i in range(n)]

    def token_text(self, tid: int) -> bytes:
        buf = C.create_string_buffer(256)
        n = lib().sb_token_text(self._h, tid, buf, 256)
        return buf.raw[:min(n, 256)]

    def transcribe_batch_ptrs(self, ptrs, sizes, params: Optional[SbParams] = None):
        """ptrs: host addresses of f32 16 kHz clips (e.g. pinned torch tensors); sizes: samples per clip."""
        n = len(ptrs)
        arr_p = (C.c_void_p * n)(*ptrs)
        arr_n = (C.c_size_t * n)(*sizes)
        res = (SbResult * n)()
        p = params if params is not None else default_params()
        rc = lib().sb_transcribe_batch(self._h, arr_p, arr_n, n, C.byref(p), res)
        try:
            check(rc)
            return [ClipResult(res[i]) for i in range(n)]
        finally:
            for i in range(n):
                lib().sb_result_free(C.byref(res[i]))

    def transcribe_batch(self, clips, params: Optional[SbParams] = None):
        arrs = [np.ascontiguousarray(c, dtype=np.float32) for c in clips]
        return self.transcribe_batch_ptrs([a.ctypes.data if a.size else 0 for a in arrs], [a.size for a in arrs], params)

    def transcribe(self, pcm, params: Optional[SbParams] = None) -> "ClipResult":
        a = np.ascontiguousarray(pcm, dtype=np.float32)
        res = SbResult()
        p = params if params is not None else default_params()
        rc = lib().sb_transcribe(self._h, a.ctypes.data if a.size else None, a.size, C.byref(p), C.byref(res))
        try:
            check(rc)
            return ClipResult(res)
        finally:
            lib().sb_result_free(C.byref(res))

    def encode(self, mel_windows: np.ndarray) -> np.ndarray:
        m = np.ascontiguousarray(mel_windows, dtype=np.float32)
        W = m.shape[0]
        out = np.empty((W, self.info.n_audio_ctx, self.info.n_audio_state), np.float32)
        check(lib().sb_encode(self._h, m.ctypes.data, W, out.ctypes.data))
        return out

    def decode_trace(self, mel_windows: np.ndarray, seek_end, n_steps: int, forced=None, want_logits=True,
                     params: Optional[SbParams] = None):
        m = np.ascontiguousarray(mel_windows, dtype=np.float32)
        W = m.shape[0]
        se = np.ascontiguousarray(seek_end, dtype=np.int32)
        f = None if forced is None else np.ascontiguousarray(forced, dtype=np.int32)
        logits = np.empty((W, n_steps, self.info.n_vocab), np.float32) if want_logits else None
        toks = np.empty((W, n_steps), np.int32)
        marg = np.empty((W, n_steps), np.float32)
        p = params if params is not None else default_params()
        check(lib().sb_decode_trace(self._h, m.ctypes.data, W, se.ctypes.data, C.byref(p),
                                    f.ctypes.data if f is not None else None, n_steps,
                                    logits.ctypes.data if logits is not None else None, toks.ctypes.data,
                                    marg.ctypes.data))
        return logits, toks, marg


# ---------------------------------------------------------------------------------------
# capture front-end
# ---------------------------------------------------------------------------------------
def _declare_frontend(l: C.CDLL) -> None:
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    l.sb_downmix_mono_dev.argtypes = [vp, i32, i32, i64, sz, i32, vp, i64, vp]
    l.sb_pcm_f32_to_i16_dev.argtypes = [vp, vp, sz, vp]
    l.sb_pcm_f32_to_i16.argtypes = [vp, vp, sz]
    l.sb_resample_48k_16k.argtypes = [vp, sz, vp, sz, vp]
    l.sb_silero_v4.argtypes = [vp, vp, i32, vp, vp, vp]
    l.sb_vad_gate.argtypes = [vp, vp, i32, C.c_float, i32, i32, i32, vp, sz, vp]
    l.sb_visualiser_levels.argtypes = [vp, sz, i32, i32, vp, vp]
    l.sb_visualiser_levels_dev.argtypes = [vp, i64, i32, i32, i32, i32, vp, vp]
    l.sb_resampler_create.argtypes = [i32, i32, C.POINTER(vp)]
    l.sb_resampler_destroy.argtypes = [vp]
    l.sb_resample_geometry.argtypes = [vp, sz, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]
    l.sb_resample_dev.argtypes = [vp, vp, i64, sz, i32, vp, i64, vp]
    l.sb_vad_create.argtypes = [vp, sz, C.POINTER(vp)]
    l.sb_vad_destroy.argtypes = [vp]
    l.sb_vad_workspace_bytes.argtypes = [i32, i32]
    l.sb_vad_workspace_bytes.restype = sz
    l.sb_vad_score_dev.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, vp, vp]
    l.sb_vad_gate_workspace_bytes.argtypes = [i32, i32]
    l.sb_vad_gate_workspace_bytes.restype = sz
    l.sb_vad_gate_dev.argtypes = [vp, vp, i64, i32, i32, C.c_float, i32, i32, i32, vp, i64, vp, vp, vp]


class Resampler:
    def __init__(self, fs_in: int, fs_out: int = 16000):
        self._h = C.c_void_p()
        check(lib().sb_resampler_create(fs_in, fs_out, C.byref(self._h)))
        self.fs_in, self.fs_out = fs_in, fs_out

    def geometry(self, n_in: int):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(lib().sb_resample_geometry(self._h, n_in, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def run_dev(self, in_ptr: int, in_stride: int, n_in: int, n_streams: int, out_ptr: int, out_stride: int, stream: int = 0):
        check(lib().sb_resample_dev(self._h, in_ptr, in_stride, n_in, n_streams, out_ptr, out_stride, stream or None))

    def close(self):
        if getattr(self, "_h", None):
            lib().sb_resampler_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Vad:
    def __init__(self, blob: np.ndarray):
        b = np.ascontiguousarray(blob, dtype=np.float32)
        self._h = C.c_void_p()
        check(lib().sb_vad_create(b.ctypes.data, b.size, C.byref(self._h)))

    @staticmethod
    def workspace_bytes(n_streams: int, n_frames: int) -> int:
        return int(lib().sb_vad_workspace_bytes(n_streams, n_frames))

    def score_dev(self, pcm_ptr, pcm_stride, n_streams, n_frames, h_ptr, c_ptr, probs_ptr, ws_ptr, stream=0):
        check(lib().sb_vad_score_dev(self._h, pcm_ptr, pcm_stride, n_streams, n_frames, h_ptr, c_ptr, probs_ptr, ws_ptr,
                                     stream or None))

    def close(self):
        if getattr(self, "_h", None):
            lib().sb_vad_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


SB_SAMPLE_F32, SB_SAMPLE_I16, SB_SAMPLE_U16 = 0, 1, 2


def downmix_mono_dev(in_ptr, sample_format: int, channels: int, in_stride: int, n_frames: int, n_streams: int, out_ptr,
                     out_stride: int, stream=None) -> None:
    check(lib().sb_downmix_mono_dev(in_ptr, sample_format, channels, in_stride, n_frames, n_streams, out_ptr, out_stride,
                                    stream or None))


def pcm_f32_to_i16_dev(in_ptr, out_ptr, n: int, stream=None) -> None:
    check(lib().sb_pcm_f32_to_i16_dev(in_ptr, out_ptr, n, stream or None))


def resample_48k_16k(pcm48k: np.ndarray) -> np.ndarray:
    """Host arrays, one stream: FrameResampler push(all) + finish at 48 -> 16 kHz; returns whole 480-sample frames."""
    x = np.ascontiguousarray(pcm48k, np.float32).reshape(-1)
    out = np.empty(x.shape[0] // 3 + 2048, np.float32)
    n = C.c_size_t(0)
    check(lib().sb_resample_48k_16k(x.ctypes.data, x.shape[0], out.ctypes.data, out.shape[0], C.addressof(n)))
    return out[: n.value].copy()


def silero_v4(vad: "Vad", pcm16k: np.ndarray, h: np.ndarray, c: np.ndarray) -> np.ndarray:
    """Host arrays, one stream: probabilities of the whole 480-sample frames of pcm16k; h, c [2, 64] are updated in place."""
    x = np.ascontiguousarray(pcm16k, np.float32).reshape(-1)
    n_frames = x.shape[0] // 480
    probs = np.empty(n_frames, np.float32)
    assert h.dtype == np.float32 and c.dtype == np.float32 and h.shape == (2, 64) and c.shape == (2, 64)
    check(lib().sb_silero_v4(vad._h, x.ctypes.data, n_frames, h.ctypes.data, c.ctypes.data, probs.ctypes.data))
    return probs


def vad_gate(probs: np.ndarray, pcm16k: np.ndarray, threshold: float, prefill: int, hangover: int, onset: int) -> np.ndarray:
    """Host arrays, one stream: the samples SmoothedVad keeps (concatenated Speech slices)."""
    p = np.ascontiguousarray(probs, np.float32).reshape(-1)
    x = np.ascontiguousarray(pcm16k, np.float32).reshape(-1)
    n_frames = p.shape[0]
    cap = (n_frames * (prefill + 1 + onset) // max(1, onset) + prefill + 1) * 480
    out = np.empty(max(cap, 1), np.float32)
    cnt = C.c_int(0)
    check(lib().sb_vad_gate(p.ctypes.data, x.ctypes.data, n_frames, threshold, prefill, hangover, onset, out.ctypes.data, cap,
                            C.addressof(cnt)))
    return out[: cnt.value * 480].copy()


def pcm_f32_to_i16(samples: np.ndarray) -> np.ndarray:
    """Host arrays: (sample * 32767) as i16 (audio_toolkit/audio/utils.rs:17-20)."""
    x = np.ascontiguousarray(samples, np.float32).reshape(-1)
    out = np.empty(x.shape[0], np.int16)
    check(lib().sb_pcm_f32_to_i16(x.ctypes.data, out.ctypes.data, x.shape[0]))
    return out


def visualiser_levels(pcm: np.ndarray, chunk_len: int, sample_rate: int) -> np.ndarray:
    """Host arrays, one stream: [n_chunks, 16] levels (audio_toolkit/audio/visualizer.rs:84-149)."""
    x = np.ascontiguousarray(pcm, np.float32).reshape(-1)
    n_chunks = x.shape[0] // chunk_len
    out = np.empty((max(n_chunks, 1), 16), np.float32)
    n = C.c_int(0)
    check(lib().sb_visualiser_levels(x.ctypes.data, x.shape[0], chunk_len, sample_rate, out.ctypes.data, C.addressof(n)))
    return out[: n.value]


def visualiser_levels_dev(pcm_ptr, stream_stride: int, n_streams: int, n_chunks: int, chunk_len: int, sample_rate: int,
                          out_ptr, stream=None) -> None:
    check(lib().sb_visualiser_levels_dev(pcm_ptr, stream_stride, n_streams, n_chunks, chunk_len, sample_rate, out_ptr,
                                         stream or None))


def vad_gate_workspace_bytes(n_streams: int, n_frames: int) -> int:
    return int(lib().sb_vad_gate_workspace_bytes(n_streams, n_frames))


def vad_gate_dev(probs_ptr, pcm_ptr, pcm_stride, n_streams, n_frames, threshold, prefill, hangover, onset, out_ptr,
                 out_stride, out_frames_ptr, ws_ptr, stream=0):
    check(lib().sb_vad_gate_dev(probs_ptr, pcm_ptr, pcm_stride, n_streams, n_frames, threshold, prefill, hangover, onset,
                                out_ptr, out_stride, out_frames_ptr, ws_ptr, stream or None))
