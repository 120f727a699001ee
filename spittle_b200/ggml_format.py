"""GGML *legacy* Whisper model file (``ggml-*.bin``, magic 0x67676d6c) reader/writer.

This is the on-disk format the reference loads through
``WhisperEngine::load_model(&path)`` (src-tauri/src/managers/transcription.rs:262-263;
custom ``*.bin`` discovery in src-tauri/src/managers/model.rs:267-382).  The format is
defined by whisper.cpp (bundled by whisper-rs-sys 0.11.1, not vendored in the reference)
and restated in SURVEY.md Appendix D:

  u32 magic | 11 x i32 hparams | mel filters (i32 n_mel, i32 n_fft, f32 data) |
  vocab (i32 n, n x (u32 len, bytes)) | tensors until EOF:
      i32 n_dims, i32 name_len, i32 ttype, i32 ne[n_dims] (fastest-varying first),
      name bytes, raw data

The C++ loader in ``csrc/ggml_loader.cpp`` reads the same bytes; this module exists so the
synthetic models of SURVEY.md 8(d) can be written once and loaded bit-identically by the
oracle, the CPU baseline and the GPU engine.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field, asdict
from typing import Dict, List

import numpy as np

GGML_MAGIC = 0x67676D6C
GGML_TYPE_F32 = 0
GGML_TYPE_F16 = 1
# ggml block-quantised tensor types (32 weights per block; SURVEY 8(f) N2: the reference's catalog ships
# Medium as q4_1 and Large-v3 as q5_0, resources/model_catalog.json:157,187).  Layouts per ggml [MEM]:
#   q4_0 {f16 d; u8 qs[16]}               x = (q - 8) d        q4_1 {f16 d, m; u8 qs[16]}          x = q d + m
#   q5_0 {f16 d; u8 qh[4]; u8 qs[16]}     x = (q - 16) d       q5_1 {f16 d, m; u8 qh[4]; u8 qs[16]} x = q d + m
#   q8_0 {f16 d; i8 qs[32]}               x = q d
# element j < 16 takes the low nibble of qs[j], element j + 16 the high nibble; the fifth bit of q5 comes from
# bit j (low half) / bit j + 16 (high half) of the little-endian u32 qh.
GGML_TYPE_Q4_0, GGML_TYPE_Q4_1, GGML_TYPE_Q5_0, GGML_TYPE_Q5_1, GGML_TYPE_Q8_0 = 2, 3, 6, 7, 8
# k-quant q5_K (catalog: breeze-asr-q5_k.bin, resources/model_catalog.json:202): super-blocks of 256 weights, 176 bytes
#   {f16 d, dmin; u8 scales[12]; u8 qh[32]; u8 qs[128]}: eight sub-blocks of 32 with 6-bit scale sc_j and minimum mn_j
#   (packed by get_scale_min_k4), x = d sc_j q - dmin mn_j, q in [0, 32): low four bits in the nibbles of qs (64 weights per
#   32 bytes: sub-block 2g in the low nibbles, 2g + 1 in the high ones), fifth bit in bit 2g / 2g + 1 of qh[l].  [MEM]
GGML_TYPE_Q5_K = 13
QK_K, Q5_K_BYTES = 256, 176
QUANT_BLOCK_BYTES = {GGML_TYPE_Q4_0: 18, GGML_TYPE_Q4_1: 20, GGML_TYPE_Q5_0: 22, GGML_TYPE_Q5_1: 24, GGML_TYPE_Q8_0: 34}
GGML_FTYPE_OF_TYPE = {GGML_TYPE_Q4_0: 2, GGML_TYPE_Q4_1: 3, GGML_TYPE_Q8_0: 7, GGML_TYPE_Q5_0: 8, GGML_TYPE_Q5_1: 9, GGML_TYPE_Q5_K: 13}


def quantize_q5_k(x: np.ndarray) -> bytes:
    """A valid q5_K encoding of a tensor whose last dim is a multiple of 256 (per sub-block min / max fit; ggml's own
    quantiser searches for better scales, the FORMAT is what matters to the loader)."""
    v = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 8, 32)
    nb = v.shape[0]
    mn = np.minimum(v.min(axis=2), 0.0)
    scale = (v.max(axis=2) - mn) / 31.0
    d = (scale.max(axis=1) / 63.0).astype("<f2")
    dmin = ((-mn).max(axis=1) / 63.0).astype("<f2")
    df, dminf = d.astype(np.float32), dmin.astype(np.float32)
    sc = np.where(df[:, None] > 0, np.rint(scale / np.where(df[:, None] > 0, df[:, None], 1)), 0).clip(0, 63).astype(np.uint8)
    mq = np.where(dminf[:, None] > 0, np.rint(-mn / np.where(dminf[:, None] > 0, dminf[:, None], 1)), 0).clip(0, 63).astype(np.uint8)
    eff = df[:, None] * sc
    q = np.where(eff[:, :, None] > 0, np.rint((v + (dminf[:, None] * mq)[:, :, None]) / np.where(eff[:, :, None] > 0, eff[:, :, None], 1)), 0)
    q = q.clip(0, 31).astype(np.uint8)
    scales = np.zeros((nb, 12), np.uint8)
    for j in range(4):
        scales[:, j] = sc[:, j] & 63
        scales[:, j + 4] = mq[:, j] & 63
    for j in range(4, 8):
        scales[:, j + 4] = (sc[:, j] & 0xF) | ((mq[:, j] & 0xF) << 4)
        scales[:, j - 4] |= (sc[:, j] >> 4) << 6
        scales[:, j] |= (mq[:, j] >> 4) << 6
    qs = np.zeros((nb, 128), np.uint8)
    qh = np.zeros((nb, 32), np.uint8)
    for g in range(4):
        lo, hi = q[:, 2 * g], q[:, 2 * g + 1]
        qs[:, 32 * g: 32 * g + 32] = (lo & 0xF) | ((hi & 0xF) << 4)
        qh |= ((lo >> 4) & 1) << (2 * g)
        qh |= ((hi >> 4) & 1) << (2 * g + 1)
    return np.concatenate([d.view(np.uint8).reshape(nb, 2), dmin.view(np.uint8).reshape(nb, 2), scales, qh, qs], axis=1).tobytes()


def dequantize_q5_k(raw: bytes, count: int) -> np.ndarray:
    """ggml dequantize_row_q5_K."""
    nb = count // QK_K
    b = np.frombuffer(raw, np.uint8, nb * Q5_K_BYTES).reshape(nb, Q5_K_BYTES)
    d = b[:, 0:2].copy().view("<f2").astype(np.float32).reshape(nb)
    dmin = b[:, 2:4].copy().view("<f2").astype(np.float32).reshape(nb)
    scales, qh, qs = b[:, 4:16].astype(np.int32), b[:, 16:48].astype(np.int32), b[:, 48:176].astype(np.int32)
    out = np.zeros((nb, 8, 32), np.float32)
    for j in range(8):
        if j < 4:
            sc, m = scales[:, j] & 63, scales[:, j + 4] & 63
        else:
            sc = (scales[:, j + 4] & 0xF) | ((scales[:, j - 4] >> 6) << 4)
            m = (scales[:, j + 4] >> 4) | ((scales[:, j] >> 6) << 4)
        g = j // 2
        nib = (qs[:, 32 * g: 32 * g + 32] & 0xF) if j % 2 == 0 else (qs[:, 32 * g: 32 * g + 32] >> 4)
        q = nib + (((qh >> j) & 1) << 4)
        out[:, j] = ((d * sc.astype(np.float32))[:, None] * q.astype(np.float32) - (dmin * m.astype(np.float32))[:, None]).astype(np.float32)
    return out.reshape(-1)


def quantize_blocks(x: np.ndarray, ttype: int) -> bytes:
    """ggml reference quantisation (quantize_row_q*_reference) of a tensor whose last dim is a multiple of 32."""
    if ttype == GGML_TYPE_Q5_K:
        return quantize_q5_k(x)
    v = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 32)
    nb = v.shape[0]
    if ttype == GGML_TYPE_Q8_0:
        amax = np.abs(v).max(axis=1)
        d = (amax / 127.0).astype(np.float32)
        idv = np.where(d > 0, 1.0 / np.where(d > 0, d, 1), 0.0).astype(np.float32)
        q = np.rint(v * idv[:, None]).astype(np.int8)
        out = np.zeros((nb, 34), np.uint8)
        out[:, 0:2] = d.astype("<f2").view(np.uint8).reshape(nb, 2)
        out[:, 2:] = q.view(np.uint8)
        return out.tobytes()
    five = ttype in (GGML_TYPE_Q5_0, GGML_TYPE_Q5_1)
    levels = 32 if five else 16
    if ttype in (GGML_TYPE_Q4_0, GGML_TYPE_Q5_0):
        idx = np.abs(v).argmax(axis=1)
        mx = v[np.arange(nb), idx]                                   # signed value of largest magnitude
        d = (mx / -(levels // 2)).astype(np.float32)
        idv = np.where(d != 0, 1.0 / np.where(d != 0, d, 1), 0.0).astype(np.float32)
        q = np.minimum(levels - 1, (v * idv[:, None] + (levels // 2 + 0.5)).astype(np.int32)).astype(np.uint8)
        m = None
    else:
        mn, mxv = v.min(axis=1), v.max(axis=1)
        d = ((mxv - mn) / (levels - 1)).astype(np.float32)
        idv = np.where(d != 0, 1.0 / np.where(d != 0, d, 1), 0.0).astype(np.float32)
        q = np.minimum(levels - 1, ((v - mn[:, None]) * idv[:, None] + 0.5).astype(np.int32)).astype(np.uint8)
        m = mn.astype(np.float32)
    lo, hi = q[:, :16], q[:, 16:]
    qs = ((lo & 0x0F) | ((hi & 0x0F) << 4)).astype(np.uint8)
    parts = [d.astype("<f2").view(np.uint8).reshape(nb, 2)]
    if m is not None:
        parts.append(m.astype("<f2").view(np.uint8).reshape(nb, 2))
    if five:
        qh = np.zeros(nb, np.uint32)
        for j in range(16):
            qh |= ((lo[:, j].astype(np.uint32) >> 4) & 1) << j
            qh |= ((hi[:, j].astype(np.uint32) >> 4) & 1) << (j + 16)
        parts.append(qh.astype("<u4").view(np.uint8).reshape(nb, 4))
    parts.append(qs)
    return np.concatenate(parts, axis=1).tobytes()


def dequantize_blocks(raw: bytes, ttype: int, count: int) -> np.ndarray:
    if ttype == GGML_TYPE_Q5_K:
        return dequantize_q5_k(raw, count)
    bs = QUANT_BLOCK_BYTES[ttype]
    nb = count // 32
    b = np.frombuffer(raw, np.uint8, nb * bs).reshape(nb, bs)
    d = b[:, 0:2].copy().view("<f2").astype(np.float32)             # [nb, 1]
    off = 2
    m = None
    if ttype in (GGML_TYPE_Q4_1, GGML_TYPE_Q5_1):
        m = b[:, 2:4].copy().view("<f2").astype(np.float32)
        off = 4
    if ttype == GGML_TYPE_Q8_0:
        q = b[:, 2:].copy().view(np.int8).astype(np.float32)
        return (q * d).astype(np.float32).reshape(-1)
    qh = None
    if ttype in (GGML_TYPE_Q5_0, GGML_TYPE_Q5_1):
        qh = b[:, off:off + 4].copy().view("<u4").reshape(nb)
        off += 4
    qs = b[:, off:off + 16]
    lo = (qs & 0x0F).astype(np.int32)
    hi = (qs >> 4).astype(np.int32)
    if qh is not None:
        j = np.arange(16)
        lo |= (((qh[:, None] >> j) & 1) << 4).astype(np.int32)
        hi |= (((qh[:, None] >> (j + 16)) & 1) << 4).astype(np.int32)
    q = np.concatenate([lo, hi], axis=1).astype(np.float32)
    if ttype == GGML_TYPE_Q4_0:
        return ((q - 8.0) * d).astype(np.float32).reshape(-1)
    if ttype == GGML_TYPE_Q5_0:
        return ((q - 16.0) * d).astype(np.float32).reshape(-1)
    return (q * d + m).astype(np.float32).reshape(-1)


@dataclass
class WhisperHParams:
    n_vocab: int = 51865
    n_audio_ctx: int = 1500
    n_audio_state: int = 768
    n_audio_head: int = 12
    n_audio_layer: int = 12
    n_text_ctx: int = 448
    n_text_state: int = 768
    n_text_head: int = 12
    n_text_layer: int = 12
    n_mels: int = 80
    ftype: int = 1

    def as_list(self) -> List[int]:
        return [self.n_vocab, self.n_audio_ctx, self.n_audio_state, self.n_audio_head,
                self.n_audio_layer, self.n_text_ctx, self.n_text_state, self.n_text_head,
                self.n_text_layer, self.n_mels, self.ftype]


# Named architectures (SURVEY.md 8(d) "Synthetic weights", Appendix E).
ARCHS: Dict[str, WhisperHParams] = {
    # fast CPU/GPU test shape: same structure, d_head = 64
    "nano": WhisperHParams(n_audio_state=128, n_audio_head=2, n_audio_layer=2,
                           n_text_state=128, n_text_head=2, n_text_layer=2, n_mels=80),
    # English-only vocabulary (ggml-*.en.bin): no language / task tokens in the prompt, ids shifted down (App. C.5)
    "nano.en": WhisperHParams(n_vocab=51864, n_audio_state=128, n_audio_head=2, n_audio_layer=2,
                              n_text_state=128, n_text_head=2, n_text_layer=2, n_mels=80),
    "micro": WhisperHParams(n_audio_state=256, n_audio_head=4, n_audio_layer=3,
                            n_text_state=256, n_text_head=4, n_text_layer=3, n_mels=128,
                            n_vocab=51866),
    "small": WhisperHParams(),
    "large-v3": WhisperHParams(n_vocab=51866, n_audio_state=1280, n_audio_head=20,
                               n_audio_layer=32, n_text_state=1280, n_text_head=20,
                               n_text_layer=32, n_mels=128),
    "large-v3-turbo": WhisperHParams(n_vocab=51866, n_audio_state=1280, n_audio_head=20,
                                     n_audio_layer=32, n_text_state=1280, n_text_head=20,
                                     n_text_layer=4, n_mels=128),
}


@dataclass
class SpecialTokens:
    """Special token ids derived from n_vocab exactly as whisper.cpp does."""
    eot: int
    sot: int
    translate: int
    transcribe: int
    solm: int
    prev: int
    nosp: int
    not_: int
    beg: int
    lang_first: int
    num_languages: int
    blank: int  # id of the token whose text is " " (suppress_blank)

    @staticmethod
    def from_n_vocab(n_vocab: int, blank: int = 220) -> "SpecialTokens":
        eot, sot, translate, transcribe, solm, prev, nosp, not_, beg = (
            50256, 50257, 50357, 50358, 50359, 50360, 50361, 50362, 50363)
        # whisper.cpp: num_languages = n_vocab - 51765 - (multilingual ? 1 : 0): English-only vocabularies keep the 99
        # language ids 50258..50356, which the logits filter always suppresses
        num_languages = max(0, n_vocab - 51765)
        if n_vocab >= 51865:  # multilingual
            num_languages = n_vocab - 51765 - 1
            eot += 1
            sot += 1
            dt = num_languages - 98
            translate += dt
            transcribe += dt
            solm += dt
            prev += dt
            nosp += dt
            not_ += dt
            beg += dt
        return SpecialTokens(eot, sot, translate, transcribe, solm, prev, nosp, not_, beg,
                             lang_first=sot + 1, num_languages=num_languages, blank=blank)


@dataclass
class GgmlModel:
    hparams: WhisperHParams
    mel_filters: np.ndarray                      # [n_mels, 201] f32
    vocab: List[bytes]                           # id -> bytes (entries present in the file)
    tensors: Dict[str, np.ndarray] = field(default_factory=dict)  # torch-shaped, f32 or f16

    @property
    def special(self) -> SpecialTokens:
        blank = 220
        for i, w in enumerate(self.vocab):
            if w == b" ":
                blank = i
                break
        return SpecialTokens.from_n_vocab(self.hparams.n_vocab, blank)

    def token_bytes(self, tid: int) -> bytes:
        """whisper_token_to_str: file vocab, else the synthesised special names."""
        if tid < len(self.vocab):
            return self.vocab[tid]
        sp = self.special
        names = {sp.eot: b"[_EOT_]", sp.sot: b"[_SOT_]", sp.translate: b"[_TRANSLATE_]",
                 sp.transcribe: b"[_TRANSCRIBE_]", sp.solm: b"[_SOLM_]", sp.prev: b"[_PREV_]",
                 sp.nosp: b"[_NOSP_]", sp.not_: b"[_NOT_]", sp.beg: b"[_BEG_]"}
        if tid in names:
            return names[tid]
        if tid > sp.beg:
            return b"[_TT_%d]" % (tid - sp.beg)
        if sp.lang_first <= tid < sp.lang_first + sp.num_languages:
            return b"[_LANG_%d]" % (tid - sp.lang_first)
        return b"[_extra_token_%d]" % tid


def write_ggml(path: str, model: GgmlModel, quant_type: int = None) -> None:
    """quant_type: store every 2-D f16 weight matrix in that ggml block type (what whisper.cpp's `quantize` tool does:
    matrices are quantised, 1-D tensors, conv kernels and positional embeddings keep their float type)."""
    hp = model.hparams
    if quant_type is not None:
        hp = WhisperHParams(**{**asdict(hp), "ftype": GGML_FTYPE_OF_TYPE[quant_type]})
    with open(path, "wb") as f:
        f.write(struct.pack("<I", GGML_MAGIC))
        f.write(struct.pack("<11i", *hp.as_list()))
        mf = np.ascontiguousarray(model.mel_filters, dtype=np.float32)
        assert mf.shape == (hp.n_mels, 201)
        f.write(struct.pack("<ii", mf.shape[0], mf.shape[1]))
        f.write(mf.tobytes())
        f.write(struct.pack("<i", len(model.vocab)))
        for w in model.vocab:
            f.write(struct.pack("<I", len(w)))
            f.write(w)
        for name, arr in model.tensors.items():
            if arr.dtype == np.float32:
                ttype = GGML_TYPE_F32
            elif arr.dtype == np.float16:
                ttype = GGML_TYPE_F16
            else:
                raise ValueError(f"{name}: unsupported dtype {arr.dtype}")
            payload = None
            if (quant_type is not None and arr.ndim == 2 and arr.dtype == np.float16
                    and arr.shape[1] % (QK_K if quant_type == GGML_TYPE_Q5_K else 32) == 0 and "positional_embedding" not in name):
                ttype = quant_type
                payload = quantize_blocks(arr.astype(np.float32), quant_type)
            nb = name.encode()
            ne = list(reversed(arr.shape))  # fastest-varying first
            f.write(struct.pack("<iii", len(ne), len(nb), ttype))
            f.write(struct.pack("<%di" % len(ne), *ne))
            f.write(nb)
            f.write(payload if payload is not None else np.ascontiguousarray(arr).tobytes())


def read_ggml(path: str) -> GgmlModel:
    with open(path, "rb") as f:
        data = f.read()
    off = 0

    def take(fmt):
        nonlocal off
        v = struct.unpack_from(fmt, data, off)
        off += struct.calcsize(fmt)
        return v

    (magic,) = take("<I")
    if magic != GGML_MAGIC:
        raise ValueError("bad magic 0x%08x (not a GGML legacy whisper file)" % magic)
    hp = WhisperHParams(*take("<11i"))
    n_mel, n_fft = take("<ii")
    mel = np.frombuffer(data, dtype="<f4", count=n_mel * n_fft, offset=off).reshape(n_mel, n_fft).copy()
    off += 4 * n_mel * n_fft
    (nv,) = take("<i")
    vocab = []
    for _ in range(nv):
        (ln,) = take("<I")
        vocab.append(bytes(data[off:off + ln]))
        off += ln
    tensors: Dict[str, np.ndarray] = {}
    while off < len(data):
        n_dims, name_len, ttype = take("<iii")
        ne = take("<%di" % n_dims)
        name = data[off:off + name_len].decode()
        off += name_len
        shape = tuple(reversed(ne))
        count = int(np.prod(shape))
        if ttype == GGML_TYPE_F32:
            arr = np.frombuffer(data, dtype="<f4", count=count, offset=off)
            off += 4 * count
        elif ttype == GGML_TYPE_F16:
            arr = np.frombuffer(data, dtype="<f2", count=count, offset=off)
            off += 2 * count
        elif ttype in QUANT_BLOCK_BYTES or ttype == GGML_TYPE_Q5_K:
            nbytes = count // QK_K * Q5_K_BYTES if ttype == GGML_TYPE_Q5_K else count // 32 * QUANT_BLOCK_BYTES[ttype]
            arr = dequantize_blocks(data[off:off + nbytes], ttype, count)        # f32, like the engine's loader
            off += nbytes
        else:
            raise ValueError(f"{name}: tensor type {ttype} (k-quants other than q5_K) not supported")
        tensors[name] = arr.reshape(shape).copy()
    return GgmlModel(hp, mel, vocab, tensors)


def hparams_dict(hp: WhisperHParams) -> dict:
    return asdict(hp)
