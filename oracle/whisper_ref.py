"""Oracle: whisper.cpp encoder / decoder / logits filter / greedy loop.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED.

Restates, in numpy f32 on the CPU, what the reference executes at
``managers/transcription.rs:501-503`` (``whisper_engine.transcribe_samples``) through
transcribe-rs 0.2.3 -> whisper-rs 0.13.2 -> whisper.cpp ``whisper_full_with_state``:

  encode      App. C.2   conv1d+GELU stem, pre-LN blocks, ln_post
  cross_kv    App. C.2   per decoder layer K = Wk enc, V = Wv enc + bv
  decode      App. C.3   KV-cached self-attn, cross-attn, MLP, tied-embedding logits
  filter      App. C.4   whisper_process_logits
  sample      App. C.4   whisper_sample_token(best=true): lowest index wins ties
  full        App. C.4   the seek/window loop with the pinned greedy configuration of
                         SURVEY.md 8(d): language "en", transcribe, timestamps on,
                         suppress_blank, no_context, max_initial_ts 1.0, temperature 0,
                         temperature_inc 0 (NO fallback), n_max = n_text_ctx/2 - 4.

``act_f16=True`` reproduces ggml's CPU rounding points (f16 weights x f16-rounded activation,
f32 accumulate; K/V stored f16; GELU through an f16 table); ``act_f16=False`` is the plain f32
model used as "truth" for stating the bf16 tolerance of the GPU path.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import logmel as _logmel

F32 = np.float32


def _h(x: np.ndarray) -> np.ndarray:
    return x.astype(np.float16).astype(np.float32)


def gelu_tanh(x: np.ndarray) -> np.ndarray:
    x = x.astype(F32)
    c = F32(0.79788456080286535587989211986876)
    return (F32(0.5) * x * (F32(1.0) + np.tanh(c * x * (F32(1.0) + F32(0.044715) * x * x)))).astype(F32)


def gelu_erf(x: np.ndarray) -> np.ndarray:
    """Exact GELU (x/2)(1 + erf(x/sqrt 2)): what OpenAI/HF Whisper use; only for the HF cross-check
    (tests/test_oracle_hf_cpu.py) -- whisper.cpp/ggml uses the tanh form above."""
    from scipy.special import erf
    x = x.astype(np.float64)
    return (0.5 * x * (1.0 + erf(x / np.sqrt(2.0)))).astype(F32)


def layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    x = x.astype(F32)
    mu = x.mean(axis=-1, keepdims=True, dtype=np.float64).astype(F32)
    xc = x - mu
    var = (xc.astype(np.float64) ** 2).mean(axis=-1, keepdims=True).astype(F32)
    return (xc / np.sqrt(var + F32(eps)) * w + b).astype(F32)


def softmax_rows(s: np.ndarray) -> np.ndarray:
    m = s.max(axis=-1, keepdims=True)
    e = np.exp(s - m)
    return (e / e.sum(axis=-1, keepdims=True)).astype(F32)


@dataclass
class DecodeConfig:
    language_id: int = 0          # "en" -> <|en|> = sot + 1 + 0; -1 = auto-detect on the first window (reference default "auto")
    translate: bool = False
    no_timestamps: bool = False
    suppress_blank: bool = True
    max_initial_ts: float = 1.0
    single_segment: bool = False
    n_max_override: Optional[int] = None   # benches/tests may cap decode length
    # Text conditioning [MEM -- oracle/ASSUMPTIONS.md]: whisper.cpp keeps `prompt_past` inside one whisper_full call: the
    # tokens of `initial_prompt` first, then after every window the context that window used + its kept tokens; a window's
    # prompt is [prev] + the last min(n_max_text_ctx, n_text_ctx/2) tokens of it + [sot, lang, task].  `no_context` (true in
    # whisper-rs' defaults) only clears what an EARLIER call left behind.  n_max_text_ctx = 0 switches the prefix off.
    initial_prompt_tokens: Optional[List[int]] = None
    n_max_text_ctx: int = 16384
    # Temperature fallback [MEM]: whisper_full decodes a window at `temperature`; if the decode failed, or the mean
    # log-probability of its kept tokens is below logprob_thold, or (more than 32 kept tokens) the entropy of the last 32
    # is below entropy_thold, it decodes it again at temperature + temperature_inc, ... up to 1.0.  At temperature > 0 the
    # token is drawn with std::discrete_distribution over the decoder's std::mt19937(0) (best_of = 1); from 0.5 up the text
    # context is dropped from the prompt.  temperature_inc = 0 is the pinned parity configuration (no fallback).
    temperature: float = 0.0
    temperature_inc: float = 0.0
    logprob_thold: float = -1.0
    entropy_thold: float = 2.4
    # whisper_full_params.suppress_nst [MEM]: whisper_process_logits sets the tokens of `non_speech_tokens` (with and without a
    # leading space) and " -", " '" to -inf.  whisper.cpp's default is false; transcribe-rs exposes it as suppress_non_speech_tokens.
    suppress_nst: bool = False


# whisper.cpp `non_speech_tokens` = OpenAI tokenizer.non_speech_tokens: symbols, bracket / dash runs, music notes
NON_SPEECH_TOKENS = (list('"#()*+/:;<=>@[\\]^_`{|}~\u300c\u300d\u300e\u300f')
                     + "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} \u266a\u266a \u266a\u266a\u266a".split()
                     + list("\u2669\u266a\u266b\u266c\u266d\u266e\u266f"))


def non_speech_token_ids(vocab) -> List[int]:
    """ids of the listed strings in this vocabulary (id -> bytes; the last id of a duplicate wins, like the loader's map)."""
    t2i = {bytes(w): i for i, w in enumerate(vocab)}
    ids = set()
    for t in NON_SPEECH_TOKENS:
        for cand in (t, " " + t):
            i = t2i.get(cand.encode("utf-8"))
            if i is not None:
                ids.add(i)
    for cand in (b" -", b" '"):
        if cand in t2i:
            ids.add(t2i[cand])
    return sorted(ids)


@dataclass
class WindowResult:
    tokens: List[int]
    result_len: int
    seek_delta: int
    failed: bool
    logits_trace: Optional[List[np.ndarray]] = None  # raw logits at each sampling step
    margins: List[float] = field(default_factory=list)  # top1-top2 of filtered logits
    plogs: List[float] = field(default_factory=list)    # log-probability of every sampled token
    temperature: float = 0.0
    attempts: int = 1
    avg_logprob: float = float("nan")
    # temperature > 0: distance (in probability mass) of the uniform number from the nearer edge of the drawn token's
    # interval of the cumulative distribution -- a draw with a small margin may legitimately differ between two
    # implementations whose logits differ by rounding
    draw_margins: List[float] = field(default_factory=list)


class WhisperOracle:
    def __init__(self, model, act_f16: bool = True, gelu: str = "tanh"):
        assert gelu in ("tanh", "erf") and not (act_f16 and gelu == "erf")
        self.gelu_kind = gelu
        self.m = model
        self.hp = model.hparams
        self.sp = model.special
        self.act_f16 = act_f16
        self.t = {k: v.astype(F32) for k, v in model.tensors.items()}
        self._r = _h if act_f16 else (lambda x: x.astype(F32))
        self.nst_ids = np.asarray(non_speech_token_ids(model.vocab), np.int64)

    # -- building blocks ---------------------------------------------------------------
    def _mm(self, x: np.ndarray, wname: str, bname: Optional[str] = None) -> np.ndarray:
        y = self._r(x) @ self.t[wname].T
        if bname is not None:
            y = y + self.t[bname].reshape(-1)
        return y.astype(F32)

    def _gelu(self, x: np.ndarray) -> np.ndarray:
        if self.act_f16:  # ggml_vec_gelu_f32: f16 table lookup inside (-10, 10)
            y = _h(gelu_tanh(_h(x)))
            y = np.where(x <= -10.0, F32(0), np.where(x >= 10.0, x, y))
            return y.astype(F32)
        return gelu_erf(x) if self.gelu_kind == "erf" else gelu_tanh(x)

    def _heads(self, k, v, n_head):
        """K, V [Tk,d] -> f16-rounded per-head views (K^T [H,dh,Tk], V [H,Tk,dh])."""
        dh = k.shape[1] // n_head
        kh = np.ascontiguousarray(self._r(k).reshape(-1, n_head, dh).transpose(1, 2, 0))
        vh = np.ascontiguousarray(self._r(v).reshape(-1, n_head, dh).transpose(1, 0, 2))
        return kh, vh

    def _mha_pre(self, q, kh, vh):
        """q:[Tq,d], kh:[H,dh,Tk], vh:[H,Tk,dh] -> [Tq,d]; probs rounded before PV (ggml mul_mat)."""
        Tq, d = q.shape
        n_head, dh = kh.shape[0], kh.shape[1]
        qh = self._r(q).reshape(Tq, n_head, dh).transpose(1, 0, 2)
        s = (qh @ kh) * F32(1.0 / np.sqrt(dh))
        p = self._r(softmax_rows(s))
        o = p @ vh
        return o.transpose(1, 0, 2).reshape(Tq, d).astype(F32)

    def _mha(self, q, k, v, n_head, mask=None):
        """q:[Tq,d] k,v:[Tk,d] -> [Tq,d]; K, V f16-stored, probs rounded before PV."""
        assert mask is None
        kh, vh = self._heads(k, v, n_head)
        return self._mha_pre(q, kh, vh)

    # -- encoder -----------------------------------------------------------------------
    def conv_stem(self, mel_win: np.ndarray) -> np.ndarray:
        """mel_win [n_mel, 3000] -> [1500, d] (conv1 k3 s1 p1 + GELU, conv2 k3 s2 p1 + GELU, + pos)."""
        hp = self.hp
        x = self._r(mel_win.T)                                    # [3000, n_mel]
        w1 = self.t["encoder.conv1.weight"]                       # [d, n_mel, 3]
        xp = np.concatenate([np.zeros((1, x.shape[1]), F32), x, np.zeros((1, x.shape[1]), F32)], 0)
        cols = np.concatenate([xp[0:-2], xp[1:-1], xp[2:]], axis=1)          # [3000, 3*n_mel] (k-major)
        wk = w1.transpose(0, 2, 1).reshape(w1.shape[0], -1)                  # [d, 3*n_mel]
        y = cols @ wk.T + self.t["encoder.conv1.bias"].reshape(-1)
        y = self._gelu(y)                                                    # [3000, d]
        w2 = self.t["encoder.conv2.weight"]
        y = self._r(y)
        yp = np.concatenate([np.zeros((1, y.shape[1]), F32), y, np.zeros((1, y.shape[1]), F32)], 0)
        cols2 = np.concatenate([yp[0:-2:2], yp[1:-1:2], yp[2::2]], axis=1)   # [1500, 3*d]
        wk2 = w2.transpose(0, 2, 1).reshape(w2.shape[0], -1)
        z = cols2 @ wk2.T + self.t["encoder.conv2.bias"].reshape(-1)
        z = self._gelu(z)
        return (z + self.t["encoder.positional_embedding"][: hp.n_audio_ctx]).astype(F32)

    def encoder_block(self, x: np.ndarray, i: int) -> np.ndarray:
        p = f"encoder.blocks.{i}"
        h = layer_norm(x, self.t[p + ".attn_ln.weight"], self.t[p + ".attn_ln.bias"])
        q = self._mm(h, p + ".attn.query.weight", p + ".attn.query.bias")
        k = self._mm(h, p + ".attn.key.weight")
        v = self._mm(h, p + ".attn.value.weight", p + ".attn.value.bias")
        a = self._mha(q, k, v, self.hp.n_audio_head)
        x = x + self._mm(a, p + ".attn.out.weight", p + ".attn.out.bias")
        h = layer_norm(x, self.t[p + ".mlp_ln.weight"], self.t[p + ".mlp_ln.bias"])
        h = self._gelu(self._mm(h, p + ".mlp.0.weight", p + ".mlp.0.bias"))
        x = x + self._mm(h, p + ".mlp.2.weight", p + ".mlp.2.bias")
        return x.astype(F32)

    def encode(self, mel_win: np.ndarray, n_layers: Optional[int] = None) -> np.ndarray:
        x = self.conv_stem(mel_win)
        L = self.hp.n_audio_layer if n_layers is None else n_layers
        for i in range(L):
            x = self.encoder_block(x, i)
        return layer_norm(x, self.t["encoder.ln_post.weight"], self.t["encoder.ln_post.bias"])

    def cross_kv(self, enc: np.ndarray):
        out = []
        for i in range(self.hp.n_text_layer):
            p = f"decoder.blocks.{i}.cross_attn"
            k = self._mm(enc, p + ".key.weight")
            v = self._mm(enc, p + ".value.weight", p + ".value.bias")
            out.append(self._heads(k, v, self.hp.n_text_head))
        return out

    # -- decoder -----------------------------------------------------------------------
    def new_kv(self):
        return [([], []) for _ in range(self.hp.n_text_layer)]

    def decode_step(self, token: int, n_past: int, kv_self, kv_cross) -> np.ndarray:
        """One token at position n_past -> f32 logits [n_vocab]."""
        hp = self.hp
        x = (self.t["decoder.token_embedding.weight"][token]
             + self.t["decoder.positional_embedding"][n_past]).astype(F32)[None, :]
        for i in range(hp.n_text_layer):
            p = f"decoder.blocks.{i}"
            h = layer_norm(x, self.t[p + ".attn_ln.weight"], self.t[p + ".attn_ln.bias"])
            q = self._mm(h, p + ".attn.query.weight", p + ".attn.query.bias")
            k = self._mm(h, p + ".attn.key.weight")
            v = self._mm(h, p + ".attn.value.weight", p + ".attn.value.bias")
            ks, vs = kv_self[i]
            del ks[n_past:], vs[n_past:]
            ks.append(self._r(k[0]))
            vs.append(self._r(v[0]))
            a = self._mha(q, np.stack(ks), np.stack(vs), hp.n_text_head)
            x = x + self._mm(a, p + ".attn.out.weight", p + ".attn.out.bias")
            h = layer_norm(x, self.t[p + ".cross_attn_ln.weight"], self.t[p + ".cross_attn_ln.bias"])
            q = self._mm(h, p + ".cross_attn.query.weight", p + ".cross_attn.query.bias")
            kc, vc = kv_cross[i]
            a = self._mha_pre(q, kc, vc)
            x = x + self._mm(a, p + ".cross_attn.out.weight", p + ".cross_attn.out.bias")
            h = layer_norm(x, self.t[p + ".mlp_ln.weight"], self.t[p + ".mlp_ln.bias"])
            h = self._gelu(self._mm(h, p + ".mlp.0.weight", p + ".mlp.0.bias"))
            x = x + self._mm(h, p + ".mlp.2.weight", p + ".mlp.2.bias")
        x = layer_norm(x, self.t["decoder.ln.weight"], self.t["decoder.ln.bias"])
        return (self._r(x) @ self.t["decoder.token_embedding.weight"].T)[0].astype(F32)

    # -- logits filter + sampler (App. C.4) -------------------------------------------
    def process_logits(self, logits: np.ndarray, tokens_cur: List[int], has_ts: bool,
                       seek_delta: int, cfg: DecodeConfig, temperature: float = 0.0):
        sp = self.sp
        n = logits.shape[0]
        lg = logits.astype(F32).copy()
        if temperature > 0.0:
            lg = (lg / F32(temperature)).astype(F32)
        NEG = F32(-np.inf)
        is_initial = len(tokens_cur) == 0
        if cfg.suppress_blank and is_initial:
            lg[sp.eot] = NEG
            lg[sp.blank] = NEG
        if cfg.suppress_nst and self.nst_ids.size:
            lg[self.nst_ids] = NEG
        lg[sp.not_] = NEG
        if cfg.no_timestamps:
            lg[sp.beg:] = NEG
        lg[sp.sot] = NEG
        lg[sp.nosp] = NEG
        lg[sp.solm] = NEG
        lg[sp.translate] = NEG
        lg[sp.transcribe] = NEG
        lg[sp.prev] = NEG
        lg[sp.lang_first: sp.lang_first + sp.num_languages] = NEG
        # timestamps have to appear in pairs, except directly before EOT
        last_ts = len(tokens_cur) > 0 and tokens_cur[-1] >= sp.beg
        pen_ts = len(tokens_cur) < 2 or tokens_cur[-2] >= sp.beg
        if last_ts:
            if pen_ts:
                lg[sp.beg:] = NEG
            else:
                lg[: sp.eot] = NEG
        if is_initial and cfg.max_initial_ts > 0.0:
            precision = 30.0 / self.hp.n_audio_ctx
            tid0 = int(round(cfg.max_initial_ts / precision))
            lg[sp.beg + tid0 + 1:] = NEG
        if has_ts:
            tid0 = seek_delta // 2
            lg[sp.beg: sp.beg + tid0] = NEG
        # log_softmax in f32
        logit_max = lg.max()
        fin = lg > NEG
        lse = F32(np.log(np.exp(lg[fin] - logit_max, dtype=F32).sum(dtype=F32))) + logit_max
        logprobs = np.where(fin, lg - lse, NEG).astype(F32)
        # timestamp-mass rule
        ts_lp = logprobs[sp.beg:]
        tfin = ts_lp > NEG
        timestamp_logprob = NEG
        if tfin.any():
            mx = ts_lp.max()
            s = np.exp(ts_lp[tfin] - mx, dtype=F32).sum(dtype=F32)
            if s > 0:
                timestamp_logprob = F32(np.log(s)) + mx
        max_text = logprobs[: sp.beg].max()
        if timestamp_logprob > max_text:
            lg[: sp.beg] = NEG
            logprobs[: sp.beg] = NEG
        probs = np.where(lg == NEG, F32(0), np.exp(logprobs, dtype=F32)).astype(F32)
        return lg, logprobs, probs

    @staticmethod
    def sample_best(probs: np.ndarray) -> int:
        # strict '<' scan from index 0: first (lowest-index) maximum wins; np.argmax does the same
        return int(np.argmax(probs))

    # -- one 30 s window ---------------------------------------------------------------
    def decode_window(self, enc: np.ndarray, seek: int, seek_end: int, cfg: DecodeConfig,
                      trace: bool = False, forced: Optional[List[int]] = None,
                      prompt_past: Optional[List[int]] = None, temperature: float = 0.0, rng=None) -> WindowResult:
        hp, sp = self.hp, self.sp
        kv_cross = self.cross_kv(enc)
        kv_self = self.new_kv()
        prompt = []
        if prompt_past and cfg.n_max_text_ctx > 0 and temperature < 0.5:
            n_take = min(cfg.n_max_text_ctx, hp.n_text_ctx // 2, len(prompt_past))
            prompt = [sp.prev] + list(prompt_past[len(prompt_past) - n_take:])
        prompt.append(sp.sot)
        if hp.n_vocab >= 51865:
            prompt.append(sp.lang_first + cfg.language_id)
            prompt.append(sp.translate if cfg.translate else sp.transcribe)
        if cfg.no_timestamps:
            prompt.append(sp.not_)
        n_past = 0
        logits = None
        for tok in prompt:
            logits = self.decode_step(tok, n_past, kv_self, kv_cross)
            n_past += 1
        n_max = hp.n_text_ctx // 2 - 4
        if cfg.n_max_override is not None:
            n_max = min(n_max, cfg.n_max_override)
        tokens: List[int] = []
        seek_delta = 100 * 30
        result_len = 0
        has_ts = False
        failed = False
        tr: List[np.ndarray] = []
        margins: List[float] = []
        plogs: List[float] = []
        draw_margins: List[float] = []
        for i in range(n_max):
            lg, logprobs, probs = self.process_logits(logits, tokens, has_ts, seek_delta, cfg, temperature)
            if trace:
                tr.append(logits.copy())
            top2 = np.partition(lg, -2)[-2:]
            margins.append(float(top2[1] - top2[0]))
            if temperature > 0.0 and rng is not None:
                # whisper_sample_token(best = false): std::discrete_distribution over probs -- normalised in double,
                # cumulative sums with the last one pinned to 1.0, first index whose cumulative probability is >= u
                p64 = probs.astype(np.float64)
                cp = np.cumsum(p64 / p64.sum())
                cp[-1] = 1.0
                u = rng.uniform()
                tid = int(np.searchsorted(cp, u, side="left"))
                draw_margins.append(float(min(cp[tid] - u, u - (cp[tid - 1] if tid > 0 else 0.0))))
            else:
                tid = self.sample_best(probs)
            if forced is not None and i < len(forced):
                tid = forced[i]
            tokens.append(tid)
            plogs.append(float(logprobs[tid]))
            # bookkeeping
            completed = False
            if tid > sp.beg:
                sd_new = 2 * (tid - sp.beg)
                if has_ts and seek_delta > sd_new and result_len < i:
                    failed = True
                    break
                seek_delta = sd_new
                result_len = i + 1
                has_ts = True
            if tid == sp.eot or (has_ts and seek + seek_delta + 100 >= seek_end):
                if result_len == 0 and not cfg.no_timestamps:
                    if seek + seek_delta + 100 >= seek_end:
                        result_len = i + 1
                    else:
                        failed = True
                        break
                if cfg.single_segment or cfg.no_timestamps:
                    result_len = i + 1
                    seek_delta = 100 * 30
                completed = True
            if completed:
                break
            if i == n_max - 1 and (result_len == 0 or seek_delta < 100 * 30 // 2):
                failed = True
                break
            logits = self.decode_step(tid, n_past, kv_self, kv_cross)
            n_past += 1
        n = min(result_len, len(tokens))
        avg = float(np.sum(np.asarray(plogs[:n], np.float64)) / n) if n > 0 else float("nan")
        return WindowResult(tokens, result_len, seek_delta, failed, tr if trace else None, margins, plogs, temperature, 1, avg,
                            draw_margins)

    def detect_language(self, enc: np.ndarray):
        """whisper.cpp whisper_lang_auto_detect_with_state (App. C.4 / reference default selected_language
        "auto", settings.rs:427-429): decode [sot] alone, softmax over the language tokens only, pick the
        most probable (lowest id on ties).  Returns (language id, probabilities)."""
        hp, sp = self.hp, self.sp
        assert hp.n_vocab >= 51865, "English-only models have no language tokens"
        logits = self.decode_step(sp.sot, 0, self.new_kv(), self.cross_kv(enc))
        lg = logits[sp.lang_first: sp.lang_first + sp.num_languages].astype(np.float64)
        p = np.exp(lg - lg.max())
        p /= p.sum()
        return int(np.argmax(lg)), p

    # -- whisper_full ------------------------------------------------------------------
    def full(self, samples: np.ndarray, cfg: Optional[DecodeConfig] = None, max_windows: int = 64):
        """Returns (text_bytes, all_tokens_kept, per_window WindowResult list)."""
        cfg = cfg or DecodeConfig()
        mel, n_len_org = _logmel.logmel_f32_faithful(samples, self.m.mel_filters)
        return self.full_from_mel(mel, n_len_org, cfg, max_windows)

    def full_from_mel(self, mel, n_len_org, cfg, max_windows: int = 64):
        seek = 0
        seek_end = n_len_org
        windows: List[WindowResult] = []
        kept: List[int] = []
        text = b""
        prompt_past: List[int] = list(cfg.initial_prompt_tokens or [])
        from .mt19937 import Mt19937
        rng = Mt19937(0)
        if seek_end < seek + 100:
            return b"", kept, windows
        while len(windows) < max_windows:
            if seek + 100 >= seek_end:
                break
            enc = self.encode(_logmel.mel_window(mel, seek, self.hp.n_audio_ctx))
            if cfg.language_id < 0:                       # whisper_full detects once, at offset 0, before the seek loop
                lang, _ = self.detect_language(enc) if self.hp.n_vocab >= 51865 else (0, None)
                cfg = dataclasses.replace(cfg, language_id=lang)
                self.last_detected_language = lang
            if seek > 0 and seek + 500 >= seek_end:      # a very short tail: whisper.cpp drops the text context
                prompt_past = []
            # temperature schedule of whisper_full (one decoder, best_of = 1)
            t_cur, attempts = max(0.0, cfg.temperature), 0
            while True:
                w = self.decode_window(enc, seek, seek_end, cfg, prompt_past=prompt_past, temperature=t_cur, rng=rng)
                attempts += 1
                n = min(w.result_len, len(w.tokens))
                failed = w.failed
                if n > 32:
                    _, counts = np.unique(np.asarray(w.tokens[n - 32:n]), return_counts=True)
                    pr = counts / 32.0
                    if float(-(pr * np.log(pr)).sum()) < cfg.entropy_thold:
                        failed = True
                success = not (failed or w.avg_logprob < cfg.logprob_thold)
                t_next = np.float32(t_cur) + np.float32(cfg.temperature_inc)
                if success or cfg.temperature_inc <= 0.0 or not (t_next < 1.0 + 1e-6):
                    break
                t_cur = float(t_next)
            w.attempts = attempts
            windows.append(w)
            toks = w.tokens[: w.result_len]
            # prompt_past = the part of it this window used + this window's kept tokens
            n_take = min(cfg.n_max_text_ctx, self.hp.n_text_ctx // 2, len(prompt_past)) \
                if (cfg.n_max_text_ctx > 0 and w.temperature < 0.5) else 0
            prompt_past = list(prompt_past[len(prompt_past) - n_take:]) + list(toks)
            kept.extend(toks)
            for t in toks:
                if t < self.sp.eot:
                    text += self.m.token_bytes(t)
            seek += w.seek_delta
        return text.strip(), kept, windows
