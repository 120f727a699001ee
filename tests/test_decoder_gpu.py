"""GPU parity: decoder logits (teacher forced) and greedy token sequences vs the oracle."""
import numpy as np
import pytest

from oracle import logmel, whisper_ref
from spittle_b200 import capi, synth, ggml_format

pytestmark = pytest.mark.gpu

# Stated logit tolerances (raw logits have std ~ 5 with the "sharp" recipe) = 1.5 x the error measured on B200 in round 2
# (nano: f16 8.5e-2, bf16 6.1e-1 over 3 windows x 20 teacher-forced steps; Small / Turbo f16: 2.0e-2, test_parity_sizes_gpu.py)
LOGIT_TOL = {capi.SB_DTYPE_F16: 0.13, capi.SB_DTYPE_BF16: 0.92}
# A token mismatch is only acceptable where the oracle's own top-1/top-2 margin is below this (two logits move against
# each other: 1.5 x the logit tolerance would be the worst case; the largest margin at a measured divergence was 0.060 / f16):
MARGIN_TOL = {capi.SB_DTYPE_F16: 0.15, capi.SB_DTYPE_BF16: 1.4}


def _setup(model_dir, arch, dtype, clip_ids, n_steps):
    path = synth.ensure_model_file(arch, model_dir)
    model = ggml_format.read_ggml(path)
    oracle = whisper_ref.WhisperOracle(model, act_f16=(dtype == capi.SB_DTYPE_F16))
    mels, ends, traces = [], [], []
    for i in clip_ids:
        x = synth.make_clip(i, seconds=30.0)
        mel, n_len_org = logmel.logmel_f64(x, model.mel_filters)
        win = logmel.mel_window(mel, 0)
        enc = oracle.encode(win)
        w = oracle.decode_window(enc, 0, n_len_org, whisper_ref.DecodeConfig(n_max_override=n_steps), trace=True)
        mels.append(win); ends.append(n_len_org); traces.append(w)
    return path, model, oracle, np.stack(mels), ends, traces


@pytest.mark.parametrize("dtype", [capi.SB_DTYPE_F16, capi.SB_DTYPE_BF16])
def test_teacher_forced_logits_match_oracle(cuda_dev, model_dir, dtype):
    n_steps = 20
    path, model, oracle, mels, ends, traces = _setup(model_dir, "nano", dtype, [1, 2, 5], n_steps)
    eng = capi.Engine(path, dtype=dtype, max_batch=4)
    forced = np.full((len(traces), n_steps), -1, np.int32)
    for w, tr in enumerate(traces):
        forced[w, :len(tr.tokens)] = tr.tokens
    logits, toks, marg = eng.decode_trace(mels, ends, n_steps, forced=forced)
    for w, tr in enumerate(traces):
        for s in range(len(tr.tokens)):
            ref = tr.logits_trace[s]
            err = float(np.abs(logits[w, s] - ref).max())
            assert err <= LOGIT_TOL[dtype], (w, s, err)
            # argmax of the *filtered* logits must agree wherever the oracle margin is decisive
            if tr.margins[s] > 1.5 * LOGIT_TOL[dtype]:
                assert toks[w, s] == tr.tokens[s]
        print(f"dtype={dtype} window {w}: {len(tr.tokens)} steps, max logit err "
              f"{max(float(np.abs(logits[w, s] - tr.logits_trace[s]).max()) for s in range(len(tr.tokens))):.3e}")
    eng.close()


@pytest.mark.parametrize("dtype", [capi.SB_DTYPE_F16, capi.SB_DTYPE_BF16])
@pytest.mark.parametrize("graph", [False, True])
def test_free_running_tokens_match_oracle(cuda_dev, model_dir, dtype, graph):
    n_steps = 32
    path, model, oracle, mels, ends, traces = _setup(model_dir, "nano", dtype, [1, 2, 3, 5], n_steps)
    eng = capi.Engine(path, dtype=dtype, max_batch=4, use_cuda_graph=graph)
    _, toks, marg = eng.decode_trace(mels, ends, n_steps, forced=None, want_logits=not graph)
    exact = 0
    for w, tr in enumerate(traces):
        n = len(tr.tokens)
        got = list(toks[w, :n])
        if got == tr.tokens:
            exact += 1
            continue
        first = next(i for i in range(n) if got[i] != tr.tokens[i])
        print(f"dtype={dtype} window {w}: first divergence at step {first}, oracle margin {tr.margins[first]:.4f}")
        assert tr.margins[first] < MARGIN_TOL[dtype], (w, first, tr.margins[first])
    print(f"dtype={dtype} graph={graph}: {exact}/{len(traces)} windows token-exact over {n_steps} steps")
    # every divergence above was checked against MARGIN_TOL; on top of that most windows must be identical
    if dtype == capi.SB_DTYPE_F16:
        assert exact >= (3 * len(traces)) // 4
    eng.close()


@pytest.mark.parametrize("dtype", [capi.SB_DTYPE_F16])
def test_transcribe_matches_oracle_full(cuda_dev, model_dir, dtype):
    """whole path through sb_transcribe_batch: log-mel -> seek loop -> text, ragged clip lengths,
    empty and sub-second inputs (reference: transcription.rs:412-416, managers/audio.rs:466-475)."""
    path = synth.ensure_model_file("nano", model_dir)
    model = ggml_format.read_ggml(path)
    oracle = whisper_ref.WhisperOracle(model, act_f16=True)
    eng = capi.Engine(path, dtype=dtype, max_batch=8)
    clips = [synth.make_clip(1, 30.0), synth.make_clip(2, 7.3), np.zeros(0, np.float32), synth.make_clip(3, 0.5),
             synth.make_clip(5, 30.0), synth.make_clip(4, 12.0)]
    params = capi.default_params(n_max_tokens=24, max_windows=3)
    res = eng.transcribe_batch(clips, params)
    assert res[2].text == b"" and res[2].windows == []
    assert res[3].text == b"" and res[3].windows == []          # < 1 s: whisper.cpp returns nothing
    cfg = whisper_ref.DecodeConfig(n_max_override=24)
    n_exact = 0
    for i in (0, 1, 4, 5):
        text, kept, wins = oracle.full(clips[i], cfg, max_windows=3)
        r = res[i]
        ok = True
        for wi, (w_ref, w_got) in enumerate(zip(wins, r.windows)):
            got = r.sampled[w_got["token_offset"]: w_got["token_offset"] + w_got["n_tokens"]]
            if got != w_ref.tokens:
                first = next((k for k in range(min(len(got), len(w_ref.tokens))) if got[k] != w_ref.tokens[k]), None)
                assert first is not None and w_ref.margins[first] < MARGIN_TOL[dtype], (i, wi, first)
                ok = False
                break
            assert w_got["result_len"] == w_ref.result_len and w_got["seek_delta"] == w_ref.seek_delta
            assert bool(w_got["failed"]) == w_ref.failed
        if ok:
            assert len(r.windows) == len(wins)
            assert r.tokens == kept and r.text == text
            n_exact += 1
    print(f"transcribe: {n_exact}/4 clips identical to the oracle end to end")
    assert n_exact >= 3
    # single-clip API routes a batch of one
    one = eng.transcribe(clips[1], params)
    assert one.text == res[1].text and one.tokens == res[1].tokens
    eng.close()


def test_language_auto_detect_matches_oracle(cuda_dev, model_dir):
    """params.language = NULL (the reference's default selected_language "auto", settings.rs:427-429):
    whisper_full detects the language from the first window ([sot] step, arg-max over the language tokens)
    and uses it for every window.  Detected ids and the resulting tokens must match the oracle."""
    path = synth.ensure_model_file("nano", model_dir)
    model = ggml_format.read_ggml(path)
    oracle = whisper_ref.WhisperOracle(model, act_f16=True)
    eng = capi.Engine(path, dtype=capi.SB_DTYPE_F16, max_batch=8)
    clips = [synth.make_clip(i, s) for i, s in ((1, 30.0), (2, 9.0), (5, 30.0), (4, 12.0))]
    params = capi.default_params(n_max_tokens=24, max_windows=2)
    params.language = None
    res = eng.transcribe_batch(clips, params)
    n_exact = 0
    for x, r in zip(clips, res):
        mel, n_len_org = logmel.logmel_f32_faithful(x, model.mel_filters)
        enc = oracle.encode(logmel.mel_window(mel, 0, oracle.hp.n_audio_ctx))
        lang, probs = oracle.detect_language(enc)
        top2 = np.sort(probs)[-2:]
        if r.lang_id != lang:
            assert top2[1] / top2[0] < 1.05, ("language mismatch at a decisive probability ratio", r.lang_id, lang, top2)
            continue
        text, kept, wins = oracle.full(x, whisper_ref.DecodeConfig(language_id=-1, n_max_override=24), max_windows=2)
        got0 = r.sampled[: r.windows[0]["n_tokens"]]
        if got0 == wins[0].tokens:
            n_exact += 1
        else:
            first = next(k for k in range(min(len(got0), len(wins[0].tokens))) if got0[k] != wins[0].tokens[k])
            assert wins[0].margins[first] < MARGIN_TOL[capi.SB_DTYPE_F16]
    assert n_exact >= 2
    # an explicit language still wins over detection, "auto" spelled out behaves like NULL
    params.language = b"auto"
    res2 = eng.transcribe_batch(clips[:2], params)
    assert [r.lang_id for r in res2] == [r.lang_id for r in res[:2]] and res2[0].sampled == res[0].sampled
    params.language = b"de"
    assert eng.transcribe(clips[1], params).lang_id == 2          # whisper language table: en, zh, de, ...
    eng.close()


@pytest.mark.parametrize("qt,arch", [(3, "nano"), (6, "nano"), (13, "micro")])   # q4_1 (catalog: Medium), q5_0 (Large-v3), q5_K (Breeze)
def test_quantised_model_file_matches_oracle(cuda_dev, model_dir, qt, arch):
    """block-quantised GGML files load (dequantised on load, then stored in the engine's 16-bit
    operand type); the oracle runs on the same dequantised weights."""
    path = synth.ensure_model_file(arch, model_dir, quant_type=qt)
    model = ggml_format.read_ggml(path)
    oracle = whisper_ref.WhisperOracle(model, act_f16=True)
    eng = capi.Engine(path, dtype=capi.SB_DTYPE_F16, max_batch=4)
    x = synth.make_clip(1, 30.0)
    mel, n_len_org = logmel.logmel_f64(x, model.mel_filters)
    win = logmel.mel_window(mel, 0)
    enc = oracle.encode(win)
    got = eng.encode(win[None])[0]
    rel = float(np.sqrt(((got - enc) ** 2).mean()) / np.sqrt((enc ** 2).mean()))
    assert rel < 2e-3, rel
    tr = oracle.decode_window(enc, 0, n_len_org, whisper_ref.DecodeConfig(n_max_override=12), trace=True)
    forced = np.full((1, 12), -1, np.int32)
    forced[0, :len(tr.tokens)] = tr.tokens
    logits, toks, _ = eng.decode_trace(win[None], [n_len_org], 12, forced=forced)
    for s_ in range(len(tr.tokens)):
        assert float(np.abs(logits[0, s_] - tr.logits_trace[s_]).max()) <= LOGIT_TOL[capi.SB_DTYPE_F16]
    eng.close()
