"""CPU: oracle/cpu_ref (the multi-threaded C++ restatement that bench.py times as the CPU baseline) against the numpy
oracle it follows -- log-mel, encoder output, teacher-forced logits, greedy tokens with text context, thread-count
invariance.  Two independent restatements of the same rounding points agree to the f16 flip noise stated below."""
import numpy as np
import pytest

from oracle import cpu_ref, logmel, whisper_ref
from spittle_b200 import ggml_format, synth


@pytest.fixture(scope="module")
def built():
    return cpu_ref.build()


def _load(arch, model_dir, n_threads=4):
    path = synth.ensure_model_file(arch, model_dir)
    model = ggml_format.read_ggml(path)
    return path, model, whisper_ref.WhisperOracle(model, act_f16=True), cpu_ref.CpuRef(path, n_threads=n_threads)


def test_logmel_matches_f32_faithful_oracle(built, model_dir):
    path, model, oracle, ref = _load("nano", model_dir)
    for i, secs in ((1, 30.0), (2, 7.3), (4, 1.2)):
        x = synth.make_clip(i, secs)
        got, n_len_org = ref.logmel(x)
        want, n_len_org_o = logmel.logmel_f32_faithful(x, model.mel_filters)
        assert got.shape == want.shape and n_len_org == n_len_org_o
        assert float(np.abs(got - want).max()) <= 5e-6      # same operation order; libm sin/cos/log10 differ in the last ulp
        f64, _ = logmel.logmel_f64(x, model.mel_filters)
        assert float(np.abs(got - f64).max()) / float(np.abs(f64).max()) <= 1e-4


@pytest.mark.parametrize("arch", ["nano", "micro"])
def test_encoder_and_logits_match_numpy_oracle(built, model_dir, arch):
    path, model, oracle, ref = _load(arch, model_dir)
    x = synth.make_clip(1, 30.0)
    mel, n_len_org = logmel.logmel_f32_faithful(x, model.mel_filters)
    win = logmel.mel_window(mel, 0)
    enc_o = oracle.encode(win)
    enc_c = ref.encode(win)
    rel = float(np.sqrt(((enc_c - enc_o) ** 2).mean()) / np.sqrt((enc_o ** 2).mean()))
    print(f"{arch}: cpu_ref vs numpy encoder rel-RMS {rel:.3e}")
    assert rel <= 1e-3          # f16 rounding flips from the different summation order (measured 3-4e-4)
    n = 20
    wo = oracle.decode_window(enc_o, 0, n_len_org, whisper_ref.DecodeConfig(n_max_override=n), trace=True)
    wc = ref.decode_window(enc_o, 0, n_len_org, n_max_override=n, trace=True, forced=wo.tokens)
    assert wc["tokens"] == wo.tokens
    errs = [float(np.abs(wc["logits"][i] - wo.logits_trace[i]).max()) for i in range(len(wo.tokens))]
    print(f"{arch}: teacher-forced logit max err {max(errs):.3e}")
    assert max(errs) <= 3e-2
    assert np.allclose(wc["margins"], wo.margins, atol=6e-2)
    free = ref.decode_window(enc_o, 0, n_len_org, n_max_override=n)
    if free["tokens"] != wo.tokens:
        first = next(i for i in range(min(len(free["tokens"]), len(wo.tokens))) if free["tokens"][i] != wo.tokens[i])
        assert wo.margins[first] < 6e-2
    else:
        assert (free["result_len"], free["seek_delta"], free["failed"]) == (wo.result_len, wo.seek_delta, wo.failed)


def test_full_with_text_context_matches_numpy_oracle(built, model_dir):
    path, model, oracle, ref = _load("nano", model_dir)
    x = np.concatenate([synth.make_clip(1, 30.0), synth.make_clip(2, 30.0), synth.make_clip(3, 12.0)])
    prompt = [401, 4002, 14001]
    n_exact = n_win = 0
    for kw in (dict(), dict(initial_prompt_tokens=prompt), dict(n_max_text_ctx=0), dict(language_id=-1)):
        cfg = whisper_ref.DecodeConfig(n_max_override=16, **kw)
        text, kept, wins = oracle.full(x, cfg, max_windows=3)
        got = ref.full(x, max_windows=3, n_max_override=16, **kw)
        assert len(got["windows"]) == len(wins) == 3
        if kw.get("language_id") == -1:
            assert got["lang"] == oracle.last_detected_language
        for w_o, w_c in zip(wins, got["windows"]):
            n_win += 1
            if w_c["tokens"] != w_o.tokens:
                first = next(i for i in range(min(len(w_c["tokens"]), len(w_o.tokens))) if w_c["tokens"][i] != w_o.tokens[i])
                assert w_o.margins[first] < 6e-2, (kw, first)
                break
            assert (w_c["result_len"], w_c["seek_delta"], w_c["failed"]) == (w_o.result_len, w_o.seek_delta, w_o.failed)
            n_exact += 1
        else:
            assert got["kept"] == kept
    print(f"cpu_ref full: {n_exact}/{n_win} windows token-exact vs the numpy oracle")
    assert n_exact >= n_win - 3


def test_results_do_not_depend_on_the_thread_count(built, model_dir):
    """every output element is summed by one thread in a fixed order: 1, 3 and 8 threads give bit-identical results."""
    path = synth.ensure_model_file("nano", model_dir)
    x = synth.make_clip(5, 11.0)
    outs = []
    for nt in (1, 3, 8):
        ref = cpu_ref.CpuRef(path, n_threads=nt)
        mel, _ = ref.logmel(x)
        enc = ref.encode(logmel.mel_window(mel, 0))
        r = ref.full(x, max_windows=1, n_max_override=12)
        outs.append((mel.tobytes(), enc.tobytes(), r["windows"][0]["tokens"], r["windows"][0]["margins"]))
        ref.close()
    assert outs[0] == outs[1] == outs[2]


def test_english_only_model(built, model_dir):
    path, model, oracle, ref = _load("nano.en", model_dir)
    x = synth.make_clip(2, 9.0)
    _, kept, wins = oracle.full(x, whisper_ref.DecodeConfig(n_max_override=16), max_windows=1)
    got = ref.full(x, max_windows=1, n_max_override=16)
    assert got["windows"][0]["n_prompt"] == 1
    if got["windows"][0]["tokens"] != wins[0].tokens:
        first = next(i for i in range(len(wins[0].tokens)) if got["windows"][0]["tokens"][i] != wins[0].tokens[i])
        assert wins[0].margins[first] < 6e-2


def test_suppress_nst_numpy_oracle_and_cpu_ref(built, model_dir, tmp_path):
    """whisper_full_params.suppress_nst: the ids of whisper.cpp's non-speech strings get -inf in whisper_process_logits.  The
    synthetic vocabulary has none of them, so they are planted on the tokens an unconstrained decode emits; with the rule on
    the decode must avoid them, and the two restatements must agree."""
    path, model, oracle, ref = _load("nano", model_dir)
    assert whisper_ref.non_speech_token_ids(model.vocab) == []
    toy = [b"a", b"(", b" (", b" -", b"-", b" '", "\u266a".encode(), b" \xe2\x99\xaa", b"))", b"x))"]
    assert whisper_ref.non_speech_token_ids(toy) == [1, 2, 3, 5, 6, 7, 8]
    x = synth.make_clip(1, 30.0)
    mel, n_len_org = logmel.logmel_f32_faithful(x, model.mel_filters)
    enc = oracle.encode(logmel.mel_window(mel, 0))
    n = 24
    base = oracle.decode_window(enc, 0, n_len_org, whisper_ref.DecodeConfig(n_max_override=n))
    sp = model.special
    planted = sorted({t for t in base.tokens if t < sp.eot})
    assert len(planted) >= 8
    m2 = synth.with_non_speech_vocab(model, planted)
    p2 = str(tmp_path / "nano-nst.bin")
    ggml_format.write_ggml(p2, m2)
    o2 = whisper_ref.WhisperOracle(m2, act_f16=True)
    assert set(o2.nst_ids.tolist()) == set(planted[:110])
    off = o2.decode_window(enc, 0, n_len_org, whisper_ref.DecodeConfig(n_max_override=n))
    on = o2.decode_window(enc, 0, n_len_org, whisper_ref.DecodeConfig(n_max_override=n, suppress_nst=True))
    assert off.tokens == base.tokens                         # the vocabulary strings alone change nothing
    assert on.tokens != off.tokens and not (set(on.tokens) & set(planted))
    r2 = cpu_ref.CpuRef(p2, n_threads=4)
    c_on = r2.decode_window(enc, 0, n_len_org, n_max_override=n, suppress_nst=True)
    c_off = r2.decode_window(enc, 0, n_len_org, n_max_override=n)
    assert c_off["tokens"] == off.tokens or min(off.margins) < 6e-2
    if c_on["tokens"] != on.tokens:
        first = next(i for i in range(min(len(c_on["tokens"]), len(on.tokens))) if c_on["tokens"][i] != on.tokens[i])
        assert on.margins[first] < 6e-2
    assert not (set(c_on["tokens"]) & set(planted))
