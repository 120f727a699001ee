"""GPU parity of whisper_full's temperature schedule: decoding at a temperature > 0 draws tokens with a
std::mt19937(0) / std::discrete_distribution stream, a failed or improbable window is decoded again one temperature up.
The pinned parity configuration keeps temperature_inc = 0; this is the path the shipped app takes when whisper.cpp's own
default (0.2) is left on."""
import numpy as np
import pytest

from oracle import whisper_ref
from oracle.mt19937 import Mt19937
from spittle_b200 import capi, ggml_format, synth

pytestmark = pytest.mark.gpu

MARGIN_TOL = 0.15          # greedy steps: oracle top-1 / top-2 logit margin below which a token may differ
DRAW_TOL = 0.05            # drawn steps: probability mass between the uniform number and the nearer interval edge


def test_mt19937_known_answer():
    """C++ standard [rand.predef]: the 10000th output of a default-constructed std::mt19937 (seed 5489) is 4123659995."""
    r = Mt19937(5489)
    for _ in range(9999):
        r.next_u32()
    assert r.next_u32() == 4123659995


def _check(res, wins, what):
    """windows identical, or first difference at an indecisive greedy margin / draw margin; returns (exact windows,
    identical tokens up to the first difference).  Drawing amplifies rounding noise: a logit difference of a few 1e-2
    moves the edges of the cumulative distribution by ~1 % of the mass, so over 40 draws most windows meet one marginal
    draw -- what the test pins is that every draw BEFORE it is identical (same stream, same filter, same CDF)."""
    n_exact = n_tok = 0
    for wi, w_ref in enumerate(wins):
        w = res.windows[wi]
        got = res.sampled[w["token_offset"]: w["token_offset"] + w["n_tokens"]]
        if w["n_attempts"] != w_ref.attempts or abs(w["temperature"] - w_ref.temperature) > 1e-6:
            # the schedule itself differs: only legitimate when the accept / reject decision was marginal
            return n_exact, n_tok
        if got != w_ref.tokens:
            k = next((i for i in range(min(len(got), len(w_ref.tokens))) if got[i] != w_ref.tokens[i]), None)
            assert k is not None, (what, wi, len(got), len(w_ref.tokens))
            if w_ref.temperature > 0:
                assert w_ref.draw_margins[k] < DRAW_TOL, (what, wi, k, w_ref.draw_margins[k])
            else:
                assert w_ref.margins[k] < MARGIN_TOL, (what, wi, k, w_ref.margins[k])
            return n_exact, n_tok + k
        assert w["result_len"] == w_ref.result_len and w["seek_delta"] == w_ref.seek_delta and bool(w["failed"]) == w_ref.failed
        assert abs(w["avg_logprob"] - w_ref.avg_logprob) < 0.05 or (np.isnan(w["avg_logprob"]) and np.isnan(w_ref.avg_logprob))
        n_exact += 1
        n_tok += len(got)
    return n_exact, n_tok


def test_sampling_and_fallback_match_oracle(cuda_dev, model_dir):
    path = synth.ensure_model_file("nano", model_dir)
    model = ggml_format.read_ggml(path)
    oracle = whisper_ref.WhisperOracle(model, act_f16=True)
    eng = capi.Engine(path, dtype=capi.SB_DTYPE_F16, max_batch=4)
    clips = [synth.make_clip(1, 30.0), synth.make_clip(2, 30.0), synth.make_clip(5, 30.0), synth.make_clip(4, 12.0)]
    total = {}
    for name, kw in (("T=0.4", dict(temperature=0.4)), ("fallback", dict(temperature_inc=0.2)),
                     ("fallback strict", dict(temperature_inc=0.2, logprob_thold=-0.3)), ("T=1", dict(temperature=1.0))):
        params = capi.default_params(n_max_tokens=40, max_windows=2, **kw)
        eng.stats(reset=True)
        res = eng.transcribe_batch(clips, params)
        st = eng.stats()
        n_exact = n_win = n_tok = 0
        for x, r in zip(clips, res):
            _, _, wins = oracle.full(x, whisper_ref.DecodeConfig(n_max_override=40, **kw), max_windows=2)
            a, b = _check(r, wins, name)
            n_exact += a
            n_tok += b
            n_win += len(wins)
        total[name] = (n_exact, n_win, st["fallbacks"], n_tok)
        print(name, "windows exact", n_exact, "/", n_win, "identical tokens before the first marginal draw", n_tok,
              "fallback decodes", st["fallbacks"])
    assert total["T=0.4"][3] >= 24 and total["T=1"][3] >= 4 and total["fallback"][0] >= 2
    assert total["fallback"][2] >= 3 and total["fallback strict"][2] > total["fallback"][2]
    assert total["T=0.4"][2] == 0
    # a fallback decode changes the tokens, and the result reports it
    plain = eng.transcribe_batch(clips, capi.default_params(n_max_tokens=40, max_windows=2))
    fb = eng.transcribe_batch(clips, capi.default_params(n_max_tokens=40, max_windows=2, temperature_inc=0.2))
    assert any(w["n_attempts"] > 1 for r in fb for w in r.windows)
    assert all(w["n_attempts"] == 1 and w["temperature"] == 0.0 for r in plain for w in r.windows)
    assert [r.sampled for r in plain] != [r.sampled for r in fb]
    # the draws are reproducible: every call starts its clips from std::mt19937(0)
    again = eng.transcribe_batch(clips, capi.default_params(n_max_tokens=40, max_windows=2, temperature_inc=0.2))
    assert [r.sampled for r in again] == [r.sampled for r in fb]
    # and they do not depend on what else is in the batch
    alone = eng.transcribe(clips[2], capi.default_params(n_max_tokens=40, max_windows=2, temperature_inc=0.2))
    assert alone.sampled == fb[2].sampled
    eng.close()


def test_sampling_at_full_size(cuda_dev, model_dir):
    """the same on Whisper Small, where the engine's logits are within 2e-2 of the oracle's (nano: 8e-2): the drawn
    sequences stay identical for longer before the first marginal draw."""
    path = synth.ensure_model_file("small", model_dir)
    model = ggml_format.read_ggml(path)
    oracle = whisper_ref.WhisperOracle(model, act_f16=True)
    eng = capi.Engine(path, dtype=capi.SB_DTYPE_F16, max_batch=2)
    x = synth.make_clip(0, 30.0)
    n_tok = 0
    for t in (0.6, 1.0):
        kw = dict(temperature=t)
        r = eng.transcribe(x, capi.default_params(n_max_tokens=32, max_windows=1, **kw))
        _, _, wins = oracle.full(x, whisper_ref.DecodeConfig(n_max_override=32, **kw), max_windows=1)
        a, b = _check(r, wins, f"small T={t}")
        print(f"small T={t}: {b} identical drawn tokens, window exact: {a}")
        n_tok += b
    assert n_tok >= 16
    eng.close()
