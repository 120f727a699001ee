"""GPU: the host-side mirrors of the reference TranscriptionManager (C++ CLI and Python) behave like
the reference surface: empty audio -> "", not loaded -> Err, background load + wait, unload."""
import os
import subprocess

import numpy as np
import pytest

from spittle_b200 import build, capi, ggml_format, synth, transcription

pytestmark = pytest.mark.gpu


def test_cpp_transcription_manager_cli(cuda_dev, model_dir, tmp_path):
    cli = build.build_host()
    path = synth.ensure_model_file("nano", model_dir)
    clip = synth.make_clip(2, 7.3)
    f = tmp_path / "clip.f32"
    clip.tofile(str(f))
    r = subprocess.run([cli, path, str(f), "en"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = dict(l.split(": ", 1) for l in r.stdout.strip().splitlines() if ": " in l)
    assert lines["before load"] == "Model is not loaded for transcription."
    assert lines["empty audio"] == "ok=1 text=''"
    assert lines["model"] == "cli-model" and lines["loaded after unload"] == "0"
    eng = capi.Engine(path, dtype=capi.SB_DTYPE_F16)
    from spittle_b200 import text_filters
    # the manager applies the reference's post-filter to the engine text (transcription.rs:549)
    assert lines["text"] == text_filters.filter_transcription_output(eng.transcribe(clip).text.decode()).strip()
    eng.close()


def test_python_transcription_manager(cuda_dev, model_dir):
    path = synth.ensure_model_file("nano", model_dir)
    st = transcription.Settings(selected_model="nano", selected_language="en")
    events = []
    tm = transcription.TranscriptionManager({"nano": path}, lambda: st, on_model_state=lambda *a: events.append(a))
    assert tm.transcribe(np.zeros(0, np.float32)) == ""
    with pytest.raises(transcription.TranscriptionError, match="Model is not loaded for transcription."):
        tm.transcribe(synth.make_clip(1, 2.0))
    tm.initiate_model_load()
    clip = synth.make_clip(1, 5.0)
    text = tm.transcribe(clip)                      # waits for the background load
    assert tm.is_model_loaded() and tm.get_current_model() == "nano"
    assert tm.transcribe_batch([clip, clip[:40000]])[0] == text
    st.selected_language = "auto"                  # the reference default (settings.rs:427-429): detected per clip
    assert isinstance(tm.transcribe(clip), str)
    st.selected_language = "xx"
    with pytest.raises(transcription.TranscriptionError, match="Whisper transcription failed"):
        tm.transcribe(clip)                         # unknown language code: loud error
    st.selected_language = "en"
    st.model_unload_timeout = "immediately"
    tm.transcribe(clip)
    assert not tm.is_model_loaded()
    with pytest.raises(transcription.TranscriptionError, match="Model not found"):
        tm.load_model("missing")
    # model-state-changed events (domain/events.rs): background load, then the immediate unload after the last call
    kinds = [e[0] for e in events]
    assert kinds[:2] == ["loading_started", "loaded"] and events[1][1] == "nano" and kinds.count("unloaded") == 1
    # jargon: the dictionary's terms become Whisper's initial_prompt (transcription.rs:461-492) and steer the decoder
    from spittle_b200 import jargon
    st.model_unload_timeout = "never"
    tm.load_model("nano")
    plain = tm.transcribe(clip)
    st.jargon_custom_terms = ["kubectl", "TypeScript", "Next.js"]
    assert tm._initial_prompt(st) == jargon.build_initial_prompt(jargon.ActiveDictionary(st.jargon_custom_terms, []))
    with_prompt = tm.transcribe(clip)
    assert with_prompt != plain
    tm.unload_model()


def test_full_size_batch_invariance_and_determinism(cuda_dev, model_dir):
    """Whisper Small (the BASELINE.json configs[1] architecture) at full size, where the numpy oracle is too slow to
    be the checker: size-independent properties instead.  (1) the result of a clip does not depend on what else is
    in the batch or on its position (clips do not interact: no_context, per-sequence state) -- except that lanes
    regroup sequences, which must not change any token; (2) two runs are bit-identical; (3) the single-clip API
    equals a batch of one; (4) empty and sub-second clips inside a large batch give "" without disturbing others."""
    path = synth.ensure_model_file("small", model_dir)
    eng = capi.Engine(path, dtype=capi.SB_DTYPE_F16, max_batch=16)
    clips = [synth.make_clip(i, 30.0 if i % 3 else 17.5) for i in range(10)]
    params = capi.default_params(n_max_tokens=40, max_windows=2)
    full = eng.transcribe_batch(clips, params)
    again = eng.transcribe_batch(clips, params)
    assert [r.sampled for r in full] == [r.sampled for r in again] and [r.text for r in full] == [r.text for r in again]
    mixed = [clips[7], np.zeros(0, np.float32), clips[2], synth.make_clip(3, 0.4), clips[0]]
    part = eng.transcribe_batch(mixed, params)
    assert part[1].text == b"" and part[3].text == b"" and part[1].windows == [] and part[3].windows == []
    for got, idx in ((part[0], 7), (part[2], 2), (part[4], 0)):
        assert got.sampled == full[idx].sampled and got.text == full[idx].text
    one = eng.transcribe(clips[5], params)
    assert one.sampled == full[5].sampled and one.text == full[5].text
    assert len({tuple(r.sampled) for r in full}) >= 8          # the audio steers the tokens: clips differ
    eng.close()


def test_long_audio_many_windows_and_batch_chunking(cuda_dev, model_dir):
    """Edge sizes: a 3.5-minute clip (the seek loop runs many windows, whisper_full semantics), more clips than
    max_batch (the engine cuts the call into groups), and a 128-sequence batch (two 64-row chunks per projection)."""
    from oracle import whisper_ref
    path = synth.ensure_model_file("nano", model_dir)
    model = ggml_format.read_ggml(path)
    params = capi.default_params(n_max_tokens=16)
    eng = capi.Engine(path, dtype=capi.SB_DTYPE_F16, max_batch=4)
    long_clip = np.concatenate([synth.make_clip(i, 30.0) for i in (1, 2, 3, 4, 5, 6, 7)])          # 210 s
    r = eng.transcribe(long_clip, params)
    assert len(r.windows) >= 7 and all(w["seek"] >= 0 for w in r.windows)
    seeks = [w["seek"] for w in r.windows]
    assert seeks == sorted(seeks) and seeks[0] == 0 and seeks[-1] + 100 < 21000 + 3000
    # the oracle on the first 3 windows of the same long clip
    oracle = whisper_ref.WhisperOracle(model, act_f16=True)
    _, _, wins = oracle.full(long_clip, whisper_ref.DecodeConfig(n_max_override=16), max_windows=3)
    for wi, w_ref in enumerate(wins):
        got = r.sampled[r.windows[wi]["token_offset"]: r.windows[wi]["token_offset"] + r.windows[wi]["n_tokens"]]
        if got != w_ref.tokens:
            first = next(k for k in range(min(len(got), len(w_ref.tokens))) if got[k] != w_ref.tokens[k])
            assert w_ref.margins[first] < 0.4, (wi, first)
            break
        assert r.windows[wi]["seek_delta"] == w_ref.seek_delta
    # 10 clips through max_batch = 4 -> groups of 4, 4, 2; identical to one-at-a-time
    clips = [synth.make_clip(i, 8.0 + i) for i in range(10)]
    batch = eng.transcribe_batch(clips, params)
    for i in (0, 3, 4, 9):
        assert batch[i].sampled == eng.transcribe(clips[i], params).sampled
    eng.close()
    big = capi.Engine(path, dtype=capi.SB_DTYPE_F16, max_batch=128)
    many = [synth.make_clip(i % 16, 6.0 + (i % 5)) for i in range(128)]
    out = big.transcribe_batch(many, params)
    for i in (0, 17, 64, 127):
        assert out[i].sampled == big.transcribe(many[i], params).sampled
    big.close()


def test_engine_tokenizer_matches_the_mirror(cuda_dev, model_dir):
    """sb_tokenize (C++: std::regex word split + greedy longest vocabulary match) == spittle_b200/tokenizer.py on the
    synthetic vocabulary, for the text jargon.build_initial_prompt produces and for odd input."""
    from spittle_b200 import ggml_format, jargon, tokenizer
    path = synth.ensure_model_file("nano", model_dir)
    model = ggml_format.read_ggml(path)
    eng = capi.Engine(path, max_batch=2)
    t2i = tokenizer.token_to_id(model.vocab)
    prompt = jargon.build_initial_prompt(jargon.ActiveDictionary(["TypeScript", "Next.js", "kubectl", "EC2", "O'Reilly's"], []))
    texts = [prompt, "", "   leading spaces and\ttabs\n", "it's 42 degrees, isn't it?", "café naïve — done", " " .join(
        model.vocab[i].decode("utf-8", errors="ignore") for i in (400, 4000, 14000, 30000))]
    for t in texts:
        got = eng.tokenize(t)
        want = tokenizer.tokenize(t2i, t)
        assert got == want, (t, got[:8], want[:8])
    assert len(eng.tokenize(prompt)) > 0
    eng.close()
