"""GPU: the host-side mirrors of the reference TranscriptionManager (C++ CLI and Python) behave like
the reference surface: empty audio -> "", not loaded -> Err, background load + wait, unload."""
import os
import subprocess

import numpy as np
import pytest

from spittle_b200 import build, capi, synth, transcription

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
    assert lines["text"] == eng.transcribe(clip).text.decode()
    eng.close()


def test_python_transcription_manager(cuda_dev, model_dir):
    path = synth.ensure_model_file("nano", model_dir)
    st = transcription.Settings(selected_model="nano", selected_language="en")
    tm = transcription.TranscriptionManager({"nano": path}, lambda: st)
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
