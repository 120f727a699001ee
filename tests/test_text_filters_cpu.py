"""Text post-filters of transcribe() against the reference's OWN unit-test vectors
(src-tauri/src/audio_toolkit/text.rs:398-673, transcribed into tests/golden/text_filters.json): the one
place on this path where the reference pins results."""
import json
import os

import pytest

from spittle_b200 import text_filters as tf

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "text_filters.json"), encoding="utf-8"))


@pytest.mark.parametrize("v", G["apply_custom_words_eq"], ids=lambda v: f"text.rs:{v['line']}")
def test_apply_custom_words_exact(v):
    assert tf.apply_custom_words(v["text"], v["words"], v["threshold"]) == v["expect"]


@pytest.mark.parametrize("v", G["apply_custom_words_contains"], ids=lambda v: f"text.rs:{v['line']}")
def test_apply_custom_words_contains(v):
    got = tf.apply_custom_words(v["text"], v["words"], v["threshold"])
    for s in v["contains"]:
        assert s in got, got
    for s in v["not_contains"]:
        assert s not in got, got


@pytest.mark.parametrize("v", G["preserve_case_pattern"], ids=lambda v: v["original"])
def test_preserve_case_pattern(v):
    assert tf.preserve_case_pattern(v["original"], v["replacement"]) == v["expect"]


@pytest.mark.parametrize("v", G["extract_punctuation"], ids=lambda v: v["word"])
def test_extract_punctuation(v):
    assert list(tf.extract_punctuation(v["word"])) == v["expect"]


@pytest.mark.parametrize("v", G["filter_transcription_output"], ids=lambda v: f"text.rs:{v['line']}:{v['text'][:16]}")
def test_filter_transcription_output(v):
    assert tf.filter_transcription_output(v["text"]) == v["expect"]


@pytest.mark.parametrize("v", G["filter_transcription_output_nonempty"], ids=lambda v: v["text"][:20])
def test_filter_keeps_legitimate_text(v):
    assert tf.filter_transcription_output(v["text"]) != ""


@pytest.mark.parametrize("v", G["clean_segment_boundaries"], ids=lambda v: f"text.rs:{v['line']}")
def test_clean_segment_boundaries(v):
    assert tf.clean_segment_boundaries(v["segments"], v["remaining"]) == v["expect"]


def test_soundex_and_levenshtein_known_answers():
    # published Soundex examples (first letter kept as typed; h/w transparent; adjacent codes merged)
    assert tf.soundex_code("robert") == tf.soundex_code("rupert") == "r163"
    assert tf.soundex_code("ashcraft") == "a261"
    assert tf.soundex_code("tymczak")[:1] == "t"
    assert tf.levenshtein("kitten", "sitting") == 3 and tf.levenshtein("", "abc") == 3 and tf.levenshtein("abc", "abc") == 0


# ---- the same vectors through the compiled C++ host mirror (host/text_filters.cpp via host/sb_transcribe_cli) ----
def _cli():
    import subprocess
    from spittle_b200 import build
    exe = build.build_host()
    def run(args, text):
        r = subprocess.run([exe] + args, input=text.encode(), capture_output=True, timeout=30)
        assert r.returncode == 0, r.stderr
        return r.stdout.decode()
    return run


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(__file__)), "spittle_b200", "libspittle_b200.so")):
        pytest.skip("libspittle_b200.so not built")
    return _cli()


def test_cpp_filter_transcription_output(cli):
    for v in G["filter_transcription_output"]:
        assert cli(["--filter"], v["text"]) == v["expect"], v
    for v in G["filter_transcription_output_nonempty"]:
        assert cli(["--filter"], v["text"]) != ""


def test_cpp_apply_custom_words(cli):
    for v in G["apply_custom_words_eq"]:
        if v["words"]:
            assert cli(["--custom-words", str(v["threshold"])] + v["words"], v["text"]) == v["expect"], v
    for v in G["apply_custom_words_contains"]:
        got = cli(["--custom-words", str(v["threshold"])] + v["words"], v["text"])
        for s_ in v["contains"]:
            assert s_ in got, (v, got)
        for s_ in v["not_contains"]:
            assert s_ not in got, (v, got)
