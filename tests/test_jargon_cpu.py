"""The reference's own unit tests of src-tauri/src/jargon.rs (:741-961), replayed against spittle_b200/jargon.py.
The tests that use builtin_profiles() run here on stand-in profiles holding exactly the entries their assertions name
(the built-in table is settings content, not algorithm)."""
from spittle_b200 import jargon as j
from spittle_b200.jargon import JargonCorrection as C

PROFILES = {
    "web_dev": j.JargonProfile("Web", ["TypeScript", "Next.js", "React"], [C("next js", "Next.js"), C("type script", "TypeScript")]),
    "devops": j.JargonProfile("DevOps", ["Terraform", "Kubernetes"], [C("cube control", "kubectl")]),
}
TS = [C("type script", "TypeScript")]


def settings(profiles, terms, corrections):
    return j.JargonSettings(list(profiles), list(terms), [C(a, b) for a, b in corrections])


def test_profile_merging():                                  # jargon.rs:742-749
    d = j.compute_active_dictionary(settings(["web_dev", "devops"], [], []), PROFILES)
    assert "TypeScript" in d.terms and "Terraform" in d.terms


def test_correction_override_priority():                     # :752-763
    d = j.compute_active_dictionary(settings(["web_dev"], [], [("next js", "NextJS")]), PROFILES)
    c = [x for x in d.corrections if x.from_.lower() == "next js"]
    assert len(c) == 1 and c[0].to == "NextJS"


def test_case_insensitive_dedup():                           # :766-778
    d = j.compute_active_dictionary(settings(["web_dev"], ["typescript"], []), PROFILES)
    assert [t for t in d.terms if t.lower() == "typescript"] == ["typescript"]


def test_protected_spans():                                  # :781-843
    for text, keep in (("Check @file.rs for type script code", "@file.rs"),
                       ("Run `type script build` with type script", "`type script build`"),
                       ("Visit https://type-script.org for type script docs", "https://type-script.org"),
                       ("Open /usr/local/bin/app and type script", "/usr/local/bin/app"),
                       ("Use --verbose and type script", "--verbose")):
        r = j.apply_corrections(text, TS)
        assert keep in r and "TypeScript" in r, (text, r)


def test_multi_word_boundary_safety():                       # :846-856
    assert j.apply_corrections("This script is good", TS) == "This script is good"


def test_stable_initial_prompt():                            # :859-870
    d = j.compute_active_dictionary(settings(["web_dev"], ["MyCustomTerm"], []), PROFILES)
    p = j.build_initial_prompt(d)
    assert p.startswith("Technical dictation. Common terms: ") and p.endswith(".") and len(p.encode()) <= 1000
    assert p.find("MyCustomTerm") < p.find("TypeScript")


def test_initial_prompt_char_limit():                        # :873-884
    d = j.ActiveDictionary(["VeryLongTermNumber%d" % i for i in range(200)], [])
    p = j.build_initial_prompt(d)
    assert 900 < len(p.encode()) <= 1000 and p.endswith(".")


def test_longest_first_ordering():                           # :887-894
    d = j.compute_active_dictionary(settings([], [], [("E C", "EC"), ("E C two", "EC2")]), PROFILES)
    assert [c.from_ for c in d.corrections] == ["E C two", "E C"]


def test_empty_and_no_corrections():                         # :897-912
    assert j.apply_corrections("", [C("test", "Test")]) == ""
    assert j.apply_corrections("Hello world", []) == "Hello world"


def test_case_insensitive_and_multiple_corrections():        # :915-942
    assert j.apply_corrections("I use Type Script and TYPE SCRIPT", TS) == "I use TypeScript and TypeScript"
    assert j.apply_corrections("I use type script with next js", TS + [C("next js", "Next.js")]) == "I use TypeScript with Next.js"


def test_empty_dictionary_prompt():                          # :945-952
    assert j.build_initial_prompt(j.ActiveDictionary([], [])) == ""


def test_manager_applies_jargon_after_the_filters():
    """managers/transcription.rs:551-580: corrections run on the filtered text, only when jargon is configured."""
    from spittle_b200.transcription import Settings, TranscriptionManager
    s = Settings(jargon_custom_corrections=[C("type script", "TypeScript")])
    assert TranscriptionManager._post_filter("um I use type script", s) == "I use TypeScript"
    assert TranscriptionManager._post_filter("um I use type script", Settings()) == "I use type script"


# ---- the same vectors through the compiled C++ host mirror (host/jargon.cpp via host/sb_transcribe_cli) ----
import os

import pytest


@pytest.fixture(scope="module")
def cli():
    import subprocess
    from spittle_b200 import build
    if not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spittle_b200", "libspittle_b200.so")):
        pytest.skip("libspittle_b200.so not built")
    exe = build.build_host()

    def run(args, text=""):
        r = subprocess.run([exe] + args, input=text.encode(), capture_output=True, timeout=30)
        assert r.returncode == 0, r.stderr
        return r.stdout.decode()
    return run


def test_cpp_apply_corrections(cli):
    ts = ["type script", "TypeScript"]
    for text, keep in (("Check @file.rs for type script code", "@file.rs"),
                       ("Run `type script build` with type script", "`type script build`"),
                       ("Visit https://type-script.org for type script docs", "https://type-script.org"),
                       ("Open /usr/local/bin/app and type script", "/usr/local/bin/app"),
                       ("Use --verbose and type script", "--verbose")):
        r = cli(["--jargon-correct"] + ts, text)
        assert keep in r and "TypeScript" in r and r == j.apply_corrections(text, TS), (text, r)
    assert cli(["--jargon-correct"] + ts, "This script is good") == "This script is good"
    assert cli(["--jargon-correct"] + ts, "I use Type Script and TYPE SCRIPT") == "I use TypeScript and TypeScript"
    assert cli(["--jargon-correct"] + ts + ["next js", "Next.js"], "I use type script with next js") == "I use TypeScript with Next.js"
    assert cli(["--jargon-correct"] + ts, "") == ""
    # longest phrase first (jargon.rs:887-894): "E C two" must win over "E C"
    assert cli(["--jargon-correct", "E C", "EC", "E C two", "EC2"], "launch an E C two box on E C") == "launch an EC2 box on EC"
    assert j.apply_corrections("launch an E C two box on E C", j.compute_active_dictionary(
        settings([], [], [("E C", "EC"), ("E C two", "EC2")]), {}).corrections) == "launch an EC2 box on EC"


def test_cpp_build_initial_prompt(cli):
    assert cli(["--jargon-prompt"]) == ""
    assert cli(["--jargon-prompt", "MyCustomTerm", "TypeScript", "typescript"]) == "Technical dictation. Common terms: MyCustomTerm, typescript."
    assert j.build_initial_prompt(j.compute_active_dictionary(settings([], ["MyCustomTerm", "TypeScript", "typescript"], []), {})) == \
        "Technical dictation. Common terms: MyCustomTerm, typescript."
    terms = ["VeryLongTermNumber%d" % i for i in range(200)]
    p = cli(["--jargon-prompt"] + terms)
    assert p == j.build_initial_prompt(j.ActiveDictionary(terms, [])) and 900 < len(p) <= 1000
