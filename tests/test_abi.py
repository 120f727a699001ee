"""CPU: the C-ABI library loads and exports every symbol include/spittle_b200.h declares."""
import ctypes
import os
import re


def declared_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "include", "spittle_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"SB_API\s+[\w\s\*]+?\b(sb_\w+)\s*\(", src)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "sb_last_error" in syms and "sb_logmel" in syms and len(syms) >= 8


def test_library_exports_every_declared_symbol(lib_built):
    l = ctypes.CDLL(lib_built)
    missing = [s for s in declared_symbols() if not hasattr(l, s)]
    assert not missing, f"declared in include/spittle_b200.h but not exported: {missing}"


def test_version_and_geometry(lib_built):
    from spittle_b200 import capi
    assert "sm_100a" in capi.version()
    # whisper.cpp geometry for a 30 s clip
    assert capi.logmel_geometry(480000) == (6000, 2999, 3002)
    assert capi.logmel_geometry(20000) == (3125, 124, 127)


def test_no_device_fails_loudly(lib_built):
    """Without a GPU the product must raise, never fall back to a CPU path."""
    import numpy as np
    import pytest
    import torch
    from spittle_b200 import capi, synth
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(capi.SbError):
        capi.MelPlan(synth.mel_filterbank(80))


# ---- struct layout: library == checked-in table == ctypes mirror == Rust bindings ------------------------------
def _golden_layout():
    import json
    root = os.path.dirname(os.path.abspath(__file__))
    return [tuple(r) for r in json.load(open(os.path.join(root, "golden", "abi_layout.json")))]


def test_abi_layout_matches_checked_in_table_and_ctypes(lib_built):
    """sizeof / offsetof of every ABI struct as the LIBRARY reports them (sb_abi_layout) against the checked-in
    table and against the ctypes mirror the tests call through (round 1 shipped a Rust sb_result 8 bytes short)."""
    import ctypes as C
    from spittle_b200 import capi
    rows = capi.abi_layout()
    assert rows == _golden_layout(), "include/spittle_b200.h changed: regenerate tests/golden/abi_layout.json AND the bindings"
    for sname, field, size, off in rows:
        st = capi.ABI_STRUCTS[sname]
        assert C.sizeof(st) == size, (sname, C.sizeof(st), size)
        assert getattr(st, field).offset == off, (sname, field)
    # every field of the mirrored structs that carry results / parameters is covered by the table
    for sname in ("sb_config", "sb_params", "sb_window_info", "sb_segment", "sb_result"):
        covered = {f for s_, f, _, _ in rows if s_ == sname}
        assert covered == {n for n, _ in capi.ABI_STRUCTS[sname]._fields_}, sname


def test_rust_bindings_assert_the_same_layout():
    """rust/spittle-b200-sys cannot be compiled here (no cargo): its const layout asserts are parsed and compared
    with the table, and its struct field lists with the header's."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "rust", "spittle-b200-sys", "src", "lib.rs")).read()
    table = {(s_, f): (size, off) for s_, f, size, off in _golden_layout()}
    sizes = {s_: size for s_, _, size, _ in _golden_layout()}
    n = 0
    for s_, size in re.findall(r"assert!\(size_of::<(\w+)>\(\) == (\d+)\)", src):
        assert sizes[s_] == int(size), s_
        n += 1
    for s_, f, off in re.findall(r"assert!\(offset_of!\((\w+), (\w+)\) == (\d+)\)", src):
        assert table[(s_, f)][1] == int(off), (s_, f)
        n += 1
    assert n >= 15
    # field order / names of the Rust structs equal the table's (fields are declared `pub name: type`)
    for sname in ("sb_config", "sb_params", "sb_window_info", "sb_segment", "sb_result"):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % sname, src, re.S).group(1)
        body = re.sub(r"//[^\n]*", "", body)
        rust_fields = re.findall(r"pub (\w+):", body)
        want = [f for s_, f, _, _ in _golden_layout() if s_ == sname]
        assert rust_fields == want, (sname, rust_fields, want)


def test_rust_drop_in_declares_the_reference_surface():
    """rust/transcription_b200.rs must offer what every caller of managers::transcription uses (reference
    transcription.rs:89-624 / transcription_mock.rs:25-55) -- it cannot be compiled here, so the surface is checked textually."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "rust", "transcription_b200.rs")).read()
    for sig in ("pub fn new(app_handle: &AppHandle, model_manager: Arc<ModelManager>) -> Result<Self>",
                "pub fn is_model_loaded(&self) -> bool", "pub fn unload_model(&self) -> Result<()>",
                "pub fn maybe_unload_immediately(&self, context: &str)", "pub fn load_model(&self, model_id: &str) -> Result<()>",
                "pub fn initiate_model_load(&self)", "pub fn get_current_model(&self) -> Option<String>",
                "pub fn transcribe(&self, audio: Vec<f32>) -> Result<String>", "impl Drop for TranscriptionManager",
                "#[derive(Clone)]\npub struct TranscriptionManager"):
        assert sig in src, sig
    # the pieces round 1 lacked: idle watcher, model-state events, last_activity, jargon prompt + corrections
    for needle in ("fn spawn_idle_watcher", "\"model-state-changed\"", "ModelStateKind::LoadingFailed", "last_activity.store",
                   "build_initial_prompt", "apply_corrections", "p.initial_prompt =", "handle.join()"):
        assert needle in src, needle
    # every sb_* symbol the drop-in calls is declared by the sys crate, and every one of those by the header
    sys_src = open(os.path.join(root, "rust", "spittle-b200-sys", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(sb_\w+)\(", src))
    declared = set(re.findall(r"pub fn (sb_\w+)\(", sys_src))
    assert used and used <= declared, used - declared
    assert declared <= set(declared_symbols()), declared - set(declared_symbols())
