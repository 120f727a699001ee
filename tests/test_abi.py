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
