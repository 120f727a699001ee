"""oracle/ -- CPU restatement of the reference's algorithm for the Whisper hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it, and only as the checker.  Nothing under ``spittle_b200/`` imports it.

PARITY UNPINNED.  The reference (tchamp1912/Spittle) holds no golden vector, known-answer
test or fixture for any function on this path, and its own
implementation cannot be built or imported here: the arithmetic lives in un-vendored
third-party crates --

    rubato 0.16.2            (src-tauri/Cargo.toml:58,  Cargo.lock:5384)
    vad-rs 0.1.5 @88b3a01    (src-tauri/Cargo.toml:63,  Cargo.lock:7737)   + ort 2.0.0-rc.10
    transcribe-rs 0.2.3      (src-tauri/Cargo.toml:76,  Cargo.lock:7471)
    whisper-rs 0.13.2 / whisper-rs-sys 0.11.1 (Cargo.lock:8156,8165; bundles whisper.cpp)

-- and no Rust toolchain exists in this image.  The oracle therefore restates the
*published* algorithms of those crates (SURVEY.md Appendices A-D) and anchors them on the
reference's call sites:

    audio_toolkit/audio/resampler.rs:24,51-56  -> oracle/resample.py
    audio_toolkit/vad/silero.rs:25,41-50       -> oracle/silero.py   (weights: first-hand
                                                  from resources/models/silero_vad_v4.onnx)
    audio_toolkit/vad/smoothed.rs:41-96        -> oracle/vad_gate.py
    managers/audio.rs:466-475                  -> oracle/vad_gate.py (short-clip pad rule)
    managers/transcription.rs:494-503          -> oracle/whisper_ref.py, oracle/logmel.py

Every assumption restated from memory of an absent source is listed in
``oracle/ASSUMPTIONS.md``.  Golden vectors under ``tests/golden/`` are generated from THIS
oracle (script: ``tests/golden/make_golden.py``) so later sessions detect oracle drift;
none come from the reference.
"""
