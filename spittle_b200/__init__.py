"""spittle_b200 -- B200-native (sm_100a) Whisper transcription hot path.

Drop-in for the path behind the reference's
``TranscriptionManager::transcribe(Vec<f32>) -> Result<String>``
(reference: src-tauri/src/managers/transcription.rs:398-605).

Layout (only what the path needs):

* ``csrc/``        hand-written CUDA kernels + the C-ABI (``include/spittle_b200.h``)
* ``capi.py``      ctypes binding of ``libspittle_b200.so`` (fails loudly if the
                   library is missing: there is NO CPU fallback in the product)
* ``transcription.py`` host-side mirror of the reference ``TranscriptionManager``
* ``audio_toolkit.py`` host-side mirror of FrameResampler / SileroVad / SmoothedVad
* ``ggml_format.py``   GGML legacy ``ggml-*.bin`` reader/writer (the model file format
                   the reference loads through whisper-rs)
* ``synth.py``     synthetic audio + random-init weight recipes

Nothing in this package imports ``oracle/``.
"""

__version__ = "0.1.0"
