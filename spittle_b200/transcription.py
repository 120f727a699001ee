"""Python mirror of the reference ``TranscriptionManager`` over the C ABI (test/bench convenience;
the compiled host side is host/transcription_manager.{hpp,cpp}, the Rust drop-in is rust/).

Reference surface (src-tauri/src/managers/transcription.rs): new :89, is_model_loaded :170,
unload_model :175, maybe_unload_immediately :211, load_model :223, initiate_model_load :374,
get_current_model :393, transcribe :398.  Error strings are the reference's.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np

from . import capi, jargon, text_filters


class TranscriptionError(RuntimeError):
    """anyhow::Error analogue."""


@dataclass
class Settings:
    """Hot-path subset of AppSettings (settings.rs:427-429, 925)."""
    selected_model: str = ""
    selected_language: str = "auto"
    translate_to_english: bool = False
    model_unload_timeout: str = "never"      # "never" | "immediately" | seconds as str
    custom_words: List[str] = field(default_factory=list)
    word_correction_threshold: float = 0.18      # default_settings.json / settings.rs
    # jargon (settings.rs; all empty by default): profile ids, user terms / corrections, and the profile table itself
    jargon_enabled_profiles: List[str] = field(default_factory=list)
    jargon_custom_terms: List[str] = field(default_factory=list)
    jargon_custom_corrections: List["jargon.JargonCorrection"] = field(default_factory=list)
    jargon_profiles: Dict[str, "jargon.JargonProfile"] = field(default_factory=dict)
    # ids of user jargon packs (settings.jargon_packs; their profiles are part of jargon_profiles, like build_profiles_map
    # transcription.rs:50-63 merges them): only their presence matters for the gates at :462-464 / :553-555
    jargon_packs: List[str] = field(default_factory=list)
    # DomainSelectorManager stand-in (transcription.rs:65-87): (settings, context_text) -> profile ids or None
    profile_selector: Optional[Callable[["Settings", str], Optional[List[str]]]] = None
    domain_selector_blend_manual_profiles: bool = False
    device: int = 0
    max_batch: int = 64
    dtype: int = capi.SB_DTYPE_F16


class TranscriptionManager:
    def __init__(self, model_paths: Dict[str, str], get_settings: Callable[[], Settings],
                 on_model_state: Optional[Callable[[str, Optional[str], Optional[str]], None]] = None):
        """on_model_state(event_type, model_id, error): stand-in for app_handle.emit("model-state-changed", ModelStateEvent)
        (domain/events.rs:3-43): "loading_started" | "loading_failed" | "loaded" | "unloaded"."""
        self._paths = model_paths                  # ModelManager::get_model_path stand-in
        self._get_settings = get_settings
        self._emit = on_model_state or (lambda *a: None)
        self._engine: Optional[capi.Engine] = None
        self._engine_lock = threading.Lock()
        self._current_model_id: Optional[str] = None
        self._last_activity = time.time()
        self._loading = False
        self._loading_cv = threading.Condition()

    def is_model_loaded(self) -> bool:
        with self._engine_lock:
            return self._engine is not None

    def unload_model(self) -> None:
        with self._engine_lock:
            if self._engine is not None:
                self._engine.close()
            self._engine = None
        self._current_model_id = None
        self._emit("unloaded", None, None)

    def maybe_unload_immediately(self, context: str) -> None:
        if self._get_settings().model_unload_timeout == "immediately" and self.is_model_loaded():
            self.unload_model()

    def load_model(self, model_id: str) -> None:
        self._emit("loading_started", model_id, None)
        path = self._paths.get(model_id)
        if path is None:
            raise TranscriptionError(f"Model not found: {model_id}")
        s = self._get_settings()
        try:
            eng = capi.Engine(path, device=s.device, max_batch=s.max_batch, dtype=s.dtype)
        except capi.SbError as e:
            msg = f"Failed to load whisper model {model_id}: {e}"
            self._emit("loading_failed", model_id, msg)
            raise TranscriptionError(msg) from e
        with self._engine_lock:
            if self._engine is not None:
                self._engine.close()
            self._engine = eng
        self._current_model_id = model_id
        self._emit("loaded", model_id, None)

    def initiate_model_load(self) -> None:
        with self._loading_cv:
            if self._loading or self.is_model_loaded():
                return
            self._loading = True

        def work():
            try:
                self.load_model(self._get_settings().selected_model)
            except TranscriptionError:
                pass
            finally:
                with self._loading_cv:
                    self._loading = False
                    self._loading_cv.notify_all()

        threading.Thread(target=work, daemon=True).start()

    def get_current_model(self) -> Optional[str]:
        return self._current_model_id

    @staticmethod
    def _effective_profile_ids(s: Settings, context_text: str) -> List[str]:
        """transcription.rs:65-87: the enabled profiles, replaced by (or blended with) the domain selector's pick."""
        ids = list(s.jargon_enabled_profiles)
        auto = s.profile_selector(s, context_text) if s.profile_selector else None
        if auto is not None:
            if s.domain_selector_blend_manual_profiles:
                ids += [p for p in auto if p not in ids]
            else:
                ids = list(auto)
        return ids

    @classmethod
    def _initial_prompt(cls, s: Settings) -> Optional[str]:
        """transcription.rs:461-492: the jargon dictionary's terms as Whisper's initial_prompt."""
        if not (s.jargon_enabled_profiles or s.jargon_custom_terms or s.jargon_packs):
            return None
        d = jargon.compute_active_dictionary(
            jargon.JargonSettings(cls._effective_profile_ids(s, ""), s.jargon_custom_terms, s.jargon_custom_corrections),
            s.jargon_profiles)
        if not d.terms:
            return None
        return jargon.build_initial_prompt(d) or None

    def _params(self, s: Settings):
        lang = s.selected_language
        if lang in ("zh-Hans", "zh-Hant"):
            lang = "zh"
        kw = dict(translate=int(s.translate_to_english))
        p = capi.default_params(**kw)
        p.language = None if lang == "auto" else lang.encode()
        prompt = self._initial_prompt(s)
        p.initial_prompt = prompt.encode("utf-8") if prompt else None
        return p

    @classmethod
    def _post_filter(cls, text: str, s: Settings) -> str:
        """transcription.rs:537-580: custom-word correction (only when configured), the filler / stutter /
        hallucination filter, then the jargon corrections (only when profiles or custom corrections are configured)."""
        if s.custom_words:
            text = text_filters.apply_custom_words(text, s.custom_words, s.word_correction_threshold)
        text = text_filters.filter_transcription_output(text)
        if s.jargon_enabled_profiles or s.jargon_custom_corrections or s.jargon_packs:
            d = jargon.compute_active_dictionary(
                jargon.JargonSettings(cls._effective_profile_ids(s, text), s.jargon_custom_terms, s.jargon_custom_corrections),
                s.jargon_profiles)
            if d.corrections:
                text = jargon.apply_corrections(text, d.corrections)
        return text

    def transcribe(self, audio) -> str:
        self._last_activity = time.time()
        a = np.ascontiguousarray(audio, dtype=np.float32)
        if a.size == 0:                                      # transcription.rs:412-416
            self.maybe_unload_immediately("empty audio")
            return ""
        with self._loading_cv:
            while self._loading:
                self._loading_cv.wait()
        s = self._get_settings()
        with self._engine_lock:
            if self._engine is None:
                raise TranscriptionError("Model is not loaded for transcription.")
            try:
                r = self._engine.transcribe(a, self._params(s))
            except capi.SbError as e:
                raise TranscriptionError(f"Whisper transcription failed: {e}") from e
        text = self._post_filter(r.text.decode("utf-8", errors="replace"), s)
        self.maybe_unload_immediately("transcription")
        return text

    def transcribe_batch(self, clips) -> List[str]:
        self._last_activity = time.time()
        with self._loading_cv:
            while self._loading:
                self._loading_cv.wait()
        s = self._get_settings()
        with self._engine_lock:
            if self._engine is None:
                raise TranscriptionError("Model is not loaded for transcription.")
            res = self._engine.transcribe_batch(clips, self._params(s))
        return [self._post_filter(r.text.decode("utf-8", errors="replace"), s) for r in res]
