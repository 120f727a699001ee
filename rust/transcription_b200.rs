// Drop-in `managers::transcription` module backed by libspittle_b200.so (the C ABI of include/spittle_b200.h).
// NOT COMPILED IN THIS IMAGE (no cargo / rustc); the same logic is compiled and tested as host/transcription_manager.cpp
// and spittle_b200/transcription.py.  rust/spittle-b200-sys holds the raw bindings and their layout asserts.
//
// Wiring in the reference (src-tauri/src/managers/mod.rs:8-12 already swaps this module by cargo feature for the CI mock):
//
//     #[cfg(feature = "b200_transcription")]
//     #[path = "transcription_b200.rs"]
//     pub mod transcription;
//
// and in src-tauri/Cargo.toml:  b200_transcription = ["dep:spittle-b200-sys"]
//
// Same public surface as src-tauri/src/managers/transcription.rs (and transcription_mock.rs:25-55):
//   new :89 (starts the idle watcher :112-163) | is_model_loaded :170 | unload_model :175 | maybe_unload_immediately :211 |
//   load_model :223 (model-state-changed events :227-236, 263-275, 357-366) | initiate_model_load :374 |
//   get_current_model :393 | transcribe :398 (jargon initial_prompt :461-492, post-filters :538-580) | Drop :608-624
// plus the additive `transcribe_batch` (SURVEY 8(b) "Batch / multi-GPU surface").  Every catalog entry whose engine type is
// Whisper goes through the B200 engine; the other engine types of the reference (Parakeet, Moonshine, SenseVoice) are not
// part of this path and are refused with the reference's own "Failed to load" error shape.
use crate::audio_toolkit::{apply_custom_words, filter_transcription_output};
use crate::domain::events::{ModelStateEvent, ModelStateKind};
use crate::jargon::{apply_corrections, build_initial_prompt, builtin_profiles, compute_active_dictionary, JargonProfile, JargonSettings};
use crate::managers::domain_selector::{DomainContext, DomainSelectorManager};
use crate::managers::model::{EngineType, ModelManager};
use crate::settings::{get_settings, AppSettings, ModelUnloadTimeout};
use anyhow::{anyhow, Result};
use log::{debug, error, info, warn};
use spittle_b200_sys as sys;
use std::collections::HashMap;
use std::ffi::{CStr, CString};
use std::sync::atomic::{AtomicBool, AtomicU64, Ordering};
use std::sync::{Arc, Condvar, Mutex};
use std::thread::{self, JoinHandle};
use std::time::{Duration, Instant, SystemTime, UNIX_EPOCH};
use tauri::{AppHandle, Emitter, Manager};

const NOT_LOADED: &str = "Model is not loaded for transcription.";

/// Owning handle of one sb_engine (one model replica per configured CUDA device).
struct Engine(*mut sys::sb_engine);
unsafe impl Send for Engine {}
impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { sys::sb_engine_destroy(self.0) };           // frees the device memory (reference: `*engine = None`)
    }
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::sb_last_error()).to_string_lossy().into_owned() }
}

fn now_ms() -> u64 {
    SystemTime::now().duration_since(UNIX_EPOCH).map(|d| d.as_millis() as u64).unwrap_or(0)
}

/// CUDA devices the engine drives: SPITTLE_B200_DEVICES="0,1,2,3" (default: device 0 alone).
fn configured_devices() -> Vec<i32> {
    std::env::var("SPITTLE_B200_DEVICES")
        .ok()
        .map(|v| v.split(',').filter_map(|s| s.trim().parse().ok()).collect::<Vec<i32>>())
        .filter(|v| !v.is_empty())
        .unwrap_or_else(|| vec![0])
}

#[derive(Clone)]
pub struct TranscriptionManager {
    engine: Arc<Mutex<Option<Engine>>>,
    model_manager: Arc<ModelManager>,
    app_handle: AppHandle,
    current_model_id: Arc<Mutex<Option<String>>>,
    last_activity: Arc<AtomicU64>,
    shutdown_signal: Arc<AtomicBool>,
    watcher_handle: Arc<Mutex<Option<JoinHandle<()>>>>,
    is_loading: Arc<Mutex<bool>>,
    loading_condvar: Arc<Condvar>,
}

impl TranscriptionManager {
    pub fn new(app_handle: &AppHandle, model_manager: Arc<ModelManager>) -> Result<Self> {
        let manager = Self {
            engine: Arc::new(Mutex::new(None)),
            model_manager,
            app_handle: app_handle.clone(),
            current_model_id: Arc::new(Mutex::new(None)),
            last_activity: Arc::new(AtomicU64::new(now_ms())),
            shutdown_signal: Arc::new(AtomicBool::new(false)),
            watcher_handle: Arc::new(Mutex::new(None)),
            is_loading: Arc::new(Mutex::new(false)),
            loading_condvar: Arc::new(Condvar::new()),
        };
        manager.spawn_idle_watcher();
        Ok(manager)
    }

    /// Every 10 s: unload the model once it has been idle longer than settings.model_unload_timeout
    /// (`Immediately` is handled inside transcribe(), `Never` has no limit).
    fn spawn_idle_watcher(&self) {
        let this = self.clone();
        let handle = thread::spawn(move || {
            while !this.shutdown_signal.load(Ordering::Relaxed) {
                thread::sleep(Duration::from_secs(10));
                if this.shutdown_signal.load(Ordering::Relaxed) {
                    break;
                }
                let settings = get_settings(&this.app_handle);
                let Some(limit_s) = settings.model_unload_timeout.to_seconds() else { continue };
                if settings.model_unload_timeout == ModelUnloadTimeout::Immediately {
                    continue;
                }
                let idle_ms = now_ms().saturating_sub(this.last_activity.load(Ordering::Relaxed));
                if idle_ms > limit_s * 1000 && this.is_model_loaded() {
                    let t0 = Instant::now();
                    if this.unload_model().is_ok() {
                        // (the reference emits `unloaded` from unload_model AND from the watcher; kept)
                        this.emit_state(ModelStateKind::Unloaded, None, None, None);
                        debug!("Model unloaded due to inactivity (took {}ms)", t0.elapsed().as_millis());
                    }
                }
            }
            debug!("Idle watcher thread shutting down gracefully");
        });
        *self.watcher_handle.lock().unwrap() = Some(handle);
    }

    fn emit_state(&self, kind: ModelStateKind, id: Option<String>, name: Option<String>, err: Option<String>) {
        let _ = self.app_handle.emit("model-state-changed", ModelStateEvent::new(kind, id, name, err));
    }

    pub fn is_model_loaded(&self) -> bool {
        self.engine.lock().unwrap().is_some()
    }

    pub fn unload_model(&self) -> Result<()> {
        let t0 = Instant::now();
        *self.engine.lock().unwrap() = None;
        *self.current_model_id.lock().unwrap() = None;
        self.emit_state(ModelStateKind::Unloaded, None, None, None);
        debug!("Model unloaded manually (took {}ms)", t0.elapsed().as_millis());
        Ok(())
    }

    pub fn maybe_unload_immediately(&self, context: &str) {
        let settings = get_settings(&self.app_handle);
        if settings.model_unload_timeout == ModelUnloadTimeout::Immediately && self.is_model_loaded() {
            info!("Immediately unloading model after {}", context);
            if let Err(e) = self.unload_model() {
                warn!("Failed to immediately unload model: {}", e);
            }
        }
    }

    pub fn load_model(&self, model_id: &str) -> Result<()> {
        let t0 = Instant::now();
        self.emit_state(ModelStateKind::LoadingStarted, Some(model_id.to_string()), None, None);
        let info = self
            .model_manager
            .get_model_info(model_id)
            .ok_or_else(|| anyhow!("Model not found: {}", model_id))?;
        let fail = |msg: String| {
            self.emit_state(ModelStateKind::LoadingFailed, Some(model_id.to_string()), Some(info.name.clone()), Some(msg.clone()));
            anyhow!(msg)
        };
        if !info.is_downloaded {
            return Err(fail("Model not downloaded".to_string()));
        }
        if !matches!(info.engine_type, EngineType::Whisper) {
            return Err(fail(format!("Failed to load model {}: the B200 engine runs Whisper GGML models only", model_id)));
        }
        let path = self.model_manager.get_model_path(model_id)?;
        let cpath = CString::new(path.to_string_lossy().as_bytes())?;
        let devices = configured_devices();
        let cfg = sys::sb_config {
            model_path: cpath.as_ptr(),
            device: devices[0],
            max_batch: 64,
            dtype: 1,                                       // SB_DTYPE_F16: ggml's own rounding points
            use_cuda_graph: 1,
            devices: if devices.len() > 1 { devices.as_ptr() } else { std::ptr::null() },
            n_devices: if devices.len() > 1 { devices.len() as i32 } else { 0 },
        };
        let mut raw: *mut sys::sb_engine = std::ptr::null_mut();
        if unsafe { sys::sb_engine_create(&cfg, &mut raw) } != 0 {
            return Err(fail(format!("Failed to load whisper model {}: {}", model_id, last_error())));
        }
        *self.engine.lock().unwrap() = Some(Engine(raw));
        *self.current_model_id.lock().unwrap() = Some(model_id.to_string());
        self.emit_state(ModelStateKind::Loaded, Some(model_id.to_string()), Some(info.name.clone()), None);
        debug!("Successfully loaded transcription model: {} (took {}ms)", model_id, t0.elapsed().as_millis());
        Ok(())
    }

    pub fn initiate_model_load(&self) {
        let mut is_loading = self.is_loading.lock().unwrap();
        if *is_loading || self.is_model_loaded() {
            return;
        }
        *is_loading = true;
        let this = self.clone();
        thread::spawn(move || {
            let settings = get_settings(&this.app_handle);
            if let Err(e) = this.load_model(&settings.selected_model) {
                error!("Failed to load model: {}", e);
            }
            *this.is_loading.lock().unwrap() = false;
            this.loading_condvar.notify_all();
        });
    }

    pub fn get_current_model(&self) -> Option<String> {
        self.current_model_id.lock().unwrap().clone()
    }

    // ---- jargon (transcription.rs:50-87): profile table = built-ins + the user's packs; the enabled profile ids are
    //      replaced by (or blended with) the domain selector's pick for the given context text ----
    fn profiles_map(settings: &AppSettings) -> HashMap<String, JargonProfile> {
        let mut profiles = builtin_profiles();
        for pack in &settings.jargon_packs {
            profiles.insert(
                pack.id.clone(),
                JargonProfile { label: pack.label.clone(), terms: pack.terms.clone(), corrections: pack.corrections.clone() },
            );
        }
        profiles
    }

    fn effective_profile_ids(&self, settings: &AppSettings, context_text: &str) -> Vec<String> {
        let mut ids = settings.jargon_enabled_profiles.clone();
        if let Some(selector) = self.app_handle.try_state::<Arc<DomainSelectorManager>>() {
            let picked = selector.select_profiles_with_timeout(settings, &DomainContext { text: context_text.to_string() });
            if let Some(auto_profiles) = picked {
                if settings.domain_selector_blend_manual_profiles {
                    for p in auto_profiles {
                        if !ids.contains(&p) {
                            ids.push(p);
                        }
                    }
                } else {
                    ids = auto_profiles;
                }
            }
        }
        ids
    }

    fn jargon_settings(&self, settings: &AppSettings, context_text: &str) -> JargonSettings {
        JargonSettings {
            enabled_profiles: self.effective_profile_ids(settings, context_text),
            custom_terms: settings.jargon_custom_terms.clone(),
            custom_corrections: settings.jargon_custom_corrections.clone(),
        }
    }

    /// WhisperInferenceParams::initial_prompt of the reference (transcription.rs:461-492): the jargon dictionary's terms.
    fn jargon_initial_prompt(&self, settings: &AppSettings) -> Option<String> {
        let active = !settings.jargon_enabled_profiles.is_empty()
            || !settings.jargon_custom_terms.is_empty()
            || !settings.jargon_packs.is_empty();
        if !active {
            return None;
        }
        let dict = compute_active_dictionary(&self.jargon_settings(settings, ""), &Self::profiles_map(settings));
        if dict.terms.is_empty() {
            return None;
        }
        let prompt = build_initial_prompt(&dict);
        if prompt.is_empty() {
            None
        } else {
            debug!("Jargon initial_prompt ({} chars)", prompt.len());
            Some(prompt)
        }
    }

    /// transcription.rs:538-580: custom words, filler / hallucination filter, jargon corrections.
    fn post_filter(&self, text: String, settings: &AppSettings) -> String {
        let corrected = if settings.custom_words.is_empty() {
            text
        } else {
            apply_custom_words(&text, &settings.custom_words, settings.word_correction_threshold)
        };
        let filtered = filter_transcription_output(&corrected);
        let jargon_on = !settings.jargon_enabled_profiles.is_empty()
            || !settings.jargon_custom_corrections.is_empty()
            || !settings.jargon_packs.is_empty();
        if !jargon_on {
            return filtered;
        }
        let dict = compute_active_dictionary(&self.jargon_settings(settings, &filtered), &Self::profiles_map(settings));
        if dict.corrections.is_empty() {
            filtered
        } else {
            apply_corrections(&filtered, &dict.corrections)
        }
    }

    /// language / task / prompt of one call -> sb_params (the CStrings must outlive the call)
    fn params(settings: &AppSettings, lang: &Option<CString>, prompt: &Option<CString>) -> sys::sb_params {
        let mut p: sys::sb_params = unsafe { std::mem::zeroed() };
        unsafe { sys::sb_params_default(&mut p) };
        p.language = lang.as_ref().map_or(std::ptr::null(), |c| c.as_ptr());      // NULL = "auto": detected per clip
        p.translate = settings.translate_to_english as i32;
        p.initial_prompt = prompt.as_ref().map_or(std::ptr::null(), |c| c.as_ptr());
        p
    }

    fn whisper_language(settings: &AppSettings) -> Result<Option<CString>> {
        Ok(match settings.selected_language.as_str() {
            "auto" => None,
            "zh-Hans" | "zh-Hant" => Some(CString::new("zh")?),        // Whisper uses ISO 639-1 codes (:448-459)
            other => Some(CString::new(other)?),
        })
    }

    fn wait_until_loaded(&self) -> Result<()> {
        let mut is_loading = self.is_loading.lock().unwrap();
        while *is_loading {
            is_loading = self.loading_condvar.wait(is_loading).unwrap();
        }
        drop(is_loading);
        if self.engine.lock().unwrap().is_none() {
            return Err(anyhow!(NOT_LOADED));
        }
        Ok(())
    }

    unsafe fn take_text(r: &mut sys::sb_result) -> String {
        let s = if r.text.is_null() {
            String::new()
        } else {
            String::from_utf8_lossy(std::slice::from_raw_parts(r.text as *const u8, r.text_len)).into_owned()
        };
        sys::sb_result_free(r);
        s
    }

    pub fn transcribe(&self, audio: Vec<f32>) -> Result<String> {
        self.last_activity.store(now_ms(), Ordering::Relaxed);
        let st = Instant::now();
        debug!("Audio vector length: {}", audio.len());
        if audio.is_empty() {
            self.maybe_unload_immediately("empty audio");
            return Ok(String::new());
        }
        self.wait_until_loaded()?;
        let settings = get_settings(&self.app_handle);
        let raw_text = {
            let guard = self.engine.lock().unwrap();          // held for the whole inference, like the reference (:437)
            let engine = guard.as_ref().ok_or_else(|| {
                anyhow!("Model failed to load after auto-load attempt. Please check your model settings.")
            })?;
            let lang = Self::whisper_language(&settings)?;
            let prompt = match self.jargon_initial_prompt(&settings) {
                Some(p) => Some(CString::new(p)?),
                None => None,
            };
            let p = Self::params(&settings, &lang, &prompt);
            let mut r: sys::sb_result = unsafe { std::mem::zeroed() };
            let rc = unsafe { sys::sb_transcribe(engine.0, audio.as_ptr(), audio.len(), &p, &mut r) };
            if rc != 0 {
                return Err(anyhow!("Whisper transcription failed: {}", last_error()));
            }
            unsafe { Self::take_text(&mut r) }
        };
        let final_result = self.post_filter(raw_text, &settings);
        info!(
            "Transcription completed in {}ms{}",
            st.elapsed().as_millis(),
            if settings.translate_to_english { " (translated)" } else { "" }
        );
        self.maybe_unload_immediately("transcription");
        Ok(final_result)
    }

    /// Additive: independent clips batched on the engine's GPU(s) (sb_transcribe_batch splits them over the configured
    /// devices, one worker thread per device, no collective).  One Result per clip, in order.
    pub fn transcribe_batch(&self, clips: Vec<Vec<f32>>) -> Vec<Result<String>> {
        self.last_activity.store(now_ms(), Ordering::Relaxed);
        if let Err(e) = self.wait_until_loaded() {
            let msg = e.to_string();
            return clips.iter().map(|_| Err(anyhow!(msg.clone()))).collect();
        }
        let settings = get_settings(&self.app_handle);
        let texts: Result<Vec<String>> = (|| {
            let guard = self.engine.lock().unwrap();
            let engine = guard.as_ref().ok_or_else(|| anyhow!(NOT_LOADED))?;
            let lang = Self::whisper_language(&settings)?;
            let prompt = match self.jargon_initial_prompt(&settings) {
                Some(p) => Some(CString::new(p)?),
                None => None,
            };
            let p = Self::params(&settings, &lang, &prompt);
            let ptrs: Vec<*const f32> = clips.iter().map(|c| c.as_ptr()).collect();
            let lens: Vec<usize> = clips.iter().map(|c| c.len()).collect();
            let mut res: Vec<sys::sb_result> = (0..clips.len()).map(|_| unsafe { std::mem::zeroed() }).collect();
            let rc = unsafe { sys::sb_transcribe_batch(engine.0, ptrs.as_ptr(), lens.as_ptr(), clips.len(), &p, res.as_mut_ptr()) };
            if rc != 0 {
                return Err(anyhow!("Whisper transcription failed: {}", last_error()));
            }
            Ok(res.iter_mut().map(|r| unsafe { Self::take_text(r) }).collect())
        })();
        let out = match texts {
            Ok(v) => v.into_iter().map(|t| Ok(self.post_filter(t, &settings))).collect(),
            Err(e) => {
                let msg = e.to_string();
                clips.iter().map(|_| Err(anyhow!(msg.clone()))).collect()
            }
        };
        self.maybe_unload_immediately("transcription");
        out
    }
}

impl Drop for TranscriptionManager {
    fn drop(&mut self) {
        debug!("Shutting down TranscriptionManager");
        self.shutdown_signal.store(true, Ordering::Relaxed);
        if let Some(handle) = self.watcher_handle.lock().unwrap().take() {
            if handle.join().is_err() {
                warn!("Failed to join idle watcher thread");
            }
        }
    }
}
