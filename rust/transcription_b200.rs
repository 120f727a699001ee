// Drop-in `managers::transcription` module backed by libspittle_b200.so.
// NOT COMPILED IN THIS IMAGE (no cargo/rustc) -- mirrors host/transcription_manager.cpp line for line.
//
// Wiring in the reference (src-tauri/src/managers/mod.rs:8-12 already swaps this module by cargo
// feature for the CI mock):
//
//     #[cfg(feature = "b200_transcription")]
//     #[path = "transcription_b200.rs"]
//     pub mod transcription;
//
// and in src-tauri/Cargo.toml:  b200_transcription = ["dep:spittle-b200-sys"]
//
// Public surface identical to src-tauri/src/managers/transcription.rs:89-605 (and to
// transcription_mock.rs:25-55): new, is_model_loaded, unload_model, maybe_unload_immediately,
// load_model, initiate_model_load, get_current_model, transcribe.
use crate::audio_toolkit::{apply_custom_words, filter_transcription_output};
use crate::managers::model::ModelManager;
use crate::settings::{get_settings, ModelUnloadTimeout};
use anyhow::Result;
use spittle_b200_sys as sys;
use std::ffi::{CStr, CString};
use std::sync::{Arc, Condvar, Mutex};
use tauri::AppHandle;

struct Engine(*mut sys::sb_engine);
unsafe impl Send for Engine {}
impl Drop for Engine {
    fn drop(&mut self) { unsafe { sys::sb_engine_destroy(self.0); } }
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::sb_last_error()).to_string_lossy().into_owned() }
}

#[derive(Clone)]
pub struct TranscriptionManager {
    engine: Arc<Mutex<Option<Engine>>>,
    model_manager: Arc<ModelManager>,
    app_handle: AppHandle,
    current_model_id: Arc<Mutex<Option<String>>>,
    is_loading: Arc<Mutex<bool>>,
    loading_condvar: Arc<Condvar>,
}

impl TranscriptionManager {
    pub fn new(app_handle: &AppHandle, model_manager: Arc<ModelManager>) -> Result<Self> {
        Ok(Self {
            engine: Arc::new(Mutex::new(None)), model_manager, app_handle: app_handle.clone(),
            current_model_id: Arc::new(Mutex::new(None)), is_loading: Arc::new(Mutex::new(false)),
            loading_condvar: Arc::new(Condvar::new()),
        })
    }
    pub fn is_model_loaded(&self) -> bool { self.engine.lock().unwrap().is_some() }
    pub fn unload_model(&self) -> Result<()> {
        *self.engine.lock().unwrap() = None;               // Drop frees device memory
        *self.current_model_id.lock().unwrap() = None;
        Ok(())
    }
    pub fn maybe_unload_immediately(&self, _context: &str) {
        let settings = get_settings(&self.app_handle);
        if settings.model_unload_timeout == ModelUnloadTimeout::Immediately && self.is_model_loaded() {
            let _ = self.unload_model();
        }
    }
    pub fn load_model(&self, model_id: &str) -> Result<()> {
        let path = self.model_manager.get_model_path(model_id)?;
        let cpath = CString::new(path.to_string_lossy().as_bytes())?;
        let cfg = sys::sb_config { model_path: cpath.as_ptr(), device: 0, max_batch: 64, dtype: 1, use_cuda_graph: 1 };
        let mut e: *mut sys::sb_engine = std::ptr::null_mut();
        if unsafe { sys::sb_engine_create(&cfg, &mut e) } != 0 {
            return Err(anyhow::anyhow!("Failed to load whisper model {}: {}", model_id, last_error()));
        }
        *self.engine.lock().unwrap() = Some(Engine(e));
        *self.current_model_id.lock().unwrap() = Some(model_id.to_string());
        Ok(())
    }
    pub fn initiate_model_load(&self) {
        let mut is_loading = self.is_loading.lock().unwrap();
        if *is_loading || self.is_model_loaded() { return; }
        *is_loading = true;
        let this = self.clone();
        std::thread::spawn(move || {
            let settings = get_settings(&this.app_handle);
            let _ = this.load_model(&settings.selected_model);
            *this.is_loading.lock().unwrap() = false;
            this.loading_condvar.notify_all();
        });
    }
    pub fn get_current_model(&self) -> Option<String> { self.current_model_id.lock().unwrap().clone() }

    pub fn transcribe(&self, audio: Vec<f32>) -> Result<String> {
        if audio.is_empty() {
            self.maybe_unload_immediately("empty audio");
            return Ok(String::new());
        }
        {
            let mut is_loading = self.is_loading.lock().unwrap();
            while *is_loading { is_loading = self.loading_condvar.wait(is_loading).unwrap(); }
        }
        let settings = get_settings(&self.app_handle);
        let text = {
            let guard = self.engine.lock().unwrap();
            let engine = guard.as_ref().ok_or_else(|| anyhow::anyhow!("Model is not loaded for transcription."))?;
            let lang = match settings.selected_language.as_str() {
                "auto" => None,
                "zh-Hans" | "zh-Hant" => Some(CString::new("zh")?),
                other => Some(CString::new(other)?),
            };
            let mut p: sys::sb_params = unsafe { std::mem::zeroed() };
            unsafe { sys::sb_params_default(&mut p) };
            p.language = lang.as_ref().map_or(std::ptr::null(), |c| c.as_ptr());
            p.translate = settings.translate_to_english as i32;
            let mut r: sys::sb_result = unsafe { std::mem::zeroed() };
            let rc = unsafe { sys::sb_transcribe(engine.0, audio.as_ptr(), audio.len(), &p, &mut r) };
            if rc != 0 { return Err(anyhow::anyhow!("Whisper transcription failed: {}", last_error())); }
            let bytes = unsafe { std::slice::from_raw_parts(r.text as *const u8, r.text_len) };
            let s = String::from_utf8_lossy(bytes).into_owned();
            unsafe { sys::sb_result_free(&mut r) };
            s
        };
        // unchanged reference post-filters (transcription.rs:538-549)
        let corrected = if !settings.custom_words.is_empty() {
            apply_custom_words(&text, &settings.custom_words, settings.word_correction_threshold)
        } else { text };
        let filtered = filter_transcription_output(&corrected);
        self.maybe_unload_immediately("transcription");
        Ok(filtered)
    }
}
