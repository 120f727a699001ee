//! Raw FFI bindings of include/spittle_b200.h (engine subset used by the drop-in manager).
//! NOT COMPILED IN THIS IMAGE: there is no Rust toolchain; kept mechanical so a maintainer can
//! `cargo build` it next to the prebuilt libspittle_b200.so.
//!
//! Layout guard: the `const _: () = assert!(...)` lines below pin sizeof / offsetof of every struct to the numbers the
//! library itself reports through `sb_abi_layout()`; the same table is checked in as tests/golden/abi_layout.json and
//! tests/test_abi.py compares it with the built library, with the ctypes mirror and with THIS FILE (it parses the
//! asserts), so a field added to the C header without updating the bindings fails the CPU test suite.
#![allow(non_camel_case_types)]
use core::mem::{offset_of, size_of};
use libc::{c_char, c_float, c_int, size_t};

#[repr(C)]
pub struct sb_engine { _private: [u8; 0] }

#[repr(C)]
pub struct sb_config {
    pub model_path: *const c_char,
    pub device: c_int,
    pub max_batch: c_int,
    pub dtype: c_int,          // 0 bf16, 1 f16
    pub use_cuda_graph: c_int,
    pub devices: *const c_int, // NULL or n_devices CUDA ordinals: one replica per device
    pub n_devices: c_int,
}

#[repr(C)]
pub struct sb_params {
    pub language: *const c_char,        // NULL = "auto"
    pub translate: c_int,
    pub initial_prompt: *const c_char,  // NULL or UTF-8 (jargon prompt)
    pub no_timestamps: c_int,
    pub suppress_blank: c_int,
    pub single_segment: c_int,
    pub max_initial_ts: c_float,
    pub n_max_tokens: c_int,
    pub max_windows: c_int,
    pub n_max_text_ctx: c_int,
    pub temperature: c_float,
    pub temperature_inc: c_float,       // 0 = no temperature fallback (sb_params_default)
    pub logprob_thold: c_float,
    pub entropy_thold: c_float,
    pub suppress_nst: c_int,            // whisper_full_params.suppress_nst (transcribe-rs: suppress_non_speech_tokens); default 0
}

#[repr(C)]
pub struct sb_window_info {
    pub seek: i32, pub n_tokens: i32, pub result_len: i32, pub seek_delta: i32, pub failed: i32, pub token_offset: i32,
    pub n_prompt: i32,
    pub temperature: c_float, pub n_attempts: i32, pub avg_logprob: c_float,
}

#[repr(C)]
pub struct sb_segment {
    pub t0: i64, pub t1: i64,           // 10 ms units
    pub text: *const c_char, pub text_len: size_t,
    pub token_offset: i32, pub n_tokens: i32,
}

#[repr(C)]
pub struct sb_result {
    pub text: *mut c_char, pub text_len: size_t,
    pub tokens: *mut i32, pub n_tokens: size_t,
    pub sampled: *mut i32, pub n_sampled: size_t,
    pub margins: *mut c_float,
    pub tids: *mut i32,
    pub logprobs: *mut c_float,
    pub windows: *mut sb_window_info, pub n_windows: size_t,
    pub segments: *mut sb_segment, pub n_segments: size_t,
    pub segment_text: *mut c_char,
    pub ms_mel: c_float, pub ms_encode: c_float, pub ms_decode: c_float,
    pub status: c_int,
    pub lang_id: c_int,
}

#[repr(C)]
pub struct sb_abi_field { pub struct_name: *const c_char, pub field: *const c_char, pub struct_size: c_int, pub offset: c_int }

// ---- layout guard (numbers = sb_abi_layout() of the library; tests/golden/abi_layout.json) ----
const _: () = assert!(size_of::<sb_config>() == 40);
const _: () = assert!(offset_of!(sb_config, devices) == 24);
const _: () = assert!(offset_of!(sb_config, n_devices) == 32);
const _: () = assert!(size_of::<sb_params>() == 72);
const _: () = assert!(offset_of!(sb_params, initial_prompt) == 16);
const _: () = assert!(offset_of!(sb_params, max_initial_ts) == 36);
const _: () = assert!(offset_of!(sb_params, n_max_text_ctx) == 48);
const _: () = assert!(offset_of!(sb_params, suppress_nst) == 68);
const _: () = assert!(size_of::<sb_window_info>() == 40);
const _: () = assert!(offset_of!(sb_window_info, n_prompt) == 24);
const _: () = assert!(size_of::<sb_segment>() == 40);
const _: () = assert!(offset_of!(sb_segment, text) == 16);
const _: () = assert!(offset_of!(sb_segment, n_tokens) == 36);
const _: () = assert!(size_of::<sb_result>() == 136);
const _: () = assert!(offset_of!(sb_result, margins) == 48);
const _: () = assert!(offset_of!(sb_result, tids) == 56);
const _: () = assert!(offset_of!(sb_result, windows) == 72);
const _: () = assert!(offset_of!(sb_result, segments) == 88);
const _: () = assert!(offset_of!(sb_result, segment_text) == 104);
const _: () = assert!(offset_of!(sb_result, ms_mel) == 112);
const _: () = assert!(offset_of!(sb_result, status) == 124);
const _: () = assert!(offset_of!(sb_result, lang_id) == 128);

extern "C" {
    pub fn sb_last_error() -> *const c_char;
    pub fn sb_abi_layout(out: *mut sb_abi_field, cap: c_int) -> c_int;
    pub fn sb_params_default(p: *mut sb_params);
    pub fn sb_engine_create(cfg: *const sb_config, out: *mut *mut sb_engine) -> c_int;
    pub fn sb_engine_destroy(e: *mut sb_engine) -> c_int;
    pub fn sb_engine_device_count(e: *const sb_engine) -> c_int;
    pub fn sb_tokenize(e: *const sb_engine, text: *const c_char, tokens: *mut i32, cap: c_int) -> c_int;
    pub fn sb_transcribe(e: *mut sb_engine, pcm16k: *const c_float, n_samples: size_t, p: *const sb_params,
                         out: *mut sb_result) -> c_int;
    pub fn sb_transcribe_batch(e: *mut sb_engine, pcm16k: *const *const c_float, n_samples: *const size_t,
                               count: size_t, p: *const sb_params, out: *mut sb_result) -> c_int;
    pub fn sb_result_free(r: *mut sb_result);
    // capture-side formats: history WAV payload, mic-level visualiser; device pointers
    pub fn sb_pcm_f32_to_i16(samples: *const c_float, out: *mut i16, n: size_t) -> c_int;
    pub fn sb_visualiser_levels(pcm: *const c_float, n_samples: size_t, chunk_len: c_int, sample_rate: c_int,
                                out: *mut c_float, n_chunks_out: *mut c_int) -> c_int;
    pub fn sb_pcm_f32_to_i16_dev(input: *const c_float, out: *mut i16, n: size_t, stream: *mut core::ffi::c_void) -> c_int;
    pub fn sb_visualiser_levels_dev(pcm: *const c_float, stream_stride: i64, n_streams: c_int, n_chunks: c_int,
                                    chunk_len: c_int, sample_rate: c_int, out: *mut c_float,
                                    stream: *mut core::ffi::c_void) -> c_int;
}
