// transcription_manager.hpp -- C++ host-side mirror of the reference's TranscriptionManager
// (src-tauri/src/managers/transcription.rs) on top of the C ABI in include/spittle_b200.h.
//
// The reference is Rust; no Rust toolchain exists in this image, so the host side above the
// C ABI is written in C++ (rust/transcription_b200.rs shows the same thing as the drop-in
// Rust module a maintainer would add).  Same public surface, argument meaning and error
// behaviour as the reference:
//
//   new(app_handle, model_manager)      transcription.rs:89    -> TranscriptionManager(ModelResolver, Settings)
//   is_model_loaded()                   transcription.rs:170
//   unload_model()                      transcription.rs:175
//   maybe_unload_immediately(context)   transcription.rs:211
//   load_model(model_id)                transcription.rs:223
//   initiate_model_load()               transcription.rs:374
//   get_current_model()                 transcription.rs:393
//   transcribe(audio) -> Result<String> transcription.rs:398
//   Drop joins the idle watcher         transcription.rs:608-624
//
// Result<T> is modelled as sb::Result<T>{ok, value, error}; the error strings are the reference's.
#pragma once
#include "jargon.hpp"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <thread>
#include <vector>

#include "../include/spittle_b200.h"

namespace sb {

template <typename T>
struct Result {
    bool ok = false;
    T value{};
    std::string error;
    static Result Ok(T v) { Result r; r.ok = true; r.value = std::move(v); return r; }
    static Result Err(std::string e) { Result r; r.ok = false; r.error = std::move(e); return r; }
};
struct Unit {};

// settings.rs ModelUnloadTimeout
enum class ModelUnloadTimeout { Never, Immediately, Min2, Min5, Min10, Min15, Hour1, Sec5 };

// the hot-path-relevant subset of AppSettings (settings.rs:427-429,925): a plain struct replaces
// the tauri store
struct Settings {
    std::string selected_model;
    std::string selected_language = "auto";
    bool translate_to_english = false;
    ModelUnloadTimeout model_unload_timeout = ModelUnloadTimeout::Never;
    std::vector<std::string> custom_words;            // settings.rs custom_words
    double word_correction_threshold = 0.18;          // settings.rs:446-448
    // jargon (all empty by default): profile ids, user terms / corrections, and the profile table they index
    std::vector<std::string> jargon_enabled_profiles, jargon_custom_terms;
    std::vector<JargonCorrection> jargon_custom_corrections;
    std::map<std::string, JargonProfile> jargon_profiles;
    // ids of the user's jargon packs (settings.jargon_packs; build_profiles_map, transcription.rs:50-63, merges their
    // profiles into the table above): only their presence matters for the gates at :462-464 / :553-555
    std::vector<std::string> jargon_packs;
    // DomainSelectorManager stand-in (transcription.rs:65-87): (context text) -> profile ids, or nullopt
    std::function<std::optional<std::vector<std::string>>(const std::string&)> profile_selector;
    bool domain_selector_blend_manual_profiles = false;
    int device = 0;
    std::vector<int> devices;              // non-empty: one model replica per listed CUDA device (transcribe_batch)
    int max_batch = 64;
    int dtype = SB_DTYPE_F16;
};

class TranscriptionManager {
  public:
    // model_id -> path of a GGML .bin (ModelManager::get_model_path, model.rs:804-847)
    using ModelResolver = std::function<std::optional<std::string>(const std::string& model_id)>;
    using SettingsFn = std::function<Settings()>;
    // app_handle.emit("model-state-changed", ModelStateEvent{event_type, model_id, model_name, error}) stand-in
    // (domain/events.rs:3-43; emitted at transcription.rs:139-147, 195-199, 227-236, 263-275, 357-366):
    // event_type is "loading_started" | "loading_failed" | "loaded" | "unloaded"
    using EventFn = std::function<void(const std::string& event_type, const std::string& model_id, const std::string& error)>;

    TranscriptionManager(ModelResolver resolver, SettingsFn get_settings, EventFn on_model_state = nullptr);
    ~TranscriptionManager();
    TranscriptionManager(const TranscriptionManager&) = delete;
    TranscriptionManager& operator=(const TranscriptionManager&) = delete;

    bool is_model_loaded();
    Result<Unit> unload_model();
    void maybe_unload_immediately(const std::string& context);
    Result<Unit> load_model(const std::string& model_id);
    void initiate_model_load();
    std::optional<std::string> get_current_model();
    Result<std::string> transcribe(std::vector<float> audio);
    // additive (SURVEY 8(b) "Batch / multi-GPU surface"): independent clips on this manager's GPU
    std::vector<Result<std::string>> transcribe_batch(const std::vector<std::vector<float>>& clips);

  private:
    static uint64_t now_ms();
    static std::optional<uint64_t> unload_limit_seconds(ModelUnloadTimeout t);
    std::string effective_language(const Settings& s) const;

    void emit(const char* event_type, const std::string& model_id = "", const std::string& error = "") {
        if (on_model_state_) on_model_state_(event_type, model_id, error);
    }
    ModelResolver resolver_;
    SettingsFn get_settings_;
    EventFn on_model_state_;
    std::mutex engine_mu_;                 // engine: Arc<Mutex<Option<LoadedEngine>>>
    sb_engine* engine_ = nullptr;
    std::mutex model_mu_;
    std::optional<std::string> current_model_id_;
    std::atomic<uint64_t> last_activity_;
    std::atomic<bool> shutdown_{false};
    std::thread watcher_;
    std::mutex loading_mu_;                // is_loading + loading_condvar
    bool is_loading_ = false;
    std::condition_variable loading_cv_;
    std::vector<std::thread> loaders_;
};

}  // namespace sb
