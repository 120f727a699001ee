// C++ mirror of the reference TranscriptionManager (see transcription_manager.hpp).
#include "transcription_manager.hpp"
#include <algorithm>
#include "text_filters.hpp"

#include <cstring>

namespace sb {

uint64_t TranscriptionManager::now_ms() {
    using namespace std::chrono;
    return (uint64_t)duration_cast<milliseconds>(system_clock::now().time_since_epoch()).count();
}

std::optional<uint64_t> TranscriptionManager::unload_limit_seconds(ModelUnloadTimeout t) {
    switch (t) {   // settings.rs ModelUnloadTimeout -> seconds (transcription.rs:112-163)
        case ModelUnloadTimeout::Never: return std::nullopt;
        case ModelUnloadTimeout::Immediately: return 0;
        case ModelUnloadTimeout::Sec5: return 5;
        case ModelUnloadTimeout::Min2: return 120;
        case ModelUnloadTimeout::Min5: return 300;
        case ModelUnloadTimeout::Min10: return 600;
        case ModelUnloadTimeout::Min15: return 900;
        case ModelUnloadTimeout::Hour1: return 3600;
    }
    return std::nullopt;
}

TranscriptionManager::TranscriptionManager(ModelResolver resolver, SettingsFn get_settings, EventFn on_model_state)
    : resolver_(std::move(resolver)), get_settings_(std::move(get_settings)), on_model_state_(std::move(on_model_state)),
      last_activity_(now_ms()) {
    // idle watcher: checks every 10 s, unloads after the configured idle time (transcription.rs:112-163)
    watcher_ = std::thread([this] {
        while (!shutdown_.load()) {
            for (int i = 0; i < 100 && !shutdown_.load(); ++i) std::this_thread::sleep_for(std::chrono::milliseconds(100));
            if (shutdown_.load()) break;
            const Settings s = get_settings_();
            const auto limit = unload_limit_seconds(s.model_unload_timeout);
            if (!limit || s.model_unload_timeout == ModelUnloadTimeout::Immediately) continue;
            if (now_ms() - last_activity_.load() > *limit * 1000 && is_model_loaded() && unload_model().ok)
                emit("unloaded");          // (the reference emits from unload_model AND from the watcher, :139-147)
        }
    });
}

TranscriptionManager::~TranscriptionManager() {
    shutdown_.store(true);
    if (watcher_.joinable()) watcher_.join();
    for (auto& t : loaders_) if (t.joinable()) t.join();
    std::lock_guard<std::mutex> g(engine_mu_);
    if (engine_) { sb_engine_destroy(engine_); engine_ = nullptr; }
}

bool TranscriptionManager::is_model_loaded() {
    std::lock_guard<std::mutex> g(engine_mu_);
    return engine_ != nullptr;
}

Result<Unit> TranscriptionManager::unload_model() {
    {
        std::lock_guard<std::mutex> g(engine_mu_);
        if (engine_) { sb_engine_destroy(engine_); engine_ = nullptr; }   // drop the engine to free memory
    }
    {
        std::lock_guard<std::mutex> g(model_mu_);
        current_model_id_.reset();
    }
    emit("unloaded");
    return Result<Unit>::Ok({});
}

void TranscriptionManager::maybe_unload_immediately(const std::string&) {
    const Settings s = get_settings_();
    if (s.model_unload_timeout == ModelUnloadTimeout::Immediately && is_model_loaded()) unload_model();
}

Result<Unit> TranscriptionManager::load_model(const std::string& model_id) {
    emit("loading_started", model_id);
    const auto path = resolver_(model_id);
    if (!path) return Result<Unit>::Err("Model not found: " + model_id);
    const Settings s = get_settings_();
    sb_config cfg{};
    cfg.model_path = path->c_str();
    cfg.device = s.device;
    cfg.max_batch = s.max_batch;
    cfg.dtype = s.dtype;
    cfg.use_cuda_graph = 1;
    if (!s.devices.empty()) { cfg.devices = s.devices.data(); cfg.n_devices = (int)s.devices.size(); }
    sb_engine* e = nullptr;
    if (sb_engine_create(&cfg, &e) != SB_OK) {
        const std::string msg = std::string("Failed to load whisper model ") + model_id + ": " + sb_last_error();
        emit("loading_failed", model_id, msg);
        return Result<Unit>::Err(msg);
    }
    {
        std::lock_guard<std::mutex> g(engine_mu_);
        if (engine_) sb_engine_destroy(engine_);
        engine_ = e;
    }
    {
        std::lock_guard<std::mutex> g(model_mu_);
        current_model_id_ = model_id;
    }
    emit("loaded", model_id);
    return Result<Unit>::Ok({});
}

void TranscriptionManager::initiate_model_load() {
    {
        std::lock_guard<std::mutex> g(loading_mu_);
        if (is_loading_ || is_model_loaded()) return;       // transcription.rs:374-380
        is_loading_ = true;
    }
    loaders_.emplace_back([this] {
        const Settings s = get_settings_();
        load_model(s.selected_model);                        // failure leaves the engine empty
        {
            std::lock_guard<std::mutex> g(loading_mu_);
            is_loading_ = false;
        }
        loading_cv_.notify_all();
    });
}

std::optional<std::string> TranscriptionManager::get_current_model() {
    std::lock_guard<std::mutex> g(model_mu_);
    return current_model_id_;
}

std::string TranscriptionManager::effective_language(const Settings& s) const {
    // "auto" => None; zh-Hans / zh-Hant => "zh" (transcription.rs:448-459)
    if (s.selected_language == "zh-Hans" || s.selected_language == "zh-Hant") return "zh";
    return s.selected_language;
}

// transcription.rs:65-87: the enabled profiles, replaced by (or blended with) the domain selector's pick
static std::vector<std::string> effective_profile_ids(const Settings& s, const std::string& context_text) {
    std::vector<std::string> ids = s.jargon_enabled_profiles;
    if (s.profile_selector) {
        if (auto picked = s.profile_selector(context_text)) {
            if (s.domain_selector_blend_manual_profiles) {
                for (const auto& p : *picked)
                    if (std::find(ids.begin(), ids.end(), p) == ids.end()) ids.push_back(p);
            } else {
                ids = *picked;
            }
        }
    }
    return ids;
}

// transcription.rs:461-492: the jargon dictionary's terms as Whisper's initial_prompt ("" = none)
static std::string jargon_initial_prompt(const Settings& s) {
    if (s.jargon_enabled_profiles.empty() && s.jargon_custom_terms.empty() && s.jargon_packs.empty()) return "";
    JargonSettings js{effective_profile_ids(s, ""), s.jargon_custom_terms, s.jargon_custom_corrections};
    const ActiveDictionary d = compute_active_dictionary(js, s.jargon_profiles);
    if (d.terms.empty()) return "";
    return build_initial_prompt(d);
}

// transcription.rs:537-580: custom-word correction (only when configured), the filler / stutter / hallucination filter,
// then the jargon corrections (only when profiles or custom corrections are configured)
static std::string post_filter(std::string text, const Settings& s) {
    if (!s.custom_words.empty()) text = apply_custom_words(text, s.custom_words, s.word_correction_threshold);
    text = filter_transcription_output(text);
    if (!s.jargon_enabled_profiles.empty() || !s.jargon_custom_corrections.empty() || !s.jargon_packs.empty()) {
        JargonSettings js{effective_profile_ids(s, text), s.jargon_custom_terms, s.jargon_custom_corrections};
        const ActiveDictionary d = compute_active_dictionary(js, s.jargon_profiles);
        if (!d.corrections.empty()) text = apply_corrections(text, d.corrections);
    }
    return text;
}

Result<std::string> TranscriptionManager::transcribe(std::vector<float> audio) {
    last_activity_.store(now_ms());
    if (audio.empty()) {                                      // transcription.rs:412-416
        maybe_unload_immediately("empty audio");
        return Result<std::string>::Ok("");
    }
    {
        std::unique_lock<std::mutex> lk(loading_mu_);         // wait while the model is loading
        loading_cv_.wait(lk, [this] { return !is_loading_; });
    }
    const Settings s = get_settings_();
    std::string text;
    {
        std::lock_guard<std::mutex> g(engine_mu_);            // held for the whole inference (transcription.rs:437)
        if (!engine_) return Result<std::string>::Err("Model is not loaded for transcription.");
        sb_params p;
        sb_params_default(&p);
        const std::string lang = effective_language(s);
        p.language = lang == "auto" ? nullptr : lang.c_str();
        p.translate = s.translate_to_english ? 1 : 0;
        const std::string prompt = jargon_initial_prompt(s);
        p.initial_prompt = prompt.empty() ? nullptr : prompt.c_str();
        sb_result r;
        if (sb_transcribe(engine_, audio.data(), audio.size(), &p, &r) != SB_OK)
            return Result<std::string>::Err(std::string("Whisper transcription failed: ") + sb_last_error());
        text.assign(r.text ? r.text : "", r.text_len);
        sb_result_free(&r);
    }
    text = post_filter(text, s);
    maybe_unload_immediately("transcription");
    return Result<std::string>::Ok(std::move(text));
}

std::vector<Result<std::string>> TranscriptionManager::transcribe_batch(const std::vector<std::vector<float>>& clips) {
    last_activity_.store(now_ms());
    std::vector<Result<std::string>> out(clips.size());
    {
        std::unique_lock<std::mutex> lk(loading_mu_);
        loading_cv_.wait(lk, [this] { return !is_loading_; });
    }
    const Settings s = get_settings_();
    std::lock_guard<std::mutex> g(engine_mu_);
    if (!engine_) {
        for (auto& r : out) r = Result<std::string>::Err("Model is not loaded for transcription.");
        return out;
    }
    std::vector<const float*> ptrs(clips.size());
    std::vector<size_t> ns(clips.size());
    for (size_t i = 0; i < clips.size(); ++i) { ptrs[i] = clips[i].data(); ns[i] = clips[i].size(); }
    sb_params p;
    sb_params_default(&p);
    const std::string lang = effective_language(s);
    p.language = lang == "auto" ? nullptr : lang.c_str();
    p.translate = s.translate_to_english ? 1 : 0;
    const std::string prompt = jargon_initial_prompt(s);
    p.initial_prompt = prompt.empty() ? nullptr : prompt.c_str();
    std::vector<sb_result> res(clips.size());
    if (sb_transcribe_batch(engine_, ptrs.data(), ns.data(), clips.size(), &p, res.data()) != SB_OK) {
        const std::string e = std::string("Whisper transcription failed: ") + sb_last_error();
        for (auto& r : out) r = Result<std::string>::Err(e);
    } else {
        for (size_t i = 0; i < clips.size(); ++i) {
            std::string text(res[i].text ? res[i].text : "", res[i].text_len);
            out[i] = Result<std::string>::Ok(post_filter(text, s));
        }
    }
    for (auto& r : res) sb_result_free(&r);
    return out;
}

}  // namespace sb
