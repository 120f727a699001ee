// Minimal host program over the C++ TranscriptionManager mirror.
//   sb_transcribe_cli <model.bin> <clip.f32> [language]       raw little-endian f32, 16 kHz mono
// Prints the transcription (exit 0) or the error (exit 1).  Used by tests/test_host_cpp_gpu.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <iostream>

#include "jargon.hpp"
#include "text_filters.hpp"
#include "transcription_manager.hpp"

int main(int argc, char** argv) {
    // text post-filter modes (no GPU needed): stdin -> stdout
    //   sb_transcribe_cli --filter                         filter_transcription_output
    //   sb_transcribe_cli --custom-words THRESH w1 w2 ..   apply_custom_words (words are separate argv entries)
    if (argc >= 2 && (!std::strcmp(argv[1], "--filter") || !std::strcmp(argv[1], "--custom-words"))) {
        std::string text((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
        if (!std::strcmp(argv[1], "--filter")) { std::fputs(sb::filter_transcription_output(text).c_str(), stdout); return 0; }
        if (argc < 3) return 2;
        std::vector<std::string> words(argv + 3, argv + argc);
        std::fputs(sb::apply_custom_words(text, words, std::atof(argv[2])).c_str(), stdout);
        return 0;
    }
    //   sb_transcribe_cli --jargon-correct from1 to1 from2 to2 ..   compute_active_dictionary (custom corrections only) +
    //                                                               apply_corrections on stdin
    //   sb_transcribe_cli --jargon-prompt term1 term2 ..            build_initial_prompt of the custom terms
    if (argc >= 2 && !std::strcmp(argv[1], "--jargon-correct")) {
        std::string text((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
        sb::JargonSettings js;
        for (int i = 2; i + 1 < argc; i += 2) js.custom_corrections.push_back({argv[i], argv[i + 1]});
        const sb::ActiveDictionary d = sb::compute_active_dictionary(js, {});
        std::fputs(sb::apply_corrections(text, d.corrections).c_str(), stdout);
        return 0;
    }
    if (argc >= 2 && !std::strcmp(argv[1], "--jargon-prompt")) {
        sb::JargonSettings js;
        for (int i = 2; i < argc; ++i) js.custom_terms.push_back(argv[i]);
        std::fputs(sb::build_initial_prompt(sb::compute_active_dictionary(js, {})).c_str(), stdout);
        return 0;
    }
    if (argc < 3) { std::fprintf(stderr, "usage: %s model.bin clip.f32 [language]\n", argv[0]); return 2; }
    const std::string model = argv[1], clip = argv[2], lang = argc > 3 ? argv[3] : "en";
    sb::Settings st;
    st.selected_model = "cli-model";
    st.selected_language = lang;
    sb::TranscriptionManager tm([&](const std::string&) { return std::optional<std::string>(model); }, [&] { return st; });
    // not loaded yet: the reference returns Err("Model is not loaded for transcription.")
    auto early = tm.transcribe(std::vector<float>(16000, 0.f));
    std::printf("before load: %s\n", early.ok ? "ok" : early.error.c_str());
    auto empty = tm.transcribe({});
    std::printf("empty audio: ok=%d text='%s'\n", (int)empty.ok, empty.value.c_str());
    tm.initiate_model_load();                      // background load; transcribe() waits on the condvar
    std::ifstream f(clip, std::ios::binary);
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<float> pcm(raw.size() / 4);
    std::memcpy(pcm.data(), raw.data(), pcm.size() * 4);
    auto r = tm.transcribe(std::move(pcm));
    if (!r.ok) { std::printf("error: %s\n", r.error.c_str()); return 1; }
    std::printf("model: %s\n", tm.get_current_model().value_or("<none>").c_str());
    std::printf("text: %s\n", r.value.c_str());
    tm.unload_model();
    std::printf("loaded after unload: %d\n", (int)tm.is_model_loaded());
    return 0;
}
