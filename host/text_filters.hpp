// text_filters.hpp -- C++ host-side mirror of the text post-filters that run inside the reference's transcribe()
// after the engine call (src-tauri/src/managers/transcription.rs:537-549):
//   apply_custom_words            src-tauri/src/audio_toolkit/text.rs:102-156
//   filter_transcription_output   src-tauri/src/audio_toolkit/text.rs:373-396
// Behavioural re-implementation (no code shared with the reference); pinned by the reference's own unit-test vectors
// (tests/golden/text_filters.json, replayed through host/sb_transcribe_cli by tests/test_text_filters_cpu.py).
// Character classes: ASCII letters / digits plus every byte >= 0x80 (so UTF-8 letters such as "è" stay inside words;
// Rust's Unicode `is_alphanumeric` additionally excludes non-ASCII punctuation -- not reproduced).
#pragma once
#include <string>
#include <vector>

namespace sb {
std::string apply_custom_words(const std::string& text, const std::vector<std::string>& custom_words, double threshold);
std::string filter_transcription_output(const std::string& text);
}  // namespace sb
