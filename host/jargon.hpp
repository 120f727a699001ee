// jargon.hpp -- C++ host-side mirror of the jargon stage of the reference's transcribe()
// (src-tauri/src/managers/transcription.rs:461-489 initial prompt, :551-580 corrections; src-tauri/src/jargon.rs):
//   compute_active_dictionary   jargon.rs:506-592
//   build_initial_prompt        jargon.rs:594-627
//   apply_corrections           jargon.rs:682-716 (protected spans :637-676)
// Behavioural re-implementation pinned by the reference's unit tests (jargon.rs:741-961), replayed through
// host/sb_transcribe_cli by tests/test_jargon_cpu.py.  The built-in profile table (jargon.rs:39-505) is settings
// content and is passed in.  std::regex classes are ASCII: \w / \b do not see non-ASCII letters as word characters
// (the regex crate is Unicode-aware); `$` in a replacement is literal.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace sb {
struct JargonCorrection { std::string from, to; };
struct JargonProfile { std::string label; std::vector<std::string> terms; std::vector<JargonCorrection> corrections; };
struct JargonSettings {
    std::vector<std::string> enabled_profiles, custom_terms;
    std::vector<JargonCorrection> custom_corrections;
};
struct ActiveDictionary { std::vector<std::string> terms; std::vector<JargonCorrection> corrections; };

ActiveDictionary compute_active_dictionary(const JargonSettings& settings, const std::map<std::string, JargonProfile>& profiles);
std::string build_initial_prompt(const ActiveDictionary& dictionary);
std::string apply_corrections(const std::string& text, const std::vector<JargonCorrection>& corrections);
}  // namespace sb
