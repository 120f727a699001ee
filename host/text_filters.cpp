#include "text_filters.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <limits>
#include <regex>
#include <sstream>

namespace sb {
namespace {

bool is_alnum(unsigned char c) { return std::isalnum(c) || c >= 0x80; }
bool is_alpha(unsigned char c) { return std::isalpha(c) || c >= 0x80; }
std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }
std::vector<std::string> split_ws(const std::string& s) {
    std::istringstream is(s); std::vector<std::string> out; std::string w;
    while (is >> w) out.push_back(w);
    return out;
}
std::string join(const std::vector<std::string>& v) {
    std::string out;
    for (size_t i = 0; i < v.size(); ++i) { if (i) out += ' '; out += v[i]; }
    return out;
}
std::string trim(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e && std::isspace((unsigned char)s[b])) ++b;
    while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
    return s.substr(b, e - b);
}

// strsim::levenshtein over bytes (identical to the char version for ASCII; multi-byte letters count per byte)
size_t levenshtein(const std::string& a, const std::string& b) {
    std::vector<size_t> prev(b.size() + 1), cur(b.size() + 1);
    for (size_t j = 0; j <= b.size(); ++j) prev[j] = j;
    for (size_t i = 1; i <= a.size(); ++i) {
        cur[0] = i;
        for (size_t j = 1; j <= b.size(); ++j)
            cur[j] = std::min({prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (a[i - 1] != b[j - 1] ? 1 : 0)});
        std::swap(prev, cur);
    }
    return prev[b.size()];
}

// natural::phonetics::soundex code: first letter kept, h/w transparent, adjacent equal digits merged, vowels dropped
std::string soundex_code(const std::string& w) {
    if (w.empty()) return "0000";
    auto digit = [](char c) {
        switch (c) {
            case 'b': case 'f': case 'p': case 'v': return '1';
            case 'c': case 'g': case 'j': case 'k': case 'q': case 's': case 'x': case 'z': return '2';
            case 'd': case 't': return '3';
            case 'l': return '4';
            case 'm': case 'n': return '5';
            case 'r': return '6';
            case 'h': case 'w': return '9';
            default: return '0';
        }
    };
    std::string enc(1, w[0]);
    for (size_t i = 1; i < w.size(); ++i) { const char d = digit(w[i]); if (d != '9') enc += d; }
    std::string dd;
    for (char c : enc) if (dd.empty() || dd.back() != c) dd += c;
    std::string code;
    for (char c : dd) if (c != '0') code += c;
    code += "0000";
    return code.substr(0, 4);
}

std::string strip_non_alnum(const std::string& w) {
    size_t b = 0, e = w.size();
    while (b < e && !is_alnum((unsigned char)w[b])) ++b;
    while (e > b && !is_alnum((unsigned char)w[e - 1])) --e;
    return w.substr(b, e - b);
}

std::string preserve_case(const std::string& original, const std::string& replacement) {
    bool all_upper = true;
    for (unsigned char c : original) if (!std::isupper(c)) { all_upper = false; break; }
    if (all_upper) { std::string r = replacement; for (auto& c : r) c = (char)std::toupper((unsigned char)c); return r; }
    if (!original.empty() && std::isupper((unsigned char)original[0]) && !replacement.empty()) {
        std::string r = replacement; r[0] = (char)std::toupper((unsigned char)r[0]); return r;
    }
    return replacement;
}

const char* kFillers[] = {"uh", "um", "uhm", "umm", "uhh", "uhhh", "ah", "eh", "hmm", "hm", "mmm", "mm", "mh", "ha", "ehh"};
const char* kHallucinations[] = {"thank you for watching", "thanks for watching", "thank you for listening", "thanks for listening",
                                 "please subscribe", "like and subscribe", "see you next time", "see you in the next video",
                                 "bye bye", "bye", "thank you", "thanks", "subtitles by", "you"};

std::string collapse_stutters(const std::string& text) {
    const auto words = split_ws(text);
    if (words.empty()) return text;
    std::vector<std::string> out;
    for (size_t i = 0; i < words.size();) {
        const std::string wl = lower(words[i]);
        bool alpha = true;
        for (unsigned char c : wl) if (!is_alpha(c)) alpha = false;
        if (wl.size() <= 2 && alpha) {
            size_t n = 1;
            while (i + n < words.size() && lower(words[i + n]) == wl) ++n;
            out.push_back(words[i]);
            i += n >= 3 ? n : 1;
        } else { out.push_back(words[i]); ++i; }
    }
    return join(out);
}

bool is_hallucination(const std::string& text) {
    std::string stripped;
    for (unsigned char c : trim(text)) if (is_alnum(c) || std::isspace(c)) stripped += (char)c;
    const std::string norm = lower(trim(stripped));
    if (norm.empty()) return false;
    for (const char* p : kHallucinations) if (norm == p) return true;
    static const std::regex res[] = {
        std::regex(R"(^(for more information[,.]?\s*)?(visit|go to)\s+\S+(\s+(or\s+)?(visit|go to)\s+\S+)*(\s+for more information)?[.,]?\s*$)", std::regex::icase),
        std::regex(R"(^for more information[,.]?\s*(visit|go to)\s+\S+[.,]?\s*$)", std::regex::icase),
        std::regex(R"(^subtitles\s+(by|provided by|created by)\s+.*$)", std::regex::icase)};
    const std::string t = trim(text);
    for (const auto& r : res) if (std::regex_search(t, r)) return true;
    return false;
}
}  // namespace

std::string apply_custom_words(const std::string& text, const std::vector<std::string>& custom_words, double threshold) {
    if (custom_words.empty()) return text;
    std::vector<std::string> nospace;
    for (const auto& w : custom_words) { std::string l = lower(w); l.erase(std::remove(l.begin(), l.end(), ' '), l.end()); nospace.push_back(l); }
    const auto words = split_ws(text);
    std::vector<std::string> out;
    for (size_t i = 0; i < words.size();) {
        bool matched = false;
        for (size_t n = 3; n >= 1 && !matched; --n) {
            if (i + n > words.size()) continue;
            std::string gram;
            for (size_t k = 0; k < n; ++k) gram += lower(strip_non_alnum(words[i + k]));
            if (gram.empty() || gram.size() > 50) continue;
            int best = -1; double best_score = std::numeric_limits<double>::max();
            for (size_t c = 0; c < nospace.size(); ++c) {
                const double cl = (double)gram.size(), wl = (double)nospace[c].size();
                const double max_len = std::max(cl, wl);
                if (std::fabs(cl - wl) > std::max(max_len * 0.25, 2.0)) continue;
                double score = max_len > 0 ? (double)levenshtein(gram, nospace[c]) / max_len : 1.0;
                if (soundex_code(gram) == soundex_code(nospace[c])) score *= 0.3;
                if (score < threshold && score < best_score) { best = (int)c; best_score = score; }
            }
            if (best >= 0) {
                const std::string& first = words[i];
                const std::string& last = words[i + n - 1];
                size_t pre = 0; while (pre < first.size() && !is_alnum((unsigned char)first[pre])) ++pre;
                size_t suf = 0; while (suf < last.size() && !is_alnum((unsigned char)last[last.size() - 1 - suf])) ++suf;
                out.push_back(first.substr(0, pre) + preserve_case(first, custom_words[best]) + last.substr(last.size() - suf));
                i += n;
                matched = true;
            }
        }
        if (!matched) { out.push_back(words[i]); ++i; }
    }
    return join(out);
}

std::string filter_transcription_output(const std::string& text) {
    std::string out = text;
    for (const char* w : kFillers) {
        const std::regex r(std::string("\\b") + w + "\\b[,.]?", std::regex::icase);
        out = std::regex_replace(out, r, "");
    }
    out = collapse_stutters(out);
    static const std::regex multi(R"(\s{2,})");
    out = trim(std::regex_replace(out, multi, " "));
    return is_hallucination(out) ? std::string() : out;
}

}  // namespace sb
