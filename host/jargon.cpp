#include "jargon.hpp"

#include <algorithm>
#include <cctype>
#include <regex>
#include <set>
#include <unordered_map>

namespace sb {
namespace {
std::string lower(const std::string& s) {
    std::string r = s;
    for (char& c : r) c = (char)std::tolower((unsigned char)c);
    return r;
}
std::string regex_escape(const std::string& s) {
    static const std::string meta = R"(\.^$|()[]{}*+?-)";
    std::string r;
    for (char c : s) {
        if (meta.find(c) != std::string::npos) r += '\\';
        r += c;
    }
    return r;
}
}  // namespace

ActiveDictionary compute_active_dictionary(const JargonSettings& st, const std::map<std::string, JargonProfile>& profiles) {
    std::unordered_map<std::string, std::string> terms_map;
    for (const auto& t : st.custom_terms) terms_map[lower(t)] = t;                  // custom casing wins
    std::vector<std::string> ids;
    for (const auto& id : st.enabled_profiles)
        if (profiles.count(id)) ids.push_back(id);
    std::sort(ids.begin(), ids.end());
    for (const auto& id : ids)
        for (const auto& t : profiles.at(id).terms) terms_map.emplace(lower(t), t);  // only if absent
    ActiveDictionary d;
    std::set<std::string> seen;
    for (const auto& t : st.custom_terms)
        if (seen.insert(lower(t)).second) d.terms.push_back(terms_map[lower(t)]);
    for (const auto& id : ids)
        for (const auto& t : profiles.at(id).terms)
            if (seen.insert(lower(t)).second) d.terms.push_back(terms_map[lower(t)]);
    std::unordered_map<std::string, JargonCorrection> cmap;
    for (const auto& id : ids)
        for (const auto& c : profiles.at(id).corrections) cmap[lower(c.from)] = c;
    for (const auto& c : st.custom_corrections) cmap[lower(c.from)] = c;              // custom overrides profile
    for (auto& kv : cmap) d.corrections.push_back(kv.second);
    std::sort(d.corrections.begin(), d.corrections.end(), [](const JargonCorrection& a, const JargonCorrection& b) {
        if (a.from.size() != b.from.size()) return a.from.size() > b.from.size();     // longest phrase first
        return a.from < b.from;
    });
    return d;
}

std::string build_initial_prompt(const ActiveDictionary& d) {
    if (d.terms.empty()) return "";
    const std::string prefix = "Technical dictation. Common terms: ", suffix = ".";
    const size_t available = 1000 - prefix.size() - suffix.size();
    std::string body;
    size_t cur = 0, n_parts = 0;
    for (const auto& t : d.terms) {
        const size_t add = n_parts == 0 ? t.size() : t.size() + 2;
        if (cur + add > available) break;
        if (n_parts) body += ", ";
        body += t;
        cur += add;
        ++n_parts;
    }
    if (n_parts == 0) return "";
    return prefix + body + suffix;
}

// Rust regex `Replacer for &str` (re.replace_all(&masked, correction.to.as_str()), jargon.rs:697): `$$` is a literal `$`;
// `$name` / `${name}` / `$N` expand capture groups -- the correction pattern has none, so `$0` is the whole match and every
// other reference is empty; a `$` followed by anything else stays.
static std::string expand_replacement(const std::string& to, const std::string& matched) {
    std::string out;
    const size_t n = to.size();
    for (size_t i = 0; i < n;) {
        if (to[i] != '$') { out += to[i++]; continue; }
        if (i + 1 < n && to[i + 1] == '$') { out += '$'; i += 2; continue; }
        size_t j = i + 1;
        std::string name;
        if (j < n && to[j] == '{') {
            const size_t k = to.find('}', j);
            if (k == std::string::npos) { out += '$'; ++i; continue; }
            name = to.substr(j + 1, k - j - 1); j = k + 1;
        } else {
            size_t k = j;
            while (k < n && (std::isalnum((unsigned char)to[k]) || to[k] == '_')) ++k;
            if (k == j) { out += '$'; ++i; continue; }
            name = to.substr(j, k - j); j = k;
        }
        bool digits = !name.empty();
        for (char ch : name) digits = digits && std::isdigit((unsigned char)ch);
        if (digits && std::stol(name.substr(0, 9)) == 0 && name.find_first_not_of('0') == std::string::npos) out += matched;
        i = j;
    }
    return out;
}

std::string apply_corrections(const std::string& text, const std::vector<JargonCorrection>& corrections) {
    if (corrections.empty() || text.empty()) return text;
    // protected spans: @refs, `code`, URLs, file paths, CLI flags (jargon.rs:637-644)
    static const std::regex prot(R"(@[\w\-./]+|`[^`]+`|https?://[^\s]+|(?:~/|/[\w\-]+(?:/[\w\-.*]+)+)|(?:^|\s)--?[\w\-]+=?(?:[\w\-./]+)?)");
    struct Span { size_t pos, len; };
    std::vector<Span> ms;
    for (std::sregex_iterator it(text.begin(), text.end(), prot), end; it != end; ++it) ms.push_back({(size_t)it->position(), (size_t)it->length()});
    std::string masked = text;
    std::vector<std::pair<std::string, std::string>> spans(ms.size());
    for (size_t k = ms.size(); k-- > 0;) {                      // back to front so the offsets stay valid
        const std::string ph = "\xE2\x9F\xA6S" + std::to_string(k) + "\xE2\x9F\xA7";      // U+27E6 S<k> U+27E7
        spans[k] = {ph, text.substr(ms[k].pos, ms[k].len)};
        masked.replace(ms[k].pos, ms[k].len, ph);
    }
    for (const auto& c : corrections) {
        try {
            const std::regex re("\\b" + regex_escape(c.from) + "\\b", std::regex::ECMAScript | std::regex::icase);
            std::string out;
            auto begin = std::sregex_iterator(masked.begin(), masked.end(), re);
            size_t last = 0;
            for (auto it = begin; it != std::sregex_iterator(); ++it) {
                out.append(masked, last, (size_t)it->position() - last);
                out += expand_replacement(c.to, it->str());        // Rust's `$` expansion of the replacement
                last = (size_t)it->position() + (size_t)it->length();
            }
            out.append(masked, last, std::string::npos);
            masked.swap(out);
        } catch (const std::regex_error&) {
        }
    }
    std::string restored = masked;
    for (const auto& s : spans) {
        size_t p = 0;
        while ((p = restored.find(s.first, p)) != std::string::npos) { restored.replace(p, s.first.size(), s.second); p += s.second.size(); }
    }
    for (const auto& s : spans)
        if (restored.find(s.first) != std::string::npos) return text;   // a placeholder survived: keep the original
    return restored;
}
}  // namespace sb
