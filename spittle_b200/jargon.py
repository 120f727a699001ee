"""Jargon dictionary merge, initial-prompt text and correction pass of ``transcribe()`` (host string work).

Mirror of the reference's src-tauri/src/jargon.rs:
  compute_active_dictionary  :506-592   custom terms first (their casing wins), then the enabled profiles in
                                        alphabetical order of their ids; custom corrections override profile ones;
                                        corrections sorted longest phrase first, then alphabetically
  build_initial_prompt       :594-627   "Technical dictation. Common terms: a, b, c." capped at 1000 bytes
  apply_corrections          :682-716   case-insensitive whole-phrase replacement outside protected spans
                                        (@refs, `code`, URLs, paths, CLI flags -- :637-664), which are masked with
                                        U+27E6 S<idx> U+27E7 placeholders and restored afterwards
It runs after filter_transcription_output in TranscriptionManager::transcribe (managers/transcription.rs:551-580).
The built-in profile TABLE (jargon.rs:39-505) is settings content, not algorithm: profiles are passed in.
Lengths are byte lengths (Rust ``str::len``).  The reference's unit tests (jargon.rs:741-961) are replayed in
tests/test_jargon_cpu.py.  Difference: a ``$`` in a replacement is literal here (the regex crate would expand ``$name``).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Tuple


@dataclass
class JargonCorrection:
    from_: str
    to: str


@dataclass
class JargonProfile:
    label: str = ""
    terms: List[str] = field(default_factory=list)
    corrections: List[JargonCorrection] = field(default_factory=list)


@dataclass
class JargonSettings:
    enabled_profiles: List[str] = field(default_factory=list)
    custom_terms: List[str] = field(default_factory=list)
    custom_corrections: List[JargonCorrection] = field(default_factory=list)


@dataclass
class ActiveDictionary:
    terms: List[str] = field(default_factory=list)
    corrections: List[JargonCorrection] = field(default_factory=list)


def compute_active_dictionary(settings: JargonSettings, profiles: Dict[str, JargonProfile]) -> ActiveDictionary:
    terms_map: Dict[str, str] = {}
    for term in settings.custom_terms:
        terms_map[term.lower()] = term                      # a later duplicate replaces an earlier one (HashMap::insert)
    profile_ids = sorted(pid for pid in settings.enabled_profiles if pid in profiles)
    for pid in profile_ids:
        for term in profiles[pid].terms:
            terms_map.setdefault(term.lower(), term)
    terms: List[str] = []
    seen = set()
    for term in settings.custom_terms:
        key = term.lower()
        if key not in seen:
            seen.add(key)
            terms.append(terms_map[key])
    for pid in profile_ids:
        for term in profiles[pid].terms:
            key = term.lower()
            if key not in seen:
                seen.add(key)
                terms.append(terms_map[key])
    cmap: Dict[str, JargonCorrection] = {}
    for pid in profile_ids:
        for c in profiles[pid].corrections:
            cmap[c.from_.lower()] = c
    for c in settings.custom_corrections:
        cmap[c.from_.lower()] = c
    corrections = sorted(cmap.values(), key=lambda c: (-len(c.from_.encode()), c.from_.encode()))
    return ActiveDictionary(terms, corrections)


def build_initial_prompt(dictionary: ActiveDictionary) -> str:
    if not dictionary.terms:
        return ""
    prefix, suffix, max_len = "Technical dictation. Common terms: ", ".", 1000
    available = max_len - len(prefix) - len(suffix)
    parts: List[str] = []
    current = 0
    for term in dictionary.terms:
        n = len(term.encode())
        addition = n if not parts else n + 2
        if current + addition > available:
            break
        parts.append(term)
        current += addition
    if not parts:
        return ""
    return prefix + ", ".join(parts) + suffix


_PROTECTED = re.compile("|".join([
    r"@[\w\-./]+",                         # @tokens like @file.rs
    r"`[^`]+`",                            # backtick code
    r"https?://[^\s]+",                    # URLs
    r"(?:~/|/[\w\-]+(?:/[\w\-.*]+)+)",     # file paths
    r"(?:^|\s)--?[\w\-]+=?(?:[\w\-./]+)?",  # CLI flags
]))


def mask_protected_spans(text: str) -> Tuple[str, List[Tuple[str, str]]]:
    """-> (masked text, [(placeholder, original)] in forward order)."""
    matches = list(_PROTECTED.finditer(text))
    masked = text
    spans: List[Tuple[str, str]] = []
    for idx in range(len(matches) - 1, -1, -1):             # back to front so the offsets stay valid
        m = matches[idx]
        ph = "⟦S%d⟧" % idx
        spans.append((ph, m.group(0)))
        masked = masked[: m.start()] + ph + masked[m.end():]
    spans.reverse()
    return masked, spans


def restore_protected_spans(text: str, spans: List[Tuple[str, str]]) -> str:
    for ph, original in spans:
        text = text.replace(ph, original)
    return text


def expand_replacement(to: str, matched: str) -> str:
    """Rust regex `Replacer for &str` (what `re.replace_all(&masked, correction.to.as_str())` does, jargon.rs:697): `$$` is a
    literal `$`; `$name` / `${name}` / `$N` expand capture groups -- the correction pattern has none, so `$0` is the whole
    match and every other reference is the empty string; a `$` followed by anything else stays."""
    out, i, n = [], 0, len(to)
    while i < n:
        ch = to[i]
        if ch != "$":
            out.append(ch); i += 1
            continue
        if i + 1 < n and to[i + 1] == "$":
            out.append("$"); i += 2
            continue
        j = i + 1
        if j < n and to[j] == "{":
            k = to.find("}", j)
            if k < 0:
                out.append("$"); i += 1
                continue
            name, j = to[j + 1:k], k + 1
        else:
            k = j
            while k < n and (to[k].isascii() and (to[k].isalnum() or to[k] == "_")):
                k += 1
            if k == j:
                out.append("$"); i += 1
                continue
            name, j = to[j:k], k
        if name.isdigit() and int(name) == 0:
            out.append(matched)
        i = j
    return "".join(out)


def apply_corrections(text: str, corrections: List[JargonCorrection]) -> str:
    if not corrections or not text:
        return text
    masked, spans = mask_protected_spans(text)
    for c in corrections:
        try:
            pat = re.compile(r"(?i)\b" + re.escape(c.from_) + r"\b")
        except re.error:
            continue
        masked = pat.sub(lambda m, to=c.to: expand_replacement(to, m.group(0)), masked)
    restored = restore_protected_spans(masked, spans)
    if any(ph in restored for ph, _ in spans):
        return text
    return restored
