"""Text post-filters that run inside the reference ``transcribe()`` after the engine call
(SURVEY.md 8(f) N1).  Host-side string work, microseconds per call: stays on the CPU.

Mirrors, by behaviour (re-implemented, not transliterated):
  apply_custom_words            src-tauri/src/audio_toolkit/text.rs:102-156   (call site transcription.rs:538-546)
  filter_transcription_output   src-tauri/src/audio_toolkit/text.rs:373-396   (call site transcription.rs:549)
  clean_segment_boundaries      src-tauri/src/audio_toolkit/text.rs:207-240
Pinned by the reference's own unit tests (text.rs:398-673): tests/golden/text_filters.json holds every
input / expectation of those tests, tests/test_text_filters_cpu.py replays them.

Third-party pieces the reference calls, restated from their published algorithms:
  strsim 0.11 ``levenshtein``   classic edit distance over Unicode scalar values
  natural 0.5 ``phonetics::soundex(a, b)``   true when both words have the same 4-character Soundex code
     (first letter kept, h/w dropped, adjacent equal digits merged, vowels dropped, zero padded) [MEM]
"""
from __future__ import annotations

import re
from typing import List, Optional, Sequence, Tuple

# ---- filler words / hallucinations (text.rs:244-247, 304-335) -------------------------------------
FILLER_WORDS = ("uh", "um", "uhm", "umm", "uhh", "uhhh", "ah", "eh", "hmm", "hm", "mmm", "mm", "mh", "ha", "ehh")
_FILLER_RES = [re.compile(r"\b" + re.escape(w) + r"\b[,.]?", re.IGNORECASE) for w in FILLER_WORDS]
_MULTI_SPACE = re.compile(r"\s{2,}")

HALLUCINATION_PHRASES = frozenset((
    "thank you for watching", "thanks for watching", "thank you for listening", "thanks for listening",
    "please subscribe", "like and subscribe", "see you next time", "see you in the next video",
    "bye bye", "bye", "thank you", "thanks", "subtitles by", "you",
))
_HALLUCINATION_RES = [
    re.compile(r"^(for more information[,.]?\s*)?(visit|go to)\s+\S+(\s+(or\s+)?(visit|go to)\s+\S+)*(\s+for more information)?[.,]?\s*$", re.IGNORECASE),
    re.compile(r"^for more information[,.]?\s*(visit|go to)\s+\S+[.,]?\s*$", re.IGNORECASE),
    re.compile(r"^subtitles\s+(by|provided by|created by)\s+.*$", re.IGNORECASE),
]


def _blen(s: str) -> int:
    """Rust ``str::len`` is the UTF-8 byte length; the reference's length heuristics use it."""
    return len(s.encode("utf-8"))


def levenshtein(a: str, b: str) -> int:
    if a == b:
        return 0
    if not a:
        return len(b)
    if not b:
        return len(a)
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


_SOUNDEX_DIGIT = {}
for _chars, _d in (("bfpv", "1"), ("cgjkqsxz", "2"), ("dt", "3"), ("l", "4"), ("mn", "5"), ("r", "6"), ("hw", "9")):
    for _c in _chars:
        _SOUNDEX_DIGIT[_c] = _d


def soundex_code(word: str) -> str:
    if not word:
        return "0000"
    enc = [word[0]] + [_SOUNDEX_DIGIT.get(c, "0") for c in word[1:]]
    enc = [c for c in enc if c != "9"]                      # h / w do not separate equal codes
    dedup: List[str] = []
    for c in enc:
        if not dedup or dedup[-1] != c:
            dedup.append(c)
    code = [c for c in dedup if c != "0"]                   # vowels (and everything unmapped) drop out
    return ("".join(code) + "0000")[:4]


def soundex_match(a: str, b: str) -> bool:
    return soundex_code(a) == soundex_code(b)


def _strip_non_alnum(w: str) -> str:
    i, j = 0, len(w)
    while i < j and not w[i].isalnum():
        i += 1
    while j > i and not w[j - 1].isalnum():
        j -= 1
    return w[i:j]


def build_ngram(words: Sequence[str]) -> str:
    return "".join(_strip_non_alnum(w).lower() for w in words)


def find_best_match(candidate: str, custom_words: Sequence[str], custom_nospace: Sequence[str],
                    threshold: float) -> Optional[Tuple[str, float]]:
    if not candidate or _blen(candidate) > 50:
        return None
    best, best_score = None, float("inf")
    cl = _blen(candidate)
    for original, cw in zip(custom_words, custom_nospace):
        wl = _blen(cw)
        max_len = float(max(cl, wl))
        if abs(cl - wl) > max(max_len * 0.25, 2.0):
            continue
        score = levenshtein(candidate, cw) / max_len if max_len > 0 else 1.0
        if soundex_match(candidate, cw):
            score *= 0.3
        if score < threshold and score < best_score:
            best, best_score = original, score
    return None if best is None else (best, best_score)


def extract_punctuation(word: str) -> Tuple[str, str]:
    n_pre = 0
    for c in word:
        if c.isalnum():
            break
        n_pre += 1
    n_suf = 0
    for c in reversed(word):
        if c.isalnum():
            break
        n_suf += 1
    # the reference slices by these counts independently, so an all-punctuation word yields (word, word)
    return word[:n_pre], (word[len(word) - n_suf:] if n_suf else "")


def preserve_case_pattern(original: str, replacement: str) -> str:
    if all(c.isupper() for c in original):                 # vacuously true for "" like Iterator::all
        return replacement.upper()
    if original[:1].isupper():
        return replacement[:1].upper()[:1] + replacement[1:] if replacement else replacement
    return replacement


def apply_custom_words(text: str, custom_words: Sequence[str], threshold: float) -> str:
    if not custom_words:
        return text
    nospace = [w.lower().replace(" ", "") for w in custom_words]
    words = text.split()
    out: List[str] = []
    i = 0
    while i < len(words):
        for n in (3, 2, 1):                                 # greedy: longest n-gram first
            if i + n > len(words):
                continue
            gram = words[i:i + n]
            m = find_best_match(build_ngram(gram), custom_words, nospace, threshold)
            if m is not None:
                prefix, _ = extract_punctuation(gram[0])
                _, suffix = extract_punctuation(gram[-1])
                out.append(prefix + preserve_case_pattern(gram[0], m[0]) + suffix)
                i += n
                break
        else:
            out.append(words[i])
            i += 1
    return " ".join(out)


def collapse_stutters(text: str) -> str:
    words = text.split()
    if not words:
        return text
    out: List[str] = []
    i = 0
    while i < len(words):
        w = words[i]
        wl = w.lower()
        if _blen(wl) <= 2 and all(c.isalpha() for c in wl):
            n = 1
            while i + n < len(words) and words[i + n].lower() == wl:
                n += 1
            out.append(w)
            i += n if n >= 3 else 1
        else:
            out.append(w)
            i += 1
    return " ".join(out)


def is_hallucination(text: str) -> bool:
    stripped = "".join(c for c in text.strip() if c.isalnum() or c.isspace())
    normalized = stripped.strip().lower()
    if not normalized:
        return False
    if normalized in HALLUCINATION_PHRASES:
        return True
    trimmed = text.strip()
    return any(r.search(trimmed) for r in _HALLUCINATION_RES)


def filter_transcription_output(text: str) -> str:
    out = text
    for r in _FILLER_RES:
        out = r.sub("", out)
    out = collapse_stutters(out)
    out = _MULTI_SPACE.sub(" ", out).strip()
    return "" if is_hallucination(out) else out


def _trim_segment(s: str) -> str:
    s = s.strip().rstrip(".")
    while s.endswith("..."):
        s = s[:-3]
    return s.rstrip("!").rstrip("?").rstrip(",").strip()


def clean_segment_boundaries(segments: Sequence[str], remaining: str) -> str:
    parts = [t.lower() for t in (_trim_segment(s) for s in segments) if t]
    r = _trim_segment(remaining)
    if r:
        parts.append(r.lower())
    return " ".join(parts)
