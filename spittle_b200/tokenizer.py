"""whisper.cpp's text tokeniser (``tokenize`` / ``whisper_tokenize``) restated for tests [MEM]: the text is split
into words by the GPT-2 style regex over its BYTES ('s|'t|'re|'ve|'m|'ll|'d| ?letters| ?digits| ?other|whitespace),
every word is then cut greedily into the longest vocabulary entries; bytes that no entry covers are skipped.  A
vocabulary entry that occurs twice keeps its LAST id.  The engine's own implementation is sb_tokenize (csrc/engine.cu);
this mirror only checks it (initial_prompt: managers/transcription.rs:461-499, jargon.rs:594-627)."""
from __future__ import annotations

import re
from typing import Dict, List, Sequence

_WORD = re.compile(rb"'s|'t|'re|'ve|'m|'ll|'d| ?[A-Za-z]+| ?[0-9]+| ?[^\sA-Za-z0-9]+|\s+(?!\S)|\s+")


def token_to_id(vocab: Sequence[bytes]) -> Dict[bytes, int]:
    return {w: i for i, w in enumerate(vocab)}          # a later duplicate overwrites an earlier one


def tokenize(vocab: Sequence[bytes], text) -> List[int]:
    t2i = vocab if isinstance(vocab, dict) else token_to_id(vocab)
    data = text if isinstance(text, bytes) else text.encode("utf-8")
    out: List[int] = []
    for m in _WORD.finditer(data):
        word = m.group(0)
        i, n = 0, len(word)
        while i < n:
            j = n
            while j > i:
                tid = t2i.get(word[i:j])
                if tid is not None:
                    out.append(tid)
                    i = j
                    break
                j -= 1
            else:
                i += 1
    return out
