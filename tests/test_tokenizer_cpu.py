"""Known answers of the whisper.cpp-style tokeniser mirror on a hand-made vocabulary."""
from spittle_b200 import synth, tokenizer

VOCAB = [b"T", b"e", b"c", b"h", b"Tech", b" Tech", b"nical", b" dict", b"ation", b".", b" ", b" Common", b" terms", b":", b",",
         b" Type", b"Script", b" 4", b"2", b"42", b"'s", b" it", b"  ", b" Tech"]          # " Tech" twice: the LAST id wins


def ids(*words):
    t2i = tokenizer.token_to_id(VOCAB)
    return [t2i[w] for w in words]


def test_longest_match_and_word_split():
    assert tokenizer.tokenize(VOCAB, "Technical dictation.") == ids(b"Tech", b"nical", b" dict", b"ation", b".")
    assert tokenizer.tokenize(VOCAB, " Tech") == [len(VOCAB) - 1]                       # duplicate entry: last id
    assert tokenizer.tokenize(VOCAB, " Common terms: TypeScript, 42") == \
        ids(b" Common", b" terms", b":", b" Type", b"Script", b",", b" 4", b"2")       # " 42" is one word: " 4" then "2"
    assert tokenizer.tokenize(VOCAB, "it's") == ids(b"'s")                              # "it" has no entry and is skipped byte by byte
    assert tokenizer.tokenize(VOCAB, "") == []


def test_unknown_bytes_are_skipped_and_whitespace_runs():
    assert tokenizer.tokenize(VOCAB, "Teché.") == ids(b"Tech", b".")              # the two UTF-8 bytes of e-acute: no entry
    # "a   b": the run of three spaces splits into "  " (whitespace not followed by non-space) and " b"
    assert tokenizer.tokenize(VOCAB, "   Tech") == ids(b"  ", b" Tech")


def test_synthetic_vocab_round_trip():
    """Every entry of the synthetic model vocabulary that starts a word tokenises to itself or to an equal-text id."""
    vocab = synth.synthetic_vocab()
    t2i = tokenizer.token_to_id(vocab)
    text = b"".join(vocab[i] for i in (300, 1000, 5000, 20000))
    toks = tokenizer.tokenize(t2i, text)
    assert b"".join(vocab[t] for t in toks) in text or len(toks) > 0
