"""Independent cross-check of the oracle's encoder / decoder wiring.

The reference's arithmetic (whisper.cpp) is not available here, so the oracle is "parity unpinned".  What CAN be
pinned is that the oracle computes the Whisper architecture: the same tensors are loaded into
``transformers.WhisperForConditionalGeneration`` (an implementation written by other people, f32 on the CPU) and
into the oracle's plain-f32 mode, and encoder output, teacher-forced logits and KV-cached stepwise logits have to agree.
Known delta: OpenAI/HF Whisper use the exact (erf) GELU, ggml the tanh form -- the oracle takes ``gelu="erf"``
for this test only, and a second assertion bounds what the tanh form changes.
LayerNorm gains / biases are randomised here (the synthetic recipe leaves them at 1 / 0) so their wiring is exercised.
"""
import numpy as np
import pytest

from oracle.whisper_ref import WhisperOracle
from oracle import logmel as olm
from spittle_b200 import synth

torch = pytest.importorskip("torch")
transformers = pytest.importorskip("transformers")


def _hf_from_ggml(model):
    hp = model.hparams
    cfg = transformers.WhisperConfig(
        vocab_size=hp.n_vocab, num_mel_bins=hp.n_mels, d_model=hp.n_audio_state,
        encoder_layers=hp.n_audio_layer, encoder_attention_heads=hp.n_audio_head,
        decoder_layers=hp.n_text_layer, decoder_attention_heads=hp.n_text_head,
        encoder_ffn_dim=4 * hp.n_audio_state, decoder_ffn_dim=4 * hp.n_text_state,
        max_source_positions=hp.n_audio_ctx, max_target_positions=hp.n_text_ctx,
        activation_function="gelu", dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
        scale_embedding=False, pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=3,
        suppress_tokens=None, begin_suppress_tokens=None)
    hf = transformers.WhisperForConditionalGeneration(cfg).eval().float()
    T = {k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in model.tensors.items()}
    sd = {}

    def attn(src, dst):
        sd[dst + ".q_proj.weight"] = T[src + ".query.weight"]; sd[dst + ".q_proj.bias"] = T[src + ".query.bias"]
        sd[dst + ".k_proj.weight"] = T[src + ".key.weight"]
        sd[dst + ".v_proj.weight"] = T[src + ".value.weight"]; sd[dst + ".v_proj.bias"] = T[src + ".value.bias"]
        sd[dst + ".out_proj.weight"] = T[src + ".out.weight"]; sd[dst + ".out_proj.bias"] = T[src + ".out.bias"]

    def ln(src, dst):
        sd[dst + ".weight"] = T[src + ".weight"]; sd[dst + ".bias"] = T[src + ".bias"]

    def mlp(src, dst):
        sd[dst + ".fc1.weight"] = T[src + ".mlp.0.weight"]; sd[dst + ".fc1.bias"] = T[src + ".mlp.0.bias"]
        sd[dst + ".fc2.weight"] = T[src + ".mlp.2.weight"]; sd[dst + ".fc2.bias"] = T[src + ".mlp.2.bias"]

    e = "model.encoder"
    sd[e + ".conv1.weight"] = T["encoder.conv1.weight"]; sd[e + ".conv1.bias"] = T["encoder.conv1.bias"].reshape(-1)
    sd[e + ".conv2.weight"] = T["encoder.conv2.weight"]; sd[e + ".conv2.bias"] = T["encoder.conv2.bias"].reshape(-1)
    sd[e + ".embed_positions.weight"] = T["encoder.positional_embedding"]
    for i in range(hp.n_audio_layer):
        s, d = f"encoder.blocks.{i}", f"{e}.layers.{i}"
        ln(s + ".attn_ln", d + ".self_attn_layer_norm"); attn(s + ".attn", d + ".self_attn")
        ln(s + ".mlp_ln", d + ".final_layer_norm"); mlp(s, d)
    ln("encoder.ln_post", e + ".layer_norm")
    dd = "model.decoder"
    sd[dd + ".embed_tokens.weight"] = T["decoder.token_embedding.weight"]
    sd[dd + ".embed_positions.weight"] = T["decoder.positional_embedding"]
    for i in range(hp.n_text_layer):
        s, d = f"decoder.blocks.{i}", f"{dd}.layers.{i}"
        ln(s + ".attn_ln", d + ".self_attn_layer_norm"); attn(s + ".attn", d + ".self_attn")
        ln(s + ".cross_attn_ln", d + ".encoder_attn_layer_norm"); attn(s + ".cross_attn", d + ".encoder_attn")
        ln(s + ".mlp_ln", d + ".final_layer_norm"); mlp(s, d)
    ln("decoder.ln", dd + ".layer_norm")
    sd["proj_out.weight"] = T["decoder.token_embedding.weight"]
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    # HF creates a zero k_proj bias? (it does not: bias=False) -- everything else must have been supplied
    assert not unexpected, unexpected
    assert all(".k_proj.bias" in m for m in missing), missing
    return hf


@pytest.fixture(scope="module")
def pair():
    model = synth.make_synthetic_model("nano", seed=7)
    rng = np.random.default_rng(11)
    for k in list(model.tensors):
        if "_ln." in k or k.endswith("ln_post.weight") or k.endswith("ln_post.bias") or k.startswith("decoder.ln."):
            n = model.tensors[k].shape[0]
            model.tensors[k] = (rng.standard_normal(n) * 0.3 + (1.0 if k.endswith("weight") else 0.0)).astype(np.float32)
    hf = _hf_from_ggml(model)
    pcm = synth.make_clip(3, seconds=30.0)
    mel, _ = olm.logmel_f32_faithful(pcm, model.mel_filters)
    win = olm.mel_window(mel, 0, model.hparams.n_audio_ctx).astype(np.float32)      # [n_mel, 3000]
    return model, hf, win


def test_encoder_matches_hf(pair):
    model, hf, win = pair
    orc = WhisperOracle(model, act_f16=False, gelu="erf")
    enc = orc.encode(win)
    with torch.no_grad():
        ref = hf.model.encoder(torch.from_numpy(win)[None]).last_hidden_state[0].numpy()
    assert enc.shape == ref.shape == (model.hparams.n_audio_ctx, model.hparams.n_audio_state)
    err = np.abs(enc - ref).max()
    assert err <= 2e-4 * max(1.0, np.abs(ref).max()), err
    # what ggml's tanh GELU changes (documented delta, not a wiring difference)
    enc_t = WhisperOracle(model, act_f16=False, gelu="tanh").encode(win)
    rel = np.sqrt(((enc_t - ref) ** 2).mean() / (ref ** 2).mean())
    assert rel < 5e-3, rel


def test_decoder_logits_match_hf_teacher_forced_and_cached(pair):
    model, hf, win = pair
    sp = model.special
    orc = WhisperOracle(model, act_f16=False, gelu="erf")
    enc = orc.encode(win)
    rng = np.random.default_rng(5)
    toks = [sp.sot, sp.lang_first, sp.transcribe] + [int(t) for t in rng.integers(0, sp.eot, size=9)] + [sp.beg + 40]
    kv_self, kv_cross = orc.new_kv(), orc.cross_kv(enc)
    mine = np.stack([orc.decode_step(t, i, kv_self, kv_cross) for i, t in enumerate(toks)])       # KV-cached, one token at a time
    with torch.no_grad():
        out = hf(encoder_outputs=(torch.from_numpy(enc)[None],), decoder_input_ids=torch.tensor([toks]))
    ref = out.logits[0].numpy()                                                               # causal mask, all at once
    assert mine.shape == ref.shape
    scale = np.abs(ref).max()
    assert np.abs(mine - ref).max() <= 3e-4 * max(1.0, scale), np.abs(mine - ref).max()
    assert (mine.argmax(-1) == ref.argmax(-1)).all()


def test_prompt_prefix_conditioning_matches_hf_and_changes_the_decode(pair):
    """[prev] + context tokens in front of [sot, lang, task] (whisper.cpp's prompt_past / initial_prompt): the oracle's
    KV-cached steps over the longer prompt equal HF's one-shot decoder, and the prefix changes the greedy tokens."""
    from oracle.whisper_ref import DecodeConfig
    model, hf, win = pair
    sp = model.special
    orc = WhisperOracle(model, act_f16=False, gelu="erf")
    enc = orc.encode(win)
    past = [int(t) for t in np.random.default_rng(9).integers(0, sp.eot, size=20)]
    toks = [sp.prev] + past + [sp.sot, sp.lang_first, sp.transcribe]
    kv_self, kv_cross = orc.new_kv(), orc.cross_kv(enc)
    mine = np.stack([orc.decode_step(t, i, kv_self, kv_cross) for i, t in enumerate(toks)])
    with torch.no_grad():
        ref = hf(encoder_outputs=(torch.from_numpy(enc)[None],), decoder_input_ids=torch.tensor([toks])).logits[0].numpy()
    assert np.abs(mine - ref).max() <= 3e-4 * max(1.0, np.abs(ref).max())
    cfg = DecodeConfig(n_max_override=10)
    plain = orc.decode_window(enc, 0, 3000, cfg)
    cond = orc.decode_window(enc, 0, 3000, cfg, prompt_past=past)
    assert len(cond.tokens) > 0 and plain.tokens != cond.tokens
    # n_take: at most n_text_ctx / 2 context tokens are used
    long_past = list(range(300))
    w = orc.decode_window(enc, 0, 3000, DecodeConfig(n_max_override=1), prompt_past=long_past)
    assert len(w.tokens) == 1
