"""GPU parity AT THE BASELINE SIZES (BASELINE.json configs[1] "token-exact vs CPU reference" on Whisper Small,
north_star on Large-v3 Turbo): encoder output, teacher-forced logits and free-running greedy tokens of the engine
against the oracle on the full-size architectures (K = 768 / 1280 / 3072 / 5120 GEMM shapes, 12 / 20 heads)."""
import time

import numpy as np
import pytest

from oracle import logmel, whisper_ref
from spittle_b200 import capi, ggml_format, synth

pytestmark = pytest.mark.gpu

# Stated tolerances, f16 engine vs the ggml-faithful oracle (act_f16): 1.5 x the error measured on B200 in round 2
# (profiles/r2_parity_sizes.md: encoder rel-RMS 4.70e-4 / 4.76e-4, teacher-forced logits 1.84e-2 / 2.00e-2 over 16 steps).
# For scale: two CPU restatements of the same rounding points (numpy oracle vs oracle/cpu_ref) differ by 3.7e-4 / 1.5e-2.
ENC_REL_RMS = {"small": 7.1e-4, "large-v3-turbo": 7.2e-4}
LOGIT_TOL = {"small": 2.8e-2, "large-v3-turbo": 3.0e-2}   # raw logits have std ~ 5 (sharp recipe)
MARGIN_TOL = 0.06                                         # a token may only differ where the oracle margin is below 2 x LOGIT_TOL


@pytest.fixture(scope="module", params=["small", "large-v3-turbo"])
def sized(request, cuda_dev, model_dir):
    arch = request.param
    path = synth.ensure_model_file(arch, model_dir)
    model = ggml_format.read_ggml(path)
    oracle = whisper_ref.WhisperOracle(model, act_f16=True)
    eng = capi.Engine(path, dtype=capi.SB_DTYPE_F16, max_batch=4)
    x = synth.make_clip(0, 30.0)
    mel, n_len_org = logmel.logmel_f64(x, model.mel_filters)
    win = logmel.mel_window(mel, 0)
    t0 = time.perf_counter()
    enc = oracle.encode(win)
    print(f"{arch}: oracle encoder {time.perf_counter() - t0:.1f} s")
    yield arch, model, oracle, eng, x, win, n_len_org, enc
    eng.close()


def test_encoder_output_at_full_size(sized):
    arch, model, oracle, eng, x, win, n_len_org, enc = sized
    got = eng.encode(win[None])[0]
    rel = float(np.sqrt(((got - enc) ** 2).mean()) / np.sqrt((enc ** 2).mean()))
    mx = float(np.abs(got - enc).max())
    print(f"{arch}: encoder rel-RMS {rel:.3e}, max-abs {mx:.3e} (output RMS {float(np.sqrt((enc ** 2).mean())):.3f})")
    assert rel <= ENC_REL_RMS[arch], rel


def test_teacher_forced_logits_at_full_size(sized):
    arch, model, oracle, eng, x, win, n_len_org, enc = sized
    n_steps = 16
    tr = oracle.decode_window(enc, 0, n_len_org, whisper_ref.DecodeConfig(n_max_override=n_steps), trace=True)
    forced = np.full((1, n_steps), -1, np.int32)
    forced[0, :len(tr.tokens)] = tr.tokens
    logits, toks, marg = eng.decode_trace(win[None], [n_len_org], n_steps, forced=forced)
    errs = [float(np.abs(logits[0, s] - tr.logits_trace[s]).max()) for s in range(len(tr.tokens))]
    print(f"{arch}: teacher-forced logits over {len(errs)} steps: max err {max(errs):.3e}, oracle margins min "
          f"{min(tr.margins):.3f} median {float(np.median(tr.margins)):.3f}")
    assert max(errs) <= LOGIT_TOL[arch], errs
    for s in range(len(tr.tokens)):
        if tr.margins[s] > 2 * LOGIT_TOL[arch]:
            assert toks[0, s] == tr.tokens[s], (s, tr.margins[s])


def test_greedy_tokens_at_full_size(sized):
    """free-running greedy tokens of clip 0 through sb_transcribe (log-mel -> encoder -> decode loop): identical to
    the oracle, or first divergence at an oracle margin below MARGIN_TOL (reported)."""
    arch, model, oracle, eng, x, win, n_len_org, enc = sized
    n_tok = 64
    r = eng.transcribe(x, capi.default_params(n_max_tokens=n_tok, max_windows=1))
    w = oracle.decode_window(enc, 0, n_len_org, whisper_ref.DecodeConfig(n_max_override=n_tok))
    got = r.sampled[: r.windows[0]["n_tokens"]]
    if got == w.tokens:
        print(f"{arch}: {len(got)} greedy tokens identical to the oracle")
        assert r.windows[0]["result_len"] == w.result_len and r.windows[0]["seek_delta"] == w.seek_delta
    else:
        first = next(k for k in range(min(len(got), len(w.tokens))) if got[k] != w.tokens[k])
        print(f"{arch}: first divergence at step {first}, oracle margin {w.margins[first]:.4f}")
        assert w.margins[first] < MARGIN_TOL, (first, w.margins[first])
        assert first >= 8, "divergence this early means a real difference, not an indecisive margin"
