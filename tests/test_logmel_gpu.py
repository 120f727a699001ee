"""GPU parity: k_logmel vs the oracle (SURVEY 8(d) tolerances), through the C ABI."""
import numpy as np
import pytest

from oracle import logmel
from spittle_b200 import capi, synth

pytestmark = pytest.mark.gpu

# tolerance vs the f64 oracle (SURVEY.md 8(d) "Parity tolerances"):
#   max|a-b| / max|b| <= 1e-4  and  |a-b| <= 1e-4*|b| + 1e-4 element-wise
RANGE_REL = 1e-4
ABS_REL = 1e-4


def _check(a, b):
    assert a.shape == b.shape
    err = np.abs(a.astype(np.float64) - b)
    assert err.max() / np.abs(b).max() <= RANGE_REL, f"range-relative error {err.max() / np.abs(b).max():.3e}"
    assert np.all(err <= ABS_REL * np.abs(b) + ABS_REL)
    return err.max()


@pytest.mark.parametrize("n_mel", [80, 128])
def test_logmel_host_api_matches_oracle(cuda_dev, n_mel):
    filt = synth.mel_filterbank(n_mel)
    plan = capi.MelPlan(filt)
    for i, secs in [(0, 30.0), (1, 30.0), (2, 30.0), (3, 30.0), (4, 11.3), (5, 1.25), (7, 0.05)]:
        x = synth.make_clip(i, seconds=secs)
        got, n_len_org = plan.logmel(x)
        ref, ref_org = logmel.logmel_f64(x, filt)
        assert n_len_org == ref_org
        _check(got, ref)


def test_logmel_silence_and_edge_lengths(cuda_dev):
    filt = synth.mel_filterbank(80)
    plan = capi.MelPlan(filt)
    for n in (201, 399, 400, 401, 20000, 479999):
        x = np.zeros(n, np.float32)
        got, _ = plan.logmel(x)
        assert np.all(got == np.float32(-1.5))           # (max(-10, -18) + 4) / 4
        rng = np.random.default_rng(n)
        x = rng.uniform(-1, 1, n).astype(np.float32)
        got, _ = plan.logmel(x)
        ref, _ = logmel.logmel_f64(x, filt)
        _check(got, ref)
    with pytest.raises(capi.SbError):
        plan.logmel(np.zeros(100, np.float32))


def test_logmel_batch_dev_matches_host_api(cuda_dev):
    import torch
    filt = synth.mel_filterbank(80)
    plan = capi.MelPlan(filt)
    clips = np.stack([synth.make_clip(i, seconds=30.0) for i in range(4)])
    n_len, n_len_org, n_calc = capi.logmel_geometry(clips.shape[1])
    stride = (n_calc + 31) // 32 * 32
    pcm = torch.from_numpy(clips).to(cuda_dev)
    mel = torch.zeros((4, 80, stride), dtype=torch.float32, device=cuda_dev)
    cmax = torch.zeros(4, dtype=torch.int32, device=cuda_dev)
    floor = torch.zeros(4, dtype=torch.float32, device=cuda_dev)
    capi.logmel_batch_dev(plan, pcm.data_ptr(), 4, clips.shape[1], mel.data_ptr(), stride, cmax.data_ptr(),
                          floor.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    mel = mel.cpu().numpy()
    for i in range(4):
        ref, _ = logmel.logmel_f64(clips[i], filt)
        _check(mel[i, :, :n_calc], ref[:, :n_calc])
        assert abs(float(floor[i]) - float(ref[0, -1])) < 1e-6


def test_logmel_linearity_property_full_size(cuda_dev):
    """Size-independent property at the full 30 s size: scaling the input by g shifts the raw
    log10 spectrum by 2*log10(g), which the max-8 clamp and (x+4)/4 turn into a constant
    offset of log10(g)/2 on every non-floor cell."""
    filt = synth.mel_filterbank(80)
    plan = capi.MelPlan(filt)
    x = synth.make_clip(3, seconds=30.0) * 0.5
    a, _ = plan.logmel(x)
    b, _ = plan.logmel(x * 0.5)
    d = a - b
    assert np.abs(d - np.log10(2.0) / 2).max() < 2e-4
