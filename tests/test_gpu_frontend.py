"""Log-mel frontend through the C ABI against the float64 oracle (north_star: 1e-4 relative)."""
import numpy as np
import pytest
import torch

from helpers import logmel_close, rel_l2
from oracle import frontend as ofe
from bsed_b200.utilities import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-4   # |a - b| <= 1e-4 * max(1, |b|) on the dB features


def test_melspec_and_logmel_full_clips():
    from bsed_b200 import engine
    clips = synth.make_clips(4, seed=2023)
    a = torch.from_numpy(clips).cuda()
    mel = engine.melspec(a)
    out = engine.amp_to_db(mel, 1255)
    torch.cuda.synchronize()
    assert mel.shape == (4, 1255, 128) and out.shape == (4, 1255, 128)
    for i in range(4):
        ref_mel = ofe.preprocess(clips[i])
        assert rel_l2(mel[i].cpu().numpy(), ref_mel) < 2e-6
        worst, frac = logmel_close(out[i].cpu().numpy(), ofe.transform(ref_mel)[0])
        assert worst <= RTOL, (i, worst, frac)


def test_edge_clips():
    """all-zero clip (every bin -100 dB, clamp inactive), single impulse, full-scale 1 kHz sine."""
    from bsed_b200 import engine
    clips = synth.make_clips(3, seed=1, edge_cases=True)
    out = engine.amp_to_db(engine.melspec(torch.from_numpy(clips).cuda()), 1255).cpu().numpy()
    assert np.allclose(out[0], -100.0, atol=2e-5)
    ref1 = ofe.logmel(clips[1])
    worst, frac = logmel_close(out[1], ref1)
    assert worst <= RTOL, ("impulse", worst, frac)
    # pure tone: the reference's float64 FFT resolves side-lobes 100+ dB below the peak that an fp32 FFT cannot;
    # everything above the 80 dB top_db clamp floor that matters is compared on the clamped output
    ref2 = ofe.logmel(clips[2])
    worst, frac = logmel_close(out[2], ref2)
    assert frac < 0.02 and worst < 5e-3, ("sine", worst, frac)
    peak = np.unravel_index(ref2.argmax(), ref2.shape)
    assert abs(out[2][peak] - ref2[peak]) < 1e-4


@pytest.mark.parametrize("n", [1025, 5000, 31999, 32000, 255 * 40, 255 * 40 + 254])
def test_ragged_lengths(n):
    from bsed_b200 import engine, _lib
    y = synth.make_clips(2, seed=n, n_samples=n)
    mel = engine.melspec(torch.from_numpy(y).cuda()).cpu().numpy()
    assert mel.shape[1] == 1 + n // 255 == _lib.load().bsed_frontend_n_frames(n)
    for i in range(2):
        assert rel_l2(mel[i], ofe.preprocess(y[i])) < 3e-6


def test_too_short_clip_is_rejected():
    from bsed_b200 import engine, _lib
    with pytest.raises(_lib.BsedError):
        engine.melspec(torch.zeros(1, 1000).cuda())


def test_transform_noise_pad_trunc_scaler():
    from bsed_b200 import engine
    rng = np.random.default_rng(5)
    clips = synth.make_clips(2, seed=9)
    mel = np.stack([ofe.preprocess(c) for c in clips])                       # (2, 1255, 128)
    noise = rng.standard_normal(mel.shape).astype(np.float32)
    mean = rng.standard_normal(128).astype(np.float32) * 5 - 20
    std = (rng.random(128).astype(np.float32) + 0.5) * 10
    md, nd = torch.from_numpy(mel).cuda(), torch.from_numpy(noise).cuda()
    # noisy branch (teacher input), float64 in the reference
    got = engine.amp_to_db(md, 1255, nd, 30.0).cpu().numpy()
    for i in range(2):
        _, ref_noisy = ofe.transform(mel[i], unit_noise=noise[i])
        worst, frac = logmel_close(got[i], ref_noisy[0])
        assert worst <= RTOL, (i, worst, frac)
    # padding: 1000 frames in, rows >= 1000 are literal 0; truncation: 1255 in, 900 out
    got = engine.amp_to_db(md[:, :1000].contiguous(), 1255).cpu().numpy()
    ref = ofe.transform(mel[0][:1000])[0]
    assert (got[0][1000:] == 0).all() and logmel_close(got[0], ref)[0] <= RTOL
    got = engine.amp_to_db(md, 900).cpu().numpy()
    ref = ofe.transform(mel[1], frames=900)[0]
    assert got.shape == (2, 900, 128) and logmel_close(got[1], ref)[0] <= RTOL
    # scaler
    got = engine.amp_to_db(md, 1255, None, 0.0, torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda()).cpu().numpy()
    ref = ofe.transform(mel[0], mean=mean.astype(np.float64), std=std.astype(np.float64))[0]
    assert np.abs(got[0] - ref).max() < 1e-4


def test_reference_signature_preprocess_and_transforms():
    """preprocess(audio) -> (T,128) float32 ndarray and get_transforms(...)((mel, label)) as in the reference."""
    from bsed_b200.data.preprocess import preprocess
    from bsed_b200.data import Transforms as T
    y = synth.make_clips(1, seed=4)[0]
    mel = preprocess(y)
    assert isinstance(mel, np.ndarray) and mel.dtype == np.float32 and mel.shape == (1255, 128)
    assert rel_l2(mel, ofe.preprocess(y)) < 2e-6
    db = preprocess(y, compute_log=True)
    assert logmel_close(db, ofe.amplitude_to_db(ofe.preprocess(y)))[0] <= RTOL
    tr = T.get_transforms(1255, noise_dict_params={"mean": 0., "snr": 30})
    label = np.zeros((313, 20))
    np.random.seed(2023)
    (clean, noisy), lab = tr((mel, label))
    assert clean.shape == noisy.shape == (1, 1255, 128) and lab.shape == (313, 20) and lab.dtype == torch.float32
    np.random.seed(2023)
    unit = np.random.standard_normal(mel.shape)
    ref_clean, ref_noisy = ofe.transform(mel, unit_noise=unit.astype(np.float32))
    assert logmel_close(clean.numpy(), ref_clean)[0] <= RTOL
    assert logmel_close(noisy.numpy(), ref_noisy)[0] <= RTOL


def test_fused_logmel_is_bit_identical_to_the_two_calls():
    """bsed_logmel (STFT + mel with the clip maximum, then one dB pass) == bsed_melspec -> bsed_amp_to_db, with and without
    the scaler, for full and ragged lengths."""
    from bsed_b200 import engine
    from bsed_b200.utilities import synth
    clips = torch.from_numpy(synth.make_clips(3, seed=5)).cuda()
    mean, std = torch.randn(128, device="cuda"), torch.rand(128, device="cuda") + 0.5
    for n in (320000, 200001):
        a = clips[:, :n].contiguous()
        mel = engine.melspec(a)
        for kw in ({}, dict(scaler_mean=mean, scaler_std=std)):
            ref = engine.amp_to_db(mel, 1255, **kw)
            got, mel2 = engine.logmel(a, 1255, return_mel=True, **kw)
            assert torch.equal(mel2, mel) and torch.equal(got, ref)
