"""The frontend oracle is PARITY-UNPINNED by the reference (librosa absent, no reference tests), so it
is cross-checked against independent implementations present in the container: torch.stft and
torchaudio's Slaney filterbank, plus hand-derived known answers."""
import numpy as np
import pytest
import torch

from oracle import frontend as fe
from bsed_b200.utilities import synth


def test_frame_count_and_shapes():
    y = synth.make_clips(1, seed=1)[0]
    S = fe.stft(y)
    assert S.shape == (1025, 1255) and S.dtype == np.complex64
    m = fe.preprocess(y)
    assert m.shape == (1255, 128) and m.dtype == np.float32
    assert fe.n_frames_for(320000) == 1255 == fe.MAX_FRAMES


def test_stft_matches_torch_stft():
    y = synth.make_clips(1, seed=2, n_samples=32000)[0]
    S = fe.stft(y)
    win = torch.from_numpy(np.hamming(2048))
    T = torch.stft(torch.from_numpy(y).double(), n_fft=2048, hop_length=255, window=win, center=True,
                   pad_mode="reflect", return_complex=True).numpy()
    assert T.shape == S.shape
    assert np.abs(S - T).max() / np.abs(T).max() < 2e-7


def test_hamming_is_symmetric_numpy_window():
    assert np.array_equal(fe.hamming_window(2048), np.hamming(2048)) or \
        np.abs(fe.hamming_window(2048) - np.hamming(2048)).max() < 1e-15


def test_mel_filterbank_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    fb = fe.mel_filterbank()
    ref = ta.functional.melscale_fbanks(1025, 0.0, 16000.0, 128, 32000, norm=None, mel_scale="slaney").numpy().T
    assert fb.shape == (128, 1025) and fb.dtype == np.float32
    assert np.abs(fb - ref).max() < 2e-5
    nnz = int((fb != 0).sum())
    assert 1900 < nnz < 2100
    assert ((fb != 0).sum(axis=0) <= 2).all()       # at most two bands per FFT bin


def test_mel_scale_known_answers():
    assert fe.hz_to_mel(1000.0) == pytest.approx(15.0)
    assert fe.mel_to_hz(15.0) == pytest.approx(1000.0)
    assert fe.hz_to_mel(6400.0) == pytest.approx(15.0 + 27.0)
    assert fe.mel_to_hz(fe.hz_to_mel(12345.0)) == pytest.approx(12345.0)


def test_amplitude_to_db_rules():
    x = np.array([[1.0, 10.0], [1e-7, 0.0]], dtype=np.float32)
    d = fe.amplitude_to_db(x)
    assert d.dtype == np.float32
    assert d[0, 0] == pytest.approx(0.0, abs=1e-6) and d[0, 1] == pytest.approx(20.0, abs=1e-5)
    assert d[1, 0] == pytest.approx(20.0 - 80.0) and d[1, 1] == pytest.approx(-60.0)   # top_db clamp
    z = fe.amplitude_to_db(np.zeros((4, 4), dtype=np.float32))
    assert np.allclose(z, -100.0, atol=2e-5) and (z == z[0, 0]).all()                       # amin floor, clamp inactive


def test_pad_trunc_and_transform():
    m = np.abs(np.random.default_rng(0).standard_normal((1000, 128))).astype(np.float32)
    out = fe.transform(m)
    assert out.shape == (1, 1255, 128) and out.dtype == np.float32
    assert (out[0, 1000:] == 0).all()
    long = np.abs(np.random.default_rng(1).standard_normal((1300, 128))).astype(np.float32)
    assert fe.transform(long).shape == (1, 1255, 128)
    clean, noisy = fe.transform(m, unit_noise=np.random.default_rng(2).standard_normal(m.shape))
    assert clean.shape == noisy.shape == (1, 1255, 128)
    assert np.array_equal(clean, out)
    assert not np.array_equal(clean, noisy)


def test_noise_std_definition():
    m = np.abs(np.random.default_rng(3).standard_normal((50, 128))) + 0.1
    std = fe.noise_std(m, 30.0)
    assert std.shape == (128,)
    assert np.allclose(std, np.sqrt((m ** 2).mean(0) * 1e-3))


def test_sine_peak_bin():
    t = np.arange(32000) / 32000.0
    y = np.sin(2 * np.pi * 1000.0 * t).astype(np.float32)
    mag = np.abs(fe.stft(y))
    assert mag[:, 60].argmax() == 64          # 1 kHz -> bin 1000 / 15.625
    assert mag[64, 60] == pytest.approx(0.54 * 2048 / 2, rel=2e-3)
