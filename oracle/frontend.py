"""Oracle (TEST INFRASTRUCTURE): float64 numpy restatement of the log-mel frontend.

PARITY UNPINNED by the reference (librosa is not installable here and the reference
ships no test vectors); cross-checked in tests/test_oracle_frontend.py against
torch.stft / torchaudio.functional.melscale_fbanks.

Follows, line by line in meaning (not in code):
  * src/data/preprocess.py:18-45      preprocess(audio, compute_log=False)
  * src/data/config.py:47-57          sr / n_window / hop_size / n_mels / f_min / f_max
  * src/data/Transforms.py:74-86      ApplyLog  -> librosa.amplitude_to_db
  * src/data/Transforms.py:89-109     pad_trunc_seq
  * src/data/Transforms.py:155-179    AugmentGaussianNoise.gaussian_noise (noise injected explicitly)
  * src/utilities/Scaler.py:104-110   Scaler.normalize
librosa semantics restated (versions 0.9-0.10; the reference pins none):
  stft:  reflect-pad n_fft//2, frame t = ypad[hop*t : hop*t+n_fft] * window (float64
         product), rfft in float64, result stored as complex64.
  mel :  Slaney scale (htk=False), norm=None, float32 basis, projection in float32.
  dB  :  10*log10(max(amin^2, x^2)) - 10*log10(max(amin^2, ref^2)), then
         max(., max_over_whole_array - top_db); ref=1, amin=1e-5, top_db=80.
"""
import numpy as np
import scipy.fft

SR = 32000
N_FFT = 2048
HOP = 255
N_MELS = 128
F_MIN = 0.0
F_MAX = 16000.0
MAX_FRAMES = 1255  # ceil(10 * 32000 / 255), src/data/config.py:59


def hamming_window(n=N_FFT):
    """np.hamming(n): symmetric, float64 (src/data/preprocess.py:19)."""
    k = np.arange(n, dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * k / (n - 1))


def n_frames_for(n_samples, hop=HOP):
    return 1 + n_samples // hop


def stft(audio, n_fft=N_FFT, hop=HOP):
    """librosa.stft(audio, n_fft, hop, window=hamming, center=True, pad_mode='reflect')
    -> complex64 (n_fft//2+1, n_frames).  src/data/preprocess.py:21-28."""
    y = np.asarray(audio)
    win = hamming_window(n_fft)
    ypad = np.pad(y, n_fft // 2, mode="reflect")
    nfr = 1 + (ypad.shape[0] - n_fft) // hop
    idx = hop * np.arange(nfr)[:, None] + np.arange(n_fft)[None, :]
    frames = ypad[idx].astype(np.float64) * win[None, :]
    spec = scipy.fft.rfft(frames, axis=-1)           # float64 FFT
    return spec.T.astype(np.complex64)               # stored as complex64


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mel = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz,
                    min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mel)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    f = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f)


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=F_MIN, fmax=F_MAX):
    """librosa.filters.mel(htk=False, norm=None) -> float32 (n_mels, n_fft//2+1)."""
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    edges = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(edges)
    ramps = edges[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, fftfreqs.shape[0]), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    return w


_FB_CACHE = {}


def preprocess(audio, compute_log=False):
    """src/data/preprocess.py:18-45 -> float32 (n_frames, 128) amplitude-mel."""
    key = (SR, N_FFT, N_MELS, F_MIN, F_MAX)
    if key not in _FB_CACHE:
        _FB_CACHE[key] = mel_filterbank()
    fb = _FB_CACHE[key]
    mag = np.abs(stft(audio))                        # float32 (1025, T)
    mel = fb @ mag                                   # float32 GEMM (128, T)
    if compute_log:
        mel = amplitude_to_db(mel)
    return np.ascontiguousarray(mel.T).astype(np.float32)


def amplitude_to_db(x, amin=1e-5, top_db=80.0):
    """librosa.amplitude_to_db(x, ref=1.0, amin=1e-5, top_db=80) in x's dtype
    (src/data/Transforms.py:86; float32 for cached features, float64 once noise was added)."""
    x = np.asarray(x)
    power = np.square(np.abs(x))
    a2 = np.asarray(amin * amin, dtype=power.dtype)
    log_spec = 10.0 * np.log10(np.maximum(a2, power))
    log_spec = log_spec - 10.0 * np.log10(np.maximum(a2, np.asarray(1.0, dtype=power.dtype)))
    if top_db is not None and log_spec.size:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec.astype(power.dtype)


def noise_std(mel, snr=30.0):
    """Per-mel-bin std of AugmentGaussianNoise.gaussian_noise (Transforms.py:172)."""
    mel = np.asarray(mel)
    return np.sqrt(np.mean((mel ** 2) * (10 ** (-snr / 10)), axis=-2))


def add_noise(mel, unit_noise, snr=30.0):
    """features + N(0, std_f): `unit_noise` is a standard-normal array of mel's shape
    supplied by the caller so both implementations see the same draw
    (np.random.normal(0, std, shape) == std * standard_normal(shape) in distribution)."""
    std = noise_std(mel, snr)
    return np.asarray(mel, dtype=np.float64) + std[None, :] * np.asarray(unit_noise, dtype=np.float64)


def pad_trunc_seq(x, max_len=MAX_FRAMES):
    """src/data/Transforms.py:89-109 (zero rows appended AFTER the log)."""
    x = np.asarray(x)
    if x.shape[-2] <= max_len:
        pad = [(0, 0)] * (x.ndim - 2) + [(0, max_len - x.shape[-2]), (0, 0)]
        return np.pad(x, pad, mode="constant")
    return x[..., :max_len, :]


def transform(mel, unit_noise=None, snr=30.0, frames=MAX_FRAMES, mean=None, std=None):
    """get_transforms(frames, scaler, add_axis=0, noise_dict_params={'snr':30}) applied to one
    cached amplitude-mel (Transforms.py:304-322).  Returns float32 (1, frames, 128), or the pair
    (clean, noisy) when a noise draw is given (student / teacher inputs, main.py:198)."""
    def tail(z):
        z = pad_trunc_seq(amplitude_to_db(z), frames).astype(np.float32)[None]
        if mean is not None:
            z = ((z - mean) / std).astype(np.float32)
        return z
    if unit_noise is None:
        return tail(np.asarray(mel, dtype=np.float32))
    return tail(np.asarray(mel, dtype=np.float32)), tail(add_noise(mel, unit_noise, snr))


def logmel(audio):
    """audio (n,) f32 -> (1255,128) f32 log-mel: preprocess() then ApplyLog + PadOrTrunc."""
    return transform(preprocess(audio))[0]
