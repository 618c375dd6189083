"""Seeded synthetic clips and labels: the workload generator for tests and
bench.py.  There is no dataset in the reference tree and no network; every measured configuration
runs on these."""
import numpy as np

SR = 32000
CLIP_SAMPLES = 320000
N_CLASS = 20


def make_clips(n, seed=2023, n_samples=CLIP_SAMPLES, edge_cases=False):
    """n clips of 0.05*N(0,1) background + 2-5 linear chirps (1-12 kHz, 0.2-2.5 s, amplitude
    U[0.05,0.5], Hann-ramped).  With edge_cases the first clips are: all zeros, a single impulse,
    a full-scale 1 kHz sine."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, n_samples), dtype=np.float32)
    t = np.arange(n_samples, dtype=np.float64) / SR
    for i in range(n):
        y = 0.05 * rng.standard_normal(n_samples)
        for _ in range(int(rng.integers(2, 6))):
            dur = rng.uniform(0.2, 2.5)
            dur = min(dur, n_samples / SR * 0.9)
            t0 = rng.uniform(0, n_samples / SR - dur)
            f0, f1 = rng.uniform(1000, 12000, size=2)
            amp = rng.uniform(0.05, 0.5)
            i0, i1 = int(t0 * SR), int((t0 + dur) * SR)
            tt = t[i0:i1] - t0
            phase = 2 * np.pi * (f0 * tt + 0.5 * (f1 - f0) / dur * tt * tt)
            y[i0:i1] += amp * np.hanning(i1 - i0) * np.sin(phase)
        out[i] = y.astype(np.float32)
    if edge_cases:
        if n > 0:
            out[0] = 0.0
        if n > 1:
            out[1] = 0.0
            out[1, n_samples // 3] = 1.0
        if n > 2:
            out[2] = np.sin(2 * np.pi * 1000.0 * t).astype(np.float32)
    return out


def make_events(n, seed=2023):
    """Per clip 1-4 events: (onset_s, offset_s, class) with class ~ U{0..19}, onset ~ U[0,9],
    duration ~ U[0.25,3] s."""
    rng = np.random.default_rng(seed + 1)
    clips = []
    for _ in range(n):
        ev = []
        for _ in range(int(rng.integers(1, 5))):
            on = rng.uniform(0, 9)
            off = min(10.0, on + rng.uniform(0.25, 3.0))
            ev.append((float(on), float(off), int(rng.integers(0, N_CLASS))))
        clips.append(ev)
    return clips


def make_targets(n, seed=2023, n_frames=313):
    """(n, 313, 20) float32 many-hot strong targets encoded like ManyHotEncoder.encode_strong_df."""
    y = np.zeros((n, n_frames, N_CLASS), dtype=np.float32)
    for i, ev in enumerate(make_events(n, seed)):
        for on_s, off_s, c in ev:
            on = int(on_s * SR // 255 // 4)
            off = int(off_s * SR // 255 // 4)
            y[i, on:off, c] = 1
    return y


def make_logmel_like(n, seed=2023, n_frames=1255, n_mels=128):
    """Cheap stand-in for log-mel inputs when only the CRNN is under test: smooth random fields in
    the dB range of real features."""
    rng = np.random.default_rng(seed + 2)
    base = rng.standard_normal((n, n_frames // 8 + 2, n_mels // 8 + 2))
    up = np.repeat(np.repeat(base, 8, axis=1), 8, axis=2)[:, :n_frames, :n_mels]
    x = -30.0 + 12.0 * up + 3.0 * rng.standard_normal((n, n_frames, n_mels))
    return np.ascontiguousarray(x[:, None], dtype=np.float32)
