"""Frontend stage A with the reference's signature (src/data/preprocess.py:18-45):

    preprocess(audio: np.ndarray[float32, (N,)], compute_log=False) -> np.ndarray[float32, (T, 128)]

STFT (n_fft 2048, hop 255, symmetric Hamming, reflect) -> |X| -> 128 Slaney mel bands, computed by
csrc/frontend.cu.  Host buffers in, host buffers out (the copies are part of the call, as they are
for a caller of the reference); `preprocess_batch` is the device-resident batched form.
"""
import numpy as np
import torch

from .. import engine
from . import config as cfg


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("bird-sound-event-detecion_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def preprocess_batch(audio, compute_log=False):
    """audio: (B, N) float32 CUDA tensor -> (B, T, 128) CUDA tensor."""
    mel = engine.melspec(audio)
    if compute_log:
        mel = engine.amp_to_db(mel, frames=mel.shape[1])
    return mel


def preprocess(audio, compute_log=False):
    a = np.ascontiguousarray(audio, dtype=np.float32)
    if a.ndim != 1:
        raise ValueError("preprocess expects a mono waveform (N,)")
    t = torch.from_numpy(a).pin_memory().to(_device(), non_blocking=True)[None]
    mel = preprocess_batch(t, compute_log)[0]
    return mel.cpu().numpy().astype(np.float32)


def segment(audio, seg_samples=cfg.sr * cfg.seg_sec):
    """librosa.util.frame(audio, seg, seg, axis=0): non-overlapping 10 s segments, tail dropped
    (src/data/preprocess.py:196)."""
    n = (len(audio) // seg_samples) * seg_samples
    return np.asarray(audio[:n]).reshape(-1, seg_samples)
