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
    if compute_log:
        return engine.logmel(audio, frames=1 + audio.shape[1] // cfg.hop_size)
    return engine.melspec(audio)


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


def clip_annotations(annotation_df, count, seg_sec=cfg.seg_sec):
    """Rows of an (onset, offset, event_label) table that lie inside 10 s segment `count`, shifted to clip time
    (src/data/preprocess.py:206-219; the reference's merge / overlap-splitting of events that cross a segment border
    is dataset curation and stays with the caller)."""
    lo, hi = count * seg_sec, (count + 1) * seg_sec
    d = annotation_df[(annotation_df["onset"] >= lo) & (annotation_df["offset"] < hi)][["onset", "offset", "event_label"]].copy()
    d["onset"] -= lo
    d["offset"] -= lo
    return d.drop_duplicates()


def write_feature_cache(audio, name, saved_path, annotation_df=None, batch_clips=64):
    """One recording -> the reference's stage-A cache (src/data/preprocess.py:196-229): non-overlapping 10 s segments
    (tail dropped), `<saved_path>/wav/<name>_<k>.npy` = (1255, 128) float32 amplitude-mel (compute_log=False) and
    `<saved_path>/annotation/<name>_<k>.txt` = TSV onset/offset/event_label.  The mel features of all segments come
    from batched bsed_melspec launches.  Returns the list of .npy paths."""
    import os
    import pandas as pd
    mel_dir, ann_dir = os.path.join(saved_path, "wav"), os.path.join(saved_path, "annotation")
    os.makedirs(mel_dir, exist_ok=True)
    os.makedirs(ann_dir, exist_ok=True)
    segs = segment(np.ascontiguousarray(audio, dtype=np.float32))
    dev = _device()
    paths = []
    for b0 in range(0, len(segs), batch_clips):
        t = torch.from_numpy(np.ascontiguousarray(segs[b0:b0 + batch_clips])).to(dev)
        mel = preprocess_batch(t, compute_log=False).cpu().numpy().astype(np.float32)
        for j in range(mel.shape[0]):
            k = b0 + j
            path = os.path.join(mel_dir, f"{name}_{k}.npy")
            np.save(path, mel[j])
            ann = clip_annotations(annotation_df, k) if annotation_df is not None else pd.DataFrame(
                columns=["onset", "offset", "event_label"])
            ann.to_csv(os.path.join(ann_dir, f"{name}_{k}.txt"), sep="\t", index=False)
            paths.append(path)
    return paths
