"""Bulk log-mel extraction + CRNN pseudo-label inference over a long recording (BASELINE.json configs[4];
reference: src/data/preprocess.py:196-204 segmentation + preprocess, src/audio_tagging.py:256-283 weak labels,
src/evaluation_measures.py:188-209 strong events).

    pseudo_label_stream(audio, model, predictor, ...) -> dict(weak_rows, events, n_clips, span)

The stream is cut into non-overlapping 10 s clips (tail dropped, as librosa.util.frame(..., 320000, 320000) does),
the clips are sharded over the ranks in contiguous blocks (independent units, no data-path
collective), and each rank runs, on its own GPU and entirely in libbsed.so kernels,
    framed STFT -> |X| -> mel -> dB (per-clip 80 dB clamp) -> CRNN + Predictor (eval) -> weak >= thr labels
    and threshold -> median(14) -> region decoding of the strong probabilities.
Only label rows and event lists cross back to the host; `gather=True` concatenates them in rank order.
"""
import numpy as np
import torch

from . import engine
from .data import config as cfg
from .data.preprocess import segment
from .evaluation_measures import decode_events
from .utilities import shard


def pseudo_label_stream(audio, model, predictor, batch_clips=48, weak_threshold=0.5, strong_threshold=0.5,
                        median_window=None, scaler=None, labels=None, rank=None, world=None, gather=True,
                        name_fmt="stream_{:05d}"):
    """audio: 1-D float32 waveform at cfg.sr (numpy, pinned or pageable host memory, or a CUDA tensor).
    Returns weak_rows [(clip name, "label,label")], events [(clip name, label, onset s, offset s)] for the clips of
    this rank (all clips if gather), n_clips (whole stream) and span (this rank's [begin, end))."""
    labels = labels or cfg.bird_list
    if median_window is None:
        median_window = max(int(cfg.median_window_s * cfg.sr / cfg.hop_size / cfg.pooling_time_ratio), 1) \
            if hasattr(cfg, "median_window_s") else 14
    dev = model._flat.device
    if dev.type != "cuda":
        raise RuntimeError("pseudo_label_stream needs the models on a CUDA device (no CPU fallback)")
    if rank is None:
        rank = torch.distributed.get_rank() if shard.world_size() > 1 else 0
    if world is None:
        world = shard.world_size()
    seg = cfg.sr * cfg.seg_sec
    n_clips = (audio.shape[0] // seg)
    begin, end = shard.clip_shard(n_clips, rank, world)
    was_training = (model.training, predictor.training)
    model.eval()
    predictor.eval()
    mean = std = None
    if scaler is not None:
        mean = torch.as_tensor(scaler.mean_, dtype=torch.float32, device=dev).contiguous()
        std = torch.as_tensor(scaler.std_, dtype=torch.float32, device=dev).contiguous()
    scale = cfg.pooling_time_ratio / (cfg.sr / cfg.hop_size)
    weak_rows, events = [], []
    with torch.no_grad():
        for b0 in range(begin, end, batch_clips):
            b1 = min(end, b0 + batch_clips)
            chunk = audio[b0 * seg:b1 * seg]
            if not torch.is_tensor(chunk):
                chunk = torch.from_numpy(np.ascontiguousarray(chunk, dtype=np.float32))
            clips = chunk.to(dev, non_blocking=True).reshape(b1 - b0, seg)
            x = engine.logmel(clips, cfg.max_frames, scaler_mean=mean, scaler_std=std)[:, None]     # fused frontend
            enc, _ = model(x)
            strong, weak = predictor(enc)
            decoded = decode_events(strong, (strong_threshold,), median_window)[strong_threshold]
            mask = (weak >= weak_threshold).cpu().numpy()
            for j in range(b1 - b0):
                name = name_fmt.format(b0 + j)
                names = [labels[c] for c in np.nonzero(mask[j])[0]]
                if names:
                    weak_rows.append((name, ",".join(names)))
                for c, on, off in decoded[j]:
                    events.append((name, labels[c], min(max(on * scale, 0.0), cfg.max_len_seconds),
                                   min(max(off * scale, 0.0), cfg.max_len_seconds)))
    model.train(was_training[0])
    predictor.train(was_training[1])
    if gather and world > 1:
        weak_rows = shard.gather_in_rank_order(weak_rows)
        events = shard.gather_in_rank_order(events)
    return dict(weak_rows=weak_rows, events=events, n_clips=n_clips, span=(begin, end))
