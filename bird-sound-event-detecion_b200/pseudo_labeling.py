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
    batches = [(b0, min(end, b0 + batch_clips)) for b0 in range(begin, end, batch_clips)]
    # Two-deep pipeline: the audio of batch k + 1 crosses PCIe on a copy stream while batch k computes, and the results of
    # batch k (event lists, weak mask) come back through pinned buffers that the host reads one batch later -- the
    # compute stream never waits for the host.
    cur = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev)
    on_device = torch.is_tensor(audio) and audio.is_cuda
    dbuf = [None, None] if on_device else [torch.empty(batch_clips * seg, dtype=torch.float32, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    for e in freed:
        e.record(cur)

    def h2d(k):
        b0, b1 = batches[k]
        slot = k & 1
        chunk = audio[b0 * seg:b1 * seg]
        if on_device:
            dbuf[slot] = chunk.float().contiguous()
            ready[slot].record(cur)
            return
        if not torch.is_tensor(chunk):
            chunk = torch.from_numpy(np.ascontiguousarray(chunk, dtype=np.float32))
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])               # the batch that last used this buffer is done with it
            dbuf[slot][:chunk.numel()].copy_(chunk, non_blocking=True)
            ready[slot].record(copy_stream)

    def finalize(job):
        b0, b1, ev_h, n_h, mask_h, done = job
        done.synchronize()
        ev, n, mask = ev_h.numpy(), n_h.numpy(), mask_h.numpy()
        for j in range(b1 - b0):
            name = name_fmt.format(b0 + j)
            names = [labels[c] for c in np.nonzero(mask[j])[0]]
            if names:
                weak_rows.append((name, ",".join(names)))
            for k in range(min(int(n[j]), ev.shape[1])):
                c, on, off = (int(v) for v in ev[j, k])
                events.append((name, labels[c], min(max(on * scale, 0.0), cfg.max_len_seconds),
                               min(max(off * scale, 0.0), cfg.max_len_seconds)))

    n_cls = len(labels)
    max_events = n_cls * ((cfg.max_frames // cfg.pooling_time_ratio + 1) // 2)
    host = [(torch.empty(batch_clips, max_events, 3, dtype=torch.int32).pin_memory(),
             torch.empty(batch_clips, dtype=torch.int32).pin_memory(),
             torch.empty(batch_clips, n_cls, dtype=torch.bool).pin_memory()) for _ in range(2)] if batches else []
    pending = None
    with torch.no_grad():
        if batches:
            h2d(0)
        for k, (b0, b1) in enumerate(batches):
            slot = k & 1
            if k + 1 < len(batches):
                h2d(k + 1)
            cur.wait_event(ready[slot])
            clips = dbuf[slot][:(b1 - b0) * seg].reshape(b1 - b0, seg)
            x = engine.logmel(clips, cfg.max_frames, scaler_mean=mean, scaler_std=std)[:, None]     # fused frontend
            freed[slot].record(cur)
            enc, _ = model(x)
            strong, weak = predictor(enc)
            ev, n = engine.median_decode(strong, strong_threshold, median_window, max_events)
            ev_h, n_h, mask_h = (t[:b1 - b0] for t in host[slot])     # free: the job that used this slot was read last turn
            ev_h.copy_(ev, non_blocking=True)
            n_h.copy_(n, non_blocking=True)
            mask_h.copy_(weak >= weak_threshold, non_blocking=True)
            done = torch.cuda.Event()
            done.record(cur)
            if pending is not None:
                finalize(pending)
            pending = (b0, b1, ev_h, n_h, mask_h, done)
        if pending is not None:
            finalize(pending)
    model.train(was_training[0])
    predictor.train(was_training[1])
    if gather and world > 1:
        weak_rows = shard.gather_in_rank_order(weak_rows)
        events = shard.gather_in_rank_order(events)
    return dict(weak_rows=weak_rows, events=events, n_clips=n_clips, span=(begin, end))
