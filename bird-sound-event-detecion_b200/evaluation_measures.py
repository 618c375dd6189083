"""Inference + decision post-processing with the reference's entry point
(src/evaluation_measures.py:123-283, prediction half):

    get_predictions(model, dataloader, decoder, pooling_time_ratio=1, thresholds=[0.5],
                    median_window=1, save_predictions=None, ..., predictor=None, fpn=False)
        -> (prediction_df | [prediction_df per threshold], groundtruth_df, duration_df)

Per batch: CRNN + Predictor forward, then threshold -> median filter over time -> contiguous-region
decoding on the device (csrc/frontend.cu: median_decode_kernel); only the event list crosses back to
the host.  Frames are converted to seconds with pooling_time_ratio / (sr / hop) and clipped to
[0, max_len_seconds] exactly as the reference does (:208-209).  The sed_eval / psds_eval metric half
of the reference file is CPU bookkeeping in un-vendored packages and is out of scope.
"""
import os
import os.path as osp

import numpy as np
import pandas as pd
import torch

from . import engine
from .data import config as cfg


def decode_events(strong, thresholds=(0.5,), median_window=1):
    """strong (B, T, C) CUDA probabilities -> {threshold: [per clip [(class, on_frame, off_frame)]]}."""
    out = {}
    for th in thresholds:
        ev, n = engine.median_decode(strong, th, median_window)
        ev, n = ev.cpu().numpy(), n.cpu().numpy()
        out[th] = [[tuple(int(v) for v in ev[b, k]) for k in range(n[b])] for b in range(strong.shape[0])]
    return out


def events_to_df(events, labels, filename, pooling_time_ratio):
    rows = [(labels[c], on, off) for c, on, off in events]
    pred = pd.DataFrame(rows, columns=["event_label", "onset", "offset"])
    if len(pred):
        scale = pooling_time_ratio / (cfg.sr / cfg.hop_size)
        pred[["onset", "offset"]] = (pred[["onset", "offset"]].astype(float) * scale).clip(0, cfg.max_len_seconds)
    pred["filename"] = filename
    return pred


def get_predictions(model, dataloader, decoder, pooling_time_ratio=1, thresholds=[0.5], median_window=1,
                    save_predictions=None, del_model=False, learned_post=False, predictor=None, fpn=False,
                    saved_feature_dir=None):
    if learned_post:
        raise NotImplementedError("class-wise (learned_post) median windows are outside the hot path")
    if predictor is None:
        raise ValueError("get_predictions needs the Predictor (the reference's predictor != None branch)")
    labels = getattr(getattr(decoder, "__self__", None), "labels", None) or cfg.bird_list
    prediction_dfs = {th: [] for th in thresholds}
    filename_list, annotation_folder_list = [], []
    dev = model._flat.device
    for i, (((input_data, _ema_input), _target), selected_file_path) in enumerate(dataloader):
        filename = [osp.splitext(osp.basename(f))[0] for f in selected_file_path]
        folders = [osp.join(osp.dirname(osp.dirname(f)), "annotation") for f in selected_file_path]
        with torch.no_grad():
            encoded_x, feature_out = model(input_data.to(dev, non_blocking=True))
            pred_strong, _ = predictor(encoded_x, inference=True) if fpn else predictor(encoded_x)
        if saved_feature_dir is not None:
            np.save(osp.join(saved_feature_dir, '{}'.format(i)), feature_out.cpu().numpy())
        decoded = decode_events(pred_strong, thresholds, median_window)
        for th in thresholds:
            for j, ev in enumerate(decoded[th]):
                prediction_dfs[th].append(events_to_df(ev, labels, filename[j], pooling_time_ratio))
        filename_list += filename
        annotation_folder_list += folders
    cols = ["event_label", "onset", "offset", "filename"]
    for th in thresholds:
        dfs = [d for d in prediction_dfs[th] if len(d)]
        prediction_dfs[th] = pd.concat(dfs, ignore_index=True) if dfs else pd.DataFrame(columns=cols)

    uniq = list(dict.fromkeys(zip(filename_list, annotation_folder_list)))
    duration_df = pd.DataFrame({"filename": [f for f, _ in uniq], "duration": 10})
    gts = []
    for f, folder in uniq:
        path = osp.join(folder, f + ".txt")
        if osp.exists(path):
            g = pd.read_csv(path, sep="\t")
            g["filename"] = f
            gts.append(g)
    groundtruth_df = pd.concat(gts, ignore_index=True) if gts else None

    if save_predictions is not None:
        if isinstance(save_predictions, str):
            if len(thresholds) == 1:
                save_predictions = [save_predictions]
            else:
                base, ext = osp.splitext(save_predictions)
                save_predictions = [osp.join(base, f"{th:.3f}{ext}") for th in thresholds]
        assert len(save_predictions) == len(thresholds)
        for path, th in zip(save_predictions, thresholds):
            if osp.dirname(path):
                os.makedirs(osp.dirname(path), exist_ok=True)
            prediction_dfs[th].to_csv(path, index=False, sep="\t", float_format="%.3f")

    res = [prediction_dfs[th] for th in thresholds]
    return (res[0] if len(res) == 1 else res), groundtruth_df, duration_df


def get_weak_predictions(model, predictor, dataloader, labels=None, threshold=0.5):
    """Weak pseudo-labelling loop of src/audio_tagging.py:256-283 (CRNN + Predictor) and of
    src/audio_tagging_inference.py:289-316 (Net_resnet, `predictor=None`: the model returns the weak probabilities
    itself): weak >= threshold -> comma-joined labels per file (rows only for files with at least one label)."""
    labels = labels or cfg.bird_list
    rows = []
    dev = next(model.parameters()).device
    for (((input_data, _e), _t), paths) in dataloader:
        with torch.no_grad():
            if predictor is None:
                weak = model(input_data.to(dev, non_blocking=True))
            else:
                enc, _ = model(input_data.to(dev, non_blocking=True))
                _, weak = predictor(enc)
        mask = (weak >= threshold).cpu().numpy()
        for j, p in enumerate(paths):
            names = [labels[c] for c in np.nonzero(mask[j])[0]]
            if names:
                rows.append((p, ",".join(names)))
    return pd.DataFrame(rows, columns=["filename", "event_labels"])
