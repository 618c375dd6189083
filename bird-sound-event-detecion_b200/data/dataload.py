"""Datasets over the reference's on-disk feature cache (reference: src/data/dataload.py:17-196).

Cache layout written by `data.preprocess.write_feature_cache` (and by the reference's ena_data_preprocess,
src/data/preprocess.py:226-229):
    <preprocess_dir>/wav/<name>_<k>.npy          (T <= 1255, 128) float32 amplitude-mel of one 10 s clip
    <preprocess_dir>/annotation/<name>_<k>.txt   TSV  onset  offset  event_label   (seconds inside the clip)
Items have the reference's shape:  (((clean, noisy), target), path)  with the transform of get_transforms.
"""
import os
from glob import glob

import numpy as np
import pandas as pd
from torch.utils.data.dataset import Dataset

from . import config as cfg


class _CacheDataset(Dataset):
    def __init__(self, preprocess_dir, encod_func, transform, compute_log=False):
        self.sample_rate = cfg.sr
        self.preprocess_dir = preprocess_dir
        self.pooling_time_ratio = cfg.pooling_time_ratio
        self.n_frames = cfg.max_frames // self.pooling_time_ratio
        self.hop_size = cfg.hop_size
        self.transform = transform
        self.encod_func = encod_func
        self.annotation_dir = os.path.join(self.preprocess_dir, "annotation")
        self.feature_dir = os.path.join(self.preprocess_dir, "wav")
        self.feature_file_list = sorted(glob(os.path.join(self.feature_dir, "*.npy")))
        self.labels = cfg.bird_list

    def __len__(self):
        return len(self.feature_file_list)

    def encode(self, label_df):
        """Strong many-hot target (dataload.py:59-82): frame = int(sec * sr // hop // pooling_time_ratio)."""
        y = np.zeros((self.n_frames, len(self.labels)))
        if isinstance(label_df, pd.DataFrame):
            for _, row in label_df.iterrows():
                i = self.labels.index(row["event_label"])
                onset = int(row["onset"] * self.sample_rate // self.hop_size // self.pooling_time_ratio)
                offset = int(row["offset"] * self.sample_rate // self.hop_size // self.pooling_time_ratio)
                y[onset:offset, i] = 1
        return y

    def _target_df(self, path):
        name = os.path.splitext(os.path.basename(path))[0]
        return pd.read_csv(os.path.join(self.annotation_dir, name + ".txt"), sep="\t")

    def __getitem__(self, index):
        path = self.feature_file_list[index]
        features = np.load(path)
        df = self._target_df(path)
        target = self.encod_func(df) if self.encod_func is not None else self.encode(df)
        sample = self.transform((features, target)) if self.transform is not None else (features, target)
        return (sample, path)


class ENA_Dataset(_CacheDataset):
    """dataload.py:17-82 (strongly labelled real clips)."""


class SYN_Dataset(_CacheDataset):
    """dataload.py:127-196 (synthetic soundscapes; same layout)."""


class ENA_Dataset_unlabeled(_CacheDataset):
    """dataload.py:84-126: targets come from a pseudo-label TSV (`filename`, `event_labels` comma-joined; the
    reference hard-codes its path) -- rows are matched on the full feature path, as in the reference."""

    def __init__(self, preprocess_dir, encod_func, transform, compute_log=False, pseudo_label_tsv=None):
        super().__init__(preprocess_dir, encod_func, transform, compute_log)
        self.annotation_dir = pseudo_label_tsv
        self._df = pd.read_csv(pseudo_label_tsv, sep="\t") if pseudo_label_tsv else pd.DataFrame(
            columns=["filename", "event_labels"])

    def _target_df(self, path):
        return self._df[self._df["filename"] == path]["event_labels"]

    def encode(self, labels):
        y = np.zeros(len(self.labels))
        for entry in labels:
            for ev in str(entry).split(","):
                if ev and ev != "nan":
                    y[self.labels.index(ev)] = 1
        return y


class ConcatDataset(Dataset):
    """dataload.py:198-254: index-concatenation of datasets."""

    def __init__(self, datasets):
        self.datasets = list(datasets)
        assert len(self.datasets) > 0, "datasets should not be an empty iterable"
        self.cumulative_sizes = list(np.cumsum([len(d) for d in self.datasets]))

    def __len__(self):
        return int(self.cumulative_sizes[-1])

    def __getitem__(self, idx):
        k = int(np.searchsorted(self.cumulative_sizes, idx, side="right"))
        prev = 0 if k == 0 else int(self.cumulative_sizes[k - 1])
        return self.datasets[k][idx - prev]
