"""Label encoding / decoding (reference: src/utilities/ManyHotEncoder.py).  encode_* are host-side
integer bookkeeping; decode_strong is backed by the device post-processing kernel when given
probabilities (see evaluation_measures.get_predictions), and has a host form for already-binary
arrays that follows dcase_util's find_contiguous_regions."""
import numpy as np

from ..data import config as cfg


class ManyHotEncoder:
    def __init__(self, labels, n_frames=None):
        if isinstance(labels, np.ndarray):
            labels = labels.tolist()
        self.labels = list(labels)
        self.n_frames = n_frames

    def encode_weak(self, labels):
        """:27-54"""
        if isinstance(labels, str):
            if labels == "empty":
                return np.zeros(len(self.labels)) - 1
            labels = [labels]
        if hasattr(labels, "columns"):  # DataFrame
            labels = [] if labels.empty else labels["event_label"]
        y = np.zeros(len(self.labels))
        for label in labels:
            if label is None or (isinstance(label, float) and np.isnan(label)):
                continue
            for event in str(label).split(","):
                if event != "" and event != "nan":
                    y[self.labels.index(event)] = 1
        return y

    def encode_strong_df(self, label_df):
        """:115-130  frame = int(sec * sr // hop // pooling_time_ratio); y[on:off, class] = 1.
        Accepts a DataFrame with onset/offset/event_label or an iterable of (onset, offset, label)."""
        y = np.zeros((self.n_frames, len(self.labels)))
        rows = label_df.itertuples(index=False) if hasattr(label_df, "itertuples") else label_df
        for row in rows:
            if hasattr(row, "onset"):
                on_s, off_s, lab = row.onset, row.offset, row.event_label
            else:
                on_s, off_s, lab = row
            i = self.labels.index(lab) if not isinstance(lab, (int, np.integer)) else int(lab)
            onset = int(on_s * cfg.sr // cfg.hop_size // cfg.pooling_time_ratio)
            offset = int(off_s * cfg.sr // cfg.hop_size // cfg.pooling_time_ratio)
            y[onset:offset, i] = 1
        return y

    def decode_weak(self, labels):
        return [self.labels[i] for i, v in enumerate(labels) if v == 1]

    def decode_strong(self, labels):
        """:148-164  binary (T, C) -> [[label, onset_frame, offset_frame], ...] class-major."""
        out = []
        arr = np.asarray(labels)
        for i, col in enumerate(arr.T):
            col = col.astype(bool)
            change = np.logical_xor(col[1:], col[:-1]).nonzero()[0] + 1
            if col.size and col[0]:
                change = np.r_[0, change]
            if col.size and col[-1]:
                change = np.r_[change, col.size]
            for on, off in change.reshape(-1, 2):
                out.append([self.labels[i], int(on), int(off)])
        return out

    def state_dict(self):
        return {"labels": self.labels, "n_frames": self.n_frames}

    @classmethod
    def load_state_dict(cls, state_dict):
        return cls(state_dict["labels"], state_dict["n_frames"])
