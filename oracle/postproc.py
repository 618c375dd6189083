"""Oracle (TEST INFRASTRUCTURE): restatement of the decision post-processing and label encoding.

PARITY UNPINNED by the reference (dcase_util is not installable here); the median rule is
cross-checked bit-exactly against scipy.ndimage.median_filter in tests/test_oracle_postproc.py.

Follows:
  * src/evaluation_measures.py:188-209       threshold -> median_filter((win,1)) -> decoder -> seconds
  * src/utilities/ManyHotEncoder.py:148-164  decode_strong (class-major, contiguous regions)
  * src/utilities/ManyHotEncoder.py:115-130  encode_strong_df (sec -> frame = int(s*sr//hop//ptr))
  * src/data/config.py:47-63                 sr, hop, pooling_time_ratio, median_window_s
dcase_util semantics restated:
  ProbabilityEncoder.binarization('global_threshold', t): 1 where p >= t  (recalled; p == t is
  tested explicitly so a drift would be visible).
  DecisionEncoder.find_contiguous_regions: XOR of neighbours -> change indices + 1, prepend 0 if
  the first frame is active, append len if the last is, reshape (-1, 2)  => runs [on, off).
"""
import numpy as np

SR = 32000
HOP = 255
POOLING_TIME_RATIO = 4
N_OUT_FRAMES = 313
MAX_LEN_SECONDS = 10.0
MEDIAN_WINDOW = max(int(0.45 * SR / HOP / POOLING_TIME_RATIO), 1)   # 14 (main.py:643-644)
FRAME_SECONDS = POOLING_TIME_RATIO / (SR / HOP)                      # 0.031875


def binarize(p, threshold=0.5):
    return (np.asarray(p) >= threshold).astype(np.int64)


def median_filter_time(b, win=MEDIAN_WINDOW):
    """scipy.ndimage.median_filter(b, (win, 1)) for binary b (T, C): mode='reflect'
    (half-sample symmetric), window covering [i - win//2, i + win - 1 - win//2]; for 0/1 data the
    median (element rank win//2 of the sorted window) is 1 iff ones >= win - win//2."""
    b = np.asarray(b).astype(np.int64)
    T = b.shape[0]
    left = win // 2
    right = win - 1 - left
    idx = np.arange(-left, T + right)
    # half-sample symmetric reflection, valid for any extension length
    period = 2 * T
    idx = np.mod(idx, period)
    idx = np.where(idx >= T, period - 1 - idx, idx)
    padded = b[idx]
    csum = np.concatenate([np.zeros((1,) + b.shape[1:], dtype=np.int64), np.cumsum(padded, axis=0)])
    ones = csum[win:win + T] - csum[0:T]
    need = win - win // 2
    return (ones >= need).astype(np.int64)


def find_contiguous_regions(col):
    col = np.asarray(col).astype(bool)
    change = np.logical_xor(col[1:], col[:-1]).nonzero()[0] + 1
    if col.size and col[0]:
        change = np.r_[0, change]
    if col.size and col[-1]:
        change = np.r_[change, col.size]
    return change.reshape((-1, 2))


def decode_strong(m):
    """-> list of (class_index, onset_frame, offset_frame), class-major then time."""
    out = []
    for c, col in enumerate(np.asarray(m).T):
        for on, off in find_contiguous_regions(col):
            out.append((c, int(on), int(off)))
    return out


def events_from_strong(strong, threshold=0.5, win=MEDIAN_WINDOW):
    """strong (T, C) probabilities -> [(class, on_frame, off_frame)] (frames, not seconds)."""
    return decode_strong(median_filter_time(binarize(strong, threshold), win))


def to_seconds(events):
    """onset/offset * pooling_time_ratio / (sr / hop), clipped to [0, 10]
    (evaluation_measures.py:208-209)."""
    res = []
    for c, on, off in events:
        res.append((c, float(np.clip(on * FRAME_SECONDS, 0, MAX_LEN_SECONDS)),
                    float(np.clip(off * FRAME_SECONDS, 0, MAX_LEN_SECONDS))))
    return res


def encode_strong(rows, n_frames=N_OUT_FRAMES, n_class=20):
    """rows: iterable of (onset_s, offset_s, class_index) -> (n_frames, n_class) float64 many-hot
    (ManyHotEncoder.py:115-130: frame = int(sec * sr // hop // pooling_time_ratio))."""
    y = np.zeros((n_frames, n_class))
    for on_s, off_s, c in rows:
        on = int(on_s * SR // HOP // POOLING_TIME_RATIO)
        off = int(off_s * SR // HOP // POOLING_TIME_RATIO)
        y[on:off, c] = 1
    return y


def weak_labels(weak, threshold=0.5):
    """src/audio_tagging.py:256-283: classes with weak probability >= threshold."""
    return [int(c) for c in np.nonzero(np.asarray(weak) >= threshold)[0]]
