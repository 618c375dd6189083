"""Stage-A feature cache written by this package (data.preprocess.write_feature_cache) and read back through the
reference-shaped datasets + transforms: the .npy holds exactly what preprocess() returns."""
import os

import numpy as np
import pandas as pd
import pytest

from bsed_b200.utilities import synth

pytestmark = pytest.mark.gpu


def test_write_cache_then_read_through_dataset(tmp_path):
    from bsed_b200.data import config as cfg
    from bsed_b200.data import dataload
    from bsed_b200.data.preprocess import preprocess, write_feature_cache
    from bsed_b200.data.Transforms import get_transforms
    from bsed_b200.utilities.ManyHotEncoder import ManyHotEncoder
    clips = synth.make_clips(2, seed=5)
    audio = np.concatenate([clips.reshape(-1), np.zeros(1234, dtype=np.float32)])
    ann = pd.DataFrame({"onset": [1.0, 12.5, 9.5], "offset": [2.0, 14.0, 10.5], "event_label": ["EATO", "WOTH", "NOCA"]})
    paths = write_feature_cache(audio, "rec", str(tmp_path), ann)
    assert [os.path.basename(p) for p in paths] == ["rec_0.npy", "rec_1.npy"]       # tail dropped
    mel0 = np.load(paths[0])
    assert mel0.dtype == np.float32 and mel0.shape == (1255, 128)
    np.testing.assert_array_equal(mel0, preprocess(clips[0]))                        # batched == one-clip entry point
    a1 = pd.read_csv(os.path.join(str(tmp_path), "annotation", "rec_1.txt"), sep="\t")
    assert list(a1.columns) == ["onset", "offset", "event_label"] and a1["event_label"].tolist() == ["WOTH"]
    assert abs(a1["onset"][0] - 2.5) < 1e-9                                          # shifted into clip time; 9.5-10.5 crosses -> dropped
    enc = ManyHotEncoder(cfg.bird_list, n_frames=313)
    ds = dataload.ENA_Dataset(str(tmp_path), enc.encode_strong_df, get_transforms(cfg.max_frames, None, 0, noise_dict_params={"mean": 0., "snr": cfg.noise_snr}))
    ((clean, noisy), target), path = ds[0]
    assert tuple(clean.shape) == (1, 1255, 128) and tuple(target.shape) == (313, 20) and path == paths[0]
    assert float(target[:, 0].sum()) > 0
