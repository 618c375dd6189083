"""On-disk formats shared with the reference: feature cache datasets, pseudo-label TSV,
checkpoint dict and the cnn. / cnn.cnn. key spellings.  CPU only (no kernels run)."""
import os

import numpy as np
import pandas as pd
import torch

from helpers import golden
from bsed_b200.data import config as cfg
from bsed_b200.data import dataload
from bsed_b200.utilities import checkpoint
from bsed_b200.utilities.ManyHotEncoder import ManyHotEncoder


def _make_cache(root, n=3):
    os.makedirs(os.path.join(root, "wav"))
    os.makedirs(os.path.join(root, "annotation"))
    rng = np.random.default_rng(0)
    for k in range(n):
        np.save(os.path.join(root, "wav", f"rec_{k}.npy"), rng.random((1255 - 5 * k, 128), dtype=np.float32))
        pd.DataFrame({"onset": [0.5 + k, 3.0], "offset": [1.5 + k, 3.4], "event_label": ["EATO", cfg.bird_list[5 + k]]}).to_csv(
            os.path.join(root, "annotation", f"rec_{k}.txt"), sep="\t", index=False)


def test_cache_datasets_return_the_reference_item_shape(tmp_path):
    root = str(tmp_path / "preprocess")
    _make_cache(root)
    enc = ManyHotEncoder(cfg.bird_list, n_frames=cfg.max_frames // cfg.pooling_time_ratio)
    ds = dataload.ENA_Dataset(root, enc.encode_strong_df, transform=None)
    assert len(ds) == 3
    (feat, target), path = ds[1]
    assert path.endswith("rec_1.npy") and feat.shape == (1250, 128) and target.shape == (313, 20)
    on = int(1.5 * cfg.sr // cfg.hop_size // cfg.pooling_time_ratio)
    off = int(2.5 * cfg.sr // cfg.hop_size // cfg.pooling_time_ratio)
    assert target[on:off, 0].all() and target[:on, 0].sum() == 0 and target[off:, 0].sum() == 0
    np.testing.assert_array_equal(dataload.ENA_Dataset(root, None, None)[1][0][1], target)   # built-in encode == encoder
    both = dataload.ConcatDataset([ds, dataload.SYN_Dataset(root, enc.encode_strong_df, None)])
    assert len(both) == 6 and both[4][1] == ds[1][1]


def test_unlabeled_dataset_reads_the_pseudo_label_tsv(tmp_path):
    root = str(tmp_path / "preprocess")
    _make_cache(root)
    files = sorted(os.listdir(os.path.join(root, "wav")))
    tsv = str(tmp_path / "pseudo.tsv")
    pd.DataFrame({"filename": [os.path.join(root, "wav", files[0]), os.path.join(root, "wav", files[2])],
                  "event_labels": ["EATO,WOTH", "BAWW"]}).to_csv(tsv, sep="\t", index=False)
    enc = ManyHotEncoder(cfg.bird_list, n_frames=313)
    ds = dataload.ENA_Dataset_unlabeled(root, enc.encode_weak, None, pseudo_label_tsv=tsv)
    (f0, t0), _ = ds[0]
    (f1, t1), _ = ds[1]
    assert t0.tolist() == [1, 1] + [0] * 18 and t1.sum() == 0
    assert ds[2][0][1][19] == 1


def test_checkpoint_round_trip_and_key_spellings(tmp_path):
    keys = [str(k) for k in golden("state_dict_keys.npz")["keys"]]
    assert len(keys) == 79 and keys[0] == "cnn.conv0.weight"
    sd = {k: torch.full((2,), float(i)) for i, k in enumerate(keys)}
    renamed = checkpoint.reference_renamed_state_dict(sd)
    assert "cnn.cnn.conv0.weight" in renamed and "rnn.rnn.weight_ih_l0" in renamed and len(renamed) == 79
    back = checkpoint.canonical_state_dict(renamed)
    assert list(back) == keys and all(torch.equal(back[k], sd[k]) for k in keys)
    assert checkpoint.canonical_state_dict(sd) == sd                      # already canonical: untouched

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.arange(3.0))
    m, p = M(), M()
    state = checkpoint.build_state(m, p, {"nclass": 20}, {"nclass": 20}, ema_model=M(), ema_predictor=M(), epoch=7)
    assert set(state) >= {"model", "model_p", "model_ema", "model_p_ema", "pooling_time_ratio", "median_window", "epoch"}
    assert set(state["model"]) == {"name", "args", "kwargs", "state_dict"}
    path = str(tmp_path / "baseline_epoch_7")
    checkpoint.save_state(state, path)
    loaded = torch.load(path, weights_only=False)
    m2, p2 = M(), M()
    with torch.no_grad():
        m2.w.zero_()
    assert checkpoint.load_models(loaded, m2, p2) == 7 and torch.equal(m2.w, m.w)
