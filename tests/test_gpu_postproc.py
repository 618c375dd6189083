"""threshold -> median -> region decode on the device: bit-exact against the oracle."""
import numpy as np
import pytest
import torch

from oracle import postproc as opp

pytestmark = pytest.mark.gpu


def _decode(strong, th=0.5, win=14, max_events=None):
    from bsed_b200 import engine
    ev, n = engine.median_decode(torch.from_numpy(strong).cuda(), th, win, max_events)
    ev, n = ev.cpu().numpy(), n.cpu().numpy()
    return [[tuple(int(v) for v in ev[b, k]) for k in range(min(n[b], ev.shape[1]))] for b in range(strong.shape[0])], n


@pytest.mark.parametrize("win", [1, 2, 7, 14, 15, 27])
def test_random_probabilities(win):
    rng = np.random.default_rng(win)
    strong = rng.random((8, 313, 20)).astype(np.float32)
    strong[1] = strong[1] ** 3            # sparse
    strong[2] = 1 - strong[2] ** 3        # dense
    for b in range(3, 8):                 # block structure like real detections
        for _ in range(6):
            c, on = rng.integers(0, 20), rng.integers(0, 300)
            strong[b, on:on + rng.integers(3, 80), c] = 0.95
    got, n = _decode(strong, win=win)
    for b in range(8):
        ref = opp.events_from_strong(strong[b], 0.5, win)
        assert got[b] == ref and n[b] == len(ref)


def test_edges():
    strong = np.zeros((4, 313, 20), dtype=np.float32)
    strong[1] = 1.0
    strong[2] = 0.5                        # exactly at the threshold: counts as active (>=)
    strong[3] = np.nextafter(np.float32(0.5), np.float32(0))
    got, n = _decode(strong)
    assert got[0] == [] and n[0] == 0
    assert got[1] == [(c, 0, 313) for c in range(20)]
    assert got[2] == [(c, 0, 313) for c in range(20)]
    assert got[3] == []


def test_small_shapes_and_truncation():
    rng = np.random.default_rng(3)
    strong = rng.random((2, 5, 3)).astype(np.float32)
    got, n = _decode(strong, win=14)
    for b in range(2):
        assert got[b] == opp.events_from_strong(strong[b], 0.5, 14)
    strong = (rng.random((1, 313, 20)) < 0.5).astype(np.float32)
    got, n = _decode(strong, win=1, max_events=10)
    ref = opp.events_from_strong(strong[0], 0.5, 1)
    assert n[0] == len(ref) and got[0] == ref[:10]


def test_reference_facing_dataframe():
    from bsed_b200 import evaluation_measures as em
    from bsed_b200.data import config as cfg
    strong = np.zeros((1, 313, 20), dtype=np.float32)
    strong[0, 31:62, 4] = 0.9
    strong[0, 300:313, 0] = 0.9
    ev = em.decode_events(torch.from_numpy(strong).cuda(), (0.5,), 14)[0.5][0]
    df = em.events_to_df(ev, cfg.bird_list, "clip_0", 4)
    ref = opp.to_seconds(opp.events_from_strong(strong[0]))
    assert list(df["event_label"]) == [cfg.bird_list[c] for c, _, _ in ref]
    assert np.allclose(df["onset"], [r[1] for r in ref]) and np.allclose(df["offset"], [r[2] for r in ref])
    assert df["offset"].max() <= 10.0 and (df["filename"] == "clip_0").all()
