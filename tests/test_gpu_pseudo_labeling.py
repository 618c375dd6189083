"""Config 5 of BASELINE.json: bulk log-mel extraction + CRNN pseudo-label inference over one long stream, sharded by
clips (bird-sound-event-detecion_b200/pseudo_labeling.py) against the CPU oracle run clip by clip."""
import numpy as np
import pytest
import torch

from helpers import bsed_models, max_abs, oracle_models
from bsed_b200.utilities import synth
from oracle import frontend as ofe
from oracle import postproc as opp

pytestmark = pytest.mark.gpu

SEG = 320000


def _stream(n_clips=5, tail=100000):
    clips = synth.make_clips(n_clips, seed=77)
    return np.concatenate([clips.reshape(-1), np.zeros(tail, dtype=np.float32) + 0.01]).astype(np.float32)


def test_stream_matches_oracle_clip_by_clip():
    from bsed_b200.data import config as cfg
    from bsed_b200.pseudo_labeling import pseudo_label_stream
    audio = _stream()
    oc, op = oracle_models(seed=12, linear_std=0.1)
    m, p = bsed_models(oc, op)
    out = pseudo_label_stream(audio, m, p, batch_clips=2, rank=0, world=1)
    assert out["n_clips"] == 5 and out["span"] == (0, 5)          # the 100000-sample tail is dropped
    # oracle, one clip at a time
    x = np.stack([ofe.logmel(audio[i * SEG:(i + 1) * SEG]) for i in range(5)])[:, None]
    with torch.no_grad():
        s_or, w_or = op(oc(torch.from_numpy(x))[0])
    # our probabilities for the same clips (same kernels the driver ran)
    from bsed_b200 import engine
    clips = torch.from_numpy(audio[:5 * SEG].reshape(5, SEG)).cuda()
    m.eval(); p.eval()
    with torch.no_grad():
        xg = engine.amp_to_db(engine.melspec(clips), 1255)[:, None]
        s, w = p(m(xg)[0])
    assert max_abs(s.cpu().numpy(), s_or.numpy()) < 1e-3 and max_abs(w.cpu().numpy(), w_or.numpy()) < 1e-3
    # weak rows: identical wherever the oracle is not within tolerance of the threshold
    got_weak = dict(out["weak_rows"])
    for i in range(5):
        name = "stream_%05d" % i
        got = set(got_weak.get(name, "").split(",")) - {""}
        for c, lab in enumerate(cfg.bird_list):
            wv = float(w_or[i, c])
            if abs(wv - 0.5) > 1e-3:
                assert (lab in got) == (wv >= 0.5), (name, lab, wv)
    # strong events: bit-exact given our probabilities
    scale = 4 / (32000 / 255)
    want = []
    for i in range(5):
        for c, on, off in opp.events_from_strong(s[i].cpu().numpy(), 0.5, 14):
            want.append(("stream_%05d" % i, cfg.bird_list[c], min(on * scale, 10.0), min(off * scale, 10.0)))
    assert len(want) > 0
    assert [(a, b) for a, b, _, _ in out["events"]] == [(a, b) for a, b, _, _ in want]
    for g, wv in zip(out["events"], want):
        assert abs(g[2] - wv[2]) < 1e-9 and abs(g[3] - wv[3]) < 1e-9


def test_rank_shards_concatenate_to_the_single_rank_result():
    from bsed_b200.pseudo_labeling import pseudo_label_stream
    audio = _stream(n_clips=5, tail=0)
    oc, op = oracle_models(seed=12, linear_std=0.1)
    m, p = bsed_models(oc, op)
    whole = pseudo_label_stream(audio, m, p, batch_clips=3, rank=0, world=1)
    parts = [pseudo_label_stream(audio, m, p, batch_clips=3, rank=r, world=2, gather=False) for r in range(2)]
    assert [q["span"] for q in parts] == [(0, 3), (3, 5)]
    assert parts[0]["weak_rows"] + parts[1]["weak_rows"] == whole["weak_rows"]
    assert parts[0]["events"] + parts[1]["events"] == whole["events"]


def test_empty_and_short_streams():
    from bsed_b200.pseudo_labeling import pseudo_label_stream
    oc, op = oracle_models(seed=12, linear_std=0.1)
    m, p = bsed_models(oc, op)
    out = pseudo_label_stream(np.zeros(1000, dtype=np.float32), m, p, rank=0, world=1)
    assert out["n_clips"] == 0 and out["weak_rows"] == [] and out["events"] == []
    out = pseudo_label_stream(np.zeros(SEG, dtype=np.float32), m, p, rank=1, world=2, gather=False)
    assert out["span"] == (1, 1) and out["events"] == []
