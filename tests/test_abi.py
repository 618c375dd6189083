"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/bsed.h
declares, fails loudly (not silently) without a GPU, and the host-side mirrors of the reference's
Python interface behave like it."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from helpers import golden, oracle_models
from oracle import postproc as opp
from oracle import train as otrain


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "bsed.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bsed_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from bsed_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libbsed.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 30
    for s in declared:
        assert hasattr(lib, s), f"{s} declared in include/bsed.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes binding list and header disagree"
    assert _lib.load().bsed_version() == 4


def test_no_torch_types_in_signatures():
    txt = open(os.path.join(ROOT, "include", "bsed.h")).read()
    code = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code and "#include <cuda" not in code


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without a GPU")
def test_fails_loudly_without_gpu():
    from bsed_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.bsed_create(0, ctypes.byref(h))
    assert rc == -2                                   # BSED_E_CUDA
    assert b"no CPU fallback" in lib.bsed_last_error()
    from bsed_b200.data.preprocess import preprocess
    with pytest.raises(RuntimeError):
        preprocess(np.zeros(32000, dtype=np.float32))
    from bsed_b200 import engine
    from bsed_b200.models import CRNN
    m = CRNN(**engine.REFERENCE_CRNN_KWARGS)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 1255, 128))


def test_frame_count_entry_point():
    from bsed_b200 import _lib
    lib = _lib.load()
    assert lib.bsed_frontend_n_frames(320000) == 1255
    assert lib.bsed_frontend_n_frames(254) == 1 and lib.bsed_frontend_n_frames(255) == 2


def test_module_state_dict_layout_matches_reference():
    from bsed_b200 import engine
    from bsed_b200.models import CRNN, Predictor
    g = golden("state_dict_keys.npz")
    m = CRNN(**engine.REFERENCE_CRNN_KWARGS)
    p = Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
    assert list(m.state_dict().keys()) == [str(k) for k in g["keys"]]
    assert [str(tuple(v.shape)) for v in m.state_dict().values()] == [str(s) for s in g["shapes"]]
    assert list(p.state_dict().keys()) == [str(k) for k in g["pred_keys"]]
    assert m._flat.numel() == 1107280 and p._flat.numel() == 10280
    # parameters are views of the flat buffer, also after load_state_dict
    oc, op = oracle_models(seed=3)
    m.load_state_dict(oc.state_dict())
    p.load_state_dict(op.state_dict())
    assert m._flat_ok() and p._flat_ok()
    assert torch.equal(m.cnn.conv3.weight, oc.state_dict()["cnn.conv3.weight"])
    assert torch.equal(m._flat[:144].view(16, 1, 3, 3), oc.state_dict()["cnn.conv0.weight"])
    assert torch.equal(m.rnn.rnn.bias_hh_l1_reverse, oc.state_dict()["rnn.rnn.bias_hh_l1_reverse"])
    assert int(m.cnn.batchnorm0.num_batches_tracked) == 0
    sd = m.state_dict()
    m2 = CRNN(**engine.REFERENCE_CRNN_KWARGS)
    m2.load_state_dict(sd)
    assert torch.equal(m2._flat, m._flat) and torch.equal(m2._flat_bn, m._flat_bn)


def test_unsupported_configs_are_rejected():
    from bsed_b200 import engine
    from bsed_b200.models import CRNN
    kw = dict(engine.REFERENCE_CRNN_KWARGS)
    with pytest.raises(NotImplementedError):
        CRNN(**{**kw, "activation": "relu"})
    with pytest.raises(NotImplementedError):
        CRNN(**{**kw, "n_RNN_cell": 64})


def test_ramps_match_oracle():
    from bsed_b200.utilities import ramps
    for cur, ln in [(0, 100), (3, 100), (50, 100), (100, 100), (250, 100), (7, 0)]:
        assert ramps.exp_rampup(cur, ln) == pytest.approx(otrain.exp_rampup(cur, ln), rel=1e-15)
        assert ramps.sigmoid_rampdown(cur, ln) == pytest.approx(otrain.sigmoid_rampdown(cur, ln), rel=1e-15)


def test_many_hot_encoder_matches_oracle():
    from bsed_b200.data import config as cfg
    from bsed_b200.utilities.ManyHotEncoder import ManyHotEncoder
    enc = ManyHotEncoder(cfg.bird_list, n_frames=313)
    rows = [(0.5, 1.25, "WOTH"), (3.0, 9.99, "BAWW"), (0.0, 0.02, "EATO")]
    y = enc.encode_strong_df(rows)
    ref = opp.encode_strong([(a, b, cfg.bird_list.index(c)) for a, b, c in rows])
    assert np.array_equal(y, ref)
    import pandas as pd
    df = pd.DataFrame(rows, columns=["onset", "offset", "event_label"])
    assert np.array_equal(enc.encode_strong_df(df), ref)
    assert enc.encode_weak(["WOTH,BAWW"]).sum() == 2
    m = (np.random.default_rng(0).random((313, 20)) < 0.3).astype(int)
    dec = enc.decode_strong(m)
    assert [(cfg.bird_list.index(l), a, b) for l, a, b in dec] == opp.decode_strong(m)


def test_scaler():
    from bsed_b200.utilities.Scaler import Scaler
    rng = np.random.default_rng(1)
    data = [(rng.standard_normal((50, 128)).astype(np.float32) * 3 + 1, None) for _ in range(4)]
    sc = Scaler()
    mean, std = sc.calculate_scaler(data)
    allx = np.stack([d[0] for d in data]).astype(np.float64)
    assert np.allclose(mean, allx.mean(axis=(0, 1)))
    assert np.allclose(std, np.sqrt((allx ** 2).mean(axis=(0, 1)) - mean ** 2))


def test_synthetic_workload_is_deterministic():
    from bsed_b200.utilities import synth
    a = synth.make_clips(3, seed=7, n_samples=16000, edge_cases=True)
    b = synth.make_clips(3, seed=7, n_samples=16000, edge_cases=True)
    assert np.array_equal(a, b) and (a[0] == 0).all() and a[1].sum() == 1.0
    t = synth.make_targets(4, seed=3)
    assert t.shape == (4, 313, 20) and set(np.unique(t)) <= {0.0, 1.0} and t.sum() > 0


def test_get_transforms_chain():
    from bsed_b200.data import Transforms as T
    c = T.get_transforms(1255, noise_dict_params={"mean": 0., "snr": 30})
    assert [type(t).__name__ for t in c.transforms] == ["AugmentGaussianNoise", "ApplyLog", "PadOrTrunc", "ToTensor"]
    with pytest.raises(NotImplementedError):
        T.Compose([T.PadOrTrunc(10), T.ApplyLog(), T.ToTensor()])
