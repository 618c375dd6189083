"""Host-side rows of SURVEY 8a that need no GPU: weights_init (a13, src/utilities/utils.py:40-63) and
adjust_learning_rate (a14, src/main.py:51-83 and the epoch-100 halving of src/main_baseline.py:72-73)."""
import math

import numpy as np
import pytest
import torch
from torch import nn

from bsed_b200 import engine
from bsed_b200.data import config as cfg
from bsed_b200.main import adjust_learning_rate
from bsed_b200.models import CRNN, CRNN_fpn, Predictor
from bsed_b200.utilities.utils import weights_init


# ---------------------------------------------------------------------------------------------
# a14: learning-rate schedule
# ---------------------------------------------------------------------------------------------
def _opt(lr=1.0):
    return torch.optim.SGD([nn.Parameter(torch.zeros(1))], lr=lr)


def _reference_lr(rampup, rampdown, c_epoch, baseline):
    """src/main.py:66 and src/main_baseline.py:66,72-73 restated"""
    lr = rampup * rampdown * cfg.max_learning_rate
    if baseline and c_epoch > 100:
        lr = lr * (0.5 ** (1 + ((c_epoch - 100) // 20)))
    return lr


@pytest.mark.parametrize("baseline", [False, True])
@pytest.mark.parametrize("c_epoch", [0, 50, 100, 101, 119, 120, 121, 140, 141, 299])
def test_adjust_learning_rate_matches_reference_formula(baseline, c_epoch):
    o, od, oc = _opt(), _opt(), _opt()
    for rampup, rampdown in ((0.0067, 1.0), (0.37, 1.0), (1.0, 1.0), (1.0, 0.5)):
        adjust_learning_rate(o, rampup, rampdown, optimizer_d=od, optimizer_crnn=oc, c_epoch=c_epoch, step_decay=baseline)
        want = _reference_lr(rampup, rampdown, c_epoch, baseline)
        assert o.param_groups[0]["lr"] == want
        assert od.param_groups[0]["lr"] == want * 0.1 and oc.param_groups[0]["lr"] == want * 0.1   # :78-84


def test_step_decay_values():
    o = _opt()
    lrs = {}
    for e in (100, 101, 120, 121, 141, 161):
        adjust_learning_rate(o, 1.0, c_epoch=e, step_decay=True)
        lrs[e] = o.param_groups[0]["lr"] / cfg.max_learning_rate
    assert lrs == {100: 1.0, 101: 0.5, 120: 0.25, 121: 0.25, 141: 0.125, 161: 0.0625}
    with pytest.raises(TypeError):
        adjust_learning_rate(o, 1.0, step_decay=True)          # the reference compares c_epoch (None > 100 raises there too)
    adjust_learning_rate(o, 0.5)                               # src/main.py form: c_epoch is not needed
    assert o.param_groups[0]["lr"] == 0.5 * cfg.max_learning_rate


# ---------------------------------------------------------------------------------------------
# a13: weights_init
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cls", [CRNN, CRNN_fpn])
def test_weights_init_distributions(cls):
    torch.manual_seed(7)
    m = cls(**engine.REFERENCE_CRNN_KWARGS)
    p = Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
    gru_bias_before = {n: v.detach().clone() for n, v in m.named_parameters() if ".bias_ih" in n or ".bias_hh" in n}
    weights_init(m)
    weights_init(p)
    n_conv = n_bn = n_gru = n_lin = 0
    for name, prm in list(m.named_parameters()) + [("pred." + n, v) for n, v in p.named_parameters()]:
        w = prm.detach()
        if w.dim() == 4:                                         # Conv2d: xavier_uniform_(gain sqrt 2), bias 0
            cout, cin, kh, kw = w.shape
            a = math.sqrt(2.0) * math.sqrt(6.0 / (cin * kh * kw + cout * kh * kw))
            assert float(w.abs().max()) <= a * (1 + 1e-6), name
            if w.numel() >= 4096:
                assert float(w.std()) == pytest.approx(a / math.sqrt(3.0), rel=0.05), name
                assert abs(float(w.mean())) < 0.05 * a, name
            bias = dict(m.named_parameters())[name[:-len("weight")] + "bias"].detach()
            assert float(bias.abs().max()) == 0.0, name
            n_conv += 1
        elif name.endswith(".bias") and ("conv" in name or "cnn_fcn" in name):   # Conv2d bias: checked with its weight
            assert float(w.abs().max()) == 0.0, name
        elif "weight_ih" in name or "weight_hh" in name:         # GRU: orthogonal_ on every matrix (384 x In)
            g = (w.T.double() @ w.double()).numpy()              # tall matrix: orthonormal columns
            assert np.abs(g - np.eye(g.shape[0])).max() < 1e-4, name
            n_gru += 1
        elif ".bias_ih" in name or ".bias_hh" in name:           # GRU biases: untouched (len(size) == 1)
            assert torch.equal(w, gru_bias_before[name]), name
        elif name.endswith("linear.weight") or name.startswith("pred.") and name.endswith("weight"):   # Linear: N(0, .01)
            assert float(w.std()) == pytest.approx(0.01, rel=0.1 if w.numel() < 2000 else 0.05), name
            assert abs(float(w.mean())) < 0.002, name
            n_lin += 1
        elif name.endswith("linear.bias") or name.startswith("pred.") and name.endswith("bias"):
            assert float(w.abs().max()) == 0.0, name
        elif "batchnorm" in name or "bn_fcn" in name:
            if name.endswith("weight"):                          # BatchNorm: N(1, .02), bias 0
                assert float(w.mean()) == pytest.approx(1.0, abs=0.02) and 0.005 < float(w.std()) < 0.04, name
                n_bn += 1
            else:
                assert float(w.abs().max()) == 0.0, name
        else:
            raise AssertionError(f"parameter {name} not classified")
    fpn = cls is CRNN_fpn
    assert n_conv == (7 + (4 if fpn else 0)) and n_bn == (8 if fpn else 7)
    assert n_gru == (24 if fpn else 8) and n_lin == (8 if fpn else 7) + 2


def test_weights_init_keeps_flat_views():
    """The modules' parameters are views into one flat buffer the kernels read: initialisation must write through."""
    m = CRNN(**engine.REFERENCE_CRNN_KWARGS)
    weights_init(m)
    flat = m.flat_tensors()[0]
    o = 0
    for _, prm in m.named_parameters():
        assert prm.data_ptr() == flat.data_ptr() + 4 * o
        assert torch.equal(prm.detach().reshape(-1), flat[o:o + prm.numel()])
        o += prm.numel()
    assert o == flat.numel()


def test_resnet_wide_gemm_chunks_and_implicit_predicate():
    """Host-side planning of the ResNet tagger's tensor-core launches (models/ResNet.py): a wide GEMM covers multiples of
    128 columns up to 1024 per launch plus a remainder below 128; the implicit-GEMM convolution takes the 3x3 / stride-1 /
    padding-1 units with 64 or 128 output channels whose width divides 128."""
    import types
    from bsed_b200.models.ResNet import Net_resnet
    ch = Net_resnet._wide_chunks
    assert ch(64) == [(0, 64)] and ch(128) == [(0, 128)] and ch(512) == [(0, 512)]
    assert ch(576) == [(0, 512), (512, 576)] and ch(1152) == [(0, 1024), (1024, 1152)]
    assert ch(4608) == [(0, 1024), (1024, 2048), (2048, 3072), (3072, 4096), (4096, 4608)]
    for n in (16, 48, 64, 576, 1152, 2304, 4608):
        c = ch(n)
        assert c[0][0] == 0 and c[-1][1] == n and all(a[1] == b[0] for a, b in zip(c, c[1:]))
        assert all((n1 - n0) % 128 == 0 and n1 - n0 <= 1024 for n0, n1 in c[:-1]) and c[-1][1] - c[-1][0] <= 1024
    unit = lambda k, s, p, cin, cout: types.SimpleNamespace(k=k, stride=s, pad=p, cin=cin, cout=cout)
    h = lambda F: types.SimpleNamespace(shape=(2, 10, F, 64))
    imp = Net_resnet._implicit
    assert imp(unit(3, 1, 1, 64, 64), h(32)) and imp(unit(3, 1, 1, 128, 128), h(16))
    assert not imp(unit(3, 2, 1, 64, 128), h(32))          # stride 2
    assert not imp(unit(1, 1, 0, 64, 128), h(32))          # 1x1 downsample
    assert not imp(unit(3, 1, 1, 256, 256), h(8))          # 256 output channels: im2col + wide GEMM
    assert not imp(unit(7, 2, 3, 1, 64), h(128))           # stem
    assert not imp(unit(3, 1, 1, 64, 64), h(24))           # width does not divide 128


def test_bench_roofline_traffic_comes_from_the_committed_capture():
    """bench.py: `roofline.traffic` is the mean DRAM bytes per conv launch of the committed ncu capture, and null when the
    live launch count per step or the precision does not match the capture."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rec = bench._conv_traffic_from_profile(18.0, "tf32x3")
    assert rec["traffic"] is not None and 2e7 < rec["traffic"] < 7e7          # 39.5 MB per launch in r02z
    assert bench._conv_traffic_from_profile(17.0, "tf32x3")["traffic"] is None
    assert bench._conv_traffic_from_profile(18.0, "tf32")["traffic"] is None
