"""oracle/train.py:isp_step (restatement of the ISP / SCT branch of src/main_baseline.py): properties the step must have
whatever the weights are.  The step assembly is parity-unpinned (the script cannot be executed here); the modules it
drives are pinned by tests/test_oracle_crnn.py."""
import pytest
import torch

from helpers import oracle_models
from oracle import train as otrain
from bsed_b200.utilities import synth


def _setup(n=2):
    oc, op = oracle_models(seed=5, linear_std=0.2, dropout=0.0, train=True)
    tc, tp = oracle_models(seed=6, linear_std=0.2, dropout=0.0, train=True)
    for prm in list(tc.parameters()) + list(tp.parameters()):
        prm.detach_()
    xs = torch.from_numpy(synth.make_logmel_like(n, seed=21))
    xr = torch.from_numpy(synth.make_logmel_like(n, seed=22))
    ts = torch.from_numpy(synth.make_targets(n, seed=24))
    tw = (torch.from_numpy(synth.make_targets(n, seed=25)).max(-2)[0] > 0).float()
    opt = torch.optim.Adam(list(oc.parameters()) + list(op.parameters()), lr=5e-4)
    return oc, op, tc, tp, opt, xr, xs, ts, tw


def test_zero_shift_collapses_the_shift_terms():
    oc, op, tc, tp, opt, xr, xs, ts, tw = _setup()
    loss, parts, outs = otrain.isp_step(oc, op, tc, tp, opt, xr, xr, tw, xs, ts, [0, 0], [0, 0], 10, 0.5)
    # without shifts (and without dropout) the shifted batches equal the plain ones: same BatchNorm batch statistics
    assert float(parts["strong_shift_class"]) == pytest.approx(float(parts["strong_class"]), rel=1e-5)
    assert float(parts["strong_freq_shift_class"]) == pytest.approx(float(parts["strong_class"]), rel=1e-5)
    assert float(parts["cons_shift"]) == pytest.approx(0.0, abs=1e-10)
    assert float(parts["cons_strong_shift"]) == pytest.approx(float(parts["cons_strong"]), rel=1e-4, abs=1e-9)
    total = sum(float(parts[k]) for k in ("strong_class", "weak_class", "cons_strong", "cons_weak", "weak_freq_shift_class",
                                          "strong_shift_class", "strong_freq_shift_class", "cons_shift"))
    total += 0.5 * (float(parts["cons_strong_shift"]) + float(parts["cons_strong_freq_shift"]))
    assert float(loss) == pytest.approx(total, rel=1e-6)
    # six student and three teacher model calls
    assert int(oc.cnn.batchnorm0.num_batches_tracked) == 6 and int(tc.cnn.batchnorm0.num_batches_tracked) >= 3


def test_time_shift_rolls_targets_by_pooled_frames():
    oc, op, tc, tp, opt, xr, xs, ts, tw = _setup()
    loss, parts, outs = otrain.isp_step(oc, op, tc, tp, opt, xr, xr, tw, xs, ts, [8, -12], [1, -1], 10, 0.5)
    assert torch.isfinite(loss) and float(parts["cons_shift"]) > 0
    assert float(parts["strong_shift_class"]) != pytest.approx(float(parts["strong_class"]), rel=1e-6)
