"""oracle/crnn.py and oracle/train.py against the fixtures produced by the reference's own modules
(tests/make_golden.py), and against the live reference when /root/reference is mounted."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import has_reference
from helpers import golden, max_abs, oracle_models, rel_l2
from oracle import crnn as ocrnn
from oracle import train as otrain
from bsed_b200.utilities import synth


def test_state_dict_keys_match_reference():
    g = golden("state_dict_keys.npz")
    oc, op = oracle_models()
    assert list(oc.state_dict().keys()) == [str(k) for k in g["keys"]]
    assert [str(tuple(v.shape)) for v in oc.state_dict().values()] == [str(s) for s in g["shapes"]]
    assert list(op.state_dict().keys()) == [str(k) for k in g["pred_keys"]]
    n = sum(p.numel() for p in oc.parameters())
    assert n == 1107280 and sum(p.numel() for p in op.parameters()) == 10280


def test_eval_forward_matches_reference_fixture():
    g = golden("crnn_eval.npz")
    x = torch.from_numpy(synth.make_logmel_like(2, seed=11))
    assert float(x.double().sum()) == pytest.approx(float(g["x_sum"]), rel=1e-12)
    oc, op = oracle_models(seed=5, linear_std=0.2)
    with torch.no_grad():
        enc, d_in = oc(x)
        strong, weak = op(enc)
    assert enc.shape == (2, 313, 256) and strong.shape == (2, 313, 20) and weak.shape == (2, 20)
    assert max_abs(enc.numpy()[:, ::8], g["enc"]) < 2e-5
    assert max_abs(strong.numpy(), g["strong"]) < 1e-5
    assert max_abs(weak.numpy(), g["weak"]) < 1e-5
    assert strong.numpy().std() > 0.05     # the fixture is sensitive (probabilities are spread)


def test_train_forward_with_hash_dropout_matches_reference_fixture():
    g = golden("crnn_train_fwd.npz")
    x = torch.from_numpy(synth.make_logmel_like(2, seed=11))
    oc, op = oracle_models(seed=5, linear_std=0.2, dropout=0.5, train=True)
    oc.set_dropout_keys(2023, 3, 0)
    with torch.no_grad():
        enc, _ = oc(x)
        strong, weak = op(enc)
    assert max_abs(strong.numpy(), g["strong"]) < 2e-5
    assert max_abs(weak.numpy(), g["weak"]) < 2e-5
    sd = oc.state_dict()
    assert max_abs(sd["cnn.batchnorm0.running_mean"].numpy(), g["rm0"]) < 1e-5
    assert rel_l2(sd["cnn.batchnorm6.running_var"].numpy(), g["rv6"]) < 1e-5
    assert int(sd["cnn.batchnorm3.num_batches_tracked"]) == int(g["nbt"]) == 1


def _run_oracle_mt(p_drop, n_steps=2):
    oc, op = oracle_models(seed=5, linear_std=0.2, dropout=p_drop, train=True)
    tc, tp = oracle_models(seed=6, linear_std=0.2, dropout=p_drop, train=True)
    for prm in list(tc.parameters()) + list(tp.parameters()):
        prm.detach_()
    xs = torch.from_numpy(synth.make_logmel_like(2, seed=21))
    xr = torch.from_numpy(synth.make_logmel_like(2, seed=22))
    xr_ema = xr + 0.5 * torch.from_numpy(synth.make_logmel_like(2, seed=23)) * 0.1
    ts = torch.from_numpy(synth.make_targets(2, seed=24))
    opt = torch.optim.Adam(list(oc.parameters()) + list(op.parameters()), lr=5e-4, betas=(0.9, 0.999))
    res = []
    for it in range(n_steps):
        gstep = 100 + it

        def hook(tag, gstep=gstep):
            if tag == "teacher":
                tc.set_dropout_keys(2023, gstep, 4)
            elif tag == "syn":
                oc.set_dropout_keys(2023, gstep, 0)
            else:
                oc.set_dropout_keys(2023, gstep, 2)

        res.append(otrain.mt_step(oc, op, tc, tp, opt, xr, xr_ema, xs, ts, gstep, rampup_length=500,
                                  ema_flavour="state_dict", dropout_hook=hook))
    return oc, op, tc, tp, res


@pytest.mark.parametrize("name,p_drop", [("mt_step_nodrop.npz", 0.0), ("mt_step_drop.npz", 0.5)])
def test_mean_teacher_step_matches_reference_fixture(name, p_drop):
    g = golden(name)
    oc, op, tc, tp, res = _run_oracle_mt(p_drop)
    for it in range(2):
        loss, parts, outs = res[it]
        assert float(loss) == pytest.approx(float(g[f"loss{it}"]), rel=2e-5)
        for k in ("weak_class", "strong_class", "cons_strong", "cons_weak"):
            assert float(parts[k]) == pytest.approx(float(g[f"{k}{it}"]), rel=5e-4, abs=1e-7)
    grads = res[0][2]["grads"]
    for k, v in grads.items():
        ref_key = k.replace("crnn.cnn.", "crnn.cnn.cnn.", 1) if k.startswith("crnn.cnn.") else k
        gn = float(g["gn_" + ref_key])
        mine = float(np.sqrt((v.numpy().astype(np.float64) ** 2).sum()))
        if gn > 1e-4:                          # conv biases in front of BatchNorm have ~0 gradient (rounding noise)
            assert mine == pytest.approx(gn, rel=2e-3), k
    ssd, tsd = oc.state_dict(), tc.state_dict()
    for k in ("cnn.conv0.weight", "cnn.glu4.linear.weight", "rnn.rnn.weight_hh_l0", "cnn.batchnorm1.running_var"):
        assert max_abs(ssd[k].numpy().reshape(-1)[:2048], g["s_" + k]) < 2e-4, k
        assert max_abs(tsd[k].numpy().reshape(-1)[:2048], g["t_" + k]) < 2e-5, k
    assert int(tsd["cnn.batchnorm0.num_batches_tracked"]) == int(g["t_nbt"])
    assert int(ssd["cnn.batchnorm0.num_batches_tracked"]) == int(g["s_nbt"]) == 4


def test_ramps_and_alpha():
    assert otrain.exp_rampup(0, 100) == pytest.approx(np.exp(-5.0))
    assert otrain.exp_rampup(100, 100) == 1.0 and otrain.exp_rampup(1000, 100) == 1.0
    assert otrain.exp_rampup(5, 0) == 1.0
    assert otrain.sigmoid_rampdown(50, 100) == pytest.approx(np.exp(-12.5 * 0.25))
    assert otrain.ema_alpha(0.999, 1) == 0.5 and otrain.ema_alpha(0.999, 10 ** 6) == 0.999


def test_hash_dropout_rate_and_determinism():
    idx = np.arange(200000)
    k = ocrnn.mix_key(2023, 7, 3)
    keep = ocrnn.keep_mask(idx, k, 0.5)
    assert abs(keep.mean() - 0.5) < 0.01
    assert np.array_equal(keep, ocrnn.keep_mask(idx, k, 0.5))
    assert ocrnn.keep_mask(idx, k, 0.0).all()
    assert abs(ocrnn.keep_mask(idx, ocrnn.mix_key(2023, 8, 3), 0.25).mean() - 0.75) < 0.01


@pytest.mark.skipif(not has_reference(), reason="/root/reference not mounted (GPU box)")
def test_oracle_equals_live_reference():
    sys.path.insert(0, "/root/reference/src")
    from models.CRNN import CRNN, Predictor
    oc, op = oracle_models(seed=9, linear_std=0.1)
    rc = CRNN(**{**ocrnn.CRNN_KWARGS, "dropout": 0.5})
    rp = Predictor(**ocrnn.PREDICTOR_KWARGS)
    sd = oc.state_dict()
    rc.cnn.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("cnn.")})
    rc.rnn.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("rnn.")})
    rp.load_state_dict(op.state_dict())
    rc.eval(); rp.eval()
    x = torch.from_numpy(synth.make_logmel_like(1, seed=3))
    with torch.no_grad():
        e1, _ = oc(x)
        e2, _ = rc(x)
        s1, w1 = op(e1)
        s2, w2 = rp(e2)
    assert torch.equal(e1, e2) and torch.equal(s1, s2) and torch.equal(w1, w2)
