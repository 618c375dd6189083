"""oracle/crnn.py:OracleCRNNfpn against the fixtures produced by the reference's own CRNN_fpn / CNN_FPN modules
(tests/make_golden_fpn.py), and against the live reference when /root/reference is mounted."""
import sys

import numpy as np
import pytest
import torch

from conftest import has_reference
from helpers import golden, max_abs, oracle_fpn_models, rel_l2
from oracle import crnn as ocrnn
from oracle import train as otrain
from bsed_b200.utilities import synth


def test_fpn_state_dict_keys_match_reference():
    g = golden("fpn_state_dict_keys.npz")
    oc, _ = oracle_fpn_models()
    assert list(oc.state_dict().keys()) == [str(k) for k in g["keys"]]
    assert [str(tuple(v.shape)) for v in oc.state_dict().values()] == [str(s) for s in g["shapes"]]
    assert [n for n, _ in oc.named_parameters()] == [str(k) for k in g["param_keys"]]
    assert sum(p.numel() for p in oc.parameters()) == int(g["n_params"]) == 2556368


def test_fpn_eval_forward_matches_reference_fixture():
    g = golden("fpn_eval.npz")
    x = torch.from_numpy(synth.make_logmel_like(2, seed=11))
    assert float(x.double().sum()) == pytest.approx(float(g["x_sum"]), rel=1e-12)
    oc, op = oracle_fpn_models(seed=5, linear_std=0.2)
    with torch.no_grad():
        enc, d_in = oc(x)
        strong, weak = op(enc)
    assert enc.shape == (2, 313, 256) and d_in is enc
    assert max_abs(enc.numpy()[:, ::8], g["enc"]) < 2e-5
    assert max_abs(strong.numpy(), g["strong"]) < 1e-5
    assert max_abs(weak.numpy(), g["weak"]) < 1e-5
    assert strong.numpy().std() > 0.05


def test_fpn_train_forward_backward_matches_reference_fixture():
    g = golden("fpn_train.npz")
    x = torch.from_numpy(synth.make_logmel_like(2, seed=11))
    oc, op = oracle_fpn_models(seed=5, linear_std=0.2, dropout=0.5, train=True)
    oc.set_dropout_keys(2023, 3, 0)
    enc, _ = oc(x)
    strong, weak = op(enc)
    w = torch.from_numpy(np.random.default_rng(7).standard_normal(strong.shape).astype(np.float32))
    ((strong * w).mean() + weak.mean()).backward()
    assert max_abs(strong.detach().numpy(), g["strong"]) < 2e-5
    assert max_abs(weak.detach().numpy(), g["weak"]) < 2e-5
    # the shared BatchNorm is applied twice per forward (src/models/CNN_FPN.py:86-96)
    assert int(oc.cnn.bn_fcn.num_batches_tracked) == int(g["nbt_fcn"]) == 2
    assert int(oc.cnn.cnn.batchnorm0.num_batches_tracked) == int(g["nbt0"]) == 1
    assert max_abs(oc.cnn.bn_fcn.running_mean.numpy(), g["rm_fcn"]) < 1e-5
    assert rel_l2(oc.cnn.bn_fcn.running_var.numpy(), g["rv_fcn"]) < 1e-5
    for n, p in oc.named_parameters():
        if n.startswith("cnn.conv1x1."):
            assert p.grad is None and ("none_" + n) in g.files      # registered but unused by the reference forward
            continue
        gn = float(g["gn_" + n])
        if gn > 1e-5:
            got = p.grad.numpy().reshape(-1)
            got = got if got.size <= 4096 else got[:: max(1, got.size // 4096)][:4096]
            assert rel_l2(got, g["g_" + n]) < 2e-3, n


def test_fpn_mean_teacher_step_matches_reference_fixture():
    g = golden("fpn_mt_step_drop.npz")
    oc, op = oracle_fpn_models(seed=5, linear_std=0.2, dropout=0.5, train=True)
    tc, tp = oracle_fpn_models(seed=6, linear_std=0.2, dropout=0.5, train=True)
    for prm in list(tc.parameters()) + list(tp.parameters()):
        prm.detach_()
    xs = torch.from_numpy(synth.make_logmel_like(2, seed=21))
    xr = torch.from_numpy(synth.make_logmel_like(2, seed=22))
    xr_ema = xr + 0.5 * torch.from_numpy(synth.make_logmel_like(2, seed=23)) * 0.1
    ts = torch.from_numpy(synth.make_targets(2, seed=24))
    opt = torch.optim.Adam(list(oc.parameters()) + list(op.parameters()), lr=5e-4, betas=(0.9, 0.999))
    for it in range(2):
        gstep = 100 + it

        def hook(tag, gstep=gstep):
            if tag == "teacher":
                tc.set_dropout_keys(2023, gstep, 4)
            elif tag == "syn":
                oc.set_dropout_keys(2023, gstep, 0)
            else:
                oc.set_dropout_keys(2023, gstep, 2)

        loss, parts, outs = otrain.mt_step(oc, op, tc, tp, opt, xr, xr_ema, xs, ts, gstep, rampup_length=500,
                                           ema_flavour="state_dict", dropout_hook=hook)
        assert float(loss) == pytest.approx(float(g[f"loss{it}"]), rel=2e-5)
        for k in ("weak_class", "strong_class", "cons_strong", "cons_weak"):
            assert float(parts[k]) == pytest.approx(float(g[f"{k}{it}"]), rel=5e-4, abs=1e-7)
    ssd, tsd = oc.state_dict(), tc.state_dict()
    for k in ("cnn.cnn_fcn.weight", "rnn_2.rnn.weight_ih_l1_reverse", "conv1x1_2.weight", "cnn.bn_fcn.running_var"):
        assert max_abs(ssd[k].numpy().reshape(-1)[:2048], g["s_" + k]) < 2e-4, k
        assert max_abs(tsd[k].numpy().reshape(-1)[:2048], g["t_" + k]) < 2e-5, k
    assert int(tsd["cnn.bn_fcn.num_batches_tracked"]) == int(g["t_nbt_fcn"]) == 4
    assert int(ssd["cnn.bn_fcn.num_batches_tracked"]) == int(g["s_nbt_fcn"]) == 8


@pytest.mark.skipif(not has_reference(), reason="/root/reference not mounted (GPU box)")
def test_fpn_oracle_equals_live_reference():
    sys.path.insert(0, "/root/reference/src")
    from models.CRNN import CRNN_fpn
    oc, _ = oracle_fpn_models(seed=9, linear_std=0.1)
    rc = CRNN_fpn(**{**ocrnn.CRNN_KWARGS, "dropout": 0.5})
    rc.load_state_dict(oc.state_dict(), strict=True)
    rc.eval()
    x = torch.from_numpy(synth.make_logmel_like(1, seed=3))
    with torch.no_grad():
        e1, _ = oc(x)
        e2, _ = rc(x)
    assert torch.equal(e1, e2)
