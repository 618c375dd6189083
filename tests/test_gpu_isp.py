"""Shift-consistency (ISP / SCT) step, src/main_baseline.py:229-277,337-584: the fused CUDA trainer against the CPU oracle
restatement (oracle/train.py:isp_step) on the same seeded clips, weights, shifts and dropout masks."""
import numpy as np
import pytest
import torch

from helpers import bsed_fpn_models, bsed_models, max_abs, oracle_fpn_models, oracle_models, rel_l2
from oracle import train as otrain
from bsed_b200.utilities import synth

pytestmark = pytest.mark.gpu


def test_roll_clips_matches_torch_roll():
    from bsed_b200 import engine
    x = torch.randn(5, 1, 1255, 128, device="cuda")
    st = [256, -128, 0, 4, -252]
    sf = [3, -4, 0, 1, -1]
    got_t = engine.roll_clips(x, torch.tensor(st, dtype=torch.int32, device="cuda"), None)
    got_f = engine.roll_clips(x, None, torch.tensor(sf, dtype=torch.int32, device="cuda"))
    for k in range(5):
        assert torch.equal(got_t[k], torch.roll(x[k], st[k], dims=1))     # bit-exact data movement
        assert torch.equal(got_f[k], torch.roll(x[k], sf[k], dims=2))


def test_loss_terms_match_torch():
    from bsed_b200 import engine
    from bsed_b200._lib import LOSS_BCE_STRONG, LOSS_BCE_WEAK, LOSS_MSE_STRONG, LOSS_MSE_WEAK
    g = torch.Generator().manual_seed(3)
    strong = torch.rand(6, 313, 20, generator=g).clamp(1e-4, 1 - 1e-4)
    weak = torch.rand(6, 20, generator=g).clamp(1e-4, 1 - 1e-4)
    tgt = (torch.rand(2, 313, 20, generator=g) > 0.8).float()
    other = torch.rand(2, 313, 20, generator=g)
    wt = (torch.rand(2, 20, generator=g) > 0.5).float()
    roll = [5, -7]
    s, w = strong.clone().requires_grad_(), weak.clone().requires_grad_()
    bce, mse = torch.nn.BCELoss(), torch.nn.MSELoss()
    rolled_t = torch.stack([torch.roll(tgt[k], roll[k], 0) for k in range(2)])
    rolled_o = torch.stack([torch.roll(s[k].detach(), roll[k], 0) for k in range(2)])
    l0 = bce(s[2:4], rolled_t)
    l1 = bce(w[0:2], tgt.max(-2)[0]) + bce(w[4:6], wt)
    l2 = 0.3 * mse(s[4:6], other)
    l3 = 0.7 * mse(w[2:4], wt)
    l4 = 0.25 * mse(s[2:4], rolled_o)
    (l0 + l1 + l2 + l4 + 0.5 * l3).backward()
    dev = "cuda"
    sd, wd = strong.to(dev), weak.to(dev)
    r = torch.tensor(roll, dtype=torch.int32, device=dev)
    terms = [dict(kind=LOSS_BCE_STRONG, pred_first=2, n=2, ref=tgt.to(dev), roll=r, slot=0),
             dict(kind=LOSS_BCE_WEAK, pred_first=0, n=2, ref=tgt.to(dev), ref_is_strong=True, slot=1),
             dict(kind=LOSS_BCE_WEAK, pred_first=4, n=2, ref=wt.to(dev), slot=1),
             dict(kind=LOSS_MSE_STRONG, pred_first=4, n=2, ref=other.to(dev), weight=0.3, slot=2),
             dict(kind=LOSS_MSE_WEAK, pred_first=2, n=2, ref=wt.to(dev), weight=0.7, grad_weight=0.35, slot=3),
             dict(kind=LOSS_MSE_STRONG, pred_first=2, n=2, ref=sd[0:2], roll=r, weight=0.25, slot=4)]
    losses, ds, dw = engine.loss_terms(sd, wd, terms, 5)
    ref = [float(v) for v in (l0, l1, l2, l3, l4)]
    assert np.allclose(losses.cpu().numpy(), ref, rtol=2e-5, atol=1e-7), (losses.cpu().numpy(), ref)
    assert rel_l2(ds.cpu().numpy(), s.grad.numpy()) < 1e-5
    assert rel_l2(dw.cpu().numpy(), w.grad.numpy()) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "default"])
@pytest.mark.parametrize("p_drop,fpn", [(0.0, False), (0.5, False), (0.5, True)])
def test_isp_step_matches_oracle(p_drop, fpn, precision, monkeypatch):
    """fpn: the same step on CRNN_fpn, the model the reference's author trains with -ISP.  precision "default" = the
    library default (3xTF32 tensor-core path), "fp32" = the CUDA-core cross-check; same tolerances."""
    from bsed_b200.main import ISP_SLOTS, ShiftConsistencyTrainer
    if precision == "default":
        monkeypatch.delenv("BSED_PRECISION", raising=False)
    else:
        monkeypatch.setenv("BSED_PRECISION", precision)
    n = 2
    omk, bmk = (oracle_fpn_models, bsed_fpn_models) if fpn else (oracle_models, bsed_models)
    oc, op = omk(seed=5, linear_std=0.2, dropout=p_drop, train=True)
    tc, tp = omk(seed=6, linear_std=0.2, dropout=p_drop, train=True)
    m, p = bmk(oc, op, dropout=p_drop)
    em, ep = bmk(tc, tp, dropout=p_drop)
    for mod in (m, p, em, ep):
        mod.train()
    for prm in list(tc.parameters()) + list(tp.parameters()) + list(em.parameters()) + list(ep.parameters()):
        prm.detach_()
    xs = torch.from_numpy(synth.make_logmel_like(n, seed=21))
    xr = torch.from_numpy(synth.make_logmel_like(n, seed=22))
    xr_ema = xr + 0.5 * torch.from_numpy(synth.make_logmel_like(n, seed=23)) * 0.1
    ts = torch.from_numpy(synth.make_targets(n, seed=24))
    target_weak = (torch.from_numpy(synth.make_targets(n, seed=25)).max(-2)[0] > 0).float()
    shifts, fshifts = [-36 * 4, 17 * 4], [3, -2]
    opt = torch.optim.Adam(list(oc.parameters()) + list(op.parameters()), lr=5e-4, betas=(0.9, 0.999))
    gstep, rampup = 40, 0.37

    # device batch offsets of the nine model calls (plan A: syn 0, real n, teacher 2n; plan B: 0, n, 2n, 3n; plan C: 0, n)
    keymap = {"syn": (0, 0), "real": (0, n), "teacher": (0, 2 * n), "real_shift": (1, 0), "real_fshift": (1, n),
              "syn_shift": (1, 2 * n), "syn_fshift": (1, 3 * n), "teacher_shift": (2, 0), "teacher_fshift": (2, n)}

    def hook(tag):
        sub, off = keymap[tag]
        (tc if tag.startswith("teacher") else oc).set_dropout_keys(2023, 3 * gstep + sub, off)

    loss, parts, outs = otrain.isp_step(oc, op, tc, tp, opt, xr, xr_ema, target_weak, xs, ts, shifts, fshifts, gstep, rampup,
                                        dropout_hook=hook)
    tr = ShiftConsistencyTrainer(m, p, em, ep, lr=5e-4, n=n, dropout_seed=2023)
    losses = tr.step(xr.cuda(), xr_ema.cuda(), target_weak.cuda(), xs.cuda(), ts.cuda(), shifts, fshifts, gstep, rampup)
    got = dict(zip(ISP_SLOTS, [float(v) for v in losses.cpu()]))
    for k in ISP_SLOTS:
        assert got[k] == pytest.approx(float(parts[k]), rel=3e-3, abs=2e-6), (k, got[k], float(parts[k]))
    assert float(ShiftConsistencyTrainer.total(losses)) == pytest.approx(float(loss), rel=1e-3)
    assert max_abs(tr.last["strong"][n:2 * n].cpu().numpy(), outs["strong"].numpy()) < 1e-3
    assert max_abs(tr.last["strong"][2 * n:3 * n].cpu().numpy(), outs["strong_shift"].numpy()) < 1e-3
    assert max_abs(tr.last["strong"][5 * n:].cpu().numpy(), outs["syn_strong_fshift"].numpy()) < 1e-3
    # gradients of the step, tensor by tensor
    o, bad = 0, []
    names = [("crnn." + k, v) for k, v in m.named_parameters()] + [("pred." + k, v) for k, v in p.named_parameters()]
    for (mod, pname, shape), (fullname, _) in zip(m._param_specs + p._param_specs, names):
        k = int(np.prod(shape))
        gg = tr.grads[o:o + k].cpu().numpy()
        o += k
        ref = outs["grads"][fullname].numpy().reshape(-1)
        if np.linalg.norm(ref) < 1e-4:
            continue                      # conv biases ahead of a train-mode BatchNorm: rounding noise
        e = rel_l2(gg, ref)
        if e > 3e-3:
            bad.append((fullname, e))
    assert not bad, bad
    # BatchNorm counters: 6 student / 3 teacher calls
    trunk, key3 = (m.cnn.cnn, "cnn.cnn.batchnorm3.running_mean") if fpn else (m.cnn, "cnn.batchnorm3.running_mean")
    assert int(trunk.batchnorm0.num_batches_tracked) == 6
    assert max_abs(trunk.batchnorm3.running_mean.cpu().numpy(), oc.state_dict()[key3].numpy()) < 1e-3
    if fpn:
        assert int(m.cnn.bn_fcn.num_batches_tracked) == 12 == int(oc.cnn.bn_fcn.num_batches_tracked)
    assert max_abs(em.rnn.rnn.weight_hh_l0.detach().cpu().numpy(), tc.state_dict()["rnn.rnn.weight_hh_l0"].numpy()) < 1e-4


def test_train_mt_isp_entry_point_runs():
    from bsed_b200 import main as bmain
    oc, op = oracle_models(seed=5, linear_std=0.2)
    tc, tp = oracle_models(seed=6, linear_std=0.2)
    m, p = bsed_models(oc, op, dropout=0.5)
    em, ep = bsed_models(tc, tp, dropout=0.5)
    for mod in (m, p, em, ep):
        mod.train()
    for prm in list(em.parameters()) + list(ep.parameters()):
        prm.detach_()
    xs = torch.from_numpy(synth.make_logmel_like(2, seed=21))
    xr = torch.from_numpy(synth.make_logmel_like(2, seed=22))
    ts = torch.from_numpy(synth.make_targets(2, seed=24))
    real = [(((xr, xr), torch.zeros(2, 20)), ["r0", "r1"])] * 2
    syn = [(((xs, xs), ts), ["s0", "s1"])]
    opt = bmain.FusedAdam(list(m.parameters()) + list(p.parameters()), lr=5e-4, betas=(0.9, 0.999))
    before = m._flat.clone()
    loss = bmain.train_mt(real, syn, m, opt, 0, ema_model=em, ema_predictor=ep, predictor=p, ISP=True)
    assert torch.isfinite(loss) and float(loss) > 0
    assert not torch.equal(before, m.flat_tensors()[0])
    assert int(m.cnn.batchnorm0.num_batches_tracked) == 12      # 6 student calls per iteration, 2 iterations
