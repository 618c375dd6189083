"""Config 3 (BASELINE.json configs[2]): the fused SCMT + adversarial-domain-adaptation iteration
(`main.AdaptationTrainer`, src/main_scmt_ada_weak_seperate.py:314-335 + :337-521, SGD-Nesterov x 3) against the CPU
oracle (`oracle/train.py:ada_step` over the reference-pinned modules of oracle/crnn.py and oracle/da.py), step level:
domain loss, the four loss terms, probabilities, all three gradient sets, the parameters after the three optimiser steps,
the teacher after the EMA.  Library default precision (3xTF32) for the CRNN, fp32 for the discriminator."""
import re

import numpy as np
import pytest
import torch

from helpers import bsed_models, max_abs, oracle_models, rel_l2
from bsed_b200.utilities import synth
from oracle import da as oda
from oracle import train as otrain

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _library_default(monkeypatch):
    monkeypatch.delenv("BSED_PRECISION", raising=False)


def _setup(n, p_drop):
    oc, op = oracle_models(seed=5, linear_std=0.2, dropout=p_drop, train=True)
    tc, tp = oracle_models(seed=6, linear_std=0.2, dropout=p_drop, train=True)
    for prm in list(tc.parameters()) + list(tp.parameters()):
        prm.detach_()
    od = oda.OracleClipDiscriminator()
    oda.seeded_disc_init(od, 3)
    od.train()
    xs = torch.from_numpy(synth.make_logmel_like(n, seed=51))
    xr = torch.from_numpy(synth.make_logmel_like(n, seed=52))
    xr_ema = xr + 0.05 * torch.from_numpy(synth.make_logmel_like(n, seed=53))
    ts = torch.from_numpy(synth.make_targets(n, seed=54))
    tw = torch.from_numpy(synth.make_targets(n, seed=55)).max(1)[0]          # weak labels of the real batch
    return oc, op, tc, tp, od, xs, xr, xr_ema, ts, tw


@pytest.mark.parametrize("disc_precision", ["fp32", None])          # None = the trainer's default (tf32x3 with the plan)
@pytest.mark.parametrize("p_drop", [0.0, 0.5])
def test_adaptation_step_matches_oracle(p_drop, disc_precision):
    from bsed_b200.DA.cdan_frame import ConditionalDomainAdversarialLoss
    from bsed_b200.main import AdaptationTrainer
    from bsed_b200.models.CRNN import Clip_Discriminator
    n, gstep, ramp, grl_iter = 3, 40, 500, 700
    lr, mom, wd = 5e-4, 0.9, 1e-4
    oc, op, tc, tp, od, xs, xr, xr_ema, ts, tw = _setup(n, p_drop)
    m, p = bsed_models(oc, op, dropout=p_drop)
    em, ep = bsed_models(tc, tp, dropout=p_drop)
    for mod in (m, p, em, ep):
        mod.train()
    d = Clip_Discriminator(256)
    d.load_state_dict(od.state_dict())
    d = d.cuda().train()
    crit = ConditionalDomainAdversarialLoss(d)
    crit.grl.iter_num = grl_iter
    tr = AdaptationTrainer(m, p, em, ep, crit, lr=lr, lr_adv=lr, momentum=mom, weight_decay=wd, n_syn=n, n_real=n,
                           disc_precision=disc_precision,
                           dropout_seed=2023)
    assert tr.plan.precision == "tf32x3" and tr.disc_precision == (disc_precision or "tf32x3")
    p_before = tr.params.clone()
    losses, dom = tr.step(xr.cuda(), xr_ema.cuda(), tw.cuda(), xs.cuda(), ts.cuda(), gstep, ramp)
    torch.cuda.synchronize()
    assert crit.grl.iter_num == grl_iter + 1

    def hook(tag):
        step = 2 * gstep + 1 if tag.startswith("adv") else 2 * gstep
        off = {"adv_syn": 0, "adv_real": n, "syn": 0, "real": n, "teacher": 2 * n}[tag]
        (tc if tag == "teacher" else oc).set_dropout_keys(2023, step, off)

    sgd = dict(lr=lr, momentum=mom, weight_decay=wd, nesterov=True)
    opt = torch.optim.SGD(list(oc.parameters()) + list(op.parameters()), **sgd)
    opt_c = torch.optim.SGD(oc.parameters(), **sgd)
    opt_d = torch.optim.SGD(od.parameters(), **sgd)
    loss_o, parts, outs = otrain.ada_step(oc, op, tc, tp, od, opt, opt_c, opt_d, xr, xr_ema, tw, xs, ts, gstep, ramp, grl_iter,
                                          dropout_hook=hook)
    # ---- losses and probabilities
    assert float(dom) == pytest.approx(float(parts["domain"]), rel=2e-5)
    got = losses.cpu().numpy()
    want = np.array([float(parts[k]) for k in ("strong_class", "weak_class", "cons_strong", "cons_weak")])
    print(f"[ada p={p_drop}] domain {float(dom):.6f} / {float(parts['domain']):.6f}; losses {got} oracle {want}")
    assert np.allclose(got, want, rtol=1e-3, atol=2e-6)
    assert max_abs(tr.last["strong"][n:].cpu().numpy(), outs["strong"].numpy()) < 1e-3
    assert max_abs(tr.last["weak"][n:].cpu().numpy(), outs["weak"].numpy()) < 1e-3
    # ---- gradients: discriminator, encoder (adversarial), encoder + predictor (main)
    names_d = [k for k, _ in d.named_parameters()]
    o, worst_d = 0, 0.0
    for nm, prm in d.named_parameters():
        k = prm.numel()
        ref = outs["d_grads"][nm].numpy().reshape(-1)
        gn = float(np.linalg.norm(ref))
        if not re.match(r"conv_\d\.bias$", nm):                  # a conv bias ahead of a train-mode BatchNorm: zero gradient
            e = rel_l2(tr.grads_d[o:o + k].cpu().numpy(), ref)
            worst_d = max(worst_d, e)
            # five train-mode BatchNorms over a 6-clip batch amplify the ~1e-5 difference of the encoder features the two
            # sides feed the discriminator (measured 2.1e-3 ... 2.8e-3 on conv_1.weight, 2e-3 with identical inputs in
            # tests/test_gpu_da.py)
            assert e < 6e-3, (nm, e, gn)
        o += k
    assert o == tr.grads_d.numel() and len(names_d) == 22
    names = ["crnn." + k for k, _ in m.named_parameters()] + ["pred." + k for k, _ in p.named_parameters()]
    sizes = [v.numel() for _, v in m.named_parameters()] + [v.numel() for _, v in p.named_parameters()]
    o, worst_a, worst_m = 0, 0.0, 0.0
    for nm, k in zip(names, sizes):
        if not re.search(r"cnn\.conv\d\.bias$", nm):
            if nm in outs["adv_grads"]:
                e = rel_l2(tr.grads_adv[o:o + k].cpu().numpy(), outs["adv_grads"][nm].numpy().reshape(-1))
                worst_a = max(worst_a, e)
                # the encoder's adversarial gradient comes back through the discriminator's five small-batch BatchNorms and
                # the single-pass tf32 conv data gradients / weight-gradient reductions (cuDNN's arithmetic for the
                # reference on a GPU): measured 4e-3 ... 5.2e-3 on the smallest tensors, stated 1.5e-2
                assert e < 1.5e-2, ("adv", nm, e)
            e = rel_l2(tr.grads[o:o + k].cpu().numpy(), outs["grads"][nm].numpy().reshape(-1))
            worst_m = max(worst_m, e)
            assert e < 5e-3, ("main", nm, e)
        o += k
    print(f"[ada p={p_drop}] worst gradient rel_l2: discriminator {worst_d:.2e}, encoder (adversarial) {worst_a:.2e}, main {worst_m:.2e}")
    # ---- parameters after the optimiser steps (two SGD steps on the encoder, one on predictor / discriminator), teacher EMA
    ref_after = torch.cat([v.detach().reshape(-1) for v in list(oc.parameters()) + list(op.parameters())]).numpy()
    moved = np.abs(ref_after - p_before.cpu().numpy()).max()
    err = np.abs(tr.params.cpu().numpy() - ref_after).max()
    print(f"[ada p={p_drop}] parameters moved by up to {moved:.2e}; max deviation from the oracle after the step {err:.2e}")
    assert err < 2e-2 * moved + 1e-7
    ref_d = torch.cat([v.detach().reshape(-1) for v in od.parameters()]).numpy()
    assert np.abs(d.flat_tensors()[0].cpu().numpy() - ref_d).max() < 1e-5
    ref_t = torch.cat([v.detach().reshape(-1) for v in list(tc.parameters()) + list(tp.parameters())]).numpy()
    assert np.abs(tr.ema_params.cpu().numpy() - ref_t).max() < 1e-5
    assert int(m.cnn.batchnorm0.num_batches_tracked) == 4 and int(d.bn_1.num_batches_tracked) == 1


def test_train_mt_drives_the_fused_adaptation_step():
    """train_mt(..., discriminator, optimizer_d, optimizer_crnn) with FusedSGD fronts = the fused iteration."""
    from bsed_b200 import main as bmain
    from bsed_b200.DA.cdan_frame import ConditionalDomainAdversarialLoss
    from bsed_b200.models.CRNN import Clip_Discriminator
    n = 2
    oc, op, tc, tp, od, xs, xr, xr_ema, ts, tw = _setup(n, 0.5)
    m, p = bsed_models(oc, op, dropout=0.5)
    em, ep = bsed_models(tc, tp, dropout=0.5)
    for mod in (m, p, em, ep):
        mod.train()
    d = Clip_Discriminator(256).cuda().train()
    crit = ConditionalDomainAdversarialLoss(d)
    real = [(((xr, xr_ema), tw), ["r0", "r1"])] * 2
    syn = [(((xs, xs), ts), ["s0", "s1"])]
    kw = dict(lr=5e-4, momentum=0.9, weight_decay=1e-4, nesterov=True)
    opt = bmain.FusedSGD(list(m.parameters()) + list(p.parameters()), **kw)
    opt_c, opt_d = bmain.FusedSGD(m.parameters(), **kw), bmain.FusedSGD(d.parameters(), **kw)
    d0, p0 = d._flat.clone(), m._flat.clone()
    loss = bmain.train_mt(real, syn, m, opt, 0, ema_model=em, ema_predictor=ep, predictor=p, discriminator=crit,
                          optimizer_d=opt_d, optimizer_crnn=opt_c)
    assert torch.isfinite(loss) and float(loss) > 0
    assert isinstance(opt._trainer, bmain.AdaptationTrainer) and opt._trainer.adv_step == 2
    assert crit.grl.iter_num == 2 and not torch.equal(d0, d.flat_tensors()[0]) and not torch.equal(p0, m.flat_tensors()[0])
    assert int(d.bn_1.num_batches_tracked) == 2 and int(m.cnn.batchnorm0.num_batches_tracked) == 8
