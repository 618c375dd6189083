"""Trainer robustness: short last batches, optimiser state in checkpoints, no-op module moves."""
import copy

import numpy as np
import pytest
import torch

from helpers import bsed_models, oracle_models
from bsed_b200.utilities import synth

pytestmark = pytest.mark.gpu


def _models(p_drop=0.5, seeds=(5, 6)):
    oc, op = oracle_models(seed=seeds[0], linear_std=0.2)
    tc, tp = oracle_models(seed=seeds[1], linear_std=0.2)
    m, p = bsed_models(oc, op, dropout=p_drop)
    em, ep = bsed_models(tc, tp, dropout=p_drop)
    for mod in (m, p, em, ep):
        mod.train()
    for prm in list(em.parameters()) + list(ep.parameters()):
        prm.detach_()
    return m, p, em, ep


def _batch(ns, nr, seed=0):
    xs = torch.from_numpy(synth.make_logmel_like(ns, seed=61 + seed)).cuda()
    xr = torch.from_numpy(synth.make_logmel_like(nr, seed=62 + seed)).cuda()
    ts = torch.from_numpy(synth.make_targets(ns, seed=63 + seed)).cuda()
    return xs, xr, (xr + 0.1).contiguous(), ts


def test_short_batches_run_in_the_plan_built_for_full_ones():
    """The reference's loaders have no drop_last: a trainer built for 3 + 3 clips must take 2 + 3, 3 + 1 ... and give what a
    trainer built for exactly that size gives."""
    from bsed_b200.main import MeanTeacherTrainer
    for ns, nr in ((2, 3), (3, 1), (1, 1)):
        xs, xr, xe, ts = _batch(ns, nr)
        big = MeanTeacherTrainer(*_models(), lr=5e-4, n_syn=3, n_real=3)
        fit = MeanTeacherTrainer(*_models(), lr=5e-4, n_syn=ns, n_real=nr)
        la, lb = big.step(xr, xe, xs, ts, 7, 100), fit.step(xr, xe, xs, ts, 7, 100)
        torch.cuda.synchronize()
        assert torch.allclose(la, lb, rtol=1e-5, atol=1e-7), (ns, nr, la, lb)
        # Adam's first step moves every weight by ~lr whatever the gradient size: an element whose gradient is rounding
        # noise (BatchNorm statistics are accumulated with atomics) may move the other way -- bounded by 2 lr, rare
        d = (big.params - fit.params).abs()
        assert float(d.max()) < 1.1e-3 and float(d.mean()) < 1e-5, (float(d.max()), float(d.mean()))
        assert big.last["strong"].shape[0] == ns + nr
    xs, xr, xe, ts = _batch(4, 3)
    with pytest.raises(ValueError, match="exceeds"):
        big.step(xr, xe, xs, ts, 8, 100)
    with pytest.raises(ValueError):
        big.step(xr, xe[:2], xs[:3], ts[:3], 8, 100)


def test_shift_consistency_trainer_takes_a_short_batch():
    from bsed_b200.main import ShiftConsistencyTrainer
    tr = ShiftConsistencyTrainer(*_models(), lr=5e-4, n=3)
    xs, xr, xe, ts = _batch(2, 2)
    tw = ts.max(1)[0]
    l = tr.step(xr, xe, tw, xs, ts, [8, -12], [1, -2], 3, 0.5)
    torch.cuda.synchronize()
    assert l.shape == (12,) and bool(torch.isfinite(l).all())
    with pytest.raises(ValueError, match="equally long"):
        tr.step(xr, xe, tw, xs[:1], ts[:1], [8, -12], [1, -2], 4, 0.5)


def test_fused_adam_state_dict_is_adams_and_resumes_exactly():
    """FusedAdam.state_dict() is torch.optim.Adam's layout (the reference resumes with optim.load_state_dict(
    state['optimizer']['state_dict']), src/main.py:831-836): a stock Adam accepts it, and a fresh trainer that loads it
    continues bit for bit like the one that was never interrupted."""
    from bsed_b200 import main as bmain
    from bsed_b200.utilities import checkpoint
    from bsed_b200 import engine
    xs, xr, xe, ts = _batch(2, 2)
    real = [(((xr.cpu(), xe.cpu()), torch.zeros(2, 313, 20)), ["r0", "r1"])] * 2
    syn = [(((xs.cpu(), xs.cpu()), ts.cpu()), ["s0", "s1"])]
    m, p, em, ep = _models()
    opt = bmain.FusedAdam(list(m.parameters()) + list(p.parameters()), lr=5e-4, betas=(0.9, 0.999))
    bmain.train_mt(real, syn, m, opt, 0, ema_model=em, ema_predictor=ep, predictor=p)
    state = checkpoint.build_state(m, p, engine.REFERENCE_CRNN_KWARGS, engine.REFERENCE_PREDICTOR_KWARGS, optimizer=opt,
                                   ema_model=em, ema_predictor=ep, epoch=1)
    state = copy.deepcopy(state)
    sd = state["optimizer"]["state_dict"]
    n_params = len(list(m.parameters())) + len(list(p.parameters()))
    assert sorted(sd["state"].keys()) == list(range(n_params)) and set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert float(sd["state"][3]["step"]) == 2.0 and sd["state"][0]["exp_avg"].shape == m.cnn.conv0.weight.shape
    assert float(sd["state"][0]["exp_avg_sq"].abs().sum()) > 0
    # a stock Adam over same-shaped parameters takes it
    clones = [torch.nn.Parameter(q.detach().clone()) for q in list(m.parameters()) + list(p.parameters())]
    stock = torch.optim.Adam(clones, lr=5e-4)
    stock.load_state_dict(sd)
    assert torch.equal(stock.state[clones[0]]["exp_avg"].cpu(), sd["state"][0]["exp_avg"].cpu())
    # resume in fresh modules / a fresh optimiser, then one more epoch on both
    m2, p2, em2, ep2 = _models(seeds=(8, 9))
    opt2 = bmain.FusedAdam(list(m2.parameters()) + list(p2.parameters()), lr=5e-4, betas=(0.9, 0.999))
    assert checkpoint.load_models(state, m2, p2, em2, ep2, optimizer=opt2) == 1
    bmain.train_mt(real, syn, m, opt, 1, ema_model=em, ema_predictor=ep, predictor=p)
    bmain.train_mt(real, syn, m2, opt2, 1, ema_model=em2, ema_predictor=ep2, predictor=p2)
    torch.cuda.synchronize()
    assert opt2._trainer.opt_step == opt._trainer.opt_step == 4
    # BatchNorm statistics are accumulated with atomics (summation order varies run to run): agreement to rounding
    d = (opt2._trainer.params - opt._trainer.params).abs()
    assert float(d.max()) < 1.1e-3 and float(d.mean()) < 1e-5, (float(d.max()), float(d.mean()))
    assert float((opt2._trainer.m - opt._trainer.m).norm() / opt._trainer.m.norm()) < 1e-2


def test_noop_module_moves_keep_the_trainer_buffers():
    from bsed_b200.main import MeanTeacherTrainer
    m, p, em, ep = _models()
    tr = MeanTeacherTrainer(m, p, em, ep, lr=5e-4, n_syn=1, n_real=1)
    ptr0 = m.flat_tensors()[0].data_ptr()
    m.cuda()
    m.to("cuda")
    m.float()
    assert m.flat_tensors()[0].data_ptr() == ptr0 == tr.params.data_ptr()
    xs, xr, xe, ts = _batch(1, 1)
    tr.step(xr, xe, xs, ts, 0, 100)
    # a real move detaches the module from the joint buffer: the trainer refuses to run on
    m.cpu()
    m.cuda()
    with pytest.raises(RuntimeError, match="no longer lives"):
        tr.step(xr, xe, xs, ts, 1, 100)


@pytest.mark.parametrize("fpn", [False, True])
def test_cuda_graph_step_equals_the_kernel_by_kernel_step(fpn):
    """MeanTeacherTrainer.step replays one CUDA graph per iteration from the second call on (device-resident step state:
    dropout keys, Adam bias corrections, EMA coefficient, consistency weight).  Five iterations against the same trainer
    enqueued kernel by kernel with host-side scalars: losses, parameters, teacher, BatchNorm statistics."""
    from bsed_b200.main import MeanTeacherTrainer
    from helpers import bsed_fpn_models, oracle_fpn_models

    def models():
        if not fpn:
            return _models()
        oc, op = oracle_fpn_models(seed=5, linear_std=0.2)
        tc, tp = oracle_fpn_models(seed=6, linear_std=0.2)
        m, p = bsed_fpn_models(oc, op, dropout=0.5)
        em, ep = bsed_fpn_models(tc, tp, dropout=0.5)
        for mod in (m, p, em, ep):
            mod.train()
        for prm in list(em.parameters()) + list(ep.parameters()):
            prm.detach_()
        return m, p, em, ep

    xs, xr, xe, ts = _batch(2, 2)
    a = MeanTeacherTrainer(*models(), lr=5e-4, n_syn=2, n_real=2, graph=True)
    b = MeanTeacherTrainer(*models(), lr=5e-4, n_syn=2, n_real=2, graph=False)
    la, lb = [], []
    for it, gstep in enumerate((40, 41, 42, 43, 60)):          # the jump exercises the state write-through
        if it == 3:
            a.lr = b.lr = 2e-4                                   # a schedule change between iterations
        la.append(a.step(xr, xe, xs, ts, gstep, 100).clone())
        lb.append(b.step(xr, xe, xs, ts, gstep, 100).clone())
    torch.cuda.synchronize()
    assert len(a._graphs) == 1 and not b._graphs and list(a.graph_launches.values())[0] > 100
    for it, (u, v) in enumerate(zip(la, lb)):
        assert torch.allclose(u, v, rtol=2e-4, atol=1e-7), (it, u, v)
    d = (a.params - b.params).abs()
    assert float(d.max()) < 2.1e-3 and float(d.mean()) < 2e-5, (float(d.max()), float(d.mean()))
    assert float((a.ema_params - b.ema_params).abs().max()) < 1e-4
    sa, sb = a.model.state_dict(), b.model.state_dict()
    for k in sa:
        if "running" in k:
            assert torch.allclose(sa[k], sb[k], rtol=1e-3, atol=1e-5), k
        if "num_batches" in k:
            assert int(sa[k]) == int(sb[k]) and int(sa[k]) in (10, 20)      # bn_fcn of CNN_FPN counts twice per call
    assert a.opt_step == b.opt_step == 5
