"""Adversarial domain-adaptation branch through the C ABI: Clip_Discriminator forward /
backward, gradient reversal and the clip-level CDAN loss against the reference fixture (tests/golden/ada.npz) and the
CPU oracle.  Tolerances: probabilities 1e-4, loss 1e-5 relative, gradients 2e-3 relative L2 per tensor (fp32 CUDA-core
GEMMs; summation order differs from the reference's cuDNN / MKL kernels)."""
import numpy as np
import pytest
import torch

from helpers import golden, max_abs, rel_l2
from oracle import da as oda

pytestmark = pytest.mark.gpu


def _disc():
    from bsed_b200.models.CRNN import Clip_Discriminator
    oc = oda.OracleClipDiscriminator()
    oda.seeded_disc_init(oc, 3)
    d = Clip_Discriminator(256)
    d.load_state_dict(oc.state_dict())
    return oc, d.cuda()


def test_state_dict_keys_and_param_count():
    g = golden("ada.npz")
    _, d = _disc()
    assert list(d.state_dict().keys()) == [str(k) for k in g["keys"]]
    from bsed_b200 import _lib
    assert d._flat.numel() == int(_lib.load().bsed_disc_param_count())


def test_forward_train_and_eval_match_reference_fixture():
    g = golden("ada.npz")
    _, d = _disc()
    f = torch.cat((oda.seeded_features(2, 31), oda.seeded_features(3, 32))).cuda()
    d.train()
    p = d(f)
    assert tuple(p.shape) == (5, 1)
    assert max_abs(p.detach().cpu().numpy().reshape(-1), g["p_train"]) < 1e-4
    sd = d.state_dict()
    for k in ("bn_1.running_mean", "bn_3.running_var", "bn_5.running_var"):
        # two train forwards happened in the fixture run for bn stats? no: the fixture's loss forward is the only one
        assert max_abs(sd[k].cpu().numpy(), g["s_" + k]) < 1e-5, k
    assert int(sd["bn_2.num_batches_tracked"]) == 1
    d.eval()
    with torch.no_grad():
        pe = d(f)
    assert max_abs(pe.cpu().numpy().reshape(-1), g["p_eval"]) < 1e-4


def test_cdan_loss_and_gradients_match_reference_fixture():
    from bsed_b200.DA.cdan_frame import ConditionalDomainAdversarialLoss
    g = golden("ada.npz")
    _, d = _disc()
    d.train()
    f_s = oda.seeded_features(2, 31).cuda().requires_grad_(True)
    f_t = oda.seeded_features(3, 32).cuda().requires_grad_(True)
    crit = ConditionalDomainAdversarialLoss(d, entropy_conditioning=False, randomized=False, reduction='mean')
    crit.grl.iter_num = 500
    loss = crit(torch.rand(2, 313, 20).cuda(), f_s, torch.rand(3, 313, 20).cuda(), f_t)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * max(1.0, float(g["loss"]))
    assert crit.grl.iter_num == int(g["iter_after"])
    assert rel_l2(f_s.grad.cpu().numpy()[:, ::7, ::5], g["df_s"]) < 2e-3
    assert rel_l2(f_t.grad.cpu().numpy()[:, ::7, ::5], g["df_t"]) < 2e-3
    assert abs(float(f_s.grad.norm()) - float(g["df_s_norm"])) < 2e-3 * float(g["df_s_norm"])
    bad = []
    for name, p in d.named_parameters():
        gn = float(g["gn_" + name])
        got = p.grad.reshape(-1).cpu().numpy()
        got = got[:: max(1, got.size // 2048)][:2048]
        if name.startswith("conv_") and name.endswith(".bias"):
            # identically zero in exact arithmetic (a train-mode BatchNorm follows): the reference holds rounding noise,
            # this library does not compute it
            assert gn < 1e-3 * float(g["gn_" + name.replace(".bias", ".weight")])
            if np.abs(got).max() > 1e-5:
                bad.append((name, float(np.abs(got).max())))
        elif rel_l2(got, g["g_" + name]) > 2e-3:
            bad.append((name, rel_l2(got, g["g_" + name])))
    assert not bad, bad


def test_discriminator_behind_the_crnn_encoder():
    """The adversarial update of src/main_scmt_ada_weak_seperate.py:314-335 in miniature: encoder features of two
    'domains' -> GRL -> D -> BCE; the encoder receives the reversed gradient, torch SGD(nesterov) steps both."""
    from helpers import bsed_models, oracle_models
    from bsed_b200.DA.cdan_frame import ConditionalDomainAdversarialLoss
    from bsed_b200.utilities import synth
    oc, op = oracle_models(seed=5, linear_std=0.2)
    m, p = bsed_models(oc, op)
    _, d = _disc()
    m.train(); d.train()
    crit = ConditionalDomainAdversarialLoss(d)
    crit.grl.iter_num = 800
    opt_c = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True)
    opt_d = torch.optim.SGD(d.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True)
    xs = torch.from_numpy(synth.make_logmel_like(2, seed=1)).cuda()
    xt = torch.from_numpy(synth.make_logmel_like(2, seed=2)).cuda()
    before_c, before_d = m._flat.clone(), d._flat.clone()
    losses = []
    for _ in range(2):
        opt_c.zero_grad(); opt_d.zero_grad()
        es, _ = m(xs)
        et, _ = m(xt)
        loss = crit(None, es, None, et)
        loss.backward()
        opt_c.step(); opt_d.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and 0.1 < losses[0] < 3.0
    assert not torch.equal(before_c, m.flat_tensors()[0]) and not torch.equal(before_d, d.flat_tensors()[0])
    assert m.cnn.conv3.weight.grad is not None and float(m.cnn.conv3.weight.grad.abs().max()) > 0


@pytest.mark.parametrize("fused", [False, True])
def test_train_mt_with_the_adaptation_branch(fused):
    """train_mt(..., discriminator, optimizer_d, optimizer_crnn) as src/main_scmt_ada_weak_seperate.py drives it."""
    from bsed_b200 import main as bmain
    from bsed_b200.DA.cdan_frame import ConditionalDomainAdversarialLoss
    from test_gpu_train import _inputs, _models
    m, p, em, ep = _models(0.5)
    _, d = _disc()
    d.train()
    crit = ConditionalDomainAdversarialLoss(d)
    crit.grl.iter_num = 300
    xs, xr, xr_ema, ts = _inputs()
    real = [(((xr, xr_ema), torch.zeros(2, 313, 20)), ["r0", "r1"])] * 2
    syn = [(((xs, xs), ts), ["s0", "s1"])]
    params = list(m.parameters()) + list(p.parameters())
    opt = bmain.FusedAdam(params, lr=5e-4, betas=(0.9, 0.999)) if fused else torch.optim.Adam(params, lr=5e-4)
    opt_c = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True)
    opt_d = torch.optim.SGD(d.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True)
    d0 = d._flat.clone()
    loss = bmain.train_mt(real, syn, m, opt, 0, ema_model=em, ema_predictor=ep, predictor=p, discriminator=crit,
                          optimizer_d=opt_d, optimizer_crnn=opt_c)
    assert torch.isfinite(loss) and float(loss) > 0
    assert crit.grl.iter_num == 302 and not torch.equal(d0, d.flat_tensors()[0])
    assert int(d.bn_1.num_batches_tracked) == 2


# (loss rel., data-gradient rel. L2, worst parameter-gradient rel. L2).  tf32x3: forward and data-gradient GEMMs error-compensated
# (loss and df at the fp32 mode's level), weight-gradient reductions single-pass tf32 as in the CRNN
TC_TOL = {"tf32": (2e-3, 5e-2, 5e-2), "tf32x3": (2e-5, 2e-3, 2e-2)}


@pytest.mark.parametrize("precision", ["tf32", "tf32x3"])
def test_tensor_core_discriminator_within_stated_tolerance(precision):
    """Opt-in tcgen05 GEMMs for the discriminator's convolutions (Clip_Discriminator.precision = "tf32" / "tf32x3"): single-pass
    tf32: loss 2e-3 relative, gradients 5e-2 relative L2 (measured 2e-2: five BatchNorms over a 5-clip batch amplify the
    rounding); 3xTF32: see TC_TOL."""
    from bsed_b200.DA.cdan_frame import ConditionalDomainAdversarialLoss
    g = golden("ada.npz")
    _, d = _disc()
    d.train()
    d.precision = precision
    tol_loss, tol_df, tol_g = TC_TOL[precision]
    f_s = oda.seeded_features(2, 31).cuda().requires_grad_(True)
    f_t = oda.seeded_features(3, 32).cuda().requires_grad_(True)
    crit = ConditionalDomainAdversarialLoss(d)
    crit.grl.iter_num = 500
    loss = crit(None, f_s, None, f_t)
    loss.backward()
    torch.cuda.synchronize()
    e_loss = abs(float(loss) - float(g["loss"])) / float(g["loss"])
    e_df = rel_l2(f_s.grad.cpu().numpy()[:, ::7, ::5], g["df_s"])
    worst = 0.0
    for name, p in d.named_parameters():
        if name.startswith("conv_") and name.endswith(".bias"):
            continue
        got = p.grad.reshape(-1).cpu().numpy()
        got = got[:: max(1, got.size // 2048)][:2048]
        worst = max(worst, rel_l2(got, g["g_" + name]))
    print("[%s] discriminator: loss rel %.2e, df rel_l2 %.2e, worst gradient rel_l2 %.3e" % (precision, e_loss, e_df, worst))
    assert e_loss < tol_loss and e_df < tol_df and worst < tol_g
