"""Net_resnet inference (src/audio_tagging_system_cnn.py:50-64, src/audio_tagging_inference.py:123-133,295) on the GPU
against the torchvision-based oracle and its fixture."""
import numpy as np
import pytest
import torch

from helpers import golden, max_abs
from oracle import resnet as ores
from bsed_b200.utilities import synth

pytestmark = pytest.mark.gpu


def test_building_blocks_match_torch():
    from bsed_b200 import engine
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 37, 20, 8, generator=g)                         # channels-last
    w = torch.randn(16, 8, 3, 3, generator=g)
    col, Ho, Wo = engine.im2col_nhwc(x.cuda(), 3, 3, 2, 2, 1, 1, 96)
    ref = torch.nn.functional.unfold(x.permute(0, 3, 1, 2), 3, padding=1, stride=2)        # (B, Cin*9, L), (ci, ky, kx) order
    ref = ref.view(2, 8, 9, -1).permute(0, 3, 2, 1).reshape(2 * Ho * Wo, 72)               # -> (ky, kx, ci)
    assert torch.equal(col[:, :72].cpu(), ref) and float(col[:, 72:].abs().max()) == 0.0
    mp = engine.maxpool_nhwc(x.cuda(), 3, 2, 1)
    assert torch.equal(mp.cpu(), torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1))
    ap = engine.avgpool_nhwc(x.cuda())
    assert max_abs(ap.cpu().numpy(), x.mean(dim=(1, 2)).numpy()) < 1e-6
    y = torch.randn(2, 5, 4, 8, generator=g)
    r = torch.randn(2, 5, 4, 8, generator=g)
    assert torch.equal(engine.add_relu(y.clone().cuda(), r.cuda()).cpu(), torch.relu(y + r))
    # 1-channel 7x7 stem geometry (scalar im2col path)
    x1 = torch.randn(1, 30, 16, 1, generator=g)
    col1, Ho1, Wo1 = engine.im2col_nhwc(x1.cuda(), 7, 7, 2, 2, 3, 3, 64)
    ref1 = torch.nn.functional.unfold(x1.permute(0, 3, 1, 2), 7, padding=3, stride=2).permute(0, 2, 1).reshape(Ho1 * Wo1, 49)
    assert torch.equal(col1[:, :49].cpu(), ref1)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("tf32", 1e-2)])
def test_eval_forward_matches_oracle_and_fixture(precision, tol):
    from bsed_b200.models.ResNet import Net_resnet
    g = golden("resnet_eval.npz")
    oc = ores.seeded_init(ores.OracleNetResnet(20), seed=17).eval()
    m = Net_resnet(pretrained=False, precision=precision)
    m.load_state_dict(oc.state_dict())                                  # identical key set (strict)
    m = m.cuda().eval()
    x = torch.from_numpy(synth.make_logmel_like(3, seed=51))
    out = m(x.cuda())
    with torch.no_grad():
        ref = oc(x)
    err_o, err_g = max_abs(out.cpu().numpy(), ref.numpy()), max_abs(out.cpu().numpy(), g["out"])
    print(f"resnet {precision}: max |p - oracle| = {err_o:.2e}, vs fixture {err_g:.2e}")
    assert out.shape == (3, 20) and err_o < tol and err_g < tol         # north-star 1e-3 in fp32; stated 1e-2 for tf32
    # clips are independent of their batch
    one = m(x[1:2].cuda())
    assert max_abs(one.cpu().numpy(), out[1:2].cpu().numpy()) < 1e-6


def test_weak_label_rows_from_the_tagger():
    """src/audio_tagging_inference.py:289-316: pred >= 0.5 -> comma-joined labels per file (the pseudo-label TSV rows)."""
    from bsed_b200.models.ResNet import Net_resnet
    from bsed_b200.data import config as cfg
    oc = ores.seeded_init(ores.OracleNetResnet(20), seed=17).eval()
    m = Net_resnet(pretrained=False, precision="fp32")
    m.load_state_dict(oc.state_dict())
    m = m.cuda().eval()
    x = torch.from_numpy(synth.make_logmel_like(3, seed=51))
    p = m(x.cuda()).cpu().numpy()
    with torch.no_grad():
        q = oc(x).numpy()
    safe = np.abs(q - 0.5) > 1e-3                                        # away from the threshold the labels must agree
    assert np.array_equal((p >= 0.5)[safe], (q >= 0.5)[safe])
    rows = [",".join(cfg.bird_list[c] for c in np.nonzero(pi >= 0.5)[0]) for pi in p]
    assert len(rows) == 3
    # the reference-facing loop: loader of (((x, x_ema), target), paths) -> DataFrame(filename, event_labels)
    from bsed_b200.evaluation_measures import get_weak_predictions
    loader = [(((x[:2], x[:2]), None), ["a.wav", "b.wav"]), (((x[2:], x[2:]), None), ["c.wav"])]
    df = get_weak_predictions(m, None, loader)
    assert list(df.columns) == ["filename", "event_labels"]
    want = {f: r for f, r in zip(["a.wav", "b.wav", "c.wav"], rows) if r}
    assert dict(zip(df["filename"], df["event_labels"])) == want


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("M,C,const_grad", [(320, 512, False), (320, 512, True), (20096, 64, False)])
def test_train_mode_batchnorm_rows_forward_backward(M, C, const_grad):
    """bsed_bn_rows_train / bsed_bn_rows_backward against torch's batch_norm + residual + ReLU autograd (fp32), including
    the worst case of the tagger: a gradient that is constant over the rows (what the global average pool sends back)."""
    from bsed_b200 import engine
    torch.manual_seed(0)
    x, res = torch.randn(M, C) * 3 + 1, torch.randn(M, C)
    gamma, beta = torch.rand(C) + 0.5, torch.randn(C) * 0.1
    dy = (torch.randn(1, C).expand(M, C) / M).contiguous() if const_grad else torch.randn(M, C)
    xr, gr, br, rr = [t.clone().requires_grad_() for t in (x, gamma, beta, res)]
    rm, rv = torch.zeros(C), torch.ones(C)
    torch.relu(torch.nn.functional.batch_norm(xr, rm, rv, gr, br, True, 0.1, 1e-5) + rr).backward(dy)
    xd = x.cuda()
    rmd, rvd, nbt = torch.zeros(C).cuda(), torch.ones(C).cuda(), torch.zeros(1, dtype=torch.int64).cuda()
    y, mr = engine.bn_rows_train(xd, gamma.cuda(), beta.cuda(), rmd, rvd, nbt, res.cuda(), True)
    dg, db = torch.zeros(C).cuda(), torch.zeros(C).cuda()
    dyd = dy.cuda().clone()
    dres = engine.bn_rows_backward(dyd, y, xd, gamma.cuda(), mr, dg, db, True)
    assert _rel(rmd.cpu(), rm) < 1e-6 and _rel(rvd.cpu(), rv) < 1e-6 and int(nbt) == 1
    assert _rel(dyd.cpu(), xr.grad) < 2e-6 and _rel(dg.cpu(), gr.grad) < 2e-6 and _rel(db.cpu(), br.grad) < 2e-6
    assert torch.equal(dres.cpu(), rr.grad)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,s,p", [(2, 40, 4, 512, 512, 3, 1, 1), (2, 79, 8, 128, 256, 3, 2, 1),
                                                  (2, 79, 8, 128, 256, 1, 2, 0), (2, 30, 16, 1, 64, 7, 2, 3)])
def test_conv_unit_gradients(B, H, W, Cin, Cout, k, s, p):
    """im2col -> GEMM forward, weight gradient (GEMM on the im2col matrix) and data gradient (GEMM -> col2im) against
    torch's conv2d autograd, for every convolution geometry of the tagger."""
    from bsed_b200 import engine
    torch.manual_seed(1)
    x, w = torch.randn(B, Cin, H, W), torch.randn(Cout, Cin, k, k) * 0.05
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    yr = torch.nn.functional.conv2d(xr, wr, None, s, p)
    dy = torch.randn_like(yr)
    yr.backward(dy)
    K = k * k * Cin
    kpad = (K + 31) // 32 * 32
    wk = torch.zeros(Cout, kpad)
    wk[:, :K] = w.permute(0, 2, 3, 1).reshape(Cout, K)
    wk = wk.cuda()
    col, Ho, Wo = engine.im2col_nhwc(x.permute(0, 2, 3, 1).contiguous().cuda(), k, k, s, s, p, p, kpad)
    y = engine.gemm_nn(col, wk.t().contiguous())
    dyh = dy.permute(0, 2, 3, 1).reshape(-1, Cout).contiguous().cuda()
    dwk = torch.zeros(Cout, kpad).cuda()
    engine.gemm_tn(dyh, col, dwk)
    dx = engine.col2im_nhwc(engine.gemm_nn(dyh, wk), (B, H, W, Cin), k, k, s, s, p, p, kpad)
    assert _rel(y.view(B, Ho, Wo, Cout).permute(0, 3, 1, 2).cpu(), yr.detach()) < 5e-6
    assert _rel(dwk[:, :K].view(Cout, k, k, Cin).permute(0, 3, 1, 2).cpu(), wr.grad) < 5e-6
    assert _rel(dx.permute(0, 3, 1, 2).cpu(), xr.grad) < 5e-6


def test_maxpool_backward_routes_to_the_first_maximum():
    from bsed_b200 import engine
    torch.manual_seed(2)
    x = torch.randn(2, 64, 61, 16).round()          # many ties
    xr = x.clone().requires_grad_()
    yr = torch.nn.functional.max_pool2d(xr, 3, 2, 1)
    dy = torch.randn_like(yr)
    yr.backward(dy)
    dx = engine.maxpool_nhwc_backward(x.permute(0, 2, 3, 1).contiguous().cuda(), dy.permute(0, 2, 3, 1).contiguous().cuda(), 3, 2, 1)
    assert torch.equal(dx.permute(0, 3, 1, 2).cpu(), xr.grad)


# Whole-network gradients: the kernels above agree with torch to ~1e-7, but 20 train-mode BatchNorms on a 2-clip batch
# behind a global average pool (a gradient that is constant over the rows of a clip) make the backward pass ill-conditioned:
# torch's own fp32 gradients differ from a float64 run of the same model by 4e-3 (tests/make_golden_resnet.py), this
# implementation's from torch's fp32 ones by 1.0-1.7e-2 on every tensor upstream of the last BatchNorm.  Stated tolerance 3e-2.
GRAD_TOL = 3e-2


def _train_inputs():
    xs = torch.from_numpy(synth.make_logmel_like(2, seed=61))
    xr = torch.from_numpy(synth.make_logmel_like(2, seed=62))
    ts = torch.from_numpy(synth.make_targets(2, seed=63))
    tw = (torch.from_numpy(synth.make_targets(2, seed=64)).max(-2)[0] > 0).float()
    return xs, xr, ts, tw


def _check_grads(g, named_grads, tol):
    bad, worst = [], 0.0
    for n, got in named_grads:
        gn = float(g["gn_" + n])
        if gn < 1e-6:
            continue
        got = got.cpu().numpy().reshape(-1)
        got = got if got.size <= 4096 else got[:: max(1, got.size // 4096)][:4096]
        e = float(np.linalg.norm(got.astype(np.float64) - g["g_" + n]) / max(np.linalg.norm(g["g_" + n]), 1e-30))
        worst = max(worst, e)
        if e > tol:
            bad.append((n, e))
    return bad, worst


def test_training_step_through_autograd_matches_fixture():
    """Reference statement order (src/audio_tagging_system_cnn.py:340-406): two model calls, BCELoss, loss.backward()."""
    from bsed_b200.models.ResNet import Net_resnet
    g = golden("resnet_train.npz")
    oc = ores.seeded_init(ores.OracleNetResnet(20), seed=17)
    m = Net_resnet(pretrained=False, precision="fp32")
    m.load_state_dict(oc.state_dict())
    m = m.cuda().train()
    xs, xr, ts, tw = [t.cuda() for t in _train_inputs()]
    loss, _ = ores.tagger_step_loss(m, xs, ts, xr, tw)                   # the oracle's loss assembly on the CUDA model
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-4)
    loss.backward()
    bad, worst = _check_grads(g, [(n, p.grad) for n, p in m.named_parameters()], GRAD_TOL)
    print(f"resnet train (autograd): loss {float(loss):.6f}, worst gradient rel_l2 {worst:.2e}")
    assert not bad, bad
    assert int(m.resnet.bn1.num_batches_tracked) == int(g["nbt"]) == 2
    for k in ("resnet.bn1.running_mean", "resnet.layer3.0.downsample.1.running_var", "resnet.layer4.1.bn2.running_var"):
        assert max_abs(m.state_dict()[k].cpu().numpy().reshape(-1)[:2048], g["s_" + k]) < 1e-3, k


def test_fused_tagger_trainer_matches_fixture():
    from bsed_b200.models.ResNet import Net_resnet, TaggerTrainer
    g = golden("resnet_train.npz")
    oc = ores.seeded_init(ores.OracleNetResnet(20), seed=17)
    m = Net_resnet(pretrained=False, precision="fp32")
    m.load_state_dict(oc.state_dict())
    m = m.cuda().train()
    xs, xr, ts, tw = [t.cuda() for t in _train_inputs()]
    tr = TaggerTrainer(m, lr=1e-3)
    loss = tr.step(xs, ts, xr, tw)
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-4)
    o, named = 0, []
    names = [n for n, _ in m.named_parameters()]
    for (mod, pname, shape), n in zip(m._param_specs, names):
        k = int(np.prod(shape))
        named.append((n, tr.grads[o:o + k]))
        o += k
    bad, worst = _check_grads(g, named, GRAD_TOL)
    print(f"resnet train (fused): worst gradient rel_l2 {worst:.2e}")
    assert not bad, bad
    sd = m.state_dict()
    for k in ("resnet.conv1.weight", "resnet.layer1.0.bn1.weight", "resnet.layer2.0.downsample.0.weight",
              "resnet.layer4.1.conv2.weight", "resnet.fc.bias"):
        d = np.abs(sd[k].cpu().numpy().reshape(-1)[:2048].astype(np.float64) - g["s_" + k])
        # one Adam step moves every weight by ~lr = 1e-3 whatever the gradient size: elements whose gradient is rounding
        # noise may move the other way
        assert d.max() < 2.2e-3 and d.mean() < 2e-4, (k, d.max(), d.mean())


@pytest.mark.parametrize("K,M,N", [(1000, 64, 64), (20096, 64, 640), (320, 512, 4608), (640, 256, 128)])
def test_gemm_tn_tensor_cores(K, M, N):
    from bsed_b200 import engine
    g = torch.Generator().manual_seed(K + M + N)
    a, b = torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)
    out = torch.zeros(M, N, device="cuda")
    engine.gemm_tn_tc(a.cuda(), b.cuda(), out)
    ref = a.double().t() @ b.double()
    assert _rel(out.cpu().double(), ref) < 2e-3           # tf32 operands, fp32 accumulation


@pytest.mark.parametrize("bwd", ["fp32", "tf32"])
def test_tf32_training_step_gradients(bwd):
    """tcgen05 tf32 GEMMs in the forward (and the backward) of the tagger against the fp32 fixture.  Stated tolerances on
    this ill-conditioned 2-clip problem: loss 1e-3; gradients 0.3 -- the tf32 rounding of the forward activations alone
    moves them by 0.19 (measured, same with fp32 or tf32 backward GEMMs); precision="fp32" is the parity mode."""
    from bsed_b200.models.ResNet import Net_resnet, TaggerTrainer
    g = golden("resnet_train.npz")
    oc = ores.seeded_init(ores.OracleNetResnet(20), seed=17)
    m = Net_resnet(pretrained=False, precision="tf32")
    m.backward_precision = bwd
    m.load_state_dict(oc.state_dict())
    m = m.cuda().train()
    xs, xr, ts, tw = [t.cuda() for t in _train_inputs()]
    tr = TaggerTrainer(m, lr=1e-3)
    loss = tr.step(xs, ts, xr, tw)
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-3)
    o, named = 0, []
    for (mod, pname, shape), n in zip(m._param_specs, [n for n, _ in m.named_parameters()]):
        k = int(np.prod(shape))
        named.append((n, tr.grads[o:o + k]))
        o += k
    bad, worst = _check_grads(g, named, 0.3)
    print(f"resnet train (fused, tf32 forward, {bwd} backward): loss {float(loss):.6f}, worst gradient rel_l2 {worst:.2e}")
    assert not bad, bad


def test_tf32_against_fp32_on_a_larger_batch():
    """The same comparison between this implementation's two precisions on 6 + 6 clips: with more rows per BatchNorm the
    backward is better conditioned and the tf32 deviation shrinks to 3.3e-3 (measured; stated bound 2e-2)."""
    from bsed_b200.models.ResNet import Net_resnet, TaggerTrainer
    oc = ores.seeded_init(ores.OracleNetResnet(20), seed=17)
    xs = torch.from_numpy(synth.make_logmel_like(6, seed=71)).cuda()
    xr = torch.from_numpy(synth.make_logmel_like(6, seed=72)).cuda()
    ts = torch.from_numpy(synth.make_targets(6, seed=73)).cuda()
    tw = (torch.from_numpy(synth.make_targets(6, seed=74)).max(-2)[0] > 0).float().cuda()
    grads, losses = {}, {}
    for prec in ("fp32", "tf32"):
        m = Net_resnet(pretrained=False, precision=prec)
        m.load_state_dict(oc.state_dict())
        m = m.cuda().train()
        tr = TaggerTrainer(m, lr=1e-3)
        losses[prec] = float(tr.step(xs, ts, xr, tw))
        grads[prec] = tr.grads.clone()
    e = _rel(grads["tf32"], grads["fp32"])
    print(f"resnet train 6 + 6 clips: tf32 vs fp32 whole-gradient rel_l2 {e:.2e}, losses {losses}")
    assert losses["tf32"] == pytest.approx(losses["fp32"], rel=1e-3) and e < 2e-2


# measured on B200: fp32 2.9e-3 (torch's own fp32: 1.6e-3); tf32x3 1.3e-2 (forward products fp32-grade, the weight-gradient
# reductions single-pass tf32); default = single-pass tf32 everywhere, cuDNN's arithmetic for the reference: 0.17 on the
# small-norm layer1 BatchNorm scales (|g| ~ 5e-4), 0.14 on conv1.weight -- 20 train-mode BatchNorms amplify the forward rounding
GRAD_TOL8 = {"fp32": 6e-3, "tf32x3": 3e-2, "default": 0.3}


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "default"])
def test_training_step_on_eight_clips_against_float64(precision, monkeypatch):
    """4 + 4 clips: four times the rows behind every train-mode BatchNorm, and the yardstick is a FLOAT64 run of torchvision's
    resnet18 assembled as the reference's Net_resnet (tests/make_golden_resnet.py: train_fixture8), from which torch's own
    fp32 gradients differ by up to 1.6e-3 (dev32_*; layer1 / layer2 tensors ~1e-3, median 4e-6)."""
    from bsed_b200.models.ResNet import Net_resnet, TaggerTrainer
    if precision == "default":
        monkeypatch.delenv("BSED_PRECISION", raising=False)
    g = golden("resnet_train8.npz")
    oc = ores.seeded_init(ores.OracleNetResnet(20), seed=17)
    m = Net_resnet(pretrained=False, precision=None if precision == "default" else precision)
    m.load_state_dict(oc.state_dict())
    m = m.cuda().train()
    xs = torch.from_numpy(synth.make_logmel_like(4, seed=81)).cuda()
    xr = torch.from_numpy(synth.make_logmel_like(4, seed=82)).cuda()
    ts = torch.from_numpy(synth.make_targets(4, seed=83)).cuda()
    tw = (torch.from_numpy(synth.make_targets(4, seed=84)).max(-2)[0] > 0).float().cuda()
    tr = TaggerTrainer(m, lr=1e-3)
    loss = tr.step(xs, ts, xr, tw)
    assert float(loss) == pytest.approx(float(g["loss64"]), rel=1e-4)
    o, bad, worst = 0, [], (0.0, "")
    for (mod, pname, shape), n in zip(m._param_specs, [n for n, _ in m.named_parameters()]):
        k = int(np.prod(shape))
        got = tr.grads[o:o + k].cpu().numpy().reshape(-1).astype(np.float64)
        o += k
        if float(g["gn_" + n]) < 1e-6:
            continue
        got = got if got.size <= 4096 else got[:: max(1, got.size // 4096)][:4096]
        ref = g["g64_" + n].astype(np.float64)
        e = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        worst = max(worst, (e, n))
        if e > GRAD_TOL8[precision]:
            bad.append((n, e, float(g["dev32_" + n])))
    print(f"resnet train 4 + 4 clips vs float64 ({precision}): loss {float(loss):.6f}, worst gradient rel_l2 {worst[0]:.2e} "
          f"({worst[1]}); torch fp32's own worst {float(g['dev32_worst']):.2e}")
    assert not bad, bad


@pytest.mark.parametrize("fused", [True, False])
def test_tagger_train_mt_entry_point(fused):
    """train_mt of src/audio_tagging_system_cnn.py:199 with the reference's three loaders, through the fused trainer and
    through a stock torch optimizer + the autograd function; a teacher copy follows by parameter EMA."""
    from bsed_b200 import audio_tagging_system_cnn as ats
    from bsed_b200.main import FusedAdam
    oc = ores.seeded_init(ores.OracleNetResnet(20), seed=17)
    m = ats.Net_resnet(pretrained=False, precision="fp32")
    m.load_state_dict(oc.state_dict())
    em = ats.Net_resnet(pretrained=False, precision="fp32")
    em.load_state_dict(oc.state_dict())
    m, em = m.cuda().train(), em.cuda().train()
    for prm in em.parameters():
        prm.detach_()
    xs, xr, ts, tw = _train_inputs()
    syn = [(((xs, xs), ts), ["s0", "s1"])]
    weak = [(((xr[:1], xr[:1]), torch.from_numpy(synth.make_targets(1, seed=65))), ["w0"])]
    unl = [(((xr[1:], xr[1:]), tw[1:]), ["u0"])]
    opt = (FusedAdam if fused else torch.optim.Adam)(m.parameters(), lr=1e-3)
    before = m._flat.clone()
    loss = ats.train_mt(unl, weak, syn, m, opt, 0, ema_model=em)
    assert torch.isfinite(loss) and float(loss) > 0
    moved = (m.flat_tensors()[0] - before).abs()
    assert float(moved.max()) > 5e-4 and float(moved.max()) < 1.1e-3            # one Adam step of lr 1e-3
    assert int(m.resnet.bn1.num_batches_tracked) == 2
    # global_step = 1: alpha = min(1 - 1/2, 0.999) = 0.5
    assert torch.allclose(em.flat_tensors()[0], 0.5 * before + 0.5 * m.flat_tensors()[0], atol=1e-7)
