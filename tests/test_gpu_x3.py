"""The library's DEFAULT precision -- error-compensated 3xTF32 on the tcgen05 tensor cores (BSED_PRECISION_TF32X3) --
against the reference-generated fixtures and the CPU oracle.

north_star: CRNN strong / weak probabilities within 1e-3 of the reference's fp32 arithmetic, event lists bit-exact given
identical probabilities.  Every tolerance below is that 1e-3 or tighter; nothing in this file forces a precision (the
environment override BSED_PRECISION is removed), so what is measured is what `get_predictions`, `pseudo_label_stream`,
`MeanTeacherTrainer` and bench.py run.

Kernel level: the split products (a*w_hi + a*w_lo + a_lo*w_hi, csrc/tc_gemm.cu, csrc/tc_conv.cu) against float64 --
relative L2 <= 1e-5 (measured 3e-7 ... 8e-6), against 3e-4 measured / 2e-3 stated for the single-pass tf32 products of
tests/test_gpu_kernels.py.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import bsed_fpn_models, bsed_models, golden, max_abs, oracle_fpn_models, oracle_models, rel_l2
from bsed_b200.utilities import synth
from oracle import postproc as opost

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-3          # north_star
X3_REL = 1e-5            # one split contraction against float64 (measured 3e-7 ... 8e-6, growing with K up to 1152)


@pytest.fixture(autouse=True)
def _library_default(monkeypatch):
    monkeypatch.delenv("BSED_PRECISION", raising=False)


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def test_default_precision_is_3xtf32():
    from bsed_b200 import engine
    assert engine.default_precision() == "tf32x3"
    plan = engine.Plan(engine.make_cfg(), max_clips=1, device="cuda", with_workspace=False)
    assert plan.precision == "tf32x3" and plan.lib.bsed_plan_get_precision(plan.p) == 2


# ---------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,K,N,bias,acc", [
    (1000, 128, 128, False, False),      # 16-wide chunks, weights on the ring (GLU of the 128-channel blocks)
    (40000, 64, 64, True, False),        # 32-wide chunks, resident (hi, lo) weights (GLU of the packed 16/32/64-channel blocks)
    (300, 16, 16, True, False), (777, 256, 64, True, False), (513, 768, 128, False, True),     # GRU data gradient shape
    (129, 64, 64, True, True), (5, 32, 32, False, False), (11268, 256, 128, True, False),      # GRU layer-1 projection
    (3000, 512, 128, True, False),       # fpn merge
    # N > 128: column blocks of 128 inside one launch (GRU projections 768, fpn merge 256, their data gradients)
    (11268, 128, 768, True, False), (1000, 768, 256, False, True), (700, 256, 512, True, False)])
def test_gemm_nt_3xtf32(M, K, N, bias, acc):
    from bsed_b200 import engine
    a, bk = _rand(M, K, seed=1), _rand(N, K, seed=2)
    bi = _rand(N, seed=3) if bias else None
    c0 = _rand(M, N, seed=4)
    ref = a.double() @ bk.double().T + (bi.double() if bias else 0) + (c0.double() if acc else 0)
    out = c0.clone().cuda() if acc else None
    got = engine.gemm_nt_tc(a.cuda(), bk.cuda(), bi.cuda() if bias else None, out=out, accumulate=acc, x3=True)
    torch.cuda.synchronize()
    e = rel_l2(got.cpu().numpy(), ref.numpy())
    print(f"[x3] gemm_nt M={M} K={K} N={N}: rel_l2 {e:.2e}")
    assert e < X3_REL


@pytest.mark.parametrize("B,T,Fq,Cin,Cout", [
    # column-tiled halo kernel (F >= 2, Cin % 32 == 0): partial t blocks, every output width, every chunk count; with 128
    # output channels one frequency bin per step (G = 1) while the tiles would not fill the SMs, two (G = 2, two issuer
    # warps) from 75 two-bin tiles up: (8, 313, 8, ...) = 96 tiles
    (2, 313, 16, 64, 128), (1, 130, 8, 128, 128), (3, 129, 64, 32, 16), (1, 257, 32, 64, 32), (2, 128, 8, 128, 64),
    (1, 1, 8, 32, 32), (2, 37, 16, 32, 32), (2, 313, 32, 32, 64), (8, 313, 8, 128, 128), (3, 11, 2, 128, 128),
    (2, 9, 4, 32, 16), (2, 100, 4, 128, 128), (13, 313, 2, 128, 128),
    # row-tiled kernel (F = 1, or 16 input channels)
    (1, 313, 1, 128, 128), (2, 5, 64, 16, 16)])
def test_conv3x3_3xtf32(B, T, Fq, Cin, Cout):
    from bsed_b200 import engine
    x = _rand(B, Cin, T, Fq, seed=10)
    w = _rand(Cout, Cin, 3, 3, seed=11, scale=0.2)
    b = _rand(Cout, seed=12)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    got = engine.conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), b.cuda(), tensor_cores="tf32x3")
    torch.cuda.synchronize()
    e = rel_l2(got.permute(0, 3, 1, 2).cpu().numpy(), ref.numpy())
    print(f"[x3] conv3x3 B={B} T={T} F={Fq} {Cin}->{Cout}: rel_l2 {e:.2e}")
    assert e < X3_REL


def test_single_pass_tf32_is_three_orders_looser():
    """The same product in the two tensor-core modes: documents what the split buys."""
    from bsed_b200 import engine
    a, bk = _rand(4096, 128, seed=5), _rand(128, 128, seed=6)
    ref = (a.double() @ bk.double().T).numpy()
    e1 = rel_l2(engine.gemm_nt_tc(a.cuda(), bk.cuda()).cpu().numpy(), ref)
    e3 = rel_l2(engine.gemm_nt_tc(a.cuda(), bk.cuda(), x3=True).cpu().numpy(), ref)
    print(f"[x3] gemm 4096x128x128: tf32 {e1:.2e}  3xtf32 {e3:.2e}")
    assert e3 < X3_REL and e1 > 30 * e3


# ---------------------------------------------------------------------------------------------
# inference parity against the reference-generated fixtures (tests/make_golden*.py run the reference's modules)
# ---------------------------------------------------------------------------------------------
def _events(strong):
    return [opost.events_from_strong(s, 0.5, 14) for s in strong]


def _check_events(strong, gold):
    """Event lists: bit-exact given identical probabilities.  The fixture holds probabilities as close as 2e-6 to the
    0.5 threshold, so elements within PROB_TOL of it are taken from the fixture before decoding; everything else must
    already agree, and the device decoder must agree with the oracle decoder on the same probabilities."""
    from bsed_b200 import engine
    near = np.abs(gold - 0.5) <= PROB_TOL
    assert np.array_equal((strong >= 0.5)[~near], (gold >= 0.5)[~near])
    snapped = np.where(near, gold, strong).astype(np.float32)
    ev, n = engine.median_decode(torch.from_numpy(snapped).cuda())
    torch.cuda.synchronize()
    want = _events(gold)
    for b in range(gold.shape[0]):
        got = [tuple(int(v) for v in e) for e in ev[b, :int(n[b])].cpu().numpy()]
        assert got == [tuple(e) for e in want[b]], f"clip {b}: event lists differ"
    return int(near.sum())


def test_crnn_eval_default_precision_matches_reference_fixture():
    g = golden("crnn_eval.npz")
    x = torch.from_numpy(synth.make_logmel_like(2, seed=11))
    oc, op = oracle_models(seed=5, linear_std=0.2)
    m, p = bsed_models(oc, op)
    m.eval()
    p.eval()
    with torch.no_grad():
        enc, _ = m(x.cuda())
        strong, weak = p(enc)
    torch.cuda.synchronize()
    es, ew = max_abs(strong.cpu().numpy(), g["strong"]), max_abs(weak.cpu().numpy(), g["weak"])
    ee = rel_l2(enc.cpu().numpy()[:, ::8], g["enc"])
    near = _check_events(strong.cpu().numpy(), g["strong"])
    print(f"[x3] crnn eval vs reference fixture: strong {es:.2e} weak {ew:.2e} enc rel_l2 {ee:.2e}; {near} probabilities "
          f"within {PROB_TOL} of the threshold")
    assert (m.precision or "tf32x3") == "tf32x3"
    assert es < PROB_TOL and ew < PROB_TOL


def test_fpn_eval_default_precision_matches_reference_fixture():
    g = golden("fpn_eval.npz")
    x = torch.from_numpy(synth.make_logmel_like(2, seed=11))
    oc, op = oracle_fpn_models(seed=5, linear_std=0.2)
    m, p = bsed_fpn_models(oc, op)
    m.eval()
    p.eval()
    with torch.no_grad():
        enc, _ = m(x.cuda())
        strong, weak = p(enc)
    torch.cuda.synchronize()
    es, ew = max_abs(strong.cpu().numpy(), g["strong"]), max_abs(weak.cpu().numpy(), g["weak"])
    near = _check_events(strong.cpu().numpy(), g["strong"])
    print(f"[x3] fpn eval vs reference fixture: strong {es:.2e} weak {ew:.2e}; {near} probabilities near the threshold")
    assert es < PROB_TOL and ew < PROB_TOL


def test_train_forward_default_precision_matches_reference_fixture():
    g = golden("crnn_train_fwd.npz")
    from bsed_b200 import engine
    x = torch.from_numpy(synth.make_logmel_like(2, seed=11))
    oc, op = oracle_models(seed=5, linear_std=0.2)
    m, p = bsed_models(oc, op, dropout=0.5)
    flat, bn, nbt = m.flat_tensors()
    plan = engine.Plan(engine.make_cfg(**m.cfg_kwargs), max_clips=2, device="cuda")
    assert plan.precision == "tf32x3"
    enc = plan.forward([dict(params=flat, bn=bn, nbt=nbt, n=2)], x.cuda(), train=True, save=False, seed=2023, step=3)
    _, strong, weak = plan.predictor_forward(p.flat_tensors()[0], enc)
    torch.cuda.synchronize()
    es, ew = max_abs(strong.cpu().numpy(), g["strong"]), max_abs(weak.cpu().numpy(), g["weak"])
    print(f"[x3] train-mode forward vs reference fixture: strong {es:.2e} weak {ew:.2e}")
    assert es < 2e-4 and ew < 2e-4


def test_get_predictions_default_precision_event_list_matches_oracle():
    """get_predictions (src/evaluation_measures.py:123-283) end to end in the default precision: the prediction frame of
    24 clips against the CPU oracle's probabilities decoded by the oracle decoder."""
    from bsed_b200.evaluation_measures import get_predictions
    from bsed_b200.utilities.ManyHotEncoder import ManyHotEncoder
    from bsed_b200.data import config as cfg
    oc, op = oracle_models(seed=9, linear_std=0.2)
    m, p = bsed_models(oc, op)
    m.eval()
    p.eval()
    n = 6
    x = torch.from_numpy(synth.make_logmel_like(n, seed=31))
    with torch.no_grad():
        so, _ = op(oc(x)[0])
    so = so.numpy()
    names = [f"clip{i}" for i in range(n)]
    loader = [(((x[i:i + 3], None), None), [f"/data/audio/{nm}.wav" for nm in names[i:i + 3]]) for i in range(0, n, 3)]
    enc = ManyHotEncoder(cfg.bird_list, n_frames=cfg.max_frames // cfg.pooling_time_ratio)
    pred, _, _ = get_predictions(m, loader, enc.decode_strong, cfg.pooling_time_ratio, median_window=14, predictor=p)
    frames = {nm: [] for nm in names}
    for _, r in pred.iterrows():
        frames[r["filename"]].append((r["event_label"], round(float(r["onset"]), 3), round(float(r["offset"]), 3)))
    # oracle: its own probabilities through its own decoder.  A (clip, class) column is compared when none of its
    # probabilities lies within PROB_TOL of the threshold (the decision is then the same on both sides)
    checked = 0
    for i, nm in enumerate(names):
        want = opost.to_seconds(opost.events_from_strong(so[i], 0.5, 14))
        for c, label in enumerate(cfg.bird_list):
            if np.any(np.abs(so[i][:, c] - 0.5) <= PROB_TOL):
                continue
            w = sorted((round(a, 3), round(b, 3)) for cc, a, b in want if cc == c)
            g_ = sorted((a, b) for l, a, b in frames[nm] if l == label)
            assert g_ == w, (nm, label, g_, w)
            checked += 1
    print(f"[x3] get_predictions: {checked} of {n * len(cfg.bird_list)} (clip, class) columns compared event by event "
          f"(the others hold a probability within {PROB_TOL} of the threshold)")
    assert checked >= n


# ---------------------------------------------------------------------------------------------
# the bench configuration itself: 12 synthetic + 12 real student clips, 12 teacher clips, against the CPU oracle
# ---------------------------------------------------------------------------------------------
def test_full_size_mean_teacher_step_matches_cpu_oracle():
    """One mean-teacher step at the size bench.py times (12 + 12 + 12 clips; src/main.py:163-527) against the CPU oracle
    (the reference's torch.nn layers, oracle/train.py): the four loss terms, the student / teacher probabilities and
    every parameter gradient.  ~10 s of CPU."""
    from oracle import crnn as ocrnn
    from oracle import train as otrain
    from bsed_b200 import engine
    from bsed_b200.main import MeanTeacherTrainer
    from bsed_b200.models import CRNN, Predictor
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ns = nr = 12
    oc, op = oracle_models(seed=5, linear_std=0.2, dropout=0.5, train=True)
    tc_, tp_ = oracle_models(seed=6, linear_std=0.2, dropout=0.5, train=True)
    for prm in list(tc_.parameters()) + list(tp_.parameters()):
        prm.detach_()
    xs = torch.from_numpy(synth.make_logmel_like(ns, seed=41))
    xr = torch.from_numpy(synth.make_logmel_like(nr, seed=42))
    xr_ema = xr + 0.05 * torch.from_numpy(synth.make_logmel_like(nr, seed=43))
    ts = torch.from_numpy(synth.make_targets(ns, seed=44))
    gstep, ramp = 100, 500

    def mk(o_c, o_p):
        kw = dict(engine.REFERENCE_CRNN_KWARGS)
        m, p = CRNN(**kw), Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
        m.load_state_dict(o_c.state_dict())
        p.load_state_dict(o_p.state_dict())
        return m.cuda().train(), p.cuda().train()

    m, p = mk(oc, op)
    em, ep = mk(tc_, tp_)
    tr = MeanTeacherTrainer(m, p, em, ep, lr=5e-4, n_syn=ns, n_real=nr, dropout_seed=2023)
    assert tr.plan.precision == "tf32x3"
    losses = tr.step(xr.cuda(), xr_ema.cuda(), xs.cuda(), ts.cuda(), gstep, ramp)
    grads = tr.grads.clone()
    torch.cuda.synchronize()

    def hook(tag):
        if tag == "teacher":
            tc_.set_dropout_keys(2023, gstep, ns + nr)
        elif tag == "syn":
            oc.set_dropout_keys(2023, gstep, 0)
        else:
            oc.set_dropout_keys(2023, gstep, ns)

    opt = torch.optim.Adam(list(oc.parameters()) + list(op.parameters()), lr=5e-4, betas=(0.9, 0.999))
    loss_o, parts, outs = otrain.mt_step(oc, op, tc_, tp_, opt, xr, xr_ema, xs, ts, gstep, rampup_length=ramp,
                                         ema_flavour="none", dropout_hook=hook)
    got = losses.cpu().numpy()
    want = np.array([float(parts[k]) for k in ("strong_class", "weak_class", "cons_strong", "cons_weak")])
    print("[x3] full-size step losses", got, "oracle", want)
    assert np.allclose(got, want, rtol=1e-3, atol=1e-6)
    es = max_abs(tr.last["strong"][ns:].cpu().numpy(), outs["strong"].numpy())
    ew = max_abs(tr.last["weak"][ns:].cpu().numpy(), outs["weak"].numpy())
    print(f"[x3] full-size step: student strong {es:.2e} weak {ew:.2e}")
    assert es < PROB_TOL and ew < PROB_TOL
    # every parameter gradient: flat buffer in named_parameters() order (CRNN then Predictor)
    import re
    names = ["crnn." + k for k, _ in m.named_parameters()] + ["pred." + k for k, _ in p.named_parameters()]
    shapes = [v.shape for _, v in m.named_parameters()] + [v.shape for _, v in p.named_parameters()]
    worst, o, n_checked = 0.0, 0, 0
    og = outs["grads"]
    for nm, shp in zip(names, shapes):
        k = int(np.prod(shp))
        ref = og[nm].numpy().reshape(-1)
        # a conv bias ahead of a train-mode BatchNorm has an identically zero gradient (the reference computes rounding
        # noise there; this library leaves it zero)
        if not re.search(r"cnn\.conv\d\.bias$", nm):
            e = rel_l2(grads[o:o + k].cpu().numpy(), ref)
            worst = max(worst, e)
            n_checked += 1
            assert e < 3e-3, (nm, e)
        o += k
    assert o == grads.numel()
    print(f"[x3] full-size step: {n_checked} gradient tensors, worst rel_l2 {worst:.2e}")
    assert n_checked >= 50          # 62 parameter tensors minus the 7 conv biases
