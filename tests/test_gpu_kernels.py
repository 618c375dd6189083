"""Generic kernels through the C ABI against torch-CPU fp32/fp64 references."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import max_abs, rel_l2

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0, scale=1.0):
    return (torch.from_numpy(np.random.default_rng(seed).standard_normal(shape)).float() * scale)


@pytest.mark.parametrize("M,K,N,bias,acc", [(1000, 128, 128, False, False), (300, 16, 16, True, False),
                                            (777, 256, 48, True, False), (513, 768, 256, False, True),
                                            (129, 64, 64, True, True), (5, 32, 32, False, False),
                                            (2049, 144, 32, True, False), (260, 288, 64, False, False),
                                            (4097, 128, 384, True, False)])
def test_gemm_nn(M, K, N, bias, acc):
    from bsed_b200 import engine
    a, b = _rand(M, K, seed=1), _rand(K, N, seed=2)
    bi = _rand(N, seed=3) if bias else None
    c0 = _rand(M, N, seed=4)
    ref = a.double() @ b.double() + (bi.double() if bias else 0) + (c0.double() if acc else 0)
    out = c0.clone().cuda() if acc else None
    got = engine.gemm_nn(a.cuda(), b.cuda(), bi.cuda() if bias else None, out=out, accumulate=acc)
    torch.cuda.synchronize()
    assert rel_l2(got.cpu().numpy(), ref.numpy()) < 2e-6


def test_gemm_nn_strided_views():
    from bsed_b200 import engine
    big = _rand(400, 768, seed=5).cuda()
    a = big[:, 384:384 + 128]          # lda = 768
    b = _rand(128, 64, seed=6).cuda()
    got = engine.gemm_nn(a, b)
    ref = a.double().cpu() @ b.double().cpu()
    assert rel_l2(got.cpu().numpy(), ref.numpy()) < 2e-6


@pytest.mark.parametrize("K,M,N", [(5000, 48, 256), (3000, 128, 128), (999, 16, 32), (70000, 64, 64),
                                   (17, 32, 16), (4000, 384, 128), (2500, 128, 16)])
def test_gemm_tn(K, M, N):
    from bsed_b200 import engine
    a, b = _rand(K, M, seed=7), _rand(K, N, seed=8)
    c0 = _rand(M, N, seed=9)
    ref = c0.double() + a.double().T @ b.double()
    got = engine.gemm_tn(a.cuda(), b.cuda(), c0.clone().cuda())
    torch.cuda.synchronize()
    assert rel_l2(got.cpu().numpy(), ref.numpy()) < 5e-6


@pytest.mark.parametrize("B,T,Fq,Cin,Cout", [(2, 37, 16, 16, 32), (1, 20, 8, 64, 128), (3, 11, 2, 128, 128),
                                             (2, 9, 4, 32, 16), (1, 313, 1, 128, 128), (2, 5, 64, 16, 16),
                                             (1, 627, 64, 16, 32)])
def test_conv3x3_channels_last(B, T, Fq, Cin, Cout):
    from bsed_b200 import engine
    x = _rand(B, Cin, T, Fq, seed=10)
    w = _rand(Cout, Cin, 3, 3, seed=11, scale=0.2)
    b = _rand(Cout, seed=12)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)             # NCHW
    got = engine.conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), b.cuda())
    torch.cuda.synchronize()
    assert got.shape == (B, T, Fq, Cout)
    assert rel_l2(got.permute(0, 3, 1, 2).cpu().numpy(), ref.numpy()) < 3e-6


def test_opt_adam_matches_torch():
    from bsed_b200 import engine
    n = 10007
    p0, ema0 = _rand(n, seed=20), _rand(n, seed=21)
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pt], lr=5e-4, betas=(0.9, 0.999), eps=1e-8)
    p, ema = p0.clone().cuda(), ema0.clone().cuda()
    m, v = torch.zeros(n).cuda(), torch.zeros(n).cuda()
    ema_ref = ema0.clone()
    for step in range(1, 6):
        g = _rand(n, seed=30 + step, scale=10.0 ** (step - 4))
        pt.grad = g.clone()
        opt.step()
        a = min(1 - 1 / (step + 1), 0.999)
        ema_ref = ema_ref * a + pt.detach() * (1. - a)
        engine.opt_ema_step(p, g.cuda(), m, v, ema, step=step, ema_step=step, lr=5e-4)
    torch.cuda.synchronize()
    assert max_abs(p.cpu().numpy(), pt.detach().numpy()) < 2e-7
    assert max_abs(ema.cpu().numpy(), ema_ref.numpy()) < 2e-7


def test_opt_sgd_nesterov_matches_torch():
    from bsed_b200 import engine
    n = 4099
    p0 = _rand(n, seed=40)
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([pt], lr=1e-2, momentum=0.9, weight_decay=1e-4, nesterov=True)
    p, m = p0.clone().cuda(), torch.zeros(n).cuda()
    for step in range(1, 5):
        g = _rand(n, seed=50 + step)
        pt.grad = g.clone()
        opt.step()
        engine.opt_ema_step(p, g.cuda(), m, None, None, step=step, kind="sgd", lr=1e-2, weight_decay=1e-4,
                            momentum=0.9)
    torch.cuda.synchronize()
    assert max_abs(p.cpu().numpy(), pt.detach().numpy()) < 1e-6


def test_ema_buffers_and_counters():
    from bsed_b200 import engine
    bn, ebn = _rand(1248, seed=60), _rand(1248, seed=61)
    nbt = torch.tensor([5, 5, 5, 5, 5, 5, 5], dtype=torch.int64)
    enbt = torch.tensor([3, 3, 3, 3, 3, 3, 3], dtype=torch.int64)
    a = min(1 - 1 / (7 + 1), 0.999)
    ref = ebn * a + bn * (1. - a)
    ref_n = (enbt * a + nbt * (1. - a)).to(torch.int64)       # what load_state_dict does to the float result
    ebn_d, enbt_d = ebn.clone().cuda(), enbt.clone().cuda()
    engine.ema_buffers(bn.cuda(), ebn_d, nbt.cuda(), enbt_d, ema_step=7)
    torch.cuda.synchronize()
    assert torch.equal(ebn_d.cpu(), ref)
    assert torch.equal(enbt_d.cpu(), ref_n)


def test_mt_loss_matches_torch():
    from bsed_b200 import engine
    rng = np.random.default_rng(70)
    B, T, C = 6, 313, 20
    strong = torch.from_numpy(rng.random((B, T, C))).float()
    strong[0, 0, 0] = 1.0           # log clamp at -100
    strong[0, 0, 1] = 0.0
    weak = torch.from_numpy(rng.random((B, C))).float()
    tgt = torch.from_numpy((rng.random((3, T, C)) < 0.1).astype(np.float32))
    tgt[0, 0, 1] = 1.0
    s_ema = torch.from_numpy(rng.random((2, T, C))).float()
    w_ema = torch.from_numpy(rng.random((2, C))).float()
    cons_w = 0.37
    st, wk = strong.clone().requires_grad_(), weak.clone().requires_grad_()
    bce, mse = torch.nn.BCELoss(), torch.nn.MSELoss()
    parts = [bce(st[0:3], tgt), bce(wk[0:3], tgt.max(-2)[0]), cons_w * mse(st[3:5], s_ema), cons_w * mse(wk[3:5], w_ema)]
    sum(parts).backward()
    losses, ds, dw = engine.mt_loss(strong.cuda(), weak.cuda(), 0, 3, tgt.cuda(), 3, 2, s_ema.cuda(), w_ema.cuda(), cons_w)
    torch.cuda.synchronize()
    for i in range(4):
        assert float(losses[i]) == pytest.approx(float(parts[i]), rel=2e-5)
    assert rel_l2(ds.cpu().numpy(), st.grad.numpy()) < 1e-5
    assert rel_l2(dw.cpu().numpy(), wk.grad.numpy()) < 1e-5
    assert float(ds[5].abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------
# tcgen05 (kind::tf32) + TMA kernels: same contracts, tf32 input rounding (10-bit mantissa)
# ---------------------------------------------------------------------------------------------
TF32_REL = 2e-3


@pytest.mark.parametrize("B,T,Fq,Cin,Cout", [(2, 37, 16, 32, 32), (1, 20, 8, 64, 128), (3, 11, 2, 128, 128),
                                             (2, 9, 4, 32, 16), (1, 313, 1, 128, 128), (2, 5, 64, 16, 16),
                                             (1, 627, 64, 16, 32), (2, 313, 32, 32, 64), (2, 40, 128, 16, 32),
                                             # column-tiled halo kernel (F >= 8, Cin % 32 == 0): partial t blocks,
                                             # every output width, weights resident / ringed
                                             (2, 313, 16, 64, 128), (1, 130, 8, 128, 128), (3, 129, 64, 32, 16),
                                             (1, 257, 32, 64, 32), (2, 128, 8, 128, 64), (1, 1, 8, 32, 32),
                                             # 128 output channels with enough tiles for two bins per step (two issuer warps)
                                             (8, 313, 8, 128, 128), (25, 313, 2, 128, 128)])
def test_conv3x3_tensor_cores(B, T, Fq, Cin, Cout):
    from bsed_b200 import engine
    x = _rand(B, Cin, T, Fq, seed=10)
    w = _rand(Cout, Cin, 3, 3, seed=11, scale=0.2)
    b = _rand(Cout, seed=12)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    got = engine.conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), b.cuda(), tensor_cores=True)
    torch.cuda.synchronize()
    assert rel_l2(got.permute(0, 3, 1, 2).cpu().numpy(), ref.numpy()) < TF32_REL


@pytest.mark.parametrize("M,K,N,bias,acc", [(1000, 128, 128, False, False), (300, 16, 16, True, False),
                                            (777, 256, 64, True, False), (513, 768, 128, False, True),
                                            (129, 64, 64, True, True), (5, 32, 32, False, False),
                                            (40000, 32, 32, True, False),
                                            # N > 128: column blocks of 128 inside one launch
                                            (2000, 256, 768, True, False), (900, 768, 256, False, True)])
def test_gemm_nt_tensor_cores(M, K, N, bias, acc):
    from bsed_b200 import engine
    a, bk = _rand(M, K, seed=1), _rand(N, K, seed=2)
    bi = _rand(N, seed=3) if bias else None
    c0 = _rand(M, N, seed=4)
    ref = a.double() @ bk.double().T + (bi.double() if bias else 0) + (c0.double() if acc else 0)
    out = c0.clone().cuda() if acc else None
    got = engine.gemm_nt_tc(a.cuda(), bk.cuda(), bi.cuda() if bias else None, out=out, accumulate=acc)
    torch.cuda.synchronize()
    assert rel_l2(got.cpu().numpy(), ref.numpy()) < TF32_REL


@pytest.mark.parametrize("tensor_cores", [False, True])
@pytest.mark.parametrize("B,T,Fq,Cin,Cout", [(2, 37, 16, 32, 32), (1, 20, 8, 64, 128), (3, 11, 2, 128, 128),
                                             (2, 313, 1, 128, 128), (2, 50, 64, 16, 32), (1, 313, 32, 32, 64),
                                             (2, 100, 4, 128, 128), (3, 21, 8, 16, 64), (1, 9, 128, 16, 32)])
def test_conv3x3_weight_gradient(B, T, Fq, Cin, Cout, tensor_cores):
    from bsed_b200 import engine
    x = _rand(B, Cin, T, Fq, seed=20).double().requires_grad_(False)
    dy = _rand(B, Cout, T, Fq, seed=21).double()
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x, w, None, padding=1).backward(dy)
    got = engine.conv3x3_wgrad(x.float().permute(0, 2, 3, 1).contiguous().cuda(),
                               dy.float().permute(0, 2, 3, 1).contiguous().cuda(), tensor_cores=tensor_cores)
    torch.cuda.synchronize()
    assert rel_l2(got.cpu().numpy(), w.grad.numpy()) < (TF32_REL if tensor_cores else 1e-5)
