"""Fused data-parallel step kernel (reduce-scatter over peer memory + optimiser + EMA + all-gather) in its one-rank form: with world = 1 the
kernel runs the same arrive / reduce / depart protocol against its own buffers and must equal bsed_opt_ema_step bit for
bit.  The multi-GPU form is checked by tests/dp_fused_check.py under torchrun on a multi-GPU box."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["adam", "sgd"])
def test_world1_equals_plain_optimizer_step(kind):
    from bsed_b200 import _lib, engine
    from bsed_b200._lib import OptCfg, check, ptr, stream_ptr
    lib = _lib.load()
    h = _lib.handle(0)
    n = 1117560
    g = torch.Generator(device="cuda").manual_seed(0)
    p0, e0 = torch.randn(n, device="cuda", generator=g), torch.randn(n, device="cuda", generator=g)
    grads = torch.empty(n, device="cuda")
    flags = torch.zeros(64, dtype=torch.int32, device="cuda")
    pa, ma, va, ea = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0), e0.clone()
    pb, mb, vb, eb = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0), e0.clone()
    pg, pf = (C.c_void_p * 1)(grads.data_ptr()), (C.c_void_p * 1)(flags.data_ptr())
    pp, pe = (C.c_void_p * 1)(pa.data_ptr()), (C.c_void_p * 1)(ea.data_ptr())
    for step in range(1, 5):
        grads.copy_(torch.randn(n, device="cuda", generator=g) * 0.01)
        cfg = OptCfg()
        cfg.kind = 0 if kind == "adam" else 1
        cfg.lr, cfg.beta1, cfg.beta2, cfg.eps = 5e-4, 0.9, 0.999, 1e-8
        cfg.weight_decay, cfg.momentum, cfg.grad_scale, cfg.ema_alpha = (0.0 if kind == "adam" else 1e-4), 0.9, 1.0, 0.999
        cfg.step, cfg.ema_step = step, step
        check(lib.bsed_dp_opt_ema_step(h, 0, 1, pg, pp, pe, pf, step, ptr(ma), ptr(va), n, C.byref(cfg), stream_ptr()),
              "bsed_dp_opt_ema_step")
        engine.opt_ema_step(pb, grads, mb, vb, eb, step=step, ema_step=step, kind=kind, lr=5e-4,
                            weight_decay=(0.0 if kind == "adam" else 1e-4), grad_scale=1.0)
    torch.cuda.synchronize()
    assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(ea, eb)
    if kind == "adam":
        assert torch.equal(va, vb)
    f = flags.cpu()
    assert int(f[0]) == 4 and int(f[16]) == 4 and int(f[32]) == 0 and int(f[33]) == 0     # arrive, depart, counter, error


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one box (run by hand with gpurun --gpus 2)")
def test_two_rank_fused_step_equals_nccl_plus_optimizer():
    """tests/dp_fused_check.py under torchrun: every rank holds different gradients; the fused kernel must give the same
    parameters / EMA as NCCL all-reduce + bsed_opt_ema_step and bit-identical replicas."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "dp_fused_check.py")],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "replicas bit-identical: True" in r.stdout or "FUSED-DP UNAVAILABLE" in r.stdout, r.stdout[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one box (run by hand with gpurun --gpus 2)")
def test_two_rank_training_steps_keep_replicas_identical():
    """tests/dp_train_check.py under torchrun: mean-teacher step (fused exchange inside the CUDA graph, and the NCCL path)
    and the config-3 adaptation step keep the replicas bit-identical."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29300 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "dp_train_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DP-TRAIN-CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
