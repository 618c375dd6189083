"""Diagnostics for a GPU run: prints the parity numbers the tests assert on (and timings), so one
gpurun call gives the full picture even when assertions fail.  Writes gpurun_out/report.txt."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import bsed_models, golden, logmel_close, max_abs, oracle_models, rel_l2  # noqa: E402
from oracle import frontend as ofe  # noqa: E402
from bsed_b200 import engine  # noqa: E402
from bsed_b200.utilities import synth  # noqa: E402

out = []


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    out.append(s)


def section(fn):
    say("==", fn.__name__)
    try:
        fn()
    except Exception:
        say("EXCEPTION", traceback.format_exc())


def frontend():
    clips = synth.make_clips(4, seed=2023)
    clips[1:4] = synth.make_clips(3, seed=1, edge_cases=True)
    a = torch.from_numpy(clips).cuda()
    mel = engine.melspec(a)
    db = engine.amp_to_db(mel, 1255)
    torch.cuda.synchronize()
    for i, name in enumerate(["chirps", "zeros", "impulse", "sine"]):
        ref_mel = ofe.preprocess(clips[i])
        ref_db = ofe.transform(ref_mel)[0]
        w, f = logmel_close(db[i].cpu().numpy(), ref_db)
        say(name, "mel rel_l2", rel_l2(mel[i].cpu().numpy(), ref_mel), "dB worst", w, "frac>1e-4", f,
            "max abs dB", max_abs(db[i].cpu().numpy(), ref_db))
    big = torch.from_numpy(synth.make_clips(8, seed=5)).cuda().repeat(32, 1)
    for _ in range(3):
        engine.melspec(big)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(10):
        m = engine.melspec(big)
    torch.cuda.synchronize()
    dt = (time.time() - t) / 10
    say("melspec 256 clips: %.3f ms -> %.0f clips/s, %.1f GB/s algorithmic" % (dt * 1e3, 256 / dt, 256 * 1922560 / dt / 1e9))


def crnn_eval():
    g = golden("crnn_eval.npz")
    x = torch.from_numpy(synth.make_logmel_like(2, seed=11))
    oc, op = oracle_models(seed=5, linear_std=0.2)
    m, p = bsed_models(oc, op, dropout=0.5)
    m.eval(); p.eval()
    with torch.no_grad():
        enc, _ = m(x.cuda())
        strong, weak = p(enc)
    say("enc err", max_abs(enc.cpu().numpy()[:, ::8], g["enc"]), "strong err", max_abs(strong.cpu().numpy(), g["strong"]),
        "weak err", max_abs(weak.cpu().numpy(), g["weak"]))
    flat, bn, nbt = m.flat_tensors()
    plan = engine.Plan(engine.make_cfg(**m.cfg_kwargs), max_clips=2, device="cuda")
    plan.forward([dict(params=flat, bn=bn, nbt=nbt, n=2)], x.cuda(), train=False, save=False)
    h = x
    with torch.no_grad():
        for i in range(7):
            for name in (f"conv{i}", f"batchnorm{i}", f"glu{i}", f"dropout{i}", f"pooling{i}"):
                h = getattr(oc.cnn, name)(h)
            ref = h.permute(0, 2, 3, 1).contiguous().numpy()
            got = plan.debug_tensor(f"pool{i}").cpu().numpy().reshape(ref.shape)
            say("  block", i, "rel_l2", rel_l2(got, ref))


def train_step_timing():
    from bsed_b200.main import MeanTeacherTrainer
    oc, op = oracle_models(seed=5, linear_std=0.2)
    tc, tp = oracle_models(seed=6, linear_std=0.2)
    m, p = bsed_models(oc, op, dropout=0.5)
    em, ep = bsed_models(tc, tp, dropout=0.5)
    for mod in (m, p, em, ep):
        mod.train()
    tr = MeanTeacherTrainer(m, p, em, ep, n_syn=12, n_real=12)
    x = torch.from_numpy(synth.make_logmel_like(12, seed=1)).cuda()
    ts = torch.from_numpy(synth.make_targets(12, seed=2)).cuda()
    for i in range(3):
        l = tr.step(x, x, x, ts, i, 500)
    torch.cuda.synchronize()
    t = time.time()
    n = 5
    for i in range(n):
        l = tr.step(x, x, x, ts, 3 + i, 500)
    torch.cuda.synchronize()
    dt = (time.time() - t) / n
    say("mean-teacher step (24 student + 12 teacher clips): %.2f ms -> %.0f clips/s; losses %s" %
        (dt * 1e3, 24 / dt, [round(float(v), 4) for v in l.cpu()]))
    say("workspace GB", tr.plan.ws_bytes / 1e9)


def ada_step_timing():
    from bsed_b200 import main as bmain
    from bsed_b200.DA.cdan_frame import ConditionalDomainAdversarialLoss
    from bsed_b200.models.CRNN import Clip_Discriminator
    oc, op = oracle_models(seed=5, linear_std=0.2)
    m, p = bsed_models(oc, op, dropout=0.5)
    m.train(); p.train()
    d = Clip_Discriminator(256).cuda().train()
    crit = ConditionalDomainAdversarialLoss(d)
    crit.grl.iter_num = 500
    opt_c = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True)
    opt_d = torch.optim.SGD(d.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True)
    x = torch.from_numpy(synth.make_logmel_like(12, seed=1)).cuda()
    xs = torch.from_numpy(synth.make_logmel_like(12, seed=2)).cuda()
    for _ in range(3):
        bmain.adversarial_step(m, p, crit, opt_c, opt_d, x, xs)
    torch.cuda.synchronize()
    t = time.time()
    n = 5
    for _ in range(n):
        l = bmain.adversarial_step(m, p, crit, opt_c, opt_d, x, xs)
    torch.cuda.synchronize()
    dt = (time.time() - t) / n
    # discriminator alone
    f = torch.randn(24, 313, 256, device="cuda").tanh().requires_grad_(True)
    for _ in range(3):
        crit(None, f[:12], None, f[12:]).backward()
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(n):
        crit(None, f[:12], None, f[12:]).backward()
    torch.cuda.synchronize()
    dd = (time.time() - t) / n
    say("adversarial step (12 syn + 12 real clips, student fwd+bwd, D fwd+bwd, 2 x SGD): %.2f ms; discriminator fwd+bwd alone "
        "(24 clips): %.2f ms; domain loss %.4f" % (dt * 1e3, dd * 1e3, float(l)))


if __name__ == "__main__":
    say(torch.cuda.get_device_name(0))
    for fn in (frontend, crnn_eval, train_step_timing, ada_step_timing):
        section(fn)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "report.txt"), "w") as fh:
        fh.write("\n".join(out) + "\n")
