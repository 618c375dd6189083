import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import bsed_models, golden, max_abs, oracle_models, rel_l2
from bsed_b200.utilities import synth
from bsed_b200 import engine
from oracle import crnn as ocrnn

def inference_flag():
    x = torch.from_numpy(synth.make_logmel_like(2, seed=51))
    oc, op = oracle_models(seed=10, linear_std=0.5)
    m, p = bsed_models(oc, op); m.eval(); p.eval()
    with torch.no_grad():
        enc, _ = m(x.cuda()); s0, w0 = p(enc); s1, w1 = p(enc, inference=True)
        e2, _ = oc(x); s2, w2 = op(e2, inference=True); s3, w3 = op(e2)
    print("weak mine", w1[0].cpu().numpy().round(3)); print("weak orac", w2[0].numpy().round(3))
    print("strong noninf diff", max_abs(s0.cpu().numpy(), s3.numpy()), "enc diff", max_abs(enc.cpu().numpy(), e2.numpy()))
    print("gated mine max per class", s1[0].cpu().numpy().max(0).round(3)); print("gated orac", s2[0].numpy().max(0).round(3))

def mt_drop():
    from bsed_b200.main import MeanTeacherTrainer
    import test_gpu_train as T
    for name, pd in (("mt_step_nodrop.npz", 0.0), ("mt_step_drop.npz", 0.5)):
        g = golden(name)
        m, p, em, ep = T._models(pd)
        xs, xr, xr_ema, ts = [t.cuda() for t in T._inputs()]
        tr = MeanTeacherTrainer(m, p, em, ep, lr=5e-4, n_syn=2, n_real=2, dropout_seed=2023)
        l = tr.step(xr, xr_ema, xs, ts, global_step=100, rampup_length=500)
        print(name, "losses", [float(v) for v in l.cpu()], "ref", [float(g[k + "0"]) for k in ("strong_class", "weak_class", "cons_strong", "cons_weak")])
        print(" strong real diff", max_abs(tr.last["strong"][2:].cpu().numpy(), g["strong0"]), "weak", max_abs(tr.last["weak"][2:].cpu().numpy(), g["weak0"]))
        o = 0
        for (mod, pname, shape), (fullname, _) in zip(m._param_specs + p._param_specs, list(m.named_parameters()) + list(p.named_parameters())):
            k = int(np.prod(shape)); got = tr.grads[o:o + k].cpu().numpy(); o += k
            key = ("g_crnn." + fullname.replace("cnn.", "cnn.cnn.", 1)) if o <= tr.n_crnn else "g_pred." + fullname
            ref = g[key]; gs = got if got.size <= 4096 else got[:: max(1, got.size // 4096)][:4096]
            print("  %-34s rel %.2e  gn %.3e" % (fullname, rel_l2(gs, ref), float(g[key.replace("g_", "gn_", 1)])))
        ssd = m.state_dict()
        for k in ("cnn.conv0.weight", "cnn.glu4.linear.weight", "rnn.rnn.weight_hh_l0"):
            d = np.abs(ssd[k].cpu().numpy().reshape(-1)[:2048] - g["s_" + k]); print("  param", k, "max", d.max(), "mean", d.mean())

inference_flag(); mt_drop()
