import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_train as T
from helpers import golden, rel_l2
from bsed_b200.main import MeanTeacherTrainer

def names(m, p):
    out = []; o = 0
    for (mod, pname, shape), (fullname, _) in zip(m._param_specs + p._param_specs, list(m.named_parameters()) + list(p.named_parameters())):
        k = int(np.prod(shape)); out.append((fullname, o, k)); o += k
    return out

def run(poison, pd=0.0):
    g = golden("mt_step_nodrop.npz" if pd == 0 else "mt_step_drop.npz")
    m, p, em, ep = T._models(pd)
    xs, xr, xr_ema, ts = [t.cuda() for t in T._inputs()]
    tr = MeanTeacherTrainer(m, p, em, ep, lr=5e-4, n_syn=2, n_real=2, dropout_seed=2023)
    if poison is not None:
        tr.plan.ws.fill_(poison)
    tr.step(xr, xr_ema, xs, ts, global_step=100, rampup_length=500)
    torch.cuda.synchronize()
    gr = tr.grads.clone()
    res = []
    for fullname, o, k in names(m, p):
        got = gr[o:o + k].cpu().numpy()
        key = ("g_crnn." + fullname.replace("cnn.", "cnn.cnn.", 1)) if o + k <= tr.n_crnn else "g_pred." + fullname
        ref = g[key]; gs = got if got.size <= 4096 else got[:: max(1, got.size // 4096)][:4096]
        res.append((fullname, rel_l2(gs, ref), bool(np.isnan(got).any())))
    return gr, res

for poison in (0, 0xFF, 0x3F):
    gr, res = run(poison)
    print("poison", hex(poison), "worst", max(r[1] for r in res if "conv" not in r[0] or "bias" not in r[0]), "nan tensors", [r[0] for r in res if r[2]][:12])
    print("   ", [(r[0], "%.1e" % r[1]) for r in res][:8], "...", [(r[0], "%.1e" % r[1]) for r in res][-8:])
g1, _ = run(0); g2, _ = run(0)
print("determinism: max abs diff between two identical runs", float((g1 - g2).abs().max()), "rel", float((g1 - g2).norm() / g1.norm()))
