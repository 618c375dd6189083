import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_train as T
from bsed_b200.main import MeanTeacherTrainer

def run():
    m, p, em, ep = T._models(0.0)
    xs, xr, xr_ema, ts = [t.cuda() for t in T._inputs()]
    tr = MeanTeacherTrainer(m, p, em, ep, lr=5e-4, n_syn=2, n_real=2, dropout_seed=2023)
    tr.plan.ws.zero_()
    tr.step(xr, xr_ema, xs, ts, global_step=100, rampup_length=500)
    torch.cuda.synchronize()
    taps = {k: tr.plan.debug_tensor(k).clone() for k in ("enc", "xg", "saved0", "saved1", "denc", "dx1", "dxg", "dgh", "dpool0", "dpool1", "dxn", "gru0", "gru1", "pool6", "xhat6")}
    taps["d_enc_in"] = tr.d_enc.clone(); taps["grads"] = tr.grads.clone()
    return taps

a = run(); b = run()
for k in a:
    d = (a[k] - b[k]).abs().max().item(); n = a[k].abs().max().item()
    print("%-10s maxdiff %.3e (max |v| %.3e) nan %s" % (k, d, n, bool(torch.isnan(a[k]).any())))
