import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_train as T
from bsed_b200.main import MeanTeacherTrainer
m, p, em, ep = T._models(0.5)
xs, xr, xr_ema, ts = [t.cuda() for t in T._inputs()]
tr = MeanTeacherTrainer(m, p, em, ep, lr=5e-4, n_syn=2, n_real=2, dropout_seed=2023)
l = tr.step(xr, xr_ema, xs, ts, global_step=100, rampup_length=500)
torch.cuda.synchronize()
print("losses", l.cpu())
