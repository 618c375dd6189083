"""Where does the host time of a graph-replayed step go?   python tests/graph_probe.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsed_b200 import engine  # noqa: E402
from bsed_b200.main import MeanTeacherTrainer  # noqa: E402
from bsed_b200.models import CRNN, Predictor  # noqa: E402
from bsed_b200.utilities import synth  # noqa: E402
from bsed_b200.utilities.utils import weights_init  # noqa: E402

dev = torch.device("cuda", 0)


def make():
    m, p = CRNN(**engine.REFERENCE_CRNN_KWARGS), Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
    weights_init(m)
    weights_init(p)
    return m.to(dev).train(), p.to(dev).train()


m, p = make()
em, ep = make()
tr = MeanTeacherTrainer(m, p, em, ep, lr=5e-4, n_syn=12, n_real=12)
x = torch.from_numpy(synth.make_logmel_like(12, seed=1)).to(dev)
xs = torch.from_numpy(synth.make_logmel_like(12, seed=2)).to(dev)
ts = torch.from_numpy(synth.make_targets(12, seed=3)).to(dev)
for i in range(4):
    tr.step(x, x, xs, ts, i, 5000)
torch.cuda.synchronize()
g, out = next(iter(tr._graphs.values()))
for name, fn in (("step()", lambda i: tr.step(x, x, xs, ts, 4 + i, 5000)), ("g.replay()", lambda i: g.replay()),
                 ("_assert_homed()", lambda i: tr._assert_homed()), ("input copies", lambda i: tr.x[:12].copy_(xs))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(20):
        fn(i)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:18s} host {1e3 * (t1 - t0) / 20:.3f} ms per call; drained after {1e3 * (t2 - t1):.1f} ms more")
# spaced launches: does the launch call itself block while the previous replay is running?
for gap in (0.0, 0.010):
    ts_ = []
    for i in range(10):
        torch.cuda.synchronize() if gap else None
        t0 = time.perf_counter()
        g.replay()
        ts_.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    print(f"replay host time, {'idle GPU' if gap else 'back to back'}: " + " ".join(f"{1e3 * t:.2f}" for t in ts_))
