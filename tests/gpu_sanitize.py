"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): one tensor-core mean-teacher step on 2+2 clips, the
frontend, the decoder and the discriminator.   compute-sanitizer --tool memcheck python tests/gpu_sanitize.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["BSED_PRECISION"] = os.environ.get("BSED_SANITIZE_PRECISION", "tf32")
import test_gpu_train as T  # noqa: E402
from bsed_b200 import engine  # noqa: E402
from bsed_b200.main import MeanTeacherTrainer  # noqa: E402
from bsed_b200.utilities import synth  # noqa: E402

m, p, em, ep = T._models(0.5)
xs, xr, xr_ema, ts = [t.cuda() for t in T._inputs()]
tr = MeanTeacherTrainer(m, p, em, ep, lr=5e-4, n_syn=2, n_real=2, dropout_seed=2023)
loss = tr.step(xr, xr_ema, xs, ts, global_step=100, rampup_length=500)
torch.cuda.synchronize()
print("precision", tr.plan.precision, "losses", loss.cpu())
clips = torch.from_numpy(synth.make_clips(2, seed=1)).cuda()
lm = engine.amp_to_db(engine.melspec(clips), 1255)
ev, n = engine.median_decode(tr.last["strong"].detach(), 0.5, 14)
from bsed_b200.DA.cdan_frame import ConditionalDomainAdversarialLoss  # noqa: E402
from bsed_b200.models.CRNN import Clip_Discriminator  # noqa: E402
d = Clip_Discriminator(256).cuda().train()
f = torch.randn(3, 313, 256, device="cuda").tanh().requires_grad_(True)
ConditionalDomainAdversarialLoss(d)(None, f[:2], None, f[2:]).backward()
torch.cuda.synchronize()
print("ok", float(lm.mean()), int(n.sum()), float(f.grad.abs().sum()))
