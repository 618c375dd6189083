"""Net_resnet inference timing (CUDA events): python tests/bench_resnet.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsed_b200.models.ResNet import Net_resnet  # noqa: E402
from bsed_b200.utilities import synth  # noqa: E402

for prec in ("tf32", "fp32"):
    m = Net_resnet(pretrained=False, precision=prec).cuda().eval()
    x = torch.from_numpy(synth.make_logmel_like(24, seed=1)).cuda()
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"Net_resnet eval, 24 clips, {prec}: {ms:.2f} ms -> {24e3 / ms:.0f} clips/s ({24 * 11.0 / ms:.1f} TFLOP/s algorithmic)")

# training step (src/audio_tagging_system_cnn.py:340-406 shapes: 12 synthetic + 12 weak/unlabeled clips)
from bsed_b200.models.ResNet import TaggerTrainer  # noqa: E402

m = Net_resnet(pretrained=False, precision="tf32").cuda().train()
tr = TaggerTrainer(m, lr=1e-3)
xs = torch.from_numpy(synth.make_logmel_like(12, seed=2)).cuda()
xr = torch.from_numpy(synth.make_logmel_like(12, seed=3)).cuda()
ts = torch.from_numpy(synth.make_targets(12, seed=4)).cuda()
tw = (torch.from_numpy(synth.make_targets(12, seed=5)).max(-2)[0] > 0).float().cuda()
for _ in range(2):
    loss = tr.step(xs, ts, xr, tw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    loss = tr.step(xs, ts, xr, tw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"Net_resnet training step, 12 + 12 clips (tf32 forward GEMMs, {m.backward_precision or "tf32"} backward GEMMs): {ms:.1f} ms -> {24e3 / ms:.0f} clips/s; "
      f"loss {float(loss):.4f}; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
