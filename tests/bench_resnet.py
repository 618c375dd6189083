"""Net_resnet (config 4, src/audio_tagging_system_cnn.py:50-64) timing with CUDA events, next to the reference's own
arithmetic on the same B200: torchvision's resnet18 in PyTorch eager mode (stock cuDNN settings), the comparison SURVEY
section 8d asks for.        python tests/bench_resnet.py            (prints one JSON line at the end)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsed_b200.models.ResNet import Net_resnet  # noqa: E402
from bsed_b200.utilities import synth  # noqa: E402

report = {}
for prec in ("tf32x3", "tf32", "fp32"):
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
    report[f"eval_{prec}_ms_per_24_clips"] = ms

# training step (src/audio_tagging_system_cnn.py:340-406 shapes: 12 synthetic + 12 weak/unlabeled clips)
from bsed_b200.models.ResNet import TaggerTrainer  # noqa: E402

m = Net_resnet(pretrained=False, precision=None).cuda().train()          # library default precision
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
print(f"Net_resnet training step, 12 + 12 clips (default precision): {ms:.1f} ms -> {24e3 / ms:.0f} clips/s; "
      f"loss {float(loss):.4f}; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
report["train_step_ms_12+12_clips"] = ms

# ---- the reference's arithmetic: torchvision resnet18 (oracle/resnet.py restates the eight lines of the class), eager
from oracle import resnet as ores  # noqa: E402

om = ores.OracleNetResnet().cuda().eval()
x = torch.from_numpy(synth.make_logmel_like(24, seed=1)).cuda()
with torch.no_grad():
    for _ in range(3):
        om(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        om(x)
    e1.record()
    torch.cuda.synchronize()
report["torchvision_eager_eval_ms_per_24_clips"] = e0.elapsed_time(e1) / 10
om.train()
opt = torch.optim.Adam(om.parameters(), lr=1e-3)


def tv_step():
    opt.zero_grad()
    l, _ = ores.tagger_step_loss(om, xs, ts, xr, tw)
    l.backward()
    opt.step()
    return l.item()


for _ in range(3):
    tv_step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    tv_step()
e1.record()
torch.cuda.synchronize()
report["torchvision_eager_train_step_ms_12+12_clips"] = e0.elapsed_time(e1) / 10
report["cudnn_allow_tf32"] = bool(torch.backends.cudnn.allow_tf32)
print(f"torchvision resnet18 eager on the same GPU: eval {report['torchvision_eager_eval_ms_per_24_clips']:.2f} ms per 24 clips, "
      f"training step {report['torchvision_eager_train_step_ms_12+12_clips']:.1f} ms")
print(json.dumps(report))
