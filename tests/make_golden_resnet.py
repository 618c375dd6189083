"""Generate tests/golden/resnet_eval.npz with torchvision's resnet18 assembled as the reference's Net_resnet
(src/audio_tagging_system_cnn.py:50-64) in the build container:   python tests/make_golden_resnet.py
Weights and inputs are regenerated from seeds by the tests; the fixture holds the outputs and stage checksums."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import resnet as ores  # noqa: E402
from bsed_b200.utilities import synth  # noqa: E402


def main():
    torch.set_num_threads(8)
    m = ores.seeded_init(ores.OracleNetResnet(20), seed=17).eval()
    x = torch.from_numpy(synth.make_logmel_like(3, seed=51))
    with torch.no_grad():
        r = m.resnet
        h = r.maxpool(r.relu(r.bn1(r.conv1(x))))
        l1 = r.layer1(h)
        l2 = r.layer2(l1)
        l4 = r.layer4(r.layer3(l2))
        out = m(x)
    sd = m.state_dict()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resnet_eval.npz"), out=out.numpy(),
                        x_sum=float(x.double().sum()), stem_sum=float(h.double().sum()), l1_sum=float(l1.double().sum()),
                        l2_sum=float(l2.double().sum()), l4_sum=float(l4.double().sum()), l4_abs=float(l4.double().abs().sum()),
                        keys=np.array(list(sd.keys())), shapes=np.array([str(tuple(v.shape)) for v in sd.values()]),
                        n_params=sum(p.numel() for p in m.parameters()))
    print("resnet golden written; out range", float(out.min()), float(out.max()), "shape", tuple(out.shape))


def train_fixture():
    """One iteration of the tagger's train_mt (src/audio_tagging_system_cnn.py:340-406) with the torchvision-based model:
    two model calls in train mode, BCE on the weak outputs, backward, Adam(lr 1e-3)."""
    torch.set_num_threads(8)
    m = ores.seeded_init(ores.OracleNetResnet(20), seed=17).train()
    xs = torch.from_numpy(synth.make_logmel_like(2, seed=61))
    xr = torch.from_numpy(synth.make_logmel_like(2, seed=62))
    ts = torch.from_numpy(synth.make_targets(2, seed=63))
    tw = (torch.from_numpy(synth.make_targets(2, seed=64)).max(-2)[0] > 0).float()
    loss, grads = ores.tagger_step_loss(m, xs, ts, xr, tw)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt.zero_grad()
    loss.backward()
    rec = {"loss": float(loss)}
    for n, p in m.named_parameters():
        g = p.grad.detach().numpy().reshape(-1)
        rec["g_" + n] = g if g.size <= 4096 else g[:: max(1, g.size // 4096)][:4096]
        rec["gn_" + n] = float(np.sqrt((g.astype(np.float64) ** 2).sum()))
    opt.step()
    sd = m.state_dict()
    for k in ("resnet.conv1.weight", "resnet.layer1.0.bn1.weight", "resnet.layer2.0.downsample.0.weight", "resnet.layer4.1.conv2.weight",
              "resnet.fc.bias", "resnet.bn1.running_mean", "resnet.layer3.0.downsample.1.running_var", "resnet.layer4.1.bn2.running_var"):
        rec["s_" + k] = sd[k].numpy().reshape(-1)[:2048]
    rec["nbt"] = int(sd["resnet.bn1.num_batches_tracked"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resnet_train.npz"), **rec)
    print("resnet train fixture: loss", rec["loss"], "nbt", rec["nbt"])


def train_fixture8():
    """The same iteration on 4 + 4 clips (better conditioned: four times the rows behind every train-mode BatchNorm), with
    the float64 run of the same model next to torch's fp32 one: `g64_*` is the yardstick, `dev32` records how far torch's
    own fp32 gradients are from it, tensor by tensor (worst and median printed)."""
    torch.set_num_threads(8)
    xs = torch.from_numpy(synth.make_logmel_like(4, seed=81))
    xr = torch.from_numpy(synth.make_logmel_like(4, seed=82))
    ts = torch.from_numpy(synth.make_targets(4, seed=83))
    tw = (torch.from_numpy(synth.make_targets(4, seed=84)).max(-2)[0] > 0).float()
    rec, grads = {}, {}
    for dt in (torch.float32, torch.float64):
        m = ores.seeded_init(ores.OracleNetResnet(20), seed=17).to(dt).train()
        loss, _ = ores.tagger_step_loss(m, xs.to(dt), ts.to(dt), xr.to(dt), tw.to(dt))
        loss.backward()
        grads[dt] = {n: p.grad.detach().double().numpy().reshape(-1) for n, p in m.named_parameters()}
        rec["loss" if dt == torch.float32 else "loss64"] = float(loss)
    devs = []
    for n, g32 in grads[torch.float32].items():
        g64 = grads[torch.float64][n]
        sl = slice(None) if g32.size <= 4096 else slice(None, None, max(1, g32.size // 4096))
        rec["g64_" + n] = g64[sl][:4096].astype(np.float32)
        rec["gn_" + n] = float(np.linalg.norm(g64))
        rec["dev32_" + n] = float(np.linalg.norm(g32 - g64) / max(np.linalg.norm(g64), 1e-30))
        if rec["gn_" + n] > 1e-6:
            devs.append(rec["dev32_" + n])
            if devs[-1] > 1e-4:
                print("   torch fp32 vs fp64", n, devs[-1], "norm", rec["gn_" + n])
    rec["dev32_worst"], rec["dev32_median"] = max(devs), float(np.median(devs))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resnet_train8.npz"), **rec)
    print("resnet 4 + 4 fixture: loss", rec["loss"], "loss64", rec["loss64"], "torch fp32 vs fp64 gradients: worst",
          rec["dev32_worst"], "median", rec["dev32_median"])


if __name__ == "__main__":
    main()
    train_fixture()
    train_fixture8()
