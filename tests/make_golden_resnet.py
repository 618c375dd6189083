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


if __name__ == "__main__":
    main()
