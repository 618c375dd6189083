"""oracle/resnet.py (Net_resnet restated over torchvision) against its fixture, and the state-dict contract of the CUDA
model's parameter holder."""
import numpy as np
import pytest
import torch

from helpers import golden, max_abs
from oracle import resnet as ores
from bsed_b200.utilities import synth


def test_oracle_matches_fixture():
    g = golden("resnet_eval.npz")
    m = ores.seeded_init(ores.OracleNetResnet(20), seed=17).eval()
    x = torch.from_numpy(synth.make_logmel_like(3, seed=51))
    assert float(x.double().sum()) == pytest.approx(float(g["x_sum"]), rel=1e-12)
    with torch.no_grad():
        out = m(x)
    assert out.shape == (3, 20)
    assert max_abs(out.numpy(), g["out"]) < 1e-5
    assert out.numpy().std() > 0.05                      # the fixture is sensitive
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])


def test_state_dict_keys_equal_the_reference_model():
    from bsed_b200.models.ResNet import Net_resnet
    g = golden("resnet_eval.npz")
    m = Net_resnet(pretrained=False)
    assert list(m.state_dict().keys()) == [str(k) for k in g["keys"]]
    assert [str(tuple(v.shape)) for v in m.state_dict().values()] == [str(s) for s in g["shapes"]]
    with pytest.raises(NotImplementedError):
        Net_resnet(pretrained=True)
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(1, 1, 1255, 128))            # CPU tensors: no fallback
