"""Oracle (TEST INFRASTRUCTURE): torch-CPU restatement of the adversarial domain-adaptation branch.

PINNED: tests/make_golden_ada.py executes the reference's own `models.CRNN_GRL.Clip_Discriminator` and
`DA.cdan_frame.ConditionalDomainAdversarialLoss` (imported from /root/reference/src, with the `np.float = float`
shim the reference needs on NumPy >= 1.24) on seeded inputs and commits tests/golden/ada.npz;
tests/test_oracle_da.py checks this restatement against it (and against the live reference when present).

Follows:
  * src/models/CRNN_GRL.py:16-52   Clip_Discriminator
  * src/DA/grl.py:12-74            GradientReverseFunction / WarmStartGradientReverseLayer
  * src/DA/cdan_frame.py:89-119    ConditionalDomainAdversarialLoss.forward (entropy_conditioning=False)
"""
import math

import numpy as np
import torch
from torch import nn
import torch.nn.functional as F


class OracleClipDiscriminator(nn.Module):
    def __init__(self, input_dim=256, dropout=0):
        super().__init__()
        ch = [1, 128, 64, 32, 16, 8]
        for l in range(5):
            setattr(self, f"conv_{l + 1}", nn.Conv2d(ch[l], ch[l + 1], kernel_size=3, stride=2))
        self.avgpool = nn.AdaptiveAvgPool2d((2, 1))
        self.dense_d = nn.Linear(16, 1)
        for l in range(5):
            setattr(self, f"bn_{l + 1}", nn.BatchNorm2d(ch[l + 1]))

    def forward(self, x):
        x = torch.unsqueeze(x.permute(0, 2, 1), 1)
        for l in range(1, 6):
            x = F.leaky_relu(getattr(self, f"bn_{l}")(getattr(self, f"conv_{l}")(x)), 0.2)
        x = self.avgpool(x)
        x = x.view(-1, x.shape[1] * x.shape[2] * x.shape[3])
        return torch.sigmoid(self.dense_d(x))


class _GRL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, coeff):
        ctx.coeff = coeff
        return x * 1.0

    @staticmethod
    def backward(ctx, g):
        return g.neg() * ctx.coeff, None


def grl_coeff(iter_num, alpha=1.0, lo=0.0, hi=1.0, max_iters=1000):
    return float(2.0 * (hi - lo) / (1.0 + math.exp(-alpha * iter_num / max_iters)) - (hi - lo) + lo)


def cdan_clip_loss(disc, f_s, f_t, iter_num):
    """BCE(D(GRL(cat(f_s, f_t))), [1]*B_s + [0]*B_t), GRL coefficient of iteration `iter_num`."""
    f = torch.cat((f_s, f_t), dim=0)
    d = torch.squeeze(disc(_GRL.apply(f, grl_coeff(iter_num))))
    label = torch.cat((torch.ones(f_s.size(0), device=f.device), torch.zeros(f_t.size(0), device=f.device)))
    return F.binary_cross_entropy(d, label), d


def seeded_disc_init(disc, seed):
    """Deterministic weights for fixtures: PCG64 normal draws scaled like PyTorch's defaults; BN affine perturbed."""
    g = np.random.default_rng(seed)
    with torch.no_grad():
        for name, p in disc.named_parameters():
            if name.startswith("conv") and name.endswith("weight"):
                fan_in = p.shape[1] * 9
                p.copy_(torch.from_numpy(g.standard_normal(tuple(p.shape)).astype(np.float32)) * (1.0 / math.sqrt(fan_in)))
            elif name.startswith("bn") and name.endswith("weight"):
                p.copy_(torch.from_numpy((1.0 + 0.1 * g.standard_normal(tuple(p.shape))).astype(np.float32)))
            elif name == "dense_d.weight":
                p.copy_(torch.from_numpy((0.5 * g.standard_normal(tuple(p.shape))).astype(np.float32)))
            else:
                p.copy_(torch.from_numpy((0.1 * g.standard_normal(tuple(p.shape))).astype(np.float32)))


def seeded_features(n, seed):
    """(n, 313, 256) encoder-like features in [-1, 1] (GRU outputs are tanh-bounded)."""
    g = np.random.default_rng(seed)
    return torch.from_numpy(np.tanh(g.standard_normal((n, 313, 256))).astype(np.float32))
