"""Oracle (TEST INFRASTRUCTURE): the ResNet-18 weak tagger of the reference, restated on the CPU.

Follows src/audio_tagging_system_cnn.py:50-64 (`Net_resnet`): `torchvision.models.resnet18`, `fc` replaced by
`Linear(512, len(cfg.bird_list))`, `conv1` by `Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)`, sigmoid on the
output; inference use: src/audio_tagging_inference.py:123-133 (`Net_resnet(pretrained=False)`, `load_state_dict`,
`eval()`), :295 (`pred_weak = model(input_data)`).

The arithmetic lives in torchvision (an un-vendored, unpinned dependency of the reference; torchvision 0.26 is present in
the build container and is what this file calls).  The reference script itself cannot be imported (tensorboardX,
matplotlib), so the eight lines of the class are restated here: PARITY UNPINNED by the reference's own tests (it has
none); pinned against torchvision by tests/golden/resnet_eval.npz (tests/make_golden_resnet.py).
`pretrained=True` (the reference's training default) needs the ImageNet checkpoint from the network: not available.
"""
import numpy as np
import torch
from torch import nn


class OracleNetResnet(nn.Module):
    def __init__(self, n_class=20):
        super().__init__()
        from torchvision import models
        self.resnet = models.resnet18(weights=None)
        self.resnet.fc = nn.Linear(self.resnet.fc.in_features, n_class)
        self.resnet.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        return self.sigmoid(self.resnet(x))


def seeded_init(model, seed, fc_std=0.0015):
    """Deterministic weights from a numpy PCG64 stream (fixtures regenerate wherever numpy's generator is the same):
    He-normal convolutions, BatchNorm weight ~ N(1, .1), bias ~ N(0, .1), running_mean ~ N(0, .1), running_var ~ U(.5, 1.5)
    (so the eval-mode folding is exercised with non-trivial statistics), fc ~ N(0, fc_std) -- small, because the un-normalised features of a log-mel (dB) input are of order 30 and the
    probabilities must stay off saturation for the fixture to be sensitive."""
    rng = np.random.default_rng(seed)

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))

    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.Conv2d):
                fan_out = m.out_channels * m.kernel_size[0] * m.kernel_size[1]
                m.weight.copy_(t(rng.normal(0.0, np.sqrt(2.0 / fan_out), m.weight.shape)))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(t(rng.normal(1.0, 0.1, m.weight.shape)))
                m.bias.copy_(t(rng.normal(0.0, 0.1, m.bias.shape)))
                m.running_mean.copy_(t(rng.normal(0.0, 0.1, m.running_mean.shape)))
                m.running_var.copy_(t(rng.uniform(0.5, 1.5, m.running_var.shape)))
            elif isinstance(m, nn.Linear):
                m.weight.copy_(t(rng.normal(0.0, fc_std, m.weight.shape)))
                m.bias.copy_(t(rng.normal(0.0, 0.1, m.bias.shape)))
    return model


def tagger_step_loss(model, syn_batch_input, syn_target, batch_input, target_weak):
    """Loss of one iteration of the tagger's train_mt (src/audio_tagging_system_cnn.py:340-352): two model calls in the
    order of the reference (synthetic batch, then the weak / unlabeled batch), BCELoss on the weak outputs; of the real
    batch only the weakly labelled first half counts.  Returns (loss, None)."""
    bce = nn.BCELoss()
    syn_weak_pred = model(syn_batch_input)
    weak_pred = model(batch_input)
    syn_target_weak = syn_target.max(-2)[0]
    widx = target_weak.shape[0] // 2
    loss = bce(syn_weak_pred, syn_target_weak) + bce(weak_pred[:widx], target_weak[:widx])
    return loss, None
