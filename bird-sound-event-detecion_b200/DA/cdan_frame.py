"""Clip-level conditional domain adversarial loss as the reference's runnable variant computes it
(src/DA/cdan_frame.py:16-119, used by src/main_scmt_ada_weak_seperate.py:331,793):

    loss = BCE( D( GRL( cat(f_s, f_t) ) ), [1]*B_s + [0]*B_t )

The class predictions g_s / g_t enter the reference's forward only through quantities that do not reach the returned
value when entropy_conditioning is False (`weight`, the max over classes), so they are accepted and ignored here.
D is this package's Clip_Discriminator (libbsed kernels); the BCE and its gradient are one kernel (bsed_disc_bce)."""
import ctypes as C

import torch
from torch import nn

from .. import _lib
from .._lib import check, ptr, stream_ptr
from .grl import WarmStartGradientReverseLayer


class _BCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, label):
        lib = _lib.load()
        h = _lib.handle(prob.device.index)
        p = prob.detach().contiguous().float().reshape(-1)
        loss = torch.empty(1, dtype=torch.float32, device=p.device)
        d_prob = torch.empty_like(p)
        check(lib.bsed_disc_bce(h, ptr(p), ptr(label.contiguous().float()), p.numel(), ptr(loss), ptr(d_prob), stream_ptr()),
              "bsed_disc_bce")
        ctx.save_for_backward(d_prob)
        ctx.shape = prob.shape
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d_prob,) = ctx.saved_tensors
        return (d_prob * g).reshape(ctx.shape), None


class ConditionalDomainAdversarialLoss(nn.Module):
    def __init__(self, domain_discriminator, entropy_conditioning=False, randomized=False, num_classes=-1, features_dim=-1,
                 randomized_dim=1024, reduction='mean'):
        super().__init__()
        if entropy_conditioning or randomized or reduction != 'mean':
            raise NotImplementedError("only the configuration the reference's runnable scripts use is supported: "
                                      "entropy_conditioning=False, randomized=False, reduction='mean'")
        self.domain_discriminator = domain_discriminator
        self.grl = WarmStartGradientReverseLayer(alpha=1., lo=0., hi=1., max_iters=1000, auto_step=True)
        self.domain_discriminator_accuracy = None

    def forward(self, g_s, f_s, g_t, f_t):
        f = torch.cat((f_s, f_t), dim=0)
        h = self.grl(f)
        d = torch.squeeze(self.domain_discriminator(h))
        d_label = torch.cat((torch.ones(f_s.size(0), device=f.device), torch.zeros(f_t.size(0), device=f.device)))
        return _BCEFunction.apply(d, d_label)
