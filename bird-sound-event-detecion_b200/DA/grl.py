"""Gradient reversal with warm start (reference: src/DA/grl.py:12-74).

forward: identity; backward: grad * (-coeff), coeff = 2(hi-lo)/(1+exp(-alpha*i/max_iters)) - (hi-lo) + lo with i
incremented on every forward when auto_step.  The scaling is a single elementwise pass on the device tensor."""
import math

import torch
from torch import nn
from torch.autograd import Function


class GradientReverseFunction(Function):
    @staticmethod
    def forward(ctx, input, coeff=1.):
        ctx.coeff = coeff
        return input.view_as(input) * 1.0

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.neg() * ctx.coeff, None


class GradientReverseLayer(nn.Module):
    def forward(self, *input):
        return GradientReverseFunction.apply(*input)


class WarmStartGradientReverseLayer(nn.Module):
    def __init__(self, alpha=1.0, lo=0.0, hi=1., max_iters=1000., auto_step=False):
        super().__init__()
        self.alpha, self.lo, self.hi = alpha, lo, hi
        self.iter_num = 0
        self.max_iters = max_iters
        self.auto_step = auto_step

    def coeff(self):
        return float(2.0 * (self.hi - self.lo) / (1.0 + math.exp(-self.alpha * self.iter_num / self.max_iters))
                     - (self.hi - self.lo) + self.lo)

    def forward(self, input):
        coeff = self.coeff()
        if self.auto_step:
            self.step()
        return GradientReverseFunction.apply(input, coeff)

    def step(self):
        self.iter_num += 1
