"""Host helpers of the training scripts (reference: src/utilities/utils.py:40-81)."""
import numpy as np
import torch
from torch import nn


def weights_init(model):
    """Effect of `model.apply(weights_init)` in the reference (utils.py:40-63) on our flat-parameter
    modules: xavier-uniform(gain sqrt 2) conv weights with zero bias, BatchNorm weight ~ N(1, .02) with
    zero bias, orthogonal GRU matrices (biases untouched), Linear weight ~ N(0, .01) with zero bias."""
    from ..models.CRNN import CRNN, Predictor
    with torch.no_grad():
        if isinstance(model, CRNN):
            for mod, name, shape in model._param_specs:
                p = getattr(mod, name)
                if name == "weight" and len(shape) == 4:
                    nn.init.xavier_uniform_(p, gain=np.sqrt(2))
                    mod.bias.fill_(0)
                elif name == "weight" and len(shape) == 1:
                    p.normal_(1.0, 0.02)
                    mod.bias.fill_(0)
                elif name == "weight" and len(shape) == 2:
                    p.normal_(0, 0.01)
                    mod.bias.zero_()
                elif name.startswith("weight_"):
                    w = torch.empty(shape)
                    nn.init.orthogonal_(w)
                    p.copy_(w)
        elif isinstance(model, Predictor):
            for m in (model.dense, model.dense_softmax):
                m.weight.normal_(0, 0.01)
                m.bias.zero_()
    return model


def to_cuda_if_available(*args):
    """utils.py:66-81"""
    res = list(args)
    if torch.cuda.is_available():
        for i, a in enumerate(res):
            res[i] = a.cuda()
    return res[0] if len(res) == 1 else res
