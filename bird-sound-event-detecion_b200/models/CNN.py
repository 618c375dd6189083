"""Parameter holder for the 7-block gated CNN (reference: src/models/CNN.py:33-84).

The arithmetic of the blocks (Conv3x3 -> BatchNorm2d(eps=1e-3, momentum=.99) -> GLU -> Dropout ->
AvgPool2d) runs inside libbsed.so, driven by models.CRNN.CRNN.forward; this module only owns the
parameters / buffers under the reference's state-dict names (`conv{i}.weight`,
`batchnorm{i}.running_mean`, `glu{i}.linear.weight`, ...), as views into the flat buffers the
kernels read.
"""
from torch import nn


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise NotImplementedError("this sub-module only holds parameters; call CRNN.forward (libbsed.so)")


class CNN(_Holder):
    def __init__(self, n_in_channel, nb_filters, pooling):
        super().__init__()
        self.nb_filters = list(nb_filters)
        self.pooling = [tuple(p) for p in pooling]
        for i in range(len(nb_filters)):
            conv = _Holder()
            bn = _Holder()
            glu = _Holder()
            glu.linear = _Holder()
            self.add_module(f"conv{i}", conv)
            self.add_module(f"batchnorm{i}", bn)
            self.add_module(f"glu{i}", glu)


class CNN_FPN(_Holder):
    """Parameter holder of src/models/CNN_FPN.py:33-100: `cnn` (the 7 blocks) plus the shared stage `cnn_fcn`, `glu`,
    `bn_fcn` that CNN_FPN.forward applies twice, and `conv1x1`, which the reference registers but never calls (it stays in
    the state dict and receives no gradient)."""

    def __init__(self, n_in_channel, nb_filters, pooling):
        super().__init__()
        self.nb_filters = list(nb_filters)
        self.cnn = CNN(n_in_channel, nb_filters, pooling)
        self.cnn_fcn = _Holder()
        self.glu = _Holder()
        self.glu.linear = _Holder()
        self.bn_fcn = _Holder()
        self.conv1x1 = _Holder()
