"""Oracle (TEST INFRASTRUCTURE): torch-CPU restatement of the reference CRNN / Predictor.

PINNED: tests/make_golden.py executes the reference's own modules (imported from
/root/reference/src in the build container) on seeded inputs and commits the outputs under
tests/golden/; tests/test_oracle_crnn.py checks this restatement against them (and against the
live reference whenever /root/reference is present).

Follows:
  * src/models/CNN.py:5-16    GLU  = Linear_C(x) * sigmoid(x)   (NOT nn.GLU)
  * src/models/CNN.py:33-84   CNN  = 7 x [Conv3x3(+bias) -> BatchNorm2d(eps=1e-3, momentum=.99)
                                        -> GLU -> Dropout -> AvgPool2d]
  * src/models/RNN.py:7-16    BidirectionalGRU = nn.GRU(bidirectional, batch_first)
  * src/models/CRNN.py:178-240   CRNN.forward -> (x, d_input)
  * src/models/CRNN.py:548-577   Predictor.forward -> (strong, weak)
  * src/main.py:632-641       crnn_kwargs / predictor_kwargs
  * src/models/CNN_FPN.py:33-100, src/models/CRNN.py:243-337   CNN_FPN / CRNN_fpn (OracleCRNNfpn below)
State-dict key names equal the reference's (79 CRNN keys `cnn.conv0.weight` ...
`rnn.rnn.bias_hh_l1_reverse`; Predictor `dense.*`, `dense_softmax.*`).

Dropout: the reference draws masks from torch's RNG; to make train-mode parity testable the
oracle (and the CUDA product) use a stateless counter hash `keep_mask` -- restated here in
numpy from csrc/common.cuh -- and tests/make_golden.py injects the same masks into the
reference modules.
"""
import numpy as np
import torch
from torch import nn

CRNN_KWARGS = dict(
    n_in_channel=1, nclass=20, attention=True, n_RNN_cell=128, n_layers_RNN=2, activation="glu",
    dropout=0.5, kernel_size=7 * [3], padding=7 * [1], stride=7 * [1],
    nb_filters=[16, 32, 64, 128, 128, 128, 128],
    pooling=[[2, 2], [2, 2], [1, 2], [1, 2], [1, 2], [1, 2], [1, 2]])
PREDICTOR_KWARGS = dict(nclass=20, attention=True, n_RNN_cell=128)

# dropout streams: one per CNN block, then the post-RNN dropout; CRNN_fpn adds the two applications of the shared
# stage (8, 9) and the rnn_2 / rnn_4 outputs (10, 11)   (csrc/engine.cu: bkeys / skeys)
STREAM_RNN_OUT = 7
STREAM_FCN = (8, 9)
STREAM_RNN_OUT_FPN = (7, 10, 11)


def mix_key(seed, step, stream):
    """Host-side 32-bit key for (seed, step, stream): splitmix64 finaliser, upper 32 bits."""
    m = (1 << 64) - 1
    z = (int(seed) * 0x9E3779B97F4A7C15 + int(step) * 0xD1B54A32D192ED03 + int(stream) * 0x8CB92BA72F3D8DD7 + 0x2545F4914F6CDD1D) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    z = z ^ (z >> 31)
    return (z >> 32) & 0xFFFFFFFF


def keep_mask(idx, key, p):
    """Device dropout rule (csrc/common.cuh: bsed_keep): murmur3 fmix32 of (idx*GOLD ^ key);
    keep iff hash >= floor(p * 2^32)."""
    h = (np.asarray(idx, dtype=np.uint64) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)
    h = h ^ np.uint64(key)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    thresh = np.uint64(min(int(p * 4294967296.0), 0xFFFFFFFF))
    return h >= thresh


class HashDropout(nn.Module):
    """Dropout whose mask is the device hash over the channels-last element index
    ((b*T + t)*F + f)*C + c of an NCHW tensor, or the flat index of a (B, T, C) tensor."""

    def __init__(self, p, stream):
        super().__init__()
        self.p = float(p)
        self.stream = int(stream)
        self.key = None          # set per forward by the caller (mix_key(seed, step, stream))
        self.batch_offset = 0    # clip index of x[0] inside the device batch

    def forward(self, x):
        if not self.training or self.p == 0.0 or self.key is None:
            return x
        if x.dim() == 4:
            B, C, T, F = x.shape
            b = np.arange(B)[:, None, None, None] + self.batch_offset
            c = np.arange(C)[None, :, None, None]
            t = np.arange(T)[None, None, :, None]
            f = np.arange(F)[None, None, None, :]
            idx = ((b * T + t) * F + f) * C + c
        else:
            B, T, C = x.shape
            idx = (np.arange(B * T * C) + self.batch_offset * T * C).reshape(B, T, C)
        keep = torch.from_numpy(keep_mask(idx, self.key, self.p)).to(x.dtype)
        return x * keep * (1.0 / (1.0 - self.p))


class _Gate(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.linear = nn.Linear(c, c)

    def forward(self, x):  # x: (B, C, T, F)
        lin = self.linear(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        return lin * torch.sigmoid(x)


class _BiGRU(nn.Module):
    def __init__(self, n_in, n_hidden, num_layers):
        super().__init__()
        self.rnn = nn.GRU(n_in, n_hidden, bidirectional=True, batch_first=True, num_layers=num_layers)

    def forward(self, x):
        return self.rnn(x)[0]


class OracleCRNN(nn.Module):
    def __init__(self, n_in_channel=1, nclass=20, attention=True, activation="glu", dropout=0.0,
                 n_RNN_cell=128, n_layers_RNN=2, kernel_size=None, padding=None, stride=None,
                 nb_filters=None, pooling=None, **_):
        super().__init__()
        assert activation.lower() == "glu"
        seq = nn.Sequential()
        cin = n_in_channel
        for i, cout in enumerate(nb_filters):
            seq.add_module(f"conv{i}", nn.Conv2d(cin, cout, kernel_size[i], stride[i], padding[i]))
            seq.add_module(f"batchnorm{i}", nn.BatchNorm2d(cout, eps=0.001, momentum=0.99))
            seq.add_module(f"glu{i}", _Gate(cout))
            seq.add_module(f"dropout{i}", HashDropout(dropout, i))
            seq.add_module(f"pooling{i}", nn.AvgPool2d(tuple(pooling[i])))
            cin = cout
        self.cnn = seq
        self.rnn = _BiGRU(cin, n_RNN_cell, n_layers_RNN)
        self.dropout = HashDropout(dropout, STREAM_RNN_OUT)

    def set_dropout_keys(self, seed, step, batch_offset=0):
        for m in self.modules():
            if isinstance(m, HashDropout):
                m.key = mix_key(seed, step, m.stream)
                m.batch_offset = batch_offset

    def forward(self, x):
        x = self.cnn(x)                       # (B, 128, 313, 1)
        x = x.squeeze(-1).permute(0, 2, 1)    # (B, 313, 128)
        x = self.rnn(x)                       # (B, 313, 256)
        x = self.dropout(x)
        return x, x


class CycleHashDropout(nn.Module):
    """One dropout MODULE called several times per forward (CNN_FPN.dropout twice, CRNN_fpn.dropout three times): call k
    uses streams[k].  `layout` "BCT" = the (B, C, T) tensors CRNN_fpn.forward drops (src/models/CRNN.py:318-322); the
    mask index is always the time-major element index ((b*T + t)*C + c) the kernels use."""

    def __init__(self, p, streams, layout="NCHW"):
        super().__init__()
        self.p, self.streams, self.layout = float(p), tuple(streams), layout
        self.keys = None
        self.batch_offset = 0
        self.calls = 0

    def set(self, seed, step, batch_offset=0):
        self.keys = [mix_key(seed, step, s) for s in self.streams]
        self.batch_offset = batch_offset
        self.calls = 0

    def forward(self, x):
        if not self.training or self.p == 0.0 or self.keys is None:
            return x
        key = self.keys[self.calls % len(self.keys)]
        self.calls += 1
        if self.layout == "BCT":
            B, C, T = x.shape
            b = np.arange(B)[:, None, None] + self.batch_offset
            idx = (b * T + np.arange(T)[None, None, :]) * C + np.arange(C)[None, :, None]
        else:
            B, C, T, F = x.shape
            b = np.arange(B)[:, None, None, None] + self.batch_offset
            c = np.arange(C)[None, :, None, None]
            t = np.arange(T)[None, None, :, None]
            f = np.arange(F)[None, None, None, :]
            idx = ((b * T + t) * F + f) * C + c
        keep = torch.from_numpy(keep_mask(idx, key, self.p)).to(x.dtype)
        return x * keep * (1.0 / (1.0 - self.p))


class _CNNFPN(nn.Module):
    """src/models/CNN_FPN.py:33-100.  Note `self.dropout = nn.Dropout(0.5)` (:79): the shared stage drops with p = 0.5
    whatever `conv_dropout` is."""

    def __init__(self, n_in_channel, dropout, kernel_size, padding, stride, nb_filters, pooling):
        super().__init__()
        seq = nn.Sequential()
        cin = n_in_channel
        for i, cout in enumerate(nb_filters):
            seq.add_module(f"conv{i}", nn.Conv2d(cin, cout, kernel_size[i], stride[i], padding[i]))
            seq.add_module(f"batchnorm{i}", nn.BatchNorm2d(cout, eps=0.001, momentum=0.99))
            seq.add_module(f"glu{i}", _Gate(cout))
            seq.add_module(f"dropout{i}", HashDropout(dropout, i))
            seq.add_module(f"pooling{i}", nn.AvgPool2d(tuple(pooling[i])))
            cin = cout
        self.cnn = seq
        self.cnn_fcn = nn.Conv2d(128, 128, 3, 1, 1)
        self.glu = _Gate(128)
        self.pool_fcn = nn.AvgPool2d([2, 1])
        self.bn_fcn = nn.BatchNorm2d(128, eps=0.001, momentum=0.99)
        self.conv1x1 = nn.Conv2d(256, 128, 1)          # registered, never called (as in the reference)
        self.dropout = CycleHashDropout(0.5, STREAM_FCN)

    def forward(self, x):
        x = self.cnn(x)
        outs = [x]
        for _ in range(2):
            x = self.pool_fcn(self.dropout(self.glu(self.bn_fcn(self.cnn_fcn(x)))))
            outs.append(x)
        return outs


class OracleCRNNfpn(nn.Module):
    """src/models/CRNN.py:243-337 (CRNN_fpn) with the hash dropout; state-dict keys equal the reference's."""

    def __init__(self, n_in_channel=1, nclass=20, attention=True, activation="glu", dropout=0.0,
                 n_RNN_cell=128, n_layers_RNN=2, kernel_size=None, padding=None, stride=None,
                 nb_filters=None, pooling=None, n_frames_out=313, **_):
        super().__init__()
        assert activation.lower() == "glu"
        self.cnn = _CNNFPN(n_in_channel, dropout, kernel_size, padding, stride, nb_filters, pooling)
        self.rnn = _BiGRU(nb_filters[-1], n_RNN_cell, n_layers_RNN)
        self.rnn_2 = _BiGRU(nb_filters[-1], n_RNN_cell, n_layers_RNN)
        self.rnn_4 = _BiGRU(nb_filters[-1], n_RNN_cell, n_layers_RNN)
        self.dropout = CycleHashDropout(dropout, STREAM_RNN_OUT_FPN, layout="BCT")
        self.upsample_2 = nn.Upsample((n_frames_out, 1), mode='bilinear', align_corners=True)
        self.upsample_4 = nn.Upsample((n_frames_out // 2, 1), mode='bilinear', align_corners=True)
        self.conv1x1_2 = nn.Conv2d(512, 256, 1)
        self.conv1x1_4 = nn.Conv2d(512, 256, 1)

    def set_dropout_keys(self, seed, step, batch_offset=0):
        for m in self.modules():
            if isinstance(m, HashDropout):
                m.key = mix_key(seed, step, m.stream)
                m.batch_offset = batch_offset
            elif isinstance(m, CycleHashDropout):
                m.set(seed, step, batch_offset)

    def forward(self, x, inference=False):
        x, x_2, x_4 = self.cnn(x)
        x = self.rnn(x.squeeze(-1).permute(0, 2, 1)).permute(0, 2, 1)          # (B, 256, 313)
        x_2 = self.rnn_2(x_2.squeeze(-1).permute(0, 2, 1)).permute(0, 2, 1)
        x_4 = self.rnn_4(x_4.squeeze(-1).permute(0, 2, 1)).permute(0, 2, 1)
        x = self.dropout(x).unsqueeze(-1)
        x_2 = self.dropout(x_2).unsqueeze(-1)
        x_4 = self.dropout(x_4).unsqueeze(-1)
        x_2 = self.conv1x1_2(torch.cat((x_2, self.upsample_4(x_4)), 1))
        x = self.conv1x1_4(torch.cat((x, self.upsample_2(x_2)), 1)).squeeze(-1)
        x = x.permute(0, 2, 1)
        return x, x


class OraclePredictor(nn.Module):
    def __init__(self, nclass=20, attention=True, n_RNN_cell=128, **_):
        super().__init__()
        assert attention
        self.dense = nn.Linear(2 * n_RNN_cell, nclass)
        self.dense_softmax = nn.Linear(2 * n_RNN_cell, nclass)

    def forward(self, x, inference=False):
        strong = torch.sigmoid(self.dense(x))
        sof = torch.softmax(self.dense_softmax(x), dim=-1)   # over the CLASS axis (CRNN.py:556)
        sof = torch.clamp(sof, min=1e-7, max=1)
        weak = (strong * sof).sum(1) / sof.sum(1)
        if inference:
            strong = strong * (weak > 0.5).to(strong.dtype).unsqueeze(1)
        return strong, weak


def reference_style_init(crnn, predictor, seed, linear_std=0.01):
    """weights_init of src/utilities/utils.py:40-63 in effect (xavier-uniform gain sqrt(2) convs,
    BN weight ~ N(1, .02), orthogonal GRU matrices, Linear ~ N(0, .01), zero biases) but drawn from a
    numpy Generator so fixtures regenerate identically wherever numpy's PCG64 stream is the same.
    `linear_std` > .01 spreads the probabilities so parity tests are sensitive."""
    rng = np.random.default_rng(seed)

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))

    with torch.no_grad():
        for mod in [crnn, predictor]:
            for m in mod.modules():
                if isinstance(m, nn.Conv2d):
                    fan_in = m.in_channels * 9
                    fan_out = m.out_channels * 9
                    a = np.sqrt(2.0) * np.sqrt(6.0 / (fan_in + fan_out))
                    m.weight.copy_(t(rng.uniform(-a, a, m.weight.shape)))
                    m.bias.zero_()
                elif isinstance(m, nn.BatchNorm2d):
                    m.weight.copy_(t(rng.normal(1.0, 0.02, m.weight.shape)))
                    m.bias.zero_()
                elif isinstance(m, nn.GRU):
                    for name, w in m.named_parameters():
                        if w.dim() > 1:
                            q, r = np.linalg.qr(rng.standard_normal((w.shape[0], w.shape[1])))
                            q = q * np.sign(np.diag(r))[None, :]
                            w.copy_(t(q))
                        else:
                            w.copy_(t(rng.uniform(-0.088, 0.088, w.shape)))
                elif isinstance(m, nn.Linear):
                    m.weight.copy_(t(rng.normal(0.0, linear_std, m.weight.shape)))
                    m.bias.zero_()
    return crnn, predictor
