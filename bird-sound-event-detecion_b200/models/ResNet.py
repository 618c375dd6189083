"""Net_resnet, the ResNet-18 weak-label tagger (src/audio_tagging_system_cnn.py:50-64): torchvision's resnet18 with
`conv1 = Conv2d(1, 64, 7, stride 2, padding 3, bias=False)`, `fc = Linear(512, len(bird_list))` and a sigmoid on top.

    model = Net_resnet(pretrained=False); model.load_state_dict(state["model"]["state_dict"]); model.eval()
    pred_weak = model(x)            x: (B, 1, 1255, 128) -> (B, 20)            (src/audio_tagging_inference.py:123-133, 295)

INFERENCE path only (what audio_tagging_inference.py runs to write the pseudo-label TSV): in eval mode every BatchNorm is
an affine map, folded here into the preceding convolution when the weights are (re)loaded; each convolution then runs in
libbsed.so as im2col -> GEMM (+ bias) -> [+ residual] -> ReLU on channels-last tensors.  State-dict keys equal the
reference's (`resnet.conv1.weight`, `resnet.bn1.running_mean`, `resnet.layer2.0.downsample.0.weight`, `resnet.fc.bias`, ...),
so its checkpoints load.  Training this model is not built (forward in train() mode raises); `pretrained=True` needs
torchvision's ImageNet weights, which are not reachable from here, and raises as well.
"""
import math

import torch
from torch import nn

from .. import engine


class _BN(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.eps = 1e-5


class _Conv(nn.Module):
    def __init__(self, cin, cout, k, stride, pad):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")     # torchvision's resnet init
        self.k, self.stride, self.pad = k, stride, pad


class _Block(nn.Module):
    """torchvision BasicBlock: conv1-bn1-relu-conv2-bn2 (+ downsample(x)) - relu."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1, self.bn1 = _Conv(cin, cout, 3, stride, 1), _BN(cout)
        self.conv2, self.bn2 = _Conv(cout, cout, 3, 1, 1), _BN(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(_Conv(cin, cout, 1, stride, 0), _BN(cout))
        else:
            self.downsample = None


class _ResNet18(nn.Module):
    def __init__(self, n_class):
        super().__init__()
        self.conv1, self.bn1 = _Conv(1, 64, 7, 2, 3), _BN(64)
        cin = 64
        for li, (cout, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
            setattr(self, f"layer{li}", nn.Sequential(_Block(cin, cout, stride), _Block(cout, cout, 1)))
            cin = cout
        self.fc = nn.Linear(512, n_class)


class Net_resnet(nn.Module):
    def __init__(self, pretrained=True, n_class=20, precision=None):
        super().__init__()
        if pretrained:
            raise NotImplementedError("Net_resnet(pretrained=True) needs torchvision's ImageNet checkpoint (no network here); "
                                      "build with pretrained=False and load a state dict")
        self.resnet = _ResNet18(n_class)
        self.n_class = n_class
        self.precision = precision          # "tf32" (tcgen05 GEMMs) / "fp32"; None = engine.default_precision()
        self._packed = None

    # ---- folded / packed operands (rebuilt whenever the weights may have changed)
    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    @staticmethod
    def _fold(conv, bn):
        """eval-mode BatchNorm folded into the convolution: (Wk [Cout][Kpad] in (ky, kx, ci) order, bias [Cout])."""
        with torch.no_grad():
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            w = conv.weight * scale[:, None, None, None]
            bias = bn.bias - bn.running_mean * scale
            cout, cin, kh, kw = w.shape
            K = kh * kw * cin
            kpad = (K + 31) // 32 * 32
            wk = torch.zeros(cout, kpad, dtype=torch.float32, device=w.device)
            wk[:, :K] = w.permute(0, 2, 3, 1).reshape(cout, K)
            return dict(wk=wk.contiguous(), wkT=wk.t().contiguous(), bias=bias.float().contiguous(), k=kh, stride=conv.stride,
                        pad=conv.pad, kpad=kpad, cout=cout)

    def _prepare(self):
        r = self.resnet
        ops = {"stem": self._fold(r.conv1, r.bn1), "blocks": []}
        for li in range(1, 5):
            for blk in getattr(r, f"layer{li}"):
                ops["blocks"].append(dict(c1=self._fold(blk.conv1, blk.bn1), c2=self._fold(blk.conv2, blk.bn2),
                                          ds=self._fold(blk.downsample[0], blk.downsample[1]) if blk.downsample is not None else None))
        with torch.no_grad():
            npad = (self.n_class + 15) // 16 * 16
            fcT = torch.zeros(512, npad, dtype=torch.float32, device=r.fc.weight.device)
            fcT[:, :self.n_class] = r.fc.weight.t()
            fb = torch.zeros(npad, dtype=torch.float32, device=r.fc.weight.device)
            fb[:self.n_class] = r.fc.bias
        ops["fcT"], ops["fb"] = fcT.contiguous(), fb
        self._packed = ops

    def _conv(self, x, op, tc):
        """x (B, H, W, Cin) channels-last -> (B, Ho, Wo, Cout) = conv + folded BatchNorm (no activation)."""
        B = x.shape[0]
        col, Ho, Wo = engine.im2col_nhwc(x, op["k"], op["k"], op["stride"], op["stride"], op["pad"], op["pad"], op["kpad"])
        cout = op["cout"]
        y = torch.empty(B * Ho * Wo, cout, dtype=torch.float32, device=x.device)
        if tc:
            for n0 in range(0, cout, 128):
                n1 = min(cout, n0 + 128)
                engine.gemm_nt_tc(col, op["wk"][n0:n1], op["bias"][n0:n1].contiguous(), out=y[:, n0:n1])
        else:
            engine.gemm_nn(col, op["wkT"], op["bias"], out=y)
        return y.view(B, Ho, Wo, cout)

    def forward(self, x):
        if self.training:
            raise NotImplementedError("libbsed Net_resnet implements the inference path (model.eval()); training it is not built")
        if not x.is_cuda or not self.resnet.fc.weight.is_cuda:
            raise RuntimeError("libbsed Net_resnet runs on CUDA tensors only (no CPU fallback)")
        if self._packed is None:
            self._prepare()
        ops = self._packed
        tc = (self.precision or engine.default_precision()).lower() == "tf32"
        with torch.no_grad():
            B = x.shape[0]
            h = x.detach().float().reshape(B, x.shape[-2], x.shape[-1], 1).contiguous()       # (B,1,T,F) -> (B,T,F,1)
            h = engine.add_relu(self._conv(h, ops["stem"], tc))
            h = engine.maxpool_nhwc(h, 3, 2, 1)
            for blk in ops["blocks"]:
                identity = h if blk["ds"] is None else self._conv(h, blk["ds"], tc)
                o = engine.add_relu(self._conv(h, blk["c1"], tc))
                o = self._conv(o, blk["c2"], tc)
                h = engine.add_relu(o, identity.contiguous())
            feat = engine.avgpool_nhwc(h)                                                       # (B, 512)
            logits = engine.gemm_nn(feat, ops["fcT"], ops["fb"])                                # (B, 32)
            return engine.sigmoid_rows(logits, self.n_class)
