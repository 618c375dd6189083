"""Net_resnet, the ResNet-18 weak-label tagger (src/audio_tagging_system_cnn.py:50-64): torchvision's resnet18 with
`conv1 = Conv2d(1, 64, 7, stride 2, padding 3, bias=False)`, `fc = Linear(512, len(bird_list))` and a sigmoid on top.

    model = Net_resnet(pretrained=False); model.load_state_dict(state["model"]["state_dict"]); model.eval()
    pred_weak = model(x)            x: (B, 1, 1255, 128) -> (B, 20)            (src/audio_tagging_inference.py:123-133, 295)

Every convolution runs in libbsed.so as im2col -> GEMM on channels-last tensors (csrc/resnet.cu):
  * eval mode (the inference script): each BatchNorm is an affine map, folded into the preceding convolution when the
    weights are (re)loaded; conv = im2col -> GEMM + bias -> [+ residual] -> ReLU;
  * train mode (src/audio_tagging_system_cnn.py:199-416): conv -> train-mode BatchNorm over the GEMM rows (batch statistics,
    running statistics with momentum 0.1) -> [+ residual] -> ReLU, with the backward pass (BatchNorm / ReLU / residual,
    weight gradient = GEMM on the recomputed im2col matrix, data gradient = GEMM + col2im, max-pool and average-pool
    backward; the GEMMs on tcgen05 tf32 or, with precision="fp32", on the CUDA cores) behind one torch.autograd.Function; `TaggerTrainer` fuses the two model calls of an iteration, the BCE terms
    and Adam over the flat parameter buffer.
State-dict keys equal the reference's (`resnet.conv1.weight`, `resnet.bn1.running_mean`,
`resnet.layer2.0.downsample.0.weight`, `resnet.fc.bias`, ...), so its checkpoints load.  `pretrained=True` needs torchvision's
ImageNet weights, which are not reachable from here, and raises.
"""
import math

import torch
from torch import nn

from .. import engine
from .CRNN import _FlatModule


class _Holder(nn.Module):
    pass


def _conv_holder(cin, cout, k, stride, pad):
    m = _Holder()
    m.cin, m.cout, m.k, m.stride, m.pad = cin, cout, k, stride, pad
    return m


def _block(cin, cout, stride):
    """torchvision BasicBlock: conv1-bn1-relu-conv2-bn2 (+ downsample(x)) - relu."""
    b = _Holder()
    b.conv1, b.bn1 = _conv_holder(cin, cout, 3, stride, 1), _Holder()
    b.conv2, b.bn2 = _conv_holder(cout, cout, 3, 1, 1), _Holder()
    b.downsample = nn.Sequential(_conv_holder(cin, cout, 1, stride, 0), _Holder()) if (stride != 1 or cin != cout) else None
    return b


class _ResNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        p, tape = module._forward_train(x)
        ctx.module, ctx.tape = module, tape
        return p

    @staticmethod
    def backward(ctx, d_p):
        module = ctx.module
        grads = torch.zeros_like(module._flat)
        module._backward_train(ctx.tape, d_p.contiguous().float(), grads)
        out, o = [], 0
        for _, _, shape in module._param_specs:
            k = math.prod(shape)
            out.append(grads[o:o + k].view(shape))
            o += k
        return (None, None, *out)


class Net_resnet(_FlatModule):
    BN_EPS, BN_MOMENTUM = 1e-5, 0.1          # torchvision's BatchNorm2d defaults

    def __init__(self, pretrained=True, n_class=20, precision=None):
        super().__init__()
        if pretrained:
            raise NotImplementedError("Net_resnet(pretrained=True) needs torchvision's ImageNet checkpoint (no network here); "
                                      "build with pretrained=False and load a state dict")
        self.n_class = n_class
        self.precision = precision          # GEMMs: "tf32" (default, see _precision) / "tf32x3" (tcgen05) / "fp32"
        # backward GEMMs: None = as the forward.  Note for tf32: train-mode BatchNorm on a small batch behind the global
        # average pool is ill-conditioned, so the tf32 rounding of the FORWARD activations already moves the gradients
        # (0.19 rel. L2 on the 2-clip fixture, where the fp32 kernels sit at 0.017 and torch fp32 vs float64 at 0.004);
        # tf32 backward GEMMs add nothing visible on top.  precision="fp32" is the parity mode.
        self.backward_precision = None
        r = self.resnet = _Holder()
        r.conv1, r.bn1 = _conv_holder(1, 64, 7, 2, 3), _Holder()
        cin = 64
        for li, (cout, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
            setattr(r, f"layer{li}", nn.Sequential(_block(cin, cout, stride), _block(cout, cout, 1)))
            cin = cout
        r.fc = _Holder()
        # (conv, bn) units in forward order, parameter / buffer specs in torchvision's state-dict order
        self._units = [(r.conv1, r.bn1)]
        ps, bs, cm = [], [], []

        def add_unit(conv, bn):
            ps.extend([(conv, "weight", (conv.cout, conv.cin, conv.k, conv.k)), (bn, "weight", (conv.cout,)), (bn, "bias", (conv.cout,))])
            bs.extend([(bn, "running_mean", (conv.cout,)), (bn, "running_var", (conv.cout,))])
            cm.append(bn)

        add_unit(r.conv1, r.bn1)
        for li in range(1, 5):
            for blk in getattr(r, f"layer{li}"):
                add_unit(blk.conv1, blk.bn1)
                add_unit(blk.conv2, blk.bn2)
                if blk.downsample is not None:
                    add_unit(blk.downsample[0], blk.downsample[1])
        ps += [(r.fc, "weight", (n_class, 512)), (r.fc, "bias", (n_class,))]
        self._counter_mods = cm
        self._build(ps, bs, len(cm))
        self._packed = None
        with torch.no_grad():                                   # torchvision's resnet initialisation
            for mod, name, shape in self._param_specs:
                prm = getattr(mod, name)
                if len(shape) == 4:
                    nn.init.kaiming_normal_(prm, mode="fan_out", nonlinearity="relu")
                elif mod is r.fc and name == "weight":
                    nn.init.kaiming_uniform_(prm, a=math.sqrt(5))
                elif mod is r.fc:
                    nn.init.uniform_(prm, -1 / math.sqrt(512), 1 / math.sqrt(512))
                elif name == "weight":
                    prm.fill_(1.0)
                else:
                    prm.zero_()
            for mod, name, _ in self._buffer_specs:
                getattr(mod, name).fill_(0.0 if name == "running_mean" else 1.0)
            self._flat_nbt.zero_()

    # ---- folded / packed operands of the eval path (rebuilt whenever the weights may have changed)
    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def train(self, mode=True):
        self._packed = None
        return super().train(mode)

    @staticmethod
    def _pack(w):
        """(Cout, Cin, kh, kw) -> [Cout][Kpad] in (ky, kx, ci) order, zero padded to a multiple of 32 (of 128 above 128: the
        tensor-core weight-gradient GEMM tiles its N = Kpad by 128)."""
        cout, cin, kh, kw = w.shape
        K = kh * kw * cin
        kpad = (K + 31) // 32 * 32 if K <= 128 else (K + 127) // 128 * 128
        wk = torch.zeros(cout, kpad, dtype=torch.float32, device=w.device)
        wk[:, :K] = w.permute(0, 2, 3, 1).reshape(cout, K)
        return wk, K, kpad

    def _fold(self, conv, bn):
        """eval-mode BatchNorm folded into the convolution."""
        with torch.no_grad():
            scale = bn.weight / torch.sqrt(bn.running_var + self.BN_EPS)
            w4 = (conv.weight * scale[:, None, None, None]).float().contiguous()
            wk, K, kpad = self._pack(w4)
            bias = (bn.bias - bn.running_mean * scale).float().contiguous()
            return dict(wk=wk, wkT=wk.t().contiguous(), bias=bias, kpad=kpad, w4=w4)

    def _fc_operands(self):
        r = self.resnet
        with torch.no_grad():
            npad = (self.n_class + 15) // 16 * 16
            fcT = torch.zeros(512, npad, dtype=torch.float32, device=r.fc.weight.device)
            fcT[:, :self.n_class] = r.fc.weight.t()
            fb = torch.zeros(npad, dtype=torch.float32, device=r.fc.weight.device)
            fb[:self.n_class] = r.fc.bias
        return fcT.contiguous(), fb, npad

    def _blocks(self):
        r = self.resnet
        return [blk for li in range(1, 5) for blk in getattr(r, f"layer{li}")]

    @staticmethod
    def _implicit(conv, h):
        """3x3 / stride 1 / padding 1 units with 64 or 128 output channels (layer1, layer2: half of the network's FLOPs) run on
        the CRNN's implicit-GEMM tcgen05 convolution: no materialised im2col (23 MB per clip and unit in layer1)."""
        return (conv.k == 3 and conv.stride == 1 and conv.pad == 1 and conv.cin % 32 == 0 and conv.cout in (64, 128)
                and h.shape[2] % 2 == 0 and 128 % h.shape[2] == 0)

    def _gemm_wide(self, a, w, bias, out):
        """out[:, n] = a @ w[n]^T (+ bias[n]) on the tensor cores: up to 1024 output columns per launch (multiples of 128; the
        kernel walks (row tile, column block) pairs), then the remainder."""
        for n0, n1 in self._wide_chunks(w.shape[0]):
            engine.gemm_nt_tc(a, w[n0:n1], bias[n0:n1].contiguous() if bias is not None else None, out=out[:, n0:n1],
                              x3=self._use_x3())

    @staticmethod
    def _wide_chunks(n):
        """Column ranges of one wide-GEMM launch each: multiples of 128 up to 1024 columns, then the remainder (< 128)."""
        chunks, n0 = [], 0
        while n0 < n:
            left = n - n0
            n1 = n0 + (min(left // 128 * 128, 1024) if left >= 128 else left)
            chunks.append((n0, n1))
            n0 = n1
        return chunks

    def _gemm(self, col, wk, wkT, bias, tc):
        M, cout = col.shape[0], wk.shape[0]
        y = torch.empty(M, cout, dtype=torch.float32, device=col.device)
        if tc:
            self._gemm_wide(col, wk, bias, y)
        else:
            engine.gemm_nn(col, wkT, bias, out=y)
        return y

    def _precision(self):
        """Every contraction of the tagger is a convolution, which the reference runs in TF32 on a GPU (cuDNN,
        torch.backends.cudnn.allow_tf32): single-pass tf32 is the default here (probabilities within 1.4e-4 of the fp32
        oracle, tests/test_gpu_resnet.py); BSED_PRECISION or `precision=` select "tf32x3" / "fp32"."""
        import os
        return (self.precision or os.environ.get("BSED_PRECISION", "tf32")).lower()

    def _use_tc(self):
        return self._precision() in ("tf32", "tf32x3")

    def _use_x3(self):
        return self._precision() == "tf32x3"

    # ------------------------------------------------------------------------------------------ eval
    def _forward_eval(self, x):
        r = self.resnet
        if self._packed is None:
            ops = {id(conv): self._fold(conv, bn) for conv, bn in self._all_units()}
            ops["fc"] = self._fc_operands()
            self._packed = ops
        ops, tc = self._packed, self._use_tc()

        def conv(h, c):
            op = ops[id(c)]
            if tc and self._implicit(c, h):
                return engine.conv3x3(h, op["w4"], op["bias"], tensor_cores=self._precision())
            col, Ho, Wo = engine.im2col_nhwc(h, c.k, c.k, c.stride, c.stride, c.pad, c.pad, op["kpad"])
            return self._gemm(col, op["wk"], op["wkT"], op["bias"], tc).view(h.shape[0], Ho, Wo, c.cout)

        B = x.shape[0]
        h = x.detach().float().reshape(B, x.shape[-2], x.shape[-1], 1).contiguous()       # (B,1,T,F) -> (B,T,F,1)
        h = engine.add_relu(conv(h, r.conv1))
        h = engine.maxpool_nhwc(h, 3, 2, 1)
        for blk in self._blocks():
            identity = h if blk.downsample is None else conv(h, blk.downsample[0])
            o = engine.add_relu(conv(h, blk.conv1))
            h = engine.add_relu(conv(o, blk.conv2), identity.contiguous())
        feat = engine.avgpool_nhwc(h)                                                       # (B, 512)
        fcT, fb, _ = ops["fc"]
        return engine.sigmoid_rows(engine.gemm_nn(feat, fcT, fb), self.n_class)

    def _all_units(self):
        r = self.resnet
        units = [(r.conv1, r.bn1)]
        for blk in self._blocks():
            units += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2)]
            if blk.downsample is not None:
                units.append((blk.downsample[0], blk.downsample[1]))
        return units

    # ------------------------------------------------------------------------------------------ train
    def _unit_forward(self, h, conv, bn, relu, residual, tape, tc):
        """conv -> train-mode BatchNorm -> [+ residual] -> [ReLU] on channels-last h; records what backward needs."""
        B = h.shape[0]
        wk, K, kpad = self._pack(conv.weight.detach())
        if tc and self._implicit(conv, h):
            Ho, Wo = h.shape[1], h.shape[2]
            z = engine.conv3x3(h, conv.weight.detach().float().contiguous(), None,
                               tensor_cores=self._precision()).view(-1, conv.cout)           # (M, Cout), becomes xhat
        else:
            col, Ho, Wo = engine.im2col_nhwc(h, conv.k, conv.k, conv.stride, conv.stride, conv.pad, conv.pad, kpad)
            z = self._gemm(col, wk, None if tc else wk.t().contiguous(), None, tc)          # (M, Cout), becomes xhat
            del col
        nbt = self._flat_nbt[self._counter_mods.index(bn):][:1]
        res = residual.reshape(-1, conv.cout) if residual is not None else None
        y, mr = engine.bn_rows_train(z, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, nbt, res, relu,
                                     self.BN_EPS, self.BN_MOMENTUM)
        tape.append(dict(conv=conv, bn=bn, inp=h, xhat=z, y=y if relu else None, mr=mr, wk=wk, K=K, kpad=kpad,
                         has_res=residual is not None,
                         tc=(self.backward_precision.lower() in ("tf32", "tf32x3")) if self.backward_precision else tc))
        return y.view(B, Ho, Wo, conv.cout)

    def _forward_train(self, x):
        r, tc = self.resnet, self._use_tc()
        tape = []
        B = x.shape[0]
        h0 = x.detach().float().reshape(B, x.shape[-2], x.shape[-1], 1).contiguous()
        s = self._unit_forward(h0, r.conv1, r.bn1, True, None, tape, tc)
        h = engine.maxpool_nhwc(s, 3, 2, 1)
        for blk in self._blocks():
            identity = h if blk.downsample is None else self._unit_forward(h, blk.downsample[0], blk.downsample[1], False, None, tape, tc)
            o = self._unit_forward(h, blk.conv1, blk.bn1, True, None, tape, tc)
            h = self._unit_forward(o, blk.conv2, blk.bn2, True, identity, tape, tc)
        feat = engine.avgpool_nhwc(h)
        fcT, fb, npad = self._fc_operands()
        p = engine.sigmoid_rows(engine.gemm_nn(feat, fcT, fb), self.n_class)
        return p, dict(units=tape, stem_out=s, last_shape=tuple(h.shape), feat=feat, p=p, npad=npad, B=B)

    def _grad_view(self, grads, mod, name):
        o = 0
        for m, n, shape in self._param_specs:
            k = math.prod(shape)
            if m is mod and n == name:
                return grads[o:o + k].view(shape)
            o += k
        raise KeyError(name)

    def _unit_backward(self, rec, dy, grads, need_dx, dx_out=None):
        """dy (M, Cout): gradient w.r.t. the unit's output.  Adds the unit's parameter gradients into `grads`; returns
        (dx or None, residual-branch gradient or None).  dx_out given: the input gradient is accumulated into it."""
        conv, bn = rec["conv"], rec["bn"]
        d_res = engine.bn_rows_backward(dy, rec["y"], rec["xhat"], bn.weight.detach(), rec["mr"],
                                        self._grad_view(grads, bn, "weight"), self._grad_view(grads, bn, "bias"), rec["has_res"])
        inp = rec["inp"]
        if rec["tc"] and self._implicit(conv, inp):
            # layer1 / layer2 units: weight and data gradient by the CRNN's tensor-core kernels on the unit's own tensors
            # (dW: MN-major tcgen05 reduction over the pixels; dX: the forward kernel with flipped, transposed weights)
            dy4 = dy.view(inp.shape[0], inp.shape[1], inp.shape[2], conv.cout)
            self._grad_view(grads, conv, "weight").add_(engine.conv3x3_wgrad(inp, dy4, tensor_cores=True))
            dx = None
            if need_dx:
                wflip = conv.weight.detach().permute(1, 0, 2, 3).flip(2, 3).float().contiguous()
                dx = engine.conv3x3(dy4, wflip, None, tensor_cores=self._precision())
                if dx_out is not None:
                    dx = dx_out.add_(dx.view(dx_out.shape))
            return dx, d_res
        col, _, _ = engine.im2col_nhwc(inp, conv.k, conv.k, conv.stride, conv.stride, conv.pad, conv.pad, rec["kpad"])
        dwk = torch.zeros(conv.cout, rec["kpad"], dtype=torch.float32, device=dy.device)
        tc = rec["tc"]
        (engine.gemm_tn_tc if tc else engine.gemm_tn)(dy, col, dwk)      # dWk[co][k] = sum_rows dconv[row][co] * col[row][k]
        del col
        self._grad_view(grads, conv, "weight").add_(dwk[:, :rec["K"]].view(conv.cout, conv.k, conv.k, conv.cin).permute(0, 3, 1, 2))
        dx = None
        if need_dx:
            if tc:                                                       # (M, Kpad) = dconv * Wk
                wkT = rec["wk"].t().contiguous()
                dcol = torch.empty(dy.shape[0], rec["kpad"], dtype=torch.float32, device=dy.device)
                self._gemm_wide(dy, wkT, None, dcol)
            else:
                dcol = engine.gemm_nn(dy, rec["wk"])
            dx = engine.col2im_nhwc(dcol, tuple(inp.shape), conv.k, conv.k, conv.stride, conv.stride, conv.pad, conv.pad,
                                    rec["kpad"], out=dx_out)
        return dx, d_res

    def _backward_train(self, tape, d_p, grads):
        r = self.resnet
        B, npad = tape["B"], tape["npad"]
        dev = d_p.device
        # head: sigmoid -> fc -> global average pool
        dl = engine.sigmoid_rows_backward(tape["p"], d_p, npad)                                # (B, npad)
        dfc = torch.zeros(npad, 512, dtype=torch.float32, device=dev)
        engine.gemm_tn(dl, tape["feat"], dfc)
        self._grad_view(grads, r.fc, "weight").add_(dfc[:self.n_class])
        dbs = torch.zeros(npad, 16, dtype=torch.float32, device=dev)
        engine.gemm_tn(dl, torch.ones(B, 16, dtype=torch.float32, device=dev), dbs)            # column sums of d_logits
        self._grad_view(grads, r.fc, "bias").add_(dbs[:self.n_class, 0])
        fc_pad = torch.zeros(npad, 512, dtype=torch.float32, device=dev)
        fc_pad[:self.n_class] = r.fc.weight.detach()
        dfeat = engine.gemm_nn(dl, fc_pad)                                                     # (B, 512)
        dh = engine.avgpool_nhwc_backward(dfeat, tape["last_shape"])
        # residual blocks, last to first
        units = list(tape["units"])
        for blk in reversed(self._blocks()):
            rec2 = units.pop()
            rec1 = units.pop()
            recd = units.pop() if blk.downsample is not None else None
            d_o1, d_id = self._unit_backward(rec2, dh.reshape(-1, blk.conv2.cout), grads, True)
            if recd is None:
                # identity branch: the block input's gradient starts as the masked output gradient
                dx0 = d_id.view(rec1["inp"].shape)
                dh, _ = self._unit_backward(rec1, d_o1.reshape(-1, blk.conv1.cout), grads, True, dx_out=dx0)
            else:
                dx0, _ = self._unit_backward(recd, d_id, grads, True)
                dh, _ = self._unit_backward(rec1, d_o1.reshape(-1, blk.conv1.cout), grads, True, dx_out=dx0)
        # stem: max-pool -> (conv1, bn1, relu)
        stem = units.pop()
        ds = engine.maxpool_nhwc_backward(tape["stem_out"], dh, 3, 2, 1)
        self._unit_backward(stem, ds.reshape(-1, 64), grads, False)
        assert not units

    # ------------------------------------------------------------------------------------------ entry point
    def forward(self, x):
        if not x.is_cuda or not self._flat.is_cuda:
            raise RuntimeError("libbsed Net_resnet runs on CUDA tensors only (no CPU fallback)")
        if not self.training:
            with torch.no_grad():
                return self._forward_eval(x)
        params = self.param_list()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _ResNetFunction.apply(self, x, *params)
        return self._forward_train(x)[0]


class TaggerTrainer:
    """One fused iteration of the tagger's train_mt (src/audio_tagging_system_cnn.py:340-406): two model calls (synthetic
    batch, weak / unlabeled batch; BatchNorm statistics per call), loss = BCE(syn_weak, max_t syn_target) +
    BCE(weak[:half], target_weak[:half]), backward, Adam over the flat parameter buffer (csrc/head.cu: opt_ema_kernel)."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        assert isinstance(model, Net_resnet) and model._flat.is_cuda
        self.model = model
        self.params = model.flat_tensors()[0]
        self.grads = torch.zeros_like(self.params)
        self.m, self.v = torch.zeros_like(self.params), torch.zeros_like(self.params)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.opt_step = 0

    def step(self, syn_batch_input, syn_target, batch_input, target_weak):
        """syn_target: (n, 313, C) strong or (n, C) weak targets; target_weak (n, C).  Returns the loss (device scalar)."""
        from .._lib import LOSS_BCE_WEAK
        m = self.model
        m.train()
        ps, ts = m._forward_train(syn_batch_input)
        pr, tr = m._forward_train(batch_input)
        n, widx = pr.shape[0], target_weak.shape[0] // 2
        weak = torch.cat([ps, pr]).contiguous()
        tgt = syn_target.float().contiguous()
        strong_dummy = torch.zeros(weak.shape[0], 1, weak.shape[1], dtype=torch.float32, device=weak.device)
        terms = [dict(kind=LOSS_BCE_WEAK, pred_first=0, n=ps.shape[0], ref=tgt, ref_is_strong=tgt.dim() == 3, slot=0)]
        if widx > 0:
            terms.append(dict(kind=LOSS_BCE_WEAK, pred_first=ps.shape[0], n=widx, ref=target_weak[:widx].float().contiguous(), slot=0))
        if tgt.dim() == 3:   # the loss kernel indexes strong references with the T of its `strong` argument
            strong_dummy = torch.zeros(weak.shape[0], tgt.shape[1], weak.shape[1], dtype=torch.float32, device=weak.device)
        losses, _, d_weak = engine.loss_terms(strong_dummy, weak, terms, 1)
        self.grads.zero_()
        m._backward_train(tr, d_weak[ps.shape[0]:].contiguous(), self.grads)
        m._backward_train(ts, d_weak[:ps.shape[0]].contiguous(), self.grads)
        self.opt_step += 1
        engine.opt_ema_step(self.params, self.grads, self.m, self.v, None, step=self.opt_step, kind="adam", lr=self.lr,
                            betas=self.betas, eps=self.eps, weight_decay=self.weight_decay)
        return losses[0]
