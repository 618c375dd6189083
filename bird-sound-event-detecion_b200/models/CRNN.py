"""CRNN / Predictor with the reference's constructor and forward signatures
(src/models/CRNN.py:178-240 and :548-577), executing in libbsed.so.

  model = CRNN(**crnn_kwargs); predictor = Predictor(**predictor_kwargs)        (src/main.py:632-641)
  encoded_x, d_input = model(x)            x: (B, 1, 1255, 128) -> (B, 313, 256) twice
  strong, weak = predictor(encoded_x)      (B, 313, 20), (B, 20)

Parameters and BatchNorm buffers are ordinary nn.Parameter / buffers with the reference's
state-dict keys, but they are views into flat device buffers that the kernels read directly, so
optimisers, state_dict() and checkpoints behave as in the reference.  There is no eager/PyTorch
fallback: forward on a non-CUDA tensor raises.
"""
import math
import weakref

import torch
from torch import nn

from .. import engine
from .CNN import CNN, CNN_FPN
from .RNN import BidirectionalGRU

_dropout_state = {"seed": 2023, "step": 0}


def set_dropout_seed(seed, step=0):
    """Seed of the stateless dropout hash (the reference seeds torch's RNG, src/main.py:573-574)."""
    _dropout_state["seed"] = int(seed)
    _dropout_state["step"] = int(step)


def _next_dropout_step():
    _dropout_state["step"] += 1
    return _dropout_state["seed"], _dropout_state["step"]


class _FlatModule(nn.Module):
    """nn.Module whose parameters / float buffers are views into flat tensors."""

    def _specs(self):
        raise NotImplementedError

    def _build(self, param_specs, buffer_specs, n_counters):
        n = sum(math.prod(s) for _, _, s in param_specs)
        self._flat = torch.zeros(n)
        self._flat_bn = torch.zeros(sum(math.prod(s) for _, _, s in buffer_specs))
        self._flat_nbt = torch.zeros(n_counters, dtype=torch.int64)
        self._param_specs, self._buffer_specs = param_specs, buffer_specs
        o = 0
        for mod, name, shape in param_specs:
            k = math.prod(shape)
            mod.register_parameter(name, nn.Parameter(self._flat[o:o + k].view(shape)))
            o += k
        o = 0
        for mod, name, shape in buffer_specs:
            k = math.prod(shape)
            mod.register_buffer(name, self._flat_bn[o:o + k].view(shape))
            o += k
        for i, mod in enumerate(self._counter_mods):
            mod.register_buffer("num_batches_tracked", self._flat_nbt[i])

    def _reflatten(self, flat=None, flat_bn=None, flat_nbt=None):
        """Re-point parameters/buffers at (new) flat storage, keeping their current values."""
        first = self._param_specs[0]
        cur = getattr(first[0], first[1])
        dev = cur.device
        flat = torch.empty(self._flat.numel(), device=dev, dtype=torch.float32) if flat is None else flat
        flat_bn = torch.empty(self._flat_bn.numel(), device=dev, dtype=torch.float32) if flat_bn is None else flat_bn
        flat_nbt = torch.empty(self._flat_nbt.numel(), device=dev, dtype=torch.int64) if flat_nbt is None else flat_nbt
        with torch.no_grad():
            o = 0
            for mod, name, shape in self._param_specs:
                k = math.prod(shape)
                p = getattr(mod, name)
                flat[o:o + k].copy_(p.detach().reshape(-1).to(torch.float32))
                p.data = flat[o:o + k].view(shape)
                p.grad = None
                o += k
            o = 0
            for mod, name, shape in self._buffer_specs:
                k = math.prod(shape)
                b = getattr(mod, name)
                flat_bn[o:o + k].copy_(b.detach().reshape(-1).to(torch.float32))
                mod._buffers[name] = flat_bn[o:o + k].view(shape)
                o += k
            for i, mod in enumerate(self._counter_mods):
                flat_nbt[i].copy_(mod._buffers["num_batches_tracked"])
                mod._buffers["num_batches_tracked"] = flat_nbt[i]
        self._flat, self._flat_bn, self._flat_nbt = flat, flat_bn, flat_nbt

    def _flat_ok(self):
        o = 0
        base = self._flat.data_ptr()
        for mod, name, shape in self._param_specs:
            p = getattr(mod, name)
            if p.data_ptr() != base + 4 * o or p.device != self._flat.device or not p.is_contiguous():
                return False
            o += math.prod(shape)
        o = 0
        base = self._flat_bn.data_ptr()
        for mod, name, shape in self._buffer_specs:
            b = getattr(mod, name)
            if b.data_ptr() != base + 4 * o or b.device != self._flat_bn.device:
                return False
            o += math.prod(shape)
        return True

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        # a no-op move (.cuda() / .to() / .float() to where the module already is) leaves every parameter the view it was:
        # keep the flat storage then -- a trainer may have re-homed it into its joint buffer (main.py: _rehome), and fresh
        # storage would silently detach the optimiser from the forward pass
        if not self._flat_ok():
            self._reflatten()
        return self

    def load_state_dict(self, state_dict, *a, **k):
        r = super().load_state_dict(state_dict, *a, **k)
        if not self._flat_ok():
            self._reflatten()
        return r

    def flat_tensors(self):
        """(flat params, flat BN running stats, int64 counters) the kernels read; re-flattened if
        something replaced a parameter's storage."""
        if not self._flat_ok():
            self._reflatten()
        return self._flat, self._flat_bn, self._flat_nbt

    def param_list(self):
        return [getattr(m, n) for m, n, _ in self._param_specs]


class _CRNNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, seed, step, need_grad, *params):
        flat, bn, nbt = module.flat_tensors()
        B = x.shape[0]
        train = module.training
        slot = module._acquire_slot(B, save=train and need_grad)
        xin = x.detach().contiguous().float()
        enc = slot.forward([dict(params=flat, bn=bn, nbt=nbt, n=B)], xin, train=train, save=train and need_grad,
                           seed=seed, step=step)
        ctx.module, ctx.slot = module, slot if (train and need_grad) else None
        if ctx.slot is None:
            module._release_slot(slot)
        else:  # a graph dropped without backward must still give its activations back
            ctx._fin = weakref.finalize(ctx, module._release_slot, slot)
        return enc

    @staticmethod
    def backward(ctx, d_enc):
        module, slot = ctx.module, ctx.slot
        if slot is None:
            raise RuntimeError("CRNN backward needs a forward in train() mode with grad enabled")
        grads = torch.empty_like(module._flat)
        slot.backward(1, d_enc.contiguous().float(), grads, accumulate=False)
        module._release_slot(slot)
        out, o = [], 0
        for _, _, shape in module._param_specs:
            k = math.prod(shape)
            out.append(grads[o:o + k].view(shape))
            o += k
        return (None, None, None, None, None, *out)


class CRNN(_FlatModule):
    _fpn = False

    def __init__(self, n_in_channel, nclass, attention=False, activation="Relu", dropout=0, train_cnn=True,
                 rnn_type='BGRU', n_RNN_cell=64, n_layers_RNN=1, dropout_recurrent=0, cnn_integration=False,
                 learned_post=False, **kwargs):
        super().__init__()
        nb_filters = list(kwargs.get("nb_filters", [64, 64, 64]))
        pooling = [tuple(p) for p in kwargs.get("pooling", [(1, 4)] * 3)]
        n = len(nb_filters)
        unsupported = []
        if n_in_channel != 1 or cnn_integration:
            unsupported.append("n_in_channel != 1 / cnn_integration")
        if activation.lower() != "glu":
            unsupported.append(f"activation={activation!r} (only 'glu')")
        if rnn_type != 'BGRU' or n_RNN_cell != 128 or dropout_recurrent != 0:
            unsupported.append("rnn (only BGRU, 128 cells, no recurrent dropout)")
        if list(kwargs.get("kernel_size", n * [3])) != n * [3] or list(kwargs.get("padding", n * [1])) != n * [1] \
                or list(kwargs.get("stride", n * [1])) != n * [1]:
            unsupported.append("conv geometry (only 3x3, stride 1, padding 1)")
        if not train_cnn:
            unsupported.append("train_cnn=False")
        if unsupported:
            raise NotImplementedError("libbsed CRNN supports the reference's crnn_kwargs (src/main.py:632-641); "
                                      "unsupported: " + "; ".join(unsupported))
        self.n_in_channel, self.attention, self.cnn_integration = n_in_channel, attention, cnn_integration
        self.rnn_type, self.train_cnn = rnn_type, train_cnn
        self.dropout_p = float(dropout)
        self.cfg_kwargs = dict(nclass=nclass, dropout=float(dropout), nb_filters=nb_filters, pooling=pooling,
                               n_RNN_cell=n_RNN_cell, n_layers_RNN=n_layers_RNN, fpn=self._fpn)
        if self._fpn:
            if nb_filters[-1] != 128:
                raise NotImplementedError("CRNN_fpn: the shared stage cnn_fcn is Conv2d(128, 128) (src/models/CNN_FPN.py:72)")
            self.cnn = CNN_FPN(1, nb_filters, pooling)
            trunk = self.cnn.cnn
        else:
            self.cnn = CNN(1, nb_filters, pooling)
            trunk = self.cnn
        self.rnn = BidirectionalGRU(nb_filters[-1], n_RNN_cell, dropout=dropout_recurrent, num_layers=n_layers_RNN)
        rnns = [self.rnn]
        if self._fpn:
            self.rnn_2 = BidirectionalGRU(nb_filters[-1], n_RNN_cell, dropout=dropout_recurrent, num_layers=n_layers_RNN)
            self.rnn_4 = BidirectionalGRU(nb_filters[-1], n_RNN_cell, dropout=dropout_recurrent, num_layers=n_layers_RNN)
            rnns += [self.rnn_2, self.rnn_4]
        self.dropout = nn.Dropout(dropout)   # placeholder with the reference's name; the mask is applied in-kernel
        ps, bs, cm = [], [], []
        cin = 1
        for i, c in enumerate(nb_filters):
            conv, bnm, glu = getattr(trunk, f"conv{i}"), getattr(trunk, f"batchnorm{i}"), getattr(trunk, f"glu{i}")
            ps += [(conv, "weight", (c, cin, 3, 3)), (conv, "bias", (c,)), (bnm, "weight", (c,)), (bnm, "bias", (c,)),
                   (glu.linear, "weight", (c, c)), (glu.linear, "bias", (c,))]
            bs += [(bnm, "running_mean", (c,)), (bnm, "running_var", (c,))]
            cm.append(bnm)
            cin = c
        if self._fpn:   # registration order of CNN_FPN.__init__ (src/models/CNN_FPN.py:71-79)
            f = self.cnn
            ps += [(f.cnn_fcn, "weight", (128, 128, 3, 3)), (f.cnn_fcn, "bias", (128,)),
                   (f.glu.linear, "weight", (128, 128)), (f.glu.linear, "bias", (128,)),
                   (f.bn_fcn, "weight", (128,)), (f.bn_fcn, "bias", (128,)),
                   (f.conv1x1, "weight", (128, 256, 1, 1)), (f.conv1x1, "bias", (128,))]
            bs += [(f.bn_fcn, "running_mean", (128,)), (f.bn_fcn, "running_var", (128,))]
            cm.append(f.bn_fcn)
        H = n_RNN_cell
        for rnn in rnns:
            for l in range(n_layers_RNN):
                n_in = cin if l == 0 else 2 * H
                for suf in ("", "_reverse"):
                    ps += [(rnn.rnn, f"weight_ih_l{l}{suf}", (3 * H, n_in)), (rnn.rnn, f"weight_hh_l{l}{suf}", (3 * H, H)),
                           (rnn.rnn, f"bias_ih_l{l}{suf}", (3 * H,)), (rnn.rnn, f"bias_hh_l{l}{suf}", (3 * H,))]
        if self._fpn:   # src/models/CRNN.py:281-282
            self.conv1x1_2, self.conv1x1_4 = nn.Module(), nn.Module()
            for m in (self.conv1x1_2, self.conv1x1_4):
                ps += [(m, "weight", (256, 512, 1, 1)), (m, "bias", (256,))]
        self._counter_mods = cm
        self._build(ps, bs, len(cm))
        self._slots, self._free = [], []
        # "tf32" (tensor cores) / "fp32" (CUDA cores); None = engine.default_precision()
        self.precision = kwargs.get("precision", None)
        self.reset_parameters()

    def reset_parameters(self):
        """PyTorch's default initialisation of the reference's layers (the reference then applies
        utilities.utils.weights_init, mirrored in utilities/utils.py)."""
        with torch.no_grad():
            for mod, name, shape in self._param_specs:
                p = getattr(mod, name)
                if name == "weight" and len(shape) == 4:
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                    bound = 1 / math.sqrt(shape[1] * shape[2] * shape[3])
                    nn.init.uniform_(getattr(mod, "bias"), -bound, bound)
                elif name == "weight" and len(shape) == 2:
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                    nn.init.uniform_(getattr(mod, "bias"), -1 / math.sqrt(shape[1]), 1 / math.sqrt(shape[1]))
                elif name == "weight" and len(shape) == 1:
                    p.fill_(1.0)
                    getattr(mod, "bias").zero_()
                elif name.startswith(("weight_ih", "weight_hh", "bias_ih", "bias_hh")):
                    nn.init.uniform_(p, -1 / math.sqrt(128), 1 / math.sqrt(128))
            for mod, name, _ in self._buffer_specs:
                getattr(mod, name).fill_(0.0 if name == "running_mean" else 1.0)
            self._flat_nbt.zero_()

    # ---- plan slots: each train-mode forward keeps its activations until its backward
    def _acquire_slot(self, B, save):
        prec = (self.precision or engine.default_precision()).lower()
        for s in self._free:
            if s.max_clips >= B and s.device == self._flat.device and s.precision == prec:
                self._free.remove(s)
                return s
        cfg = engine.make_cfg(**self.cfg_kwargs)
        s = engine.Plan(cfg, max_clips=B, device=self._flat.device, precision=prec)
        if s.n_params != self._flat.numel():
            raise RuntimeError(f"layout mismatch: library expects {s.n_params} parameters, module has {self._flat.numel()}")
        self._slots.append(s)
        return s

    def _release_slot(self, s):
        if s not in self._free:
            self._free.append(s)

    def forward(self, x, inference=False):
        # input size : (batch_size, n_channels, n_frames, n_freq); `inference` is accepted as CRNN_fpn.forward does
        # (src/models/CRNN.py:285) and, as there, has no effect on the encoder
        if not x.is_cuda:
            raise RuntimeError("libbsed CRNN runs on CUDA tensors only (no CPU fallback)")
        if not self._flat.is_cuda:
            raise RuntimeError("move the model to the GPU first (model.cuda())")
        seed, step = _next_dropout_step() if self.training else (0, 0)
        params = self.param_list()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)   # grad mode is off inside Function.forward
        enc = _CRNNFunction.apply(self, x, seed, step, need_grad, *params)
        return enc, enc


class CRNN_fpn(CRNN):
    """src/models/CRNN.py:243-337 -- feature-pyramid CRNN: the CNN_FPN trunk yields three time scales (313 / 156 / 78
    frames), each with its own BidirectionalGRU (`rnn`, `rnn_2`, `rnn_4`); the scales are merged top-down with
    bilinear upsampling and the 1x1 convolutions `conv1x1_2`, `conv1x1_4`.  Same constructor and
    forward(x, inference=False) -> (x, d_input) as the reference; state-dict keys `cnn.cnn.conv0.weight` ...
    `cnn.cnn_fcn.*`, `cnn.glu.linear.*`, `cnn.bn_fcn.*`, `cnn.conv1x1.*`, `rnn*.rnn.*`, `conv1x1_2.*`, `conv1x1_4.*`."""
    _fpn = True


class _PredictorFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, enc, inference, *params):
        flat, _, _ = module.flat_tensors()
        plan = module._plan(enc.shape[0])
        encc = enc.detach().contiguous().float()
        logits, strong, weak = plan.predictor_forward(flat, encc, inference=inference)
        ctx.module, ctx.plan = module, plan
        ctx.save_for_backward(encc, logits, strong, weak)
        ctx.inference = inference
        return strong, weak

    @staticmethod
    def backward(ctx, d_strong, d_weak):
        module, plan = ctx.module, ctx.plan
        if ctx.inference:
            raise RuntimeError("Predictor(inference=True) is not differentiable here")
        encc, logits, strong, weak = ctx.saved_tensors
        grads = torch.empty_like(module._flat)
        ds = d_strong.contiguous().float() if d_strong is not None else None
        dw = d_weak.contiguous().float() if d_weak is not None else None
        d_enc = plan.predictor_backward(module._flat, encc, logits, strong, weak, ds, dw, grads)
        out, o = [], 0
        for _, _, shape in module._param_specs:
            k = math.prod(shape)
            out.append(grads[o:o + k].view(shape))
            o += k
        return (None, d_enc, None, *out)


class Predictor(_FlatModule):
    def __init__(self, nclass, attention=False, n_RNN_cell=64, **kwargs):
        super().__init__()
        if not attention or n_RNN_cell != 128 or nclass > 20:
            raise NotImplementedError("libbsed Predictor supports attention=True, n_RNN_cell=128, nclass<=20 "
                                      "(src/main.py:641)")
        self.attention, self.nclass = attention, nclass
        self.dense = nn.Module()
        self.dense_softmax = nn.Module()
        ps = [(self.dense, "weight", (nclass, 256)), (self.dense, "bias", (nclass,)),
              (self.dense_softmax, "weight", (nclass, 256)), (self.dense_softmax, "bias", (nclass,))]
        self._counter_mods = []
        self._build(ps, [], 0)
        self._plans = {}
        with torch.no_grad():
            for m in (self.dense, self.dense_softmax):
                nn.init.kaiming_uniform_(m.weight, a=math.sqrt(5))
                nn.init.uniform_(m.bias, -1 / 16, 1 / 16)

    def _plan(self, n):
        key = self._flat.device
        if key not in self._plans:
            cfg = engine.make_cfg(nclass=self.nclass)
            self._plans[key] = engine.Plan(cfg, max_clips=1, device=self._flat.device, with_workspace=False)
        return self._plans[key]

    def forward(self, x, inference=False):
        if not x.is_cuda or not self._flat.is_cuda:
            raise RuntimeError("libbsed Predictor runs on CUDA tensors only (no CPU fallback)")
        return _PredictorFunction.apply(self, x, bool(inference), *self.param_list())


class _DiscFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        import ctypes as C
        from .. import _lib
        from .._lib import check, ptr, stream_ptr
        lib = _lib.load()
        flat, bn, nbt = module.flat_tensors()
        B = x.shape[0]
        h = _lib.handle(flat.device.index)
        xin = x.detach().contiguous().float()
        ws, wsb = module._workspace(B)
        check(lib.bsed_disc_set_precision(h, _lib.PRECISIONS[(module.precision or "fp32").lower()]), "bsed_disc_set_precision")
        prob = torch.empty(B, dtype=torch.float32, device=flat.device)
        check(lib.bsed_disc_forward(h, ptr(flat), ptr(bn), ptr(nbt), ptr(xin), B, int(module.training), ptr(prob), ptr(ws), wsb,
                                    stream_ptr()), "bsed_disc_forward")
        ctx.module, ctx.B, ctx.train = module, B, module.training
        ctx.save_for_backward(prob)
        ctx.need_dx = x.requires_grad
        return prob.reshape(B, 1)

    @staticmethod
    def backward(ctx, d_prob):
        from .. import _lib
        from .._lib import check, ptr, stream_ptr
        module = ctx.module
        if not ctx.train:
            raise RuntimeError("Clip_Discriminator backward needs a forward in train() mode (batch statistics)")
        lib = _lib.load()
        flat, _, _ = module.flat_tensors()
        (prob,) = ctx.saved_tensors
        h = _lib.handle(flat.device.index)
        ws, wsb = module._workspace(ctx.B)
        grads = torch.empty_like(flat)
        dx = torch.empty(ctx.B, 313, 256, dtype=torch.float32, device=flat.device) if ctx.need_dx else None
        check(lib.bsed_disc_backward(h, ptr(flat), ptr(prob), ptr(d_prob.contiguous().float().reshape(-1)), ctx.B, ptr(grads), 0,
                                     ptr(dx), ptr(ws), wsb, stream_ptr()), "bsed_disc_backward")
        out, o = [], 0
        for _, _, shape in module._param_specs:
            k = math.prod(shape)
            out.append(grads[o:o + k].view(shape))
            o += k
        return (None, dx, *out)


class Clip_Discriminator(_FlatModule):
    """src/models/CRNN_GRL.py:16-52 -- forward(x: (B, 313, 256)) -> (B, 1) domain probability, in libbsed kernels
    (csrc/disc.cu).  State-dict keys equal the reference's (conv_1.weight ... bn_5.num_batches_tracked, dense_d.*)."""

    def __init__(self, input_dim=256, dropout=0):
        super().__init__()
        chans = [1, 128, 64, 32, 16, 8]
        ps, bs, cm = [], [], []
        convs, bns = [], []
        for l in range(5):
            conv, bnm = nn.Module(), nn.Module()
            setattr(self, f"conv_{l + 1}", conv)
            convs.append(conv)
            ps += [(conv, "weight", (chans[l + 1], chans[l], 3, 3)), (conv, "bias", (chans[l + 1],))]
        self.dense_d = nn.Module()
        ps += [(self.dense_d, "weight", (1, 16)), (self.dense_d, "bias", (1,))]
        for l in range(5):
            bnm = nn.Module()
            setattr(self, f"bn_{l + 1}", bnm)
            bns.append(bnm)
            ps += [(bnm, "weight", (chans[l + 1],)), (bnm, "bias", (chans[l + 1],))]
            bs += [(bnm, "running_mean", (chans[l + 1],)), (bnm, "running_var", (chans[l + 1],))]
            cm.append(bnm)
        self._counter_mods = cm
        self._build(ps, bs, len(cm))
        self._ws = {}
        # "fp32" (default) or "tf32": the discriminator is dominated by its im2col / BatchNorm passes, not its GEMMs
        # (5.6 ms vs 4.8 ms per 24 clips), and five small-batch BatchNorms amplify the tf32 rounding of its gradients
        self.precision = "fp32"
        with torch.no_grad():   # PyTorch's default initialisation of Conv2d / Linear / BatchNorm2d
            for mod, name, shape in self._param_specs:
                p = getattr(mod, name)
                if name == "weight" and len(shape) == 4:
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                    nn.init.uniform_(mod.bias, -1 / math.sqrt(shape[1] * 9), 1 / math.sqrt(shape[1] * 9))
                elif name == "weight" and len(shape) == 2:
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                    nn.init.uniform_(mod.bias, -0.25, 0.25)
                elif name == "weight" and len(shape) == 1:
                    p.fill_(1.0)
                    mod.bias.zero_()
            for mod, name, _ in self._buffer_specs:
                getattr(mod, name).fill_(0.0 if name == "running_mean" else 1.0)
            self._flat_nbt.zero_()

    def _workspace(self, B):
        from .. import _lib
        key = (B, self._flat.device)
        if key not in self._ws:
            nb = int(_lib.load().bsed_disc_workspace_bytes(B))
            self._ws = {key: (torch.empty(nb, dtype=torch.uint8, device=self._flat.device), nb)}
        return self._ws[key]

    def forward(self, x):
        if not x.is_cuda or not self._flat.is_cuda:
            raise RuntimeError("libbsed Clip_Discriminator runs on CUDA tensors only (no CPU fallback)")
        if tuple(x.shape[1:]) != (313, 256):
            raise ValueError(f"Clip_Discriminator expects (B, 313, 256), got {tuple(x.shape)}")
        return _DiscFunction.apply(self, x, *self.param_list())
