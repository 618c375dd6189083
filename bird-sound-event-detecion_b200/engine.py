"""Thin Python layer over the C ABI: plans, workspaces and the flat-buffer call wrappers.

PyTorch is used here for device memory and streams only; all arithmetic happens inside libbsed.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import BSED_F_SAVE, BSED_F_TRAIN, CrnnCfg, Group, OptCfg, check, ptr, stream_ptr

# crnn_kwargs of the reference (src/main.py:632-641)
REFERENCE_CRNN_KWARGS = dict(
    n_in_channel=1, nclass=20, attention=True, n_RNN_cell=128, n_layers_RNN=2, activation="glu",
    dropout=0.5, kernel_size=7 * [3], padding=7 * [1], stride=7 * [1],
    nb_filters=[16, 32, 64, 128, 128, 128, 128],
    pooling=[[2, 2], [2, 2], [1, 2], [1, 2], [1, 2], [1, 2], [1, 2]])
REFERENCE_PREDICTOR_KWARGS = dict(nclass=20, attention=True, n_RNN_cell=128)


def make_cfg(nclass=20, dropout=0.5, nb_filters=(16, 32, 64, 128, 128, 128, 128),
             pooling=((2, 2), (2, 2), (1, 2), (1, 2), (1, 2), (1, 2), (1, 2)), n_RNN_cell=128,
             n_layers_RNN=2, n_frames=1255, n_mels=128, bn_eps=1e-3, bn_momentum=0.99, fpn=False):
    cfg = CrnnCfg()
    cfg.n_frames, cfg.n_mels, cfg.n_cnn = int(n_frames), int(n_mels), len(nb_filters)
    for i, (c, p) in enumerate(zip(nb_filters, pooling)):
        cfg.filters[i] = int(c)
        cfg.pool_t[i] = int(p[0])
        cfg.pool_f[i] = int(p[1])
    cfg.rnn_hidden, cfg.rnn_layers, cfg.n_class = int(n_RNN_cell), int(n_layers_RNN), int(nclass)
    cfg.dropout, cfg.bn_eps, cfg.bn_momentum = float(dropout), float(bn_eps), float(bn_momentum)
    cfg.fpn = int(bool(fpn))   # CRNN_fpn (src/models/CRNN.py:243-337)
    return cfg


def default_precision():
    """"tf32x3" -- error-compensated 3xTF32 on the tcgen05 tensor cores: fp32-grade products, the parity mode against the
    reference's fp32 arithmetic -- unless BSED_PRECISION says "tf32" (single-pass tf32, what cuDNN convolutions give the
    reference on a GPU; stated looser tolerance) or "fp32" (CUDA-core cross-check)."""
    import os
    return os.environ.get("BSED_PRECISION", "tf32x3").lower()


def _dev_index(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("bird-sound-event-detecion_b200 runs on a CUDA device only (no CPU fallback)")
    return device.index if device.index is not None else torch.cuda.current_device()


class Plan:
    """bsed_plan + its workspace.  One saved forward at a time (see models/CRNN.py for the slot pool).
    `with_workspace=False` builds a plan that only serves the Predictor calls."""

    def __init__(self, cfg, max_clips, device, with_workspace=True, precision=None):
        self.lib = _lib.load()
        self.precision = (precision or default_precision()).lower()
        if self.precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        self.device = torch.device("cuda", _dev_index(device))
        self.h = _lib.handle(self.device.index)
        self.cfg = cfg
        self.max_clips = int(max_clips)
        self.p = C.c_void_p()
        check(self.lib.bsed_plan_create(self.h, C.byref(cfg), self.max_clips, C.byref(self.p)), "bsed_plan_create")
        check(self.lib.bsed_plan_set_precision(self.p, _lib.PRECISIONS[self.precision]), "bsed_plan_set_precision")
        self.n_params = int(self.lib.bsed_plan_param_count(self.p))
        self.n_pred_params = int(self.lib.bsed_predictor_param_count(self.p))
        self.n_bn = int(self.lib.bsed_plan_bn_buffer_count(self.p))
        self.t_out = int(self.lib.bsed_plan_out_frames(self.p))
        self.ldl = int(self.lib.bsed_predictor_ldl())
        self.n_class = int(cfg.n_class)
        self.n_cnn = int(cfg.n_cnn)
        self.n_bn_layers = self.n_cnn + (1 if cfg.fpn else 0)   # entries of num_batches_tracked
        self.ws_bytes = int(self.lib.bsed_plan_workspace_bytes(self.p))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device) if with_workspace else None
        self._pred_ws = {}
        self._keep = None

    def __del__(self):
        try:
            if self.p:
                self.lib.bsed_plan_destroy(self.p)
                self.p = None
        except Exception:
            pass

    def param_offsets(self):
        buf = (C.c_int64 * 256)()
        n = self.lib.bsed_plan_param_offsets(self.p, buf, 256)
        return [int(buf[i]) for i in range(n)]

    def predictor_offsets(self):
        buf = (C.c_int64 * 16)()
        n = self.lib.bsed_predictor_param_offsets(self.p, buf, 16)
        return [int(buf[i]) for i in range(n)]

    # groups: list of dicts(params=flat fp32, bn=flat fp32, nbt=int64[n_cnn] or None, n=int)
    def forward(self, groups, x, train, save, seed=0, step=0, enc=None):
        B = sum(g["n"] for g in groups)
        x = x.reshape(B, self.cfg.n_frames, self.cfg.n_mels)
        assert x.dtype == torch.float32 and x.is_cuda and x.is_contiguous()
        arr = (Group * len(groups))()
        first = 0
        for i, g in enumerate(groups):
            assert g["params"].numel() >= self.n_params and g["bn"].numel() == self.n_bn
            arr[i].params = g["params"].data_ptr()
            arr[i].bn_buffers = g["bn"].data_ptr()
            arr[i].num_batches_tracked = g["nbt"].data_ptr() if g.get("nbt") is not None else None
            arr[i].first_clip = first
            arr[i].n_clips = g["n"]
            first += g["n"]
        if enc is None:
            enc = torch.empty(B, self.t_out, 256, dtype=torch.float32, device=self.device)
        flags = (BSED_F_TRAIN if train else 0) | (BSED_F_SAVE if save else 0)
        check(self.lib.bsed_crnn_forward(self.p, arr, len(groups), ptr(x), B, flags, int(seed), int(step), ptr(enc),
                                         ptr(self.ws), self.ws_bytes, stream_ptr()), "bsed_crnn_forward")
        self._keep = (x, [g["params"] for g in groups]) if save else None   # backward reads them again
        return enc

    def backward(self, group_mask, d_enc, grads, accumulate=False):
        assert d_enc.is_contiguous() and d_enc.dtype == torch.float32
        check(self.lib.bsed_crnn_backward(self.p, int(group_mask), ptr(d_enc), ptr(grads), int(bool(accumulate)),
                                          ptr(self.ws), self.ws_bytes, stream_ptr()), "bsed_crnn_backward")
        self._keep = None

    def _pws(self, n):
        if n not in self._pred_ws:
            nb = int(self.lib.bsed_predictor_workspace_bytes(self.p, n))
            self._pred_ws[n] = (torch.empty(nb, dtype=torch.uint8, device=self.device), nb)
        return self._pred_ws[n]

    def predictor_forward(self, pred_params, enc, inference=False, out=None):
        """out: optional (logits, strong, weak) buffers of n clips (views into larger tensors are fine)."""
        n = enc.shape[0]
        ws, wsb = self._pws(n)
        enc = enc.contiguous()
        if out is not None:
            logits, strong, weak = out
        else:
            logits = torch.empty(n, self.t_out, self.ldl, dtype=torch.float32, device=self.device)
            strong = torch.empty(n, self.t_out, self.n_class, dtype=torch.float32, device=self.device)
            weak = torch.empty(n, self.n_class, dtype=torch.float32, device=self.device)
        check(self.lib.bsed_predictor_forward(self.p, ptr(pred_params), ptr(enc), n, int(bool(inference)), ptr(logits),
                                              ptr(strong), ptr(weak), ptr(ws), wsb, stream_ptr()),
              "bsed_predictor_forward")
        return logits, strong, weak

    def predictor_backward(self, pred_params, enc, logits, strong, weak, d_strong, d_weak, grads, accumulate=False,
                           d_enc=None):
        n = enc.shape[0]
        ws, wsb = self._pws(n)
        if d_enc is None:
            d_enc = torch.empty(n, self.t_out, 256, dtype=torch.float32, device=self.device)
        check(self.lib.bsed_predictor_backward(self.p, ptr(pred_params), ptr(enc), ptr(logits), ptr(strong), ptr(weak),
                                               ptr(d_strong), ptr(d_weak), n, ptr(d_enc), ptr(grads),
                                               int(bool(accumulate)), ptr(ws), wsb, stream_ptr()),
              "bsed_predictor_backward")
        return d_enc

    def debug_tensor(self, name):
        p = C.c_void_p()
        n = C.c_int64()
        check(self.lib.bsed_plan_debug_tensor(self.p, ptr(self.ws), name.encode(), C.byref(p), C.byref(n)),
              "bsed_plan_debug_tensor")
        off = p.value - self.ws.data_ptr()
        return self.ws[off:off + 4 * n.value].view(torch.float32)


# ------------------------------------------------------------------------------------------------
# stateless wrappers
# ------------------------------------------------------------------------------------------------
def mt_loss(strong, weak, syn_first, syn_n, syn_target, real_first, real_n, strong_ema, weak_ema, cons_w):
    """Losses of the mean-teacher step + gradients w.r.t. strong / weak (src/main.py:376-477)."""
    lib = _lib.load()
    h = _lib.handle(strong.device.index)
    B, T, Cn = strong.shape
    losses = torch.empty(4, dtype=torch.float32, device=strong.device)
    d_strong = torch.empty_like(strong)
    d_weak = torch.empty_like(weak)
    check(lib.bsed_mt_loss(h, ptr(strong), ptr(weak), B, T, Cn, syn_first, syn_n, ptr(syn_target), real_first, real_n,
                           ptr(strong_ema), ptr(weak_ema), float(cons_w), ptr(losses), ptr(d_strong), ptr(d_weak),
                           stream_ptr()), "bsed_mt_loss")
    return losses, d_strong, d_weak


def loss_terms(strong, weak, terms, n_slots):
    """Generic BCE / MSE terms (include/bsed.h: bsed_loss_terms).  terms: list of dicts(kind, pred_first, n, ref, roll=None,
    ref_is_strong=False, weight=1.0, grad_weight=None (= weight), slot).  Returns (losses [n_slots], d_strong, d_weak)."""
    lib = _lib.load()
    h = _lib.handle(strong.device.index)
    B, T, Cn = strong.shape
    arr = (_lib.LossTerm * len(terms))()
    keep = []
    for i, t in enumerate(terms):
        ref = t["ref"]
        assert ref.is_contiguous() and ref.dtype == torch.float32 and ref.is_cuda
        arr[i].kind, arr[i].pred_first, arr[i].n_clips = int(t["kind"]), int(t["pred_first"]), int(t["n"])
        arr[i].ref = ref.data_ptr()
        roll = t.get("roll")
        if roll is not None:
            assert roll.dtype == torch.int32 and roll.is_cuda and roll.numel() >= t["n"]
        arr[i].roll = roll.data_ptr() if roll is not None else None
        arr[i].ref_is_strong = int(bool(t.get("ref_is_strong", False)))
        arr[i].weight = float(t.get("weight", 1.0))
        gw = t.get("grad_weight")
        arr[i].grad_weight = float(arr[i].weight if gw is None else gw)
        arr[i].slot = int(t["slot"])
        keep.append((ref, roll))
    losses = torch.empty(n_slots, dtype=torch.float32, device=strong.device)
    d_strong = torch.empty_like(strong)
    d_weak = torch.empty_like(weak)
    check(lib.bsed_loss_terms(h, ptr(strong), ptr(weak), B, T, Cn, arr, len(terms), ptr(losses), n_slots, ptr(d_strong),
                              ptr(d_weak), stream_ptr()), "bsed_loss_terms")
    return losses, d_strong, d_weak


def roll_clips(x, shift_t=None, shift_f=None, out=None):
    """out[b] = roll(roll(x[b], shift_t[b], time), shift_f[b], frequency); x (B, T, F) or (B, 1, T, F) fp32 cuda,
    shifts int32 device tensors (or None)."""
    lib = _lib.load()
    h = _lib.handle(x.device.index)
    x = x.contiguous()
    B, T, F = x.shape[0], x.shape[-2], x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    check(lib.bsed_roll_clips(h, ptr(x), ptr(shift_t), ptr(shift_f), ptr(out), B, T, F, stream_ptr()), "bsed_roll_clips")
    return out


def opt_ema_step(params, grads, m, v, ema, step, ema_step=None, kind="adam", lr=5e-4, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, momentum=0.9, grad_scale=1.0, ema_alpha=0.999):
    lib = _lib.load()
    h = _lib.handle(params.device.index)
    cfg = OptCfg()
    cfg.kind = 0 if kind == "adam" else 1
    cfg.lr, cfg.beta1, cfg.beta2, cfg.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
    cfg.weight_decay, cfg.momentum, cfg.grad_scale, cfg.ema_alpha = float(weight_decay), float(momentum), float(grad_scale), float(ema_alpha)
    cfg.step = int(step)
    cfg.ema_step = int(ema_step if ema_step is not None else step)
    check(lib.bsed_opt_ema_step(h, ptr(params), ptr(grads), ptr(m), ptr(v), ptr(ema), params.numel(), C.byref(cfg),
                                stream_ptr()), "bsed_opt_ema_step")


def ema_buffers(bn, ema_bn, nbt, ema_nbt, ema_step, ema_alpha=0.999):
    lib = _lib.load()
    h = _lib.handle(bn.device.index)
    n_nbt = nbt.numel() if nbt is not None and ema_nbt is not None else 0
    check(lib.bsed_ema_buffers(h, ptr(bn), ptr(ema_bn), bn.numel(), ptr(nbt), ptr(ema_nbt), n_nbt, float(ema_alpha),
                               int(ema_step), stream_ptr()), "bsed_ema_buffers")


def melspec(audio):
    """audio (B, n) fp32 cuda -> (B, 1 + n // 255, 128) amplitude-mel."""
    lib = _lib.load()
    h = _lib.handle(audio.device.index)
    audio = audio.contiguous()
    B, n = audio.shape
    nf = lib.bsed_frontend_n_frames(n)
    mel = torch.empty(B, nf, 128, dtype=torch.float32, device=audio.device)
    check(lib.bsed_melspec(h, ptr(audio), B, n, ptr(mel), stream_ptr()), "bsed_melspec")
    return mel


def amp_to_db(mel, frames, unit_noise=None, snr=30.0, scaler_mean=None, scaler_std=None, out=None):
    """(B, t_in, 128) amplitude-mel -> (B, frames, 128) log-mel (ApplyLog -> PadOrTrunc -> [Normalize])."""
    lib = _lib.load()
    h = _lib.handle(mel.device.index)
    mel = mel.contiguous()
    B, t_in, _ = mel.shape
    if out is None:
        out = torch.empty(B, frames, 128, dtype=torch.float32, device=mel.device)
    wsb = int(lib.bsed_amp_to_db_workspace_bytes(B))
    ws = torch.empty(wsb, dtype=torch.uint8, device=mel.device)
    check(lib.bsed_amp_to_db(h, ptr(mel), ptr(unit_noise), float(snr), B, t_in, int(frames), ptr(scaler_mean),
                             ptr(scaler_std), ptr(out), ptr(ws), wsb, stream_ptr()), "bsed_amp_to_db")
    return out


def logmel(audio, frames, scaler_mean=None, scaler_std=None, return_mel=False):
    """audio (B, n) fp32 cuda -> (B, frames, 128) log-mel in one call (STFT + mel with the clip maximum, then one dB pass);
    bit-identical to amp_to_db(melspec(audio), frames)."""
    lib = _lib.load()
    h = _lib.handle(audio.device.index)
    audio = audio.contiguous()
    B, n = audio.shape
    nf = lib.bsed_frontend_n_frames(n)
    mel = torch.empty(B, nf, 128, dtype=torch.float32, device=audio.device)
    out = torch.empty(B, frames, 128, dtype=torch.float32, device=audio.device)
    wsb = int(lib.bsed_amp_to_db_workspace_bytes(B))
    ws = torch.empty(wsb, dtype=torch.uint8, device=audio.device)
    check(lib.bsed_logmel(h, ptr(audio), B, n, int(frames), ptr(scaler_mean), ptr(scaler_std), ptr(mel), ptr(out), ptr(ws), wsb,
                          stream_ptr()), "bsed_logmel")
    return (out, mel) if return_mel else out


def median_decode(strong, threshold=0.5, win=14, max_events=None):
    """strong (B, T, C) -> (events int32 (B, max_events, 3), n_events int32 (B,))."""
    lib = _lib.load()
    h = _lib.handle(strong.device.index)
    strong = strong.contiguous()
    B, T, Cn = strong.shape
    if max_events is None:
        max_events = Cn * ((T + 1) // 2)
    events = torch.zeros(B, max_events, 3, dtype=torch.int32, device=strong.device)
    n_events = torch.zeros(B, dtype=torch.int32, device=strong.device)
    check(lib.bsed_median_decode(h, ptr(strong), B, T, Cn, float(threshold), int(win), ptr(events), int(max_events),
                                 ptr(n_events), stream_ptr()), "bsed_median_decode")
    return events, n_events


def _rows(t):
    """Pointer of a 2-D fp32 tensor whose rows are contiguous (row stride = leading dimension)."""
    assert t.dim() == 2 and t.stride(1) == 1 and t.dtype == torch.float32
    return C.c_void_p(t.data_ptr())


def gemm_nn(a, b, bias=None, out=None, accumulate=False):
    lib = _lib.load()
    h = _lib.handle(a.device.index)
    M, K = a.shape
    N = b.shape[1]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    check(lib.bsed_gemm_nn(h, _rows(a), a.stride(0), _rows(b), b.stride(0), _rows(out), out.stride(0), M, N, K, ptr(bias),
                           int(bool(accumulate)), stream_ptr()), "bsed_gemm_nn")
    return out


def gemm_tn(a, b, out):
    """out[M][N] += a[K][M]^T @ b[K][N]"""
    lib = _lib.load()
    h = _lib.handle(a.device.index)
    K, M = a.shape
    N = b.shape[1]
    check(lib.bsed_gemm_tn(h, _rows(a), a.stride(0), _rows(b), b.stride(0), _rows(out), out.stride(0), M, N, K,
                           stream_ptr()),
          "bsed_gemm_tn")
    return out


_tn_ws = {}


def gemm_tn_tc(a, b, out):
    """tcgen05: out[M][N] += a[K][M]^T @ b[K][N]  (M % 32 == 0; N % 32 == 0 and N % 128 == 0 from 128 up)."""
    lib = _lib.load()
    h = _lib.handle(a.device.index)
    K, M = a.shape
    N = b.shape[1]
    if a.device not in _tn_ws:
        nb = int(lib.bsed_conv3x3_wgrad_workspace_bytes(h))
        _tn_ws[a.device] = (torch.empty(nb, dtype=torch.uint8, device=a.device), nb)
    ws, nb = _tn_ws[a.device]
    check(lib.bsed_gemm_tn_tc(h, _rows(a), a.stride(0), _rows(b), b.stride(0), _rows(out), out.stride(0), M, N, K, ptr(ws), nb,
                              stream_ptr()), "bsed_gemm_tn_tc")
    return out


def gemm_nt_tc(a, bk, bias=None, out=None, accumulate=False, x3=False):
    """tcgen05: out[M][N] (+)= a[M][K] @ bk[N][K]^T (+ bias); x3 = error-compensated 3xTF32."""
    lib = _lib.load()
    h = _lib.handle(a.device.index)
    M, K = a.shape
    N = bk.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    if x3:
        ws = torch.empty(2 * N * bk.stride(0), dtype=torch.float32, device=a.device)
        check(lib.bsed_gemm_nt_tc3(h, _rows(a), a.stride(0), _rows(bk), bk.stride(0), _rows(out), out.stride(0), M, N, K,
                                   ptr(bias), int(bool(accumulate)), ptr(ws), stream_ptr()), "bsed_gemm_nt_tc3")
        return out
    check(lib.bsed_gemm_nt_tc(h, _rows(a), a.stride(0), _rows(bk), bk.stride(0), _rows(out), out.stride(0), M, N, K,
                              ptr(bias), int(bool(accumulate)), stream_ptr()), "bsed_gemm_nt_tc")
    return out


def conv3x3(x, weight, bias=None, tensor_cores=False):
    """channels-last x (B, T, F, Cin), weight (Cout, Cin, 3, 3) -> (B, T, F, Cout).
    tensor_cores: False = fp32 CUDA cores, True / "tf32" = tcgen05 kind::tf32, "tf32x3" = error-compensated 3xTF32."""
    lib = _lib.load()
    h = _lib.handle(x.device.index)
    B, T, F, Cin = x.shape
    Cout = weight.shape[0]
    y = torch.empty(B, T, F, Cout, dtype=torch.float32, device=x.device)
    wpack = torch.empty(2 * 9 * Cin * Cout, dtype=torch.float32, device=x.device)
    fn = (lib.bsed_conv3x3_tc3 if tensor_cores == "tf32x3" else lib.bsed_conv3x3_tc) if tensor_cores else lib.bsed_conv3x3
    check(fn(h, ptr(x.contiguous()), ptr(weight.contiguous()), ptr(bias), ptr(y), B, T, F, Cin, Cout, ptr(wpack),
             stream_ptr()), "bsed_conv3x3")
    return y


def conv3x3_wgrad(x, dy, tensor_cores=False):
    """channels-last x (B,T,F,Cin), dy (B,T,F,Cout) -> dw (Cout,Cin,3,3)."""
    lib = _lib.load()
    h = _lib.handle(x.device.index)
    B, T, F, Cin = x.shape
    Cout = dy.shape[-1]
    dw = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float32, device=x.device)
    wsb = int(lib.bsed_conv3x3_wgrad_workspace_bytes(h))
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    check(lib.bsed_conv3x3_wgrad(h, ptr(x.contiguous()), ptr(dy.contiguous()), ptr(dw), B, T, F, Cin, Cout,
                                 int(bool(tensor_cores)), ptr(ws), wsb, stream_ptr()), "bsed_conv3x3_wgrad")
    return dw


# ------------------------------------------------------------------------------------------------
# channels-last building blocks of the ResNet-18 tagger's inference path (csrc/resnet.cu)
# ------------------------------------------------------------------------------------------------
def im2col_nhwc(x, kh, kw, sh, sw, ph, pw, k_pad):
    """x (B, H, W, Cin) -> (col (B*Ho*Wo, k_pad), Ho, Wo); column order (ky, kx, ci), zero-filled up to k_pad."""
    lib = _lib.load()
    h = _lib.handle(x.device.index)
    x = x.contiguous()
    B, H, W, Cin = x.shape
    Ho, Wo = (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
    col = torch.empty(B * Ho * Wo, k_pad, dtype=torch.float32, device=x.device)
    check(lib.bsed_im2col_nhwc(h, ptr(x), ptr(col), B, H, W, Cin, kh, kw, sh, sw, ph, pw, Ho, Wo, k_pad, stream_ptr()),
          "bsed_im2col_nhwc")
    return col, Ho, Wo


def add_relu(y, residual=None):
    lib = _lib.load()
    check(lib.bsed_add_relu(_lib.handle(y.device.index), ptr(y), ptr(residual), y.numel(), stream_ptr()), "bsed_add_relu")
    return y


def maxpool_nhwc(x, k, s, p):
    lib = _lib.load()
    x = x.contiguous()
    B, H, W, Cn = x.shape
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    y = torch.empty(B, Ho, Wo, Cn, dtype=torch.float32, device=x.device)
    check(lib.bsed_maxpool_nhwc(_lib.handle(x.device.index), ptr(x), ptr(y), B, H, W, Cn, k, s, p, Ho, Wo, stream_ptr()),
          "bsed_maxpool_nhwc")
    return y


def avgpool_nhwc(x):
    """(B, H, W, C) -> (B, C) mean over the pixels."""
    lib = _lib.load()
    x = x.contiguous()
    B, H, W, Cn = x.shape
    y = torch.empty(B, Cn, dtype=torch.float32, device=x.device)
    check(lib.bsed_avgpool_nhwc(_lib.handle(x.device.index), ptr(x), ptr(y), B, H * W, Cn, stream_ptr()), "bsed_avgpool_nhwc")
    return y


def sigmoid_rows(logits, n_cols):
    lib = _lib.load()
    rows, ld = logits.shape
    out = torch.empty(rows, n_cols, dtype=torch.float32, device=logits.device)
    check(lib.bsed_sigmoid_rows(_lib.handle(logits.device.index), ptr(logits), ld, ptr(out), rows, n_cols, stream_ptr()),
          "bsed_sigmoid_rows")
    return out


def bn_rows_train(x, gamma, beta, run_mean, run_var, nbt, residual=None, relu=True, eps=1e-5, momentum=0.1):
    """Train-mode BatchNorm over the rows of x [M][C] (x becomes xhat in place) -> (y, mean_rstd [2][C])."""
    lib = _lib.load()
    h = _lib.handle(x.device.index)
    M, Cn = x.shape
    y = torch.empty_like(x)
    mr = torch.empty(2, Cn, dtype=torch.float32, device=x.device)
    wsb = int(lib.bsed_bn_rows_workspace_bytes(Cn))
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    check(lib.bsed_bn_rows_train(h, ptr(x), M, Cn, ptr(gamma), ptr(beta), float(eps), float(momentum), ptr(run_mean), ptr(run_var),
                                 ptr(nbt), ptr(residual), int(bool(relu)), ptr(y), ptr(mr), ptr(ws), wsb, stream_ptr()),
          "bsed_bn_rows_train")
    return y, mr


def bn_rows_backward(dy, y, xhat, gamma, mean_rstd, d_gamma, d_beta, want_residual_grad=False):
    """dy [M][C] becomes the gradient w.r.t. the BatchNorm input (in place); returns the residual-branch gradient or None."""
    lib = _lib.load()
    h = _lib.handle(dy.device.index)
    M, Cn = dy.shape
    d_res = torch.empty_like(dy) if want_residual_grad else None
    wsb = int(lib.bsed_bn_rows_workspace_bytes(Cn))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dy.device)
    check(lib.bsed_bn_rows_backward(h, ptr(dy), ptr(y), ptr(xhat), M, Cn, ptr(gamma), ptr(mean_rstd), ptr(d_gamma), ptr(d_beta),
                                    ptr(d_res), ptr(ws), wsb, stream_ptr()), "bsed_bn_rows_backward")
    return d_res


def col2im_nhwc(dcol, shape, kh, kw, sh, sw, ph, pw, k_pad, out=None):
    """Transpose of im2col_nhwc: dcol (B*Ho*Wo, k_pad) -> dx (B, H, W, Cin); `out` given: accumulate into it."""
    lib = _lib.load()
    B, H, W, Cin = shape
    Ho, Wo = (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
    acc = out is not None
    if out is None:
        out = torch.empty(B, H, W, Cin, dtype=torch.float32, device=dcol.device)
    check(lib.bsed_col2im_nhwc(_lib.handle(dcol.device.index), ptr(dcol), ptr(out), B, H, W, Cin, kh, kw, sh, sw, ph, pw, Ho, Wo,
                               k_pad, int(acc), stream_ptr()), "bsed_col2im_nhwc")
    return out


def maxpool_nhwc_backward(x, dy, k, s, p):
    lib = _lib.load()
    B, H, W, Cn = x.shape
    Ho, Wo = dy.shape[1], dy.shape[2]
    dx = torch.empty_like(x)
    check(lib.bsed_maxpool_nhwc_backward(_lib.handle(x.device.index), ptr(x), ptr(dy.contiguous()), ptr(dx), B, H, W, Cn, k, s, p,
                                         Ho, Wo, stream_ptr()), "bsed_maxpool_nhwc_backward")
    return dx


def avgpool_nhwc_backward(dy, shape):
    lib = _lib.load()
    B, H, W, Cn = shape
    dx = torch.empty(B, H, W, Cn, dtype=torch.float32, device=dy.device)
    check(lib.bsed_avgpool_nhwc_backward(_lib.handle(dy.device.index), ptr(dy.contiguous()), ptr(dx), B, H * W, Cn, stream_ptr()),
          "bsed_avgpool_nhwc_backward")
    return dx


def sigmoid_rows_backward(p, dp, ld):
    lib = _lib.load()
    rows, Cn = p.shape
    dl = torch.empty(rows, ld, dtype=torch.float32, device=p.device)
    check(lib.bsed_sigmoid_rows_backward(_lib.handle(p.device.index), ptr(p.contiguous()), ptr(dp.contiguous()), ptr(dl), rows, Cn,
                                         ld, stream_ptr()), "bsed_sigmoid_rows_backward")
    return dl
