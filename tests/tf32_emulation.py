"""CPU emulation of the tensor-core operand precisions on the reference-generated eval fixture (tests/golden/crnn_eval.npz):
which split of which operands brings the probabilities within 1e-3?   python tests/tf32_emulation.py   (~2 min, no GPU)
Results: profiles/r02_tf32_emulation.md."""
import sys, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import helpers
from oracle import crnn as ocrnn
from bsed_b200.utilities import synth
import torch.nn.functional as F
torch.set_num_threads(8)

def rn(x):  # round to nearest tf32 (ties away ~ add half then mask)
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)
def tr(x):
    i = x.contiguous().view(torch.int32) & ~0x1FFF
    return i.view(torch.float32)
def bf(x): return x.to(torch.bfloat16).to(torch.float32)

def mm(kind, op, a, w, *args):
    # op(a, w) with emulated operand precision
    if kind == "fp32": return op(a, w, *args)
    if kind == "rn": return op(rn(a), rn(w), *args)
    if kind == "tr": return op(tr(a), tr(w), *args)
    if kind == "tr_a_rn_w": return op(tr(a), rn(w), *args)
    if kind == "x3":   # a trunc split, w rn split, three tf32 products (inputs truncated by the MMA)
        ah = tr(a); al = tr(a - ah); wh = rn(w); wl = tr(w - wh)
        bias = args[0] if args else None
        rest = args[1:] if args else ()
        return op(ah, wh, bias, *rest) + op(ah, wl, None, *rest) + op(al, wh, None, *rest)
    if kind == "x3bf":  # corrections in bf16
        ah = tr(a); al = a - ah; wh = rn(w); wl = w - wh
        bias = args[0] if args else None
        rest = args[1:] if args else ()
        return op(ah, wh, bias, *rest) + op(bf(a), bf(wl), None, *rest) + op(bf(al), bf(w), None, *rest)
    if kind == "x2a":  # only split A
        ah = tr(a); al = tr(a - ah); wh = rn(w)
        bias = args[0] if args else None
        rest = args[1:] if args else ()
        return op(ah, wh, bias, *rest) + op(al, wh, None, *rest)
    raise ValueError(kind)

def forward(oc, op, x, kconv, kglu, kgru):
    cnn = oc.cnn
    for i in range(7):
        conv = getattr(cnn, f"conv{i}"); bn = getattr(cnn, f"batchnorm{i}"); glu = getattr(cnn, f"glu{i}"); pool = getattr(cnn, f"pooling{i}")
        if i == 0: y = conv(x)
        else: y = mm(kconv, F.conv2d, x, conv.weight, conv.bias, 1, 1)
        y = bn(y)
        lin = mm(kglu, F.linear, y.permute(0, 2, 3, 1), glu.linear.weight, glu.linear.bias).permute(0, 3, 1, 2)
        x = pool(lin * torch.sigmoid(y))
    x = x.squeeze(-1).permute(0, 2, 1)
    g = oc.rnn.rnn
    for l in range(2):
        outs = []
        for d, suf in enumerate(["", "_reverse"]):
            wih = getattr(g, f"weight_ih_l{l}{suf}"); whh = getattr(g, f"weight_hh_l{l}{suf}")
            bih = getattr(g, f"bias_ih_l{l}{suf}"); bhh = getattr(g, f"bias_hh_l{l}{suf}")
            xg = mm(kgru, F.linear, x, wih, bih)
            B, T, _ = x.shape
            h = torch.zeros(B, 128)
            out = torch.zeros(B, T, 128)
            ts = range(T) if d == 0 else range(T - 1, -1, -1)
            for t in ts:
                gh = F.linear(h, whh, bhh)
                r = torch.sigmoid(xg[:, t, :128] + gh[:, :128]); z = torch.sigmoid(xg[:, t, 128:256] + gh[:, 128:256])
                n = torch.tanh(xg[:, t, 256:] + r * gh[:, 256:])
                h = (1 - z) * n + z * h
                out[:, t] = h
            outs.append(out)
        x = torch.cat(outs, -1)
    strong, weak = op(x)
    return x, strong, weak

for seed, std in [(5, 0.2)]:
    oc, op = helpers.oracle_models(seed=seed, linear_std=std)
    x = torch.from_numpy(synth.make_logmel_like(2, seed=11))
    with torch.no_grad():
        e0, s0, w0 = forward(oc, op, x, "fp32", "fp32", "fp32")
        ref_enc, _ = oc(x); rs, rw = op(ref_enc)
        print("manual vs module", (s0 - rs).abs().max().item())
        g = helpers.golden("crnn_eval.npz")
        print("vs golden", np.abs(s0.numpy() - g["strong"]).max())
        for k in [("rn",)*3, ("tr",)*3, ("tr_a_rn_w",)*3, ("rn","fp32","fp32"), ("fp32","rn","fp32"), ("fp32","fp32","rn"), ("rn","x3","x3"), ("x3",)*3, ("x3bf",)*3, ("x2a",)*3]:
            e, s, w = forward(oc, op, x, *k)
            print(k, "enc", (e - e0).abs().max().item(), "strong", (s - s0).abs().max().item(), "weak", (w - w0).abs().max().item())
