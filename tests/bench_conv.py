"""Tensor-core 3x3 convolution shapes of the CRNN (24 student clips), CUDA-event time per launch.
    python tests/bench_conv.py [tf32x3|tf32]        (BSED_TC_DEBUG / BSED_COL_BSTAGES select measurement experiments)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

from bsed_b200 import _lib, engine  # noqa: E402

lib = _lib.load()

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
# (Cin, Cout, T, F): block 1 in its pixel-pair view, blocks 2-4 (column-tiled kernel), blocks 5-6 (row-tiled kernel)
SHAPES = [(32, 64, 627, 32), (32, 64, 313, 32), (64, 128, 313, 16), (128, 128, 313, 8), (128, 128, 313, 4), (128, 128, 313, 2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tag = f"{prec} debug={os.environ.get('BSED_TC_DEBUG', '0')} bstages={os.environ.get('BSED_COL_BSTAGES', '-')}"
for Cin, Cout, T, F in SHAPES:
    x = torch.randn(24, T, F, Cin, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    b = torch.randn(Cout, device="cuda")
    for _ in range(3):
        engine.conv3x3(x, w, b, tensor_cores=prec)
    ts = []
    for _ in range(10):
        flush.zero_()
        torch.cuda.synchronize()
        lib.bsed_profile_begin(1)          # class 1 = conv forward / data gradient: CUDA events around that launch only
        engine.conv3x3(x, w, b, tensor_cores=prec)
        torch.cuda.synchronize()
        pm = C.c_double()
        _lib.check(lib.bsed_profile_end(C.byref(pm), None, None, None), "profile_end")
        ts.append(pm.value * 1e3)
    ts.sort()
    gf = 2.0 * 24 * T * F * Cout * 9 * Cin
    print(f"[{tag}] conv {Cin:3d}->{Cout:3d} T={T} F={F:2d}: median {ts[5]:7.1f} us (min {ts[0]:7.1f})  {gf / ts[5] * 1e-6:6.1f} TFLOP/s")
