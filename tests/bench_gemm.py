"""Tensor-core GEMM shapes of the CRNN's GLU / GRU layers (24 student clips), CUDA-event time per launch.
    python tests/bench_gemm.py [tf32x3|tf32]        (BSED_TC_DEBUG selects measurement experiments)"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsed_b200 import _lib, engine  # noqa: E402

lib = _lib.load()
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
# (M, K, N): GLU 1x1 of blocks 0-3 (blocks 0-2 in their packed-pixel view), GRU input projection (both directions, 3 gates)
SHAPES = [(24 * 1255 * 128 // 4, 64, 64), (24 * 627 * 64 // 2, 64, 64), (24 * 313 * 32, 64, 64), (24 * 313 * 16, 128, 128),
          (24 * 313 * 8, 128, 128), (24 * 313, 128, 768), (24 * 313, 256, 768)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tag = f"{prec} debug={os.environ.get('BSED_TC_DEBUG', '0')}"
for M, K, N in SHAPES:
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    for _ in range(3):
        engine.gemm_nt_tc(a, w, b, out=out, x3=prec == "tf32x3")
    ts = []
    for _ in range(10):
        flush.zero_()
        torch.cuda.synchronize()
        lib.bsed_profile_begin(3)          # class 3 = plain GEMM launches
        engine.gemm_nt_tc(a, w, b, out=out, x3=prec == "tf32x3")
        torch.cuda.synchronize()
        pm = C.c_double()
        _lib.check(lib.bsed_profile_end(C.byref(pm), None, None, None), "profile_end")
        ts.append(pm.value * 1e3)
    ts.sort()
    gb = 4.0 * (M * K + M * N + K * N) * 1e-9
    print(f"[{tag}] gemm M={M:7d} K={K:3d} N={N:3d}: median {ts[5]:7.1f} us (min {ts[0]:7.1f})  {gb / ts[5] * 1e6:6.0f} GB/s algorithmic, "
          f"{2.0 * M * K * N / ts[5] * 1e-6:6.1f} TFLOP/s")
