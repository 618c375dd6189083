"""Per-kernel timings of the tensor-core GEMM family at the CRNN's layer shapes (CUDA events, warm-up 3,
mean of 10; inputs sized as in the 24-clip student batch, far larger than L2 in the early blocks).
    python tests/bench_kernels.py [conv|glu|wgrad|all]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsed_b200 import engine  # noqa: E402

B = 24
LAYERS = [  # Cin, Cout, T, F
    (16, 32, 627, 64), (32, 64, 313, 32), (64, 128, 313, 16), (128, 128, 313, 8), (128, 128, 313, 4), (128, 128, 313, 2)]


def kernel_us(fn, cls, n=10):
    """Mean device time of the library launches of profile class `cls` inside fn (CUDA events around each launch)."""
    import ctypes as C
    from bsed_b200 import _lib
    lib = _lib.load()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    lib.bsed_profile_begin(cls)
    for _ in range(n):
        fn()
    ms, fl, by, cnt = C.c_double(), C.c_double(), C.c_double(), C.c_int()
    lib.bsed_profile_end(C.byref(ms), C.byref(fl), C.byref(by), C.byref(cnt))
    return ms.value * 1e3 / max(cnt.value, 1)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    dev = "cuda"
    if what in ("glu", "all"):
        for C, T, F in [(16, 1255, 128), (32, 627, 64), (64, 313, 32), (128, 313, 16), (128, 313, 8),
                        (64, 1255, 32), (128, 1255, 16), (64, 627, 32)]:   # last three: packed views of blocks 0 / 1
            M = B * T * F
            a = torch.randn(M, C, device=dev)
            w = torch.randn(C, C, device=dev)
            bias = torch.randn(C, device=dev)
            out = torch.empty(M, C, device=dev)
            us = kernel_us(lambda: engine.gemm_nt_tc(a, w, bias, out=out), 3)
            us_acc = kernel_us(lambda: engine.gemm_nt_tc(a, w, None, out=out, accumulate=True), 3)
            gb = M * C * 4 * 2 / 1e9
            print(f"glu  C={C:3d} M={M:8d}: fwd {us:7.1f} us {gb / us * 1e6:7.0f} GB/s | accumulate {us_acc:7.1f} us "
                  f"{gb * 1.5 / us_acc * 1e6:7.0f} GB/s | {2.0 * M * C * C / us / 1e6:6.1f} TFLOP/s")
    if what in ("conv", "all"):
        for Cin, Cout, T, F in LAYERS:
            x = torch.randn(B, T, F, Cin, device=dev)
            w = torch.randn(Cout, Cin, 3, 3, device=dev)
            bias = torch.randn(Cout, device=dev)
            us = kernel_us(lambda: engine.conv3x3(x, w, bias, tensor_cores=True), 1)
            dy = torch.randn(B, T, F, Cout, device=dev)
            wT = w.permute(1, 0, 2, 3).flip(2, 3).contiguous()
            us_d = kernel_us(lambda: engine.conv3x3(dy, wT, None, tensor_cores=True), 1)
            fl = 2.0 * B * T * F * 9 * Cin * Cout
            gb = B * T * F * (Cin + Cout) * 4 / 1e9
            print(f"conv {Cin:3d}->{Cout:3d} T={T} F={F:2d}: fwd {us:7.1f} us {fl / us / 1e6:6.1f} TFLOP/s {gb / us * 1e6:6.0f} GB/s | "
                  f"dgrad {us_d:7.1f} us {fl / us_d / 1e6:6.1f} TFLOP/s")
    if what in ("wgrad", "all"):
        for Cin, Cout, T, F in LAYERS:
            x = torch.randn(B, T, F, Cin, device=dev)
            dy = torch.randn(B, T, F, Cout, device=dev)
            us = kernel_us(lambda: engine.conv3x3_wgrad(x, dy, tensor_cores=True), 2)
            fl = 2.0 * B * T * F * 9 * Cin * Cout
            gb = B * T * F * (Cin + Cout) * 4 / 1e9
            print(f"wgrad {Cin:3d}->{Cout:3d} T={T} F={F:2d}: {us:7.1f} us {fl / us / 1e6:6.1f} TFLOP/s {gb / us * 1e6:6.0f} GB/s (kernel only, excl. split-K reduce)")


if __name__ == "__main__":
    main()
