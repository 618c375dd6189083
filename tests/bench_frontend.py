"""Frontend kernel timing (CUDA events; 256 ten-second clips = 328 MB of audio, larger than L2):
    python tests/bench_frontend.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsed_b200 import engine  # noqa: E402


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
    return e0.elapsed_time(e1) / n


def main():
    B = 256
    a = torch.randn(B, 320000, device="cuda") * 0.1
    ms = timeit(lambda: engine.melspec(a))
    print(f"melspec {B} clips: {ms:.3f} ms -> {B / ms * 1e3:.0f} clips/s, {B * 1922560 / ms / 1e6:.1f} GB/s algorithmic, "
          f"{B * 1255 * 70000 / ms / 1e9:.2f} TFLOP/s fp32")
    mel = engine.melspec(a)
    out = torch.empty(B, 1255, 128, device="cuda")
    ms2 = timeit(lambda: engine.amp_to_db(mel, 1255, out=out))
    print(f"amp_to_db {B} clips: {ms2:.3f} ms -> {2 * B * 1255 * 128 * 4 / ms2 / 1e6:.0f} GB/s (read + write)")
    ms3 = timeit(lambda: engine.logmel(a, 1255))
    print(f"logmel (fused STFT + mel + clip max, one dB pass) {B} clips: {ms3:.3f} ms -> {B / ms3 * 1e3:.0f} clips/s, "
          f"{B * 1922560 / ms3 / 1e6:.1f} GB/s algorithmic")


if __name__ == "__main__":
    main()
