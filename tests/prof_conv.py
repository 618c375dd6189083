"""One tensor-core conv shape, a few launches (for ncu).  python tests/prof_conv.py Cin Cout T F [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsed_b200 import engine  # noqa: E402

Cin, Cout, T, F = [int(v) for v in sys.argv[1:5]]
B = int(sys.argv[5]) if len(sys.argv) > 5 else 24
x = torch.randn(B, T, F, Cin, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda")
b = torch.randn(Cout, device="cuda")
for _ in range(5):
    y = engine.conv3x3(x, w, b, tensor_cores=True)
torch.cuda.synchronize()
print("done", float(y.abs().mean()))
