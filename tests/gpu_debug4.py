import os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bsed_b200 import engine
def run(B,T,Fq,Cin,Cout):
    g = np.random.default_rng(0)
    x = torch.from_numpy(g.standard_normal((B,Cin,T,Fq))).float(); dy = torch.from_numpy(g.standard_normal((B,Cout,T,Fq))).float()
    w = torch.zeros(Cout,Cin,3,3,dtype=torch.float64,requires_grad=True)
    F.conv2d(x.double(), w, None, padding=1).backward(dy.double())
    xs = x.permute(0,2,3,1).contiguous().cuda(); ds = dy.permute(0,2,3,1).contiguous().cuda()
    a = engine.conv3x3_wgrad(xs, ds, tensor_cores=False).cpu()
    b = engine.conv3x3_wgrad(xs, ds, tensor_cores=True).cpu()
    torch.cuda.synchronize()
    ref = w.grad.float()
    print((B,T,Fq,Cin,Cout), "simt rel", float((a-ref).norm()/ref.norm()), "tc rel", float((b-ref).norm()/ref.norm()), "tc absmax", float(b.abs().max()), "ref absmax", float(ref.abs().max()))
    print("  ref[0,0]", ref[0,0].flatten()[:9].numpy().round(2)); print("  tc [0,0]", b[0,0].flatten()[:9].numpy().round(2))
    # correlation of tc with ref under permutations
    for name, t in (("same", b), ("tapflip", b.flip(-1).flip(-2)), ("cico_T", b.transpose(0,1) if Cin==Cout else b)):
        if t.shape == ref.shape: print("   corr", name, float((t*ref).sum()/ (t.norm()*ref.norm()+1e-30)))
run(1,4,16,32,32); run(2,37,16,32,32); run(1,8,8,64,64); run(1,16,64,32,64); run(2,313,2,128,128); run(3,313,4,128,128)
