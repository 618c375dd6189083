"""Multi-GPU check of the training steps under data parallelism (run under torchrun on a multi-GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tests/dp_train_check.py
Every rank trains on its own clips.  After several iterations -- the first enqueued kernel by kernel, the rest replayed from
the CUDA graph, whose data-parallel exchange kernel reads its epoch from the device-resident step state -- the replicas'
parameters, teacher and (for config 3) discriminator must be bit-identical across ranks, the fused path must agree with the
NCCL all-reduce path, and nobody may have timed out."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from bsed_b200 import engine
    from bsed_b200.DA.cdan_frame import ConditionalDomainAdversarialLoss
    from bsed_b200.main import AdaptationTrainer, MeanTeacherTrainer
    from bsed_b200.models import CRNN, Predictor
    from bsed_b200.models.CRNN import Clip_Discriminator
    from bsed_b200.utilities import synth
    from bsed_b200.utilities.utils import weights_init
    dev = torch.device("cuda", local)

    def make(seed):
        torch.manual_seed(seed)                        # identical initial replicas
        m, p = CRNN(**engine.REFERENCE_CRNN_KWARGS), Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
        weights_init(m)
        weights_init(p)
        return m.to(dev).train(), p.to(dev).train()

    def same_everywhere(t, what):
        ref = t.clone()
        dist.broadcast(ref, 0)
        ok = torch.tensor([int(torch.equal(ref, t))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{what}: replicas bit-identical: {bool(int(ok))}")
        return bool(int(ok))

    n = 2
    xs = torch.from_numpy(synth.make_logmel_like(n, seed=10 + rank)).to(dev)
    xr = torch.from_numpy(synth.make_logmel_like(n, seed=50 + rank)).to(dev)
    xe = (xr + 0.1).contiguous()
    ts = torch.from_numpy(synth.make_targets(n, seed=90 + rank)).to(dev)
    tw = ts.max(1)[0].contiguous()
    ok = True
    # ---- mean-teacher step: fused exchange + CUDA graph vs NCCL all-reduce (no graph)
    finals = {}
    for mode in ("fused", "nccl"):
        os.environ["BSED_DP"] = mode
        tr = MeanTeacherTrainer(*make(1), *make(2), lr=5e-4, n_syn=n, n_real=n, dropout_seed=2023 + rank)
        for it in range(6):
            losses = tr.step(xr, xe, xs, ts, it, 100)
        torch.cuda.synchronize()
        tr.check_health()
        ok &= bool(torch.isfinite(losses).all())
        ok &= same_everywhere(tr.params, f"mean-teacher [{mode}] parameters") and same_everywhere(tr.ema_params, f"mean-teacher [{mode}] teacher")
        if mode == "fused":
            graphed = bool(tr._graphs) and tr.dp is not None
            if rank == 0:
                print(f"mean-teacher [fused]: exchange kernel {'in' if tr.dp is not None else 'NOT in'} use, CUDA graph {'replayed' if tr._graphs else 'NOT used'}")
            ok &= graphed
        finals[mode] = tr.params.clone()
        del tr
    d = (finals["fused"] - finals["nccl"]).abs()
    if rank == 0:
        print(f"fused vs NCCL after 6 steps: max |dp| {float(d.max()):.2e}, mean {float(d.mean()):.2e} (Adam moves a weight by ~lr per step)")
    ok &= float(d.max()) < 6.5e-3 and float(d.mean()) < 5e-5
    # ---- config 3: adversarial update + mean-teacher update, three fused exchanges per iteration
    os.environ["BSED_DP"] = "fused"
    torch.manual_seed(7)
    disc = Clip_Discriminator(256).to(dev).train()
    crit = ConditionalDomainAdversarialLoss(disc)
    tr = AdaptationTrainer(*make(1), *make(2), crit, lr=5e-4, n_syn=n, n_real=n, dropout_seed=2023 + rank)
    for it in range(3):
        losses, dom = tr.step(xr, xe, tw, xs, ts, it, 100)
    torch.cuda.synchronize()
    tr.check_health()
    ok &= bool(torch.isfinite(losses).all()) and bool(torch.isfinite(dom).all())
    ok &= same_everywhere(tr.params, "adaptation parameters") and same_everywhere(disc.flat_tensors()[0], "adaptation discriminator")
    ok &= same_everywhere(tr.ema_params, "adaptation teacher")
    if rank == 0:
        print(f"adaptation: fused exchanges {'in' if tr.dp_adv is not None else 'NOT in'} use")
        print("DP-TRAIN-CHECK", "OK" if ok else "FAILED")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
