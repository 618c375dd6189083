"""Host-side logic of the N>1 path, exercised with world_size 2 on CPU (gloo): the contiguous clip
sharding of the inference path and the 'sum all-reduce + 1/N folded into the optimiser' gradient
exchange of the training path."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bsed_b200.utilities import shard


def test_clip_shards_tile_the_input():
    for n in (0, 1, 7, 8, 45, 360, 361):
        for world in (1, 2, 4, 8):
            spans = [shard.clip_shard(n, r, world) for r in range(world)]
            covered = [i for a, b in spans for i in range(a, b)]
            assert covered == list(range(n))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= -(-n // world)
            assert all(s <= -(-n // world) for s in sizes)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.arange(10, dtype=torch.float32) * (rank + 1)
        scale = shard.allreduce_gradients(g)
        torch.save((g * scale, scale), os.path.join(out, f"r{rank}.pt"))
        ev = [(rank, i, i + 1) for i in range(rank + 1)]
        allev = shard.gather_in_rank_order(ev)
        torch.save(allev, os.path.join(out, f"e{rank}.pt"))
        # generic-optimizer paths: one flat all-reduce averages the parameter gradients of several modules
        lin = torch.nn.Linear(3, 2)
        for prm in lin.parameters():
            prm.grad = torch.full_like(prm, float(rank + 1))
        shard.allreduce_module_grads([lin])
        torch.save([prm.grad.clone() for prm in lin.parameters()], os.path.join(out, f"g{rank}.pt"))
        # the fused peer-memory step cannot be set up without GPUs: every rank must get None (and none may hang), so the
        # trainers fall back to the all-reduce path together
        buf = torch.zeros(16)
        dp = shard.FusedDataParallel.create(buf, buf.clone(), None, None)
        torch.save(dp is None, os.path.join(out, f"d{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_gradient_exchange_and_gather_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        g, scale = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert scale == 0.5
        assert torch.allclose(g, torch.arange(10, dtype=torch.float32) * 1.5)
        ev = torch.load(os.path.join(str(tmp_path), f"e{r}.pt"))
        assert ev == [(0, 0, 1), (1, 0, 1), (1, 1, 2)]
        assert torch.load(os.path.join(str(tmp_path), f"d{r}.pt")) is True
        for gr in torch.load(os.path.join(str(tmp_path), f"g{r}.pt")):
            assert torch.allclose(gr, torch.full_like(gr, 1.5))
