"""Multi-GPU host logic.  The path shards by independent clips: inference
needs no data-path collective (contiguous blocks per rank, rank-ordered concatenation of results on
the host); training is data-parallel with ONE sum all-reduce of the flat gradient buffer per step
(4.47 MB), the 1/N average folded into the optimiser kernel's grad_scale."""
import torch
import torch.distributed as dist


def clip_shard(n_clips, rank, world):
    """[begin, end) of the contiguous block of ceil(n/world) clips owned by `rank`."""
    per = -(-n_clips // world) if world > 0 else n_clips
    a = min(n_clips, rank * per)
    return a, min(n_clips, a + per)


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_gradients(flat_grads, group=None):
    """Sum all-reduce in place; returns the scale (1/N) the optimiser kernel applies."""
    n = world_size(group)
    if n > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / n


def gather_in_rank_order(obj, group=None):
    """Concatenate per-rank Python lists (event lists, pseudo-label rows) in rank order."""
    n = world_size(group)
    if n == 1:
        return list(obj)
    parts = [None] * n
    dist.all_gather_object(parts, obj, group=group)
    return [x for p in parts for x in p]
