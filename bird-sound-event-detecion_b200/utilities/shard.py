"""Multi-GPU host logic.  The path shards by independent clips: inference
needs no data-path collective (contiguous blocks per rank, rank-ordered concatenation of results on
the host); training is data-parallel with ONE sum all-reduce of the flat gradient buffer per step
(4.47 MB), the 1/N average folded into the optimiser kernel's grad_scale."""
import ctypes as C
import logging
import os

import torch
import torch.distributed as dist

log = logging.getLogger("bsed_b200.shard")


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


def allreduce_module_grads(modules, group=None):
    """Average the .grad of every parameter of `modules` over the ranks with ONE all-reduce of a flat copy (the generic
    torch-optimizer paths: the adversarial update of src/main_scmt_ada_weak_seperate.py:314-335 under data parallelism)."""
    n = world_size(group)
    if n == 1:
        return
    grads = [p.grad for m in modules for p in m.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / n)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def gather_in_rank_order(obj, group=None):
    """Concatenate per-rank Python lists (event lists, pseudo-label rows) in rank order."""
    n = world_size(group)
    if n == 1:
        return list(obj)
    parts = [None] * n
    dist.all_gather_object(parts, obj, group=group)
    return [x for p in parts for x in p]


class FusedDataParallel:
    """Gradient exchange + optimiser + EMA in ONE kernel over NVLink peer memory (include/bsed.h: bsed_dp_opt_ema_step).

    Every rank exports its flat gradient / parameter / EMA buffers and a flag block through CUDA IPC and maps its
    peers'.  Each step one kernel per rank waits for the peers' gradients, reduces ITS slice of the flat buffer in rank
    order out of peer memory, updates that slice (parameters, optimiser state, EMA teacher) and stores the new values
    into every peer's buffers.  `FusedDataParallel.create` returns None (every rank alike) when peer mapping is
    unavailable, and the caller keeps the NCCL all-reduce path."""

    def __init__(self, grads, params, ema=None, group=None):
        from .. import _lib
        self.lib = _lib.load()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("fused data-parallel step supports up to 8 ranks")
        self.device = grads.device
        if self.device.type != "cuda":
            raise RuntimeError("peer-memory exchange needs CUDA buffers")
        self.h = _lib.handle(self.device.index)
        self.grads, self.params, self.ema = grads, params, ema
        assert params.numel() == grads.numel() and (ema is None or ema.numel() == grads.numel())
        self.flags = torch.zeros(64, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)

        def export(t):
            buf = C.create_string_buffer(64)
            off = C.c_uint64()
            _lib.check(self.lib.bsed_ipc_export(self.h, C.c_void_p(t.data_ptr()), buf, C.byref(off)), "bsed_ipc_export")
            return buf.raw, int(off.value)

        mine = dict(pid=os.getpid(), grads=export(grads), params=export(params), ema=export(ema) if ema is not None else None,
                    flags=export(self.flags), n=grads.numel())
        infos = [None] * self.world
        dist.all_gather_object(infos, mine, group=group)
        if any(i["n"] != grads.numel() or (i["ema"] is None) != (ema is None) for i in infos):
            raise RuntimeError("ranks disagree on the flat buffer layout")
        self._opened = {}
        mk = lambda: (C.c_void_p * self.world)()
        self.peer_grads, self.peer_params, self.peer_ema, self.peer_flags = mk(), mk(), mk(), mk()
        for r, info in enumerate(infos):
            if r == self.rank:
                self.peer_grads[r], self.peer_params[r], self.peer_flags[r] = grads.data_ptr(), params.data_ptr(), self.flags.data_ptr()
                self.peer_ema[r] = ema.data_ptr() if ema is not None else None
            else:
                self.peer_grads[r] = self._open(r, *info["grads"])
                self.peer_params[r] = self._open(r, *info["params"])
                self.peer_ema[r] = self._open(r, *info["ema"]) if ema is not None else None
                self.peer_flags[r] = self._open(r, *info["flags"])
        self.epoch = 0
        self.check_every = max(1, int(os.environ.get("BSED_DP_CHECK_EVERY", "50")))
        self._hflag = self._hflag_ev = None

    def close(self):
        """Unmap the peers' allocations (after the last step; every rank alike)."""
        from .. import _lib
        for base in getattr(self, "_opened", {}).values():
            try:
                self.lib.bsed_ipc_close(self.h, C.c_void_p(base), 0)
            except Exception:   # noqa: BLE001 -- interpreter shutdown
                pass
        self._opened = {}

    def __del__(self):
        self.close()

    def _open(self, r, handle, offset):
        key = (r, handle)
        if key not in self._opened:       # one mapping per peer allocation
            base = C.c_void_p()
            from .. import _lib
            _lib.check(self.lib.bsed_ipc_open(self.h, handle, 0, C.byref(base)), "bsed_ipc_open")
            self._opened[key] = base.value
        return self._opened[key] + offset

    @classmethod
    def create(cls, grads, params, ema=None, group=None):
        """Collective: every rank calls it; all get an instance or all get None."""
        ok, obj = 1, None
        try:
            obj = cls(grads, params, ema, group)
        except Exception as e:   # noqa: BLE001 -- any failure means 'use NCCL', decided jointly below
            log.warning("fused data-parallel step unavailable on rank %s: %s", dist.get_rank(group), e)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=grads.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if grads.is_cuda:
            torch.cuda.synchronize(grads.device)
        return obj if int(flag.item()) == 1 else None

    def note_graph_step(self):
        """One iteration ran from a captured graph (the kernel took its epoch from the device step state): keep the
        host's epoch and the health poll in step."""
        self.epoch += 1
        if self.epoch % self.check_every == 0:
            self._poll()

    def opt_ema_step(self, m, v, step, ema_step, kind="adam", lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                     momentum=0.9, ema_alpha=0.999, count=True):
        """Updates self.params / self.ema (every rank ends with identical buffers); m, v: this rank's optimiser state.
        count=False: the call is being recorded for / run under the device step state, which carries the epoch; the
        host's copy is advanced by note_graph_step."""
        from .._lib import OptCfg, check, ptr, stream_ptr
        cfg = OptCfg()
        cfg.kind = 0 if kind == "adam" else 1
        cfg.lr, cfg.beta1, cfg.beta2, cfg.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        cfg.weight_decay, cfg.momentum, cfg.ema_alpha = float(weight_decay), float(momentum), float(ema_alpha)
        cfg.grad_scale = 1.0 / self.world
        cfg.step, cfg.ema_step = int(step), int(ema_step)
        epoch = self.epoch + 1
        if count:
            self.epoch = epoch
        check(self.lib.bsed_dp_opt_ema_step(self.h, self.rank, self.world, self.peer_grads, self.peer_params,
                                            self.peer_ema if self.ema is not None else None, self.peer_flags, epoch,
                                            ptr(m), ptr(v), self.params.numel(), C.byref(cfg), stream_ptr()),
              "bsed_dp_opt_ema_step")
        if count and self.epoch % self.check_every == 0:
            self._poll()

    def owned_slice(self):
        """[lo, hi) of the flat buffers this rank reduces and updates (csrc/head.cu: dp_opt_ema_step)."""
        n = self.grads.numel()
        chunk = -(-n // self.world)
        chunk = (chunk + 3) // 4 * 4
        lo = min(n, self.rank * chunk)
        return lo, min(n, lo + chunk), chunk

    def gather_owned(self, t):
        """Full-length copy of a per-rank-sharded flat tensor (optimiser moments): every rank contributes its slice."""
        lo, hi, chunk = self.owned_slice()
        mine = torch.zeros(chunk, dtype=t.dtype, device=t.device)
        mine[:hi - lo].copy_(t[lo:hi])
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        return torch.cat(parts)[:t.numel()]

    def _poll(self):
        """Non-blocking health check: read the error flag copied to pinned memory by an EARLIER poll (if that copy has
        landed), then queue the next copy behind the work already enqueued.  Never stalls the host."""
        if self._hflag is None:
            self._hflag = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._hflag_ev = None
        if self._hflag_ev is not None and self._hflag_ev.query():
            if int(self._hflag[0]) != 0:
                self._raise()
            self._hflag_ev = None
        if self._hflag_ev is None:
            self._hflag.copy_(self.flags[33:34], non_blocking=True)
            self._hflag_ev = torch.cuda.Event()
            self._hflag_ev.record()

    def timed_out(self):
        """True if a spin gave up (a peer never arrived); host sync."""
        return bool(int(self.flags[33].item()))

    def check(self):
        """Host sync; raises if the exchange kernel ever timed out on this rank (updates after that were dropped)."""
        if self.timed_out():
            self._raise()

    def _raise(self):
        raise RuntimeError(
            f"fused data-parallel step: rank {self.rank} waited longer than BSED_DP_TIMEOUT_S for a peer (by step {self.epoch}); "
            "no update was applied after the timeout -- the replicas are no longer in step.  Restart from the last "
            "checkpoint, or set BSED_DP=nccl.")
