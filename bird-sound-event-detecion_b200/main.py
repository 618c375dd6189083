"""Mean-teacher training step with the reference's entry points (src/main.py):

    update_ema_variables(model, ema_model, alpha, global_step)                    (:86-100)
    adjust_learning_rate(optimizer, rampup_value, ...)                            (:51-83)
    train_mt(train_loader, syn_loader, model, optimizer, c_epoch, ema_model, ema_predictor,
             mask_weak, mask_strong, adjust_lr, discriminator, optimizer_d, predictor,
             optimizer_crnn, ISP) -> loss                                         (:163-527)

Two execution paths, both entirely in libbsed.so kernels:
  * fused (optimizer is a `FusedAdam`): `MeanTeacherTrainer.step` batches the three model calls of one
    iteration (student-synthetic, student-real, teacher-real; BatchNorm statistics stay per call) in
    the same launches, and runs loss -> backward -> [gradient all-reduce] -> Adam + EMA over flat
    buffers without host synchronisation.
  * generic (any torch.optim optimizer): the reference's statement order through the autograd
    wrappers of models/CRNN.py.
With `discriminator` (a DA.cdan_frame.ConditionalDomainAdversarialLoss around a Clip_Discriminator), `optimizer_d` and
`optimizer_crnn`, every iteration first runs the adversarial update of src/main_scmt_ada_weak_seperate.py:314-335
(student forward on both domains -> gradient reversal -> discriminator -> BCE -> backward -> both optimisers step).
With ISP=True (the shift-consistency / SCT branch of src/main_baseline.py:229-277,372-529) `ShiftConsistencyTrainer.step`
runs the nine model calls of one iteration (student: synthetic, real, time-shifted and frequency-shifted real and
synthetic; teacher: real, time-shifted real, frequency-shifted real), the twelve loss terms and one optimiser + EMA
update, all in libbsed.so kernels.
"""
import logging
import os
import random
import time

import torch
from torch import nn

from . import engine
from .data import config as cfg
from .models.CRNN import CRNN, Predictor, _dropout_state
from .utilities import ramps, shard

log = logging.getLogger("bsed_b200.train")


def adjust_learning_rate(optimizer, rampup_value, rampdown_value=1, optimizer_d=None, optimizer_crnn=None,
                         c_epoch=None, rampup_value_adv=None, step_decay=False):
    """lr = rampup * rampdown * max_lr; the d / crnn optimizers get 0.1 x (src/main.py:51-83).
    step_decay=True is the variant of src/main_baseline.py:53-90: after epoch 100 the rate is halved once and then
    again every 20 epochs (:72-73, `lr * 0.5 ** (1 + (c_epoch - 100) // 20)`); it needs c_epoch, as there."""
    lr = rampup_value * rampdown_value * cfg.max_learning_rate
    if step_decay:
        if c_epoch is None:
            raise TypeError("adjust_learning_rate(step_decay=True) needs c_epoch (src/main_baseline.py:72 compares it)")
        if c_epoch > 100:
            lr = lr * (0.5 ** (1 + ((c_epoch - 100) // 20)))
    for group in optimizer.param_groups:
        group['lr'] = lr
    for opt in (optimizer_d, optimizer_crnn):
        if opt is not None:
            for group in opt.param_groups:
                group['lr'] = lr * 0.1


def update_ema_variables(model, ema_model, alpha, global_step, flavour="state_dict"):
    """ema = a * ema + (1 - a) * model, a = min(1 - 1/(global_step+1), alpha).
    flavour "state_dict" (src/main.py:86-100): every state-dict entry, i.e. parameters, BatchNorm
    running statistics and the int64 num_batches_tracked (blended in fp32, truncated on load);
    flavour "params" (src/main_origin.py:85-89): parameters only."""
    p, bn, nbt = model.flat_tensors()
    ep, ebn, enbt = ema_model.flat_tensors()
    engine.ema_buffers(p, ep, None, None, global_step, alpha)
    if flavour == "state_dict" and bn.numel():
        engine.ema_buffers(bn, ebn, nbt, enbt, global_step, alpha)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam-compatible front (same constructor arguments and param_groups) whose update is
    the fused Adam(+EMA) kernel over the flat parameter buffer (csrc/head.cu: opt_ema_kernel).
    Reference: torch.optim.Adam(lr, betas=(0.9, 0.999), eps 1e-8, wd 0)  (src/main.py:823-828)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.fused_step = 0
        self._trainer = None
        self._pending_state = None      # a state dict loaded before train_mt built the trainer

    def step(self, closure=None):
        raise RuntimeError("FusedAdam is driven by MeanTeacherTrainer.step / train_mt")

    def _attach(self, trainer):
        self._trainer = trainer
        if self._pending_state is not None:
            trainer.load_optimizer_state(self._pending_state, self._flat_params())
            self._pending_state = None

    def _flat_params(self):
        return [p for g in self.param_groups for p in g["params"]]

    def state_dict(self):
        """torch.optim.Adam's format -- {'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]} -- so the
        reference's `optim.load_state_dict(state['optimizer']['state_dict'])` (src/main.py:866-869 layout) accepts it.
        The moments live in the trainer's flat buffers (sharded per rank under the fused data-parallel step: gathered here,
        a collective)."""
        sd = super().state_dict()
        if self._trainer is not None:
            sd["state"] = self._trainer.optimizer_state(self._flat_params())
        elif self._pending_state is not None:
            sd["state"] = self._pending_state["state"]
        return sd

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(g["params"]) != len(mine["params"]) for g, mine in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has a different parameter-group layout")
        for g, mine in zip(groups, self.param_groups):
            mine.update({k: v for k, v in g.items() if k != "params"})
        if self._trainer is not None:
            self._trainer.load_optimizer_state(state_dict, self._flat_params())
        else:
            self._pending_state = state_dict


def _rehome(modules, device):
    """Put the flat parameter buffers of several modules back to back in one tensor."""
    sizes = [m._flat.numel() for m in modules]
    joint = torch.empty(sum(sizes), dtype=torch.float32, device=device)
    o = 0
    for m, n in zip(modules, sizes):
        m._reflatten(flat=joint[o:o + n])
        o += n
    return joint, sizes


class _TrainerHealth:
    def check_health(self):
        """Host sync: raises if the fused data-parallel exchange ever timed out on this rank (utilities/shard.py)."""
        if getattr(self, "dp", None) is not None:
            self.dp.check()

    # ---- optimiser state in torch.optim.Adam's layout (checkpoints: utilities/checkpoint.py)
    def _param_slices(self, params):
        """(offset, numel) of each optimiser parameter inside the flat buffer (they are views into it)."""
        base = self.params.data_ptr()
        out = []
        for p in params:
            off = (p.data_ptr() - base) // 4
            if (p.data_ptr() - base) % 4 or off < 0 or off + p.numel() > self.params.numel():
                raise RuntimeError("an optimiser parameter does not live in the trainer's flat parameter buffer")
            out.append((off, p.numel()))
        return out

    def _full_moments(self):
        """(m, v) over the whole flat buffer.  Under the fused data-parallel step every rank owns the moments of its slice
        only: gathered here (collective over the process group)."""
        if getattr(self, "dp", None) is None:
            return self.m, self.v
        return self.dp.gather_owned(self.m), self.dp.gather_owned(self.v)

    def optimizer_state(self, params):
        m, v = self._full_moments()
        if self.opt_step == 0:
            return {}
        step = torch.tensor(float(self.opt_step))
        return {i: {"step": step.clone(), "exp_avg": m[o:o + k].view_as(p).clone(), "exp_avg_sq": v[o:o + k].view_as(p).clone()}
                for i, (p, (o, k)) in enumerate(zip(params, self._param_slices(params)))}

    def load_optimizer_state(self, state_dict, params):
        st = state_dict.get("state", {})
        self.m.zero_()
        self.v.zero_()
        steps = set()
        for i, (o, k) in enumerate(self._param_slices(params)):
            e = st.get(i, st.get(str(i)))
            if e is None:
                continue
            self.m[o:o + k].copy_(e["exp_avg"].reshape(-1))
            self.v[o:o + k].copy_(e["exp_avg_sq"].reshape(-1))
            steps.add(int(float(e["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter Adam steps differ ({sorted(steps)}): the fused update keeps one step counter")
        self.opt_step = steps.pop() if steps else 0


class MeanTeacherTrainer(_TrainerHealth):
    """One fused mean-teacher iteration (src/main.py:190-523, pretrain stage, -mt)."""

    def __init__(self, model, predictor, ema_model=None, ema_predictor=None, lr=cfg.default_learning_rate,
                 betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, n_syn=cfg.batch_size, n_real=cfg.batch_size,
                 ema_flavour="state_dict", dropout_seed=2023, process_group=None, precision=None, opt_kind="adam",
                 momentum=0.9, graph=None):
        """opt_kind "adam" (src/main.py:823-828) or "sgd" = SGD with Nesterov momentum, the optimiser of the adaptation
        scripts (src/main_scmt_ada_weak_seperate.py:858-866: lr, momentum .9, weight_decay 1e-4, nesterov)."""
        assert isinstance(model, CRNN) and isinstance(predictor, Predictor)
        assert opt_kind in ("adam", "sgd")
        self.opt_kind, self.momentum = opt_kind, momentum
        self.model, self.predictor, self.ema_model, self.ema_predictor = model, predictor, ema_model, ema_predictor
        dev = model._flat.device
        if dev.type != "cuda":
            raise RuntimeError("move the models to the GPU before building the trainer")
        self.device = dev
        self.params, sizes = _rehome([model, predictor], dev)
        self.n_crnn, self.n_pred = sizes
        self.has_teacher = ema_model is not None
        if self.has_teacher:
            self.ema_params, _ = _rehome([ema_model, ema_predictor], dev)
        self.grads = torch.zeros_like(self.params)
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.n_syn, self.n_real = n_syn, n_real
        self.ema_flavour = ema_flavour
        self.dropout_seed = dropout_seed
        self.opt_step = 0
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if (
            torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        # N > 1: the gradient exchange is fused with the optimiser + EMA in one kernel over NVLink peer memory (rank r
        # reduces, updates and publishes slice r of the flat buffers)
        # (utilities/shard.py: FusedDataParallel); BSED_DP=nccl keeps the NCCL all-reduce + separate optimiser kernel
        self.dp = None
        if self.world > 1 and os.environ.get("BSED_DP", "fused").lower() == "fused":
            self.dp = shard.FusedDataParallel.create(self.grads, self.params, self.ema_params if self.has_teacher else None,
                                                     process_group)
        B = n_syn + n_real + (n_real if self.has_teacher else 0)
        self.B = B
        self.plan = engine.Plan(engine.make_cfg(**model.cfg_kwargs), max_clips=B, device=dev,
                                precision=precision or model.precision)
        assert self.plan.n_params == self.n_crnn and self.plan.n_pred_params == self.n_pred
        self.x = torch.empty(B, 1, cfg.max_frames, cfg.n_mels, dtype=torch.float32, device=dev)
        self.enc = torch.empty(B, self.plan.t_out, 256, dtype=torch.float32, device=dev)
        self.d_enc = torch.zeros(B, self.plan.t_out, 256, dtype=torch.float32, device=dev)
        self.last = {}
        # CUDA-graph path (see step): static target buffer, device-resident step state, one graph per batch shape
        self.use_graph = (os.environ.get("BSED_GRAPH", "1") != "0") if graph is None else bool(graph)
        self.ts_static = torch.zeros(n_syn, self.plan.t_out, self.plan.n_class, dtype=torch.float32, device=dev)
        self._state = None
        self._dev_gs = self._dev_opt = self._dev_lr = self._dev_dp = None
        self._graphs, self._seen, self.graph_launches = {}, set(), {}

    def _assert_homed(self):
        """The modules' parameters must still be the views into this trainer's joint buffers (a later model.to() /
        .cuda() / .float() re-allocates them and would silently detach the optimiser from the forward pass)."""
        pairs = [(self.model, self.params, 0), (self.predictor, self.params, self.n_crnn)]
        if self.has_teacher:
            pairs += [(self.ema_model, self.ema_params, 0), (self.ema_predictor, self.ema_params, self.n_crnn)]
        for mod, buf, off in pairs:
            if mod.flat_tensors()[0].data_ptr() != buf.data_ptr() + 4 * off:
                raise RuntimeError(f"{type(mod).__name__} no longer lives in the trainer's parameter buffer (moved with .to() / "
                                   ".cuda() after the trainer was built?): build the trainer after the last move")

    def step(self, x, x_ema, xs, ts, global_step, rampup_length, max_consistency_cost=cfg.max_consistency_cost,
             target_weak=None, dropout_step=None):
        """x / x_ema: real batch (student / teacher inputs), xs / ts: synthetic batch and its strong
        targets; all CUDA tensors.  Returns the 4 loss terms as a device tensor
        [strong_bce, weak_bce, cons_strong, cons_weak] (no host sync).
        target_weak (n_real, C): weak labels of the real batch -- adds BCE(weak_pred, target_weak) to the weak term, the
        loss of the weak-label scripts (src/main_scmt_ada_weak_seperate.py:437-445).
        The batch sizes are read from the tensors: anything up to the (n_syn, n_real) the trainer was built for runs in
        the same plan and buffers (the reference's loaders have no drop_last, so the last batch of an epoch is short).

        With `use_graph` (default; BSED_GRAPH=0 turns it off) the iteration runs as ONE CUDA-graph launch from the second
        call with a given batch shape on: everything that changes between iterations (dropout draw, Adam bias corrections,
        EMA coefficient, consistency weight, data-parallel epoch) lives in a device-resident step state that a one-thread
        kernel at the head of the graph advances (include/bsed.h: bsed_step_state).  The returned tensors are then static
        buffers that the next call overwrites."""
        ns, nr = int(xs.shape[0]), int(x.shape[0])
        if ns < 1 or nr < 1 or x_ema.shape[0] != nr or ts.shape[0] != ns:
            raise ValueError(f"mean-teacher step: {ns} synthetic / {nr} real / {x_ema.shape[0]} teacher clips, {ts.shape[0]} targets")
        if ns > self.n_syn or nr > self.n_real:
            raise ValueError(f"mean-teacher step: batch of {ns} + {nr} clips exceeds the {self.n_syn} + {self.n_real} this trainer "
                             "was built for (build it with the loaders' batch_size)")
        self._assert_homed()
        nst = ns + nr
        B = nst + (nr if self.has_teacher else 0)
        self.x[:ns].copy_(xs.reshape(ns, 1, cfg.max_frames, cfg.n_mels))
        self.x[ns:nst].copy_(x.reshape(nr, 1, cfg.max_frames, cfg.n_mels))
        if self.has_teacher:
            self.x[nst:B].copy_(x_ema.reshape(nr, 1, cfg.max_frames, cfg.n_mels))
        graphable = (self.use_graph and target_weak is None and dropout_step is None
                     and (self.world == 1 or self.dp is not None))
        if not graphable:
            out = self._body(ns, nr, ts.contiguous().float(), global_step, rampup_length, max_consistency_cost, target_weak,
                             dropout_step, device_state=False)
        else:
            self.ts_static[:ns].copy_(ts)
            self._sync_state(global_step)
            key = (ns, nr, int(rampup_length), float(max_consistency_cost))
            if key in self._graphs:
                g, out = self._graphs[key]
                g.replay()
            elif key in self._seen:          # second iteration with this shape: capture it, then run it
                from . import _lib
                lib = _lib.load()
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                n0 = lib.bsed_launch_count()
                with torch.cuda.graph(g):
                    out = self._body(ns, nr, self.ts_static[:ns], global_step, rampup_length, max_consistency_cost, None, None,
                                     device_state=True)
                self.graph_launches[key] = int(lib.bsed_launch_count() - n0)
                self._graphs[key] = (g, out)
                g.replay()
            else:
                self._seen.add(key)
                out = self._body(ns, nr, self.ts_static[:ns], global_step, rampup_length, max_consistency_cost, None, None,
                                 device_state=True)
            self._dev_gs = global_step
            if self.dp is not None:
                self.dp.note_graph_step()
        self.opt_step += 1
        self.last = out
        return out["losses"]

    # ---- device-resident step state (CUDA-graph path)
    def _step_cfg(self, rampup_length, max_consistency_cost):
        from ._lib import StepCfg
        c = StepCfg()
        c.dropout_seed, c.key_mul, c.key_add = int(self.dropout_seed), 1, 0
        c.beta1, c.beta2, c.ema_alpha = float(self.betas[0]), float(self.betas[1]), 0.999
        c.max_consistency_cost, c.rampup_length = float(max_consistency_cost), int(rampup_length)
        return c

    def _sync_state(self, global_step):
        """Make the device counters agree with the host's view before an iteration: normally they already do (each
        iteration advances them by one on the device); a jump of global_step, a changed learning rate or an optimiser
        state loaded from a checkpoint is written through with one small copy."""
        import ctypes as C
        from ._lib import StepState
        dp_epoch = self.dp.epoch if self.dp is not None else 0
        want = (global_step - 1, self.opt_step, float(self.lr), dp_epoch)
        if self._state is None:
            self._state = torch.zeros(C.sizeof(StepState), dtype=torch.uint8, device=self.device)
            self._dev_view = None
        have = (self._dev_gs, self._dev_opt, self._dev_lr, self._dev_dp)
        if have != want:
            hs = StepState()
            hs.global_step, hs.opt_step, hs.dp_epoch, hs.lr = global_step - 1, self.opt_step, dp_epoch, float(self.lr)
            self._state.copy_(torch.frombuffer(bytearray(bytes(hs)), dtype=torch.uint8))
        # what the device will hold after this iteration's advance
        self._dev_gs, self._dev_opt, self._dev_lr = global_step, self.opt_step + 1, float(self.lr)
        self._dev_dp = dp_epoch + 1 if self.dp is not None else 0     # without data parallelism the host never counts it

    def _body(self, ns, nr, ts, global_step, rampup_length, max_consistency_cost, target_weak, dropout_step, device_state):
        """The kernel sequence of one iteration.  device_state: the per-iteration scalars come from the device-resident
        step state (advanced here), not from the host values passed along -- the form that can be captured."""
        from . import _lib
        nst = ns + nr
        B = nst + (nr if self.has_teacher else 0)
        m, p = self.model, self.predictor
        sp, sbn, snbt = m.flat_tensors()
        groups = [dict(params=sp, bn=sbn, nbt=snbt, n=ns), dict(params=sp, bn=sbn, nbt=snbt, n=nr)]
        if self.has_teacher:
            tp, tbn, tnbt = self.ema_model.flat_tensors()
            groups.append(dict(params=tp, bn=tbn, nbt=tnbt, n=nr))
        lib, h = _lib.load(), _lib.handle(self.device.index)
        if device_state:
            sptr = _lib.C.c_void_p(self._state.data_ptr())
            _lib.check(lib.bsed_step_state_advance(h, sptr, _lib.C.byref(self._step_cfg(rampup_length, max_consistency_cost)),
                                                   _lib.stream_ptr()), "bsed_step_state_advance")
            _lib.check(lib.bsed_set_step_state(h, sptr), "bsed_set_step_state")
        try:
            self.plan.forward(groups, self.x[:B], train=True, save=True, seed=self.dropout_seed,
                              step=global_step if dropout_step is None else dropout_step, enc=self.enc[:B])
            pp = self.params[self.n_crnn:]
            logits, strong, weak = self.plan.predictor_forward(pp, self.enc[:nst])
            cons_w = 0.0
            if self.has_teacher:
                _, strong_ema, weak_ema = self.plan.predictor_forward(self.ema_params[self.n_crnn:], self.enc[nst:B])
                cons_w = max_consistency_cost * ramps.exp_rampup(global_step, rampup_length)
            if target_weak is not None:
                from ._lib import LOSS_BCE_STRONG as BS, LOSS_BCE_WEAK as BW, LOSS_MSE_STRONG as MS, LOSS_MSE_WEAK as MW
                terms = [dict(kind=BS, pred_first=0, n=ns, ref=ts, slot=0),
                         dict(kind=BW, pred_first=0, n=ns, ref=ts, ref_is_strong=True, slot=1),
                         dict(kind=BW, pred_first=ns, n=nr, ref=target_weak.contiguous().float(), slot=1)]
                if self.has_teacher:
                    terms += [dict(kind=MS, pred_first=ns, n=nr, ref=strong_ema, weight=float(cons_w), slot=2),
                              dict(kind=MW, pred_first=ns, n=nr, ref=weak_ema, weight=float(cons_w), slot=3)]
                losses, d_strong, d_weak = engine.loss_terms(strong, weak, terms, 4)
            elif self.has_teacher:
                losses, d_strong, d_weak = engine.mt_loss(strong, weak, 0, ns, ts, ns, nr, strong_ema, weak_ema, cons_w)
            else:
                losses, d_strong, d_weak = engine.mt_loss(strong, weak, 0, ns, ts, 0, 0, None, None, 0.0)
            self.plan.predictor_backward(pp, self.enc[:nst], logits, strong, weak, d_strong, d_weak,
                                         self.grads[self.n_crnn:], accumulate=False, d_enc=self.d_enc[:nst])
            self.plan.backward(0b011, self.d_enc, self.grads[:self.n_crnn], accumulate=False)
            ema = self.ema_params if self.has_teacher else None
            opt_step = self.opt_step + 1
            if self.dp is not None:    # all-reduce over peer memory + Adam / SGD + EMA in one kernel
                self.dp.opt_ema_step(self.m, self.v, step=opt_step, ema_step=global_step + 1, kind=self.opt_kind, lr=self.lr,
                                     betas=self.betas, eps=self.eps, weight_decay=self.weight_decay, momentum=self.momentum,
                                     count=not device_state)
            else:
                grad_scale = shard.allreduce_gradients(self.grads, self.pg)   # NCCL sum over NVLink; 1/N folded below
                engine.opt_ema_step(self.params, self.grads, self.m, self.v, ema, step=opt_step, ema_step=global_step + 1,
                                    kind=self.opt_kind, lr=self.lr, betas=self.betas, eps=self.eps,
                                    weight_decay=self.weight_decay, momentum=self.momentum, grad_scale=grad_scale)
            if self.has_teacher and self.ema_flavour == "state_dict":
                engine.ema_buffers(sbn, tbn, snbt, tnbt, global_step + 1)
        finally:
            if device_state:
                _lib.check(lib.bsed_set_step_state(h, None), "bsed_set_step_state")
        return dict(strong=strong, weak=weak, losses=losses)


class FusedSGD(torch.optim.Optimizer):
    """torch.optim.SGD(lr, momentum, weight_decay, nesterov=True)-compatible front of the fused SGD-Nesterov(+EMA) kernel
    (csrc/head.cu: opt_ema_kernel), the optimiser of the adaptation scripts
    (src/main_scmt_ada_weak_seperate.py:858-870: optim, optim_crnn, optim_d)."""

    def __init__(self, params, lr=1e-3, momentum=0.9, weight_decay=0, nesterov=True, dampening=0):
        if not nesterov or dampening != 0:
            raise NotImplementedError("the fused kernel implements SGD with Nesterov momentum and no dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=True, dampening=0))
        self._trainer = None

    def step(self, closure=None):
        raise RuntimeError("FusedSGD is driven by AdaptationTrainer.step / train_mt")


class AdaptationTrainer(MeanTeacherTrainer):
    """One fused iteration of the SCMT + adversarial-domain-adaptation loop (BASELINE.json configs[2];
    src/main_scmt_ada_weak_seperate.py:train_mt with a discriminator):

      1. adversarial update (:314-335): student forward of the synthetic and the real batch (two model calls) ->
         gradient reversal (src/DA/grl.py) -> Clip_Discriminator -> BCE against the domain labels (src/DA/cdan_frame.py)
         -> backward through D and, reversed, through the encoder -> optimizer_crnn.step(), optimizer_d.step()
      2. the mean-teacher update (:337-521): the three model calls, strong / weak BCE on the synthetic batch, weak BCE
         of the real batch against its weak labels, MSE consistency with the teacher -> optimizer.step() -> EMA
    with all three optimisers SGD-Nesterov (:858-870).  Everything runs in libbsed.so kernels over flat buffers; under
    data parallelism the three gradient exchanges (encoder + predictor, encoder-adversarial, discriminator) each go
    through the fused peer-memory reduce + update kernel (utilities/shard.py)."""

    def __init__(self, model, predictor, ema_model, ema_predictor, domain_loss, lr=cfg.default_learning_rate, lr_adv=None,
                 momentum=0.9, weight_decay=1e-4, n_syn=cfg.batch_size, n_real=cfg.batch_size, dropout_seed=2023,
                 process_group=None, precision=None, disc_precision=None):
        super().__init__(model, predictor, ema_model, ema_predictor, lr=lr, weight_decay=weight_decay, n_syn=n_syn,
                         n_real=n_real, dropout_seed=dropout_seed, process_group=process_group, precision=precision,
                         opt_kind="sgd", momentum=momentum)
        # contractions of the discriminator: `disc_precision`, else BSED_DISC_PRECISION, else the plan's mode when that is
        # "tf32x3" (error-compensated forward / data-gradient GEMMs on the tensor cores with single-pass weight-gradient
        # reductions, as in the CRNN), else the module's own setting ("fp32" unless changed; "tf32" = single pass everywhere)
        self.disc_precision = (disc_precision or os.environ.get("BSED_DISC_PRECISION")
                               or ("tf32x3" if self.plan.precision == "tf32x3" else None))
        self.domain_loss = domain_loss                      # DA.cdan_frame.ConditionalDomainAdversarialLoss
        self.disc = domain_loss.domain_discriminator
        self.lr_adv = lr if lr_adv is None else lr_adv
        dev = self.device
        d_flat, _, _ = self.disc.flat_tensors()
        if d_flat.device != dev:
            raise RuntimeError("move the discriminator to the trainer's device before building the trainer")
        self.grads_adv = torch.zeros(self.n_crnn, dtype=torch.float32, device=dev)
        self.m_adv = torch.zeros(self.n_crnn, dtype=torch.float32, device=dev)
        self.grads_d = torch.zeros_like(d_flat)
        self.m_d = torch.zeros_like(d_flat)
        self.adv_step = 0
        nst = n_syn + n_real
        self.dx = torch.empty(nst, self.plan.t_out, 256, dtype=torch.float32, device=dev)
        self.prob = torch.empty(nst, dtype=torch.float32, device=dev)
        self.d_prob = torch.empty(nst, dtype=torch.float32, device=dev)
        self.dom_loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.dp_adv = self.dp_d = None
        if self.dp is not None:
            self.dp_adv = shard.FusedDataParallel.create(self.grads_adv, self.params[:self.n_crnn], None, process_group)
            self.dp_d = shard.FusedDataParallel.create(self.grads_d, d_flat, None, process_group)

    def adversarial_update(self, x, xs, global_step):
        """src/main_scmt_ada_weak_seperate.py:314-335.  Returns the domain loss (device scalar, no host sync)."""
        from . import _lib
        from ._lib import check, ptr, stream_ptr
        lib, h = _lib.load(), _lib.handle(self.device.index)
        ns, nr = int(xs.shape[0]), int(x.shape[0])
        nst = ns + nr
        self.x[:ns].copy_(xs.reshape(ns, 1, cfg.max_frames, cfg.n_mels))
        self.x[ns:nst].copy_(x.reshape(nr, 1, cfg.max_frames, cfg.n_mels))
        sp, sbn, snbt = self.model.flat_tensors()
        S = lambda k: dict(params=sp, bn=sbn, nbt=snbt, n=k)
        enc = self.enc[:nst]
        self.plan.forward([S(ns), S(nr)], self.x[:nst], train=True, save=True, seed=self.dropout_seed, step=2 * global_step + 1,
                          enc=enc)
        grl = self.domain_loss.grl
        coeff = grl.coeff()
        if grl.auto_step:
            grl.step()
        d = self.disc
        if not d.training:
            raise RuntimeError("the adversarial update needs the discriminator in train() mode (batch statistics)")
        d_flat, d_bn, d_nbt = d.flat_tensors()
        ws, wsb = d._workspace(nst)
        check(lib.bsed_disc_set_precision(h, _lib.PRECISIONS[(self.disc_precision or d.precision or "fp32").lower()]),
              "bsed_disc_set_precision")
        prob, d_prob = self.prob[:nst], self.d_prob[:nst]
        check(lib.bsed_disc_forward(h, ptr(d_flat), ptr(d_bn), ptr(d_nbt), ptr(enc), nst, 1, ptr(prob), ptr(ws), wsb, stream_ptr()),
              "bsed_disc_forward")
        key = (ns, nr)
        if getattr(self, "_labels_key", None) != key:
            self._labels = torch.cat((torch.ones(ns, device=self.device), torch.zeros(nr, device=self.device)))
            self._labels_key = key
        check(lib.bsed_disc_bce(h, ptr(prob), ptr(self._labels), nst, ptr(self.dom_loss), ptr(d_prob), stream_ptr()), "bsed_disc_bce")
        dx = self.dx[:nst]
        check(lib.bsed_disc_backward(h, ptr(d_flat), ptr(prob), ptr(d_prob), nst, ptr(self.grads_d), 0, ptr(dx), ptr(ws), wsb,
                                     stream_ptr()), "bsed_disc_backward")
        # gradient reversal: -coeff * grad (src/DA/grl.py:28-31)
        check(lib.bsed_scale_f32(h, ptr(self.d_enc[:nst]), ptr(dx), dx.numel(), -float(coeff), stream_ptr()), "bsed_scale_f32")
        self.plan.backward(0b011, self.d_enc, self.grads_adv, accumulate=False)
        self.adv_step += 1
        sgd = dict(kind="sgd", lr=self.lr_adv, weight_decay=self.weight_decay, momentum=self.momentum)
        if self.dp_adv is not None:
            self.dp_adv.opt_ema_step(self.m_adv, None, step=self.adv_step, ema_step=self.adv_step, **sgd)
            self.dp_d.opt_ema_step(self.m_d, None, step=self.adv_step, ema_step=self.adv_step, **sgd)
        else:
            scale = shard.allreduce_gradients(self.grads_adv, self.pg)
            shard.allreduce_gradients(self.grads_d, self.pg)
            engine.opt_ema_step(self.params[:self.n_crnn], self.grads_adv, self.m_adv, None, None, step=self.adv_step,
                                grad_scale=scale, **sgd)
            engine.opt_ema_step(d_flat, self.grads_d, self.m_d, None, None, step=self.adv_step, grad_scale=scale, **sgd)
        return self.dom_loss

    def step(self, x, x_ema, target_weak, xs, ts, global_step, rampup_length, max_consistency_cost=cfg.max_consistency_cost):
        """Returns (the 4 loss terms of the mean-teacher update [strong, weak, cons_strong, cons_weak], domain loss), device
        tensors, no host sync."""
        dom = self.adversarial_update(x, xs, global_step)
        losses = super().step(x, x_ema, xs, ts, global_step, rampup_length, max_consistency_cost, target_weak=target_weak,
                              dropout_step=2 * global_step)
        return losses, dom

    def check_health(self):
        for d in (self.dp, self.dp_adv, self.dp_d):
            if d is not None:
                d.check()


ISP_SLOTS = ("strong_class", "weak_class", "cons_strong", "cons_weak", "weak_freq_shift_class", "strong_shift_class",
             "strong_freq_shift_class", "cons_shift", "cons_strong_shift", "cons_strong_freq_shift", "cons_weak_shift",
             "cons_weak_freq_shift")


class ShiftConsistencyTrainer(_TrainerHealth):
    """One fused iteration of train_mt(ISP=True) with a teacher (src/main_baseline.py:229-277, 337-584).

    Model calls (each its own BatchNorm batch statistics, in the reference's order per network):
      plan A  [synthetic | real]                         student   + [real]                        teacher
      plan B  [real>>t | real>>f | synthetic>>t | synthetic>>f]     student
      plan C  [real>>t | real>>f]                                   teacher
    (>>t: per-clip roll along time by shift_list[k]; >>f: along frequency by freq_shift_list[k]).  Student predictions
    live in one (6n, T, C) tensor, teacher predictions in one (3n, T, C) tensor, and the loss is the term list of
    bsed_loss_terms; the rolled targets / rolled detached predictions are references with a per-clip roll."""

    def __init__(self, model, predictor, ema_model, ema_predictor, lr=cfg.default_learning_rate, betas=(0.9, 0.999),
                 eps=1e-8, weight_decay=0.0, n=cfg.batch_size, dropout_seed=2023, process_group=None, precision=None,
                 pooling_time_ratio=cfg.pooling_time_ratio):
        assert isinstance(model, CRNN) and isinstance(predictor, Predictor)
        assert ema_model is not None and ema_predictor is not None, "the ISP step needs the teacher"
        self.model, self.predictor, self.ema_model, self.ema_predictor = model, predictor, ema_model, ema_predictor
        dev = model._flat.device
        if dev.type != "cuda":
            raise RuntimeError("move the models to the GPU before building the trainer")
        self.device, self.n = dev, n
        self.params, sizes = _rehome([model, predictor], dev)
        self.n_crnn, self.n_pred = sizes
        self.ema_params, _ = _rehome([ema_model, ema_predictor], dev)
        self.grads = torch.zeros_like(self.params)
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.dropout_seed, self.opt_step, self.pg = dropout_seed, 0, process_group
        self.ptr = int(pooling_time_ratio)
        self.world = torch.distributed.get_world_size(process_group) if (
            torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        self.dp = None
        if self.world > 1 and os.environ.get("BSED_DP", "fused").lower() == "fused":
            self.dp = shard.FusedDataParallel.create(self.grads, self.params, self.ema_params, process_group)
        mk = lambda B: engine.Plan(engine.make_cfg(**model.cfg_kwargs), max_clips=B, device=dev,
                                   precision=precision or model.precision)
        self.planA, self.planB, self.planC = mk(3 * n), mk(4 * n), mk(2 * n)
        T, F, To = cfg.max_frames, cfg.n_mels, self.planA.t_out
        f32 = dict(dtype=torch.float32, device=dev)
        self.xA = torch.empty(3 * n, 1, T, F, **f32)
        self.xB = torch.empty(4 * n, 1, T, F, **f32)
        self.xC = torch.empty(2 * n, 1, T, F, **f32)
        # encoder outputs / gradients: [A: syn, real, teacher | B: 4n | C: 2n]
        self.enc = torch.empty(9 * n, To, 256, **f32)
        self.d_enc = torch.zeros(9 * n, To, 256, **f32)
        C, ldl = self.planA.n_class, self.planA.ldl
        self.s_logits, self.s_strong, self.s_weak = (torch.empty(6 * n, To, ldl, **f32), torch.empty(6 * n, To, C, **f32),
                                                     torch.empty(6 * n, C, **f32))
        self.t_logits, self.t_strong, self.t_weak = (torch.empty(3 * n, To, ldl, **f32), torch.empty(3 * n, To, C, **f32),
                                                     torch.empty(3 * n, C, **f32))
        self.last = {}

    def step(self, x, x_ema, target_weak, xs, ts, shift_list, freq_shift_list, global_step, rampup_value,
             max_consistency_cost=cfg.max_consistency_cost):
        """x / x_ema (n,1,T,F): real batch, student / teacher inputs; target_weak (n,C): its weak labels; xs / ts: synthetic
        batch and strong targets; shift_list (frames, multiples of pooling_time_ratio) / freq_shift_list (bins): per-clip
        shifts.  Returns the 12 loss terms (ISP_SLOTS order, device tensor, no host sync).  The batch size is read from
        the tensors (any n up to the one the trainer was built for; the reference's loaders have no drop_last); like the
        reference's loop (one shift list serves both batches, src/main_baseline.py:232-277) it needs as many synthetic
        as real clips."""
        n, dev = int(x.shape[0]), self.device
        if n < 1 or xs.shape[0] != n or x_ema.shape[0] != n or len(shift_list) != n or len(freq_shift_list) != n:
            raise ValueError(f"shift-consistency step: {n} real / {xs.shape[0]} synthetic / {x_ema.shape[0]} teacher clips, "
                             f"{len(shift_list)} / {len(freq_shift_list)} shifts -- the batches must be equally long")
        if n > self.n:
            raise ValueError(f"shift-consistency step: batch of {n} clips exceeds the {self.n} this trainer was built for")
        T, F = cfg.max_frames, cfg.n_mels
        st = torch.tensor([int(v) for v in shift_list], dtype=torch.int32, device=dev)
        sf = torch.tensor([int(v) for v in freq_shift_list], dtype=torch.int32, device=dev)
        pool_shift = torch.tensor([int(v / self.ptr) for v in shift_list], dtype=torch.int32, device=dev)
        xr = x.reshape(n, 1, T, F).float().contiguous()
        xe = x_ema.reshape(n, 1, T, F).float().contiguous()
        xsy = xs.reshape(n, 1, T, F).float().contiguous()
        xA, xB, xC = self.xA[:3 * n], self.xB[:4 * n], self.xC[:2 * n]
        xA[:n].copy_(xsy)
        xA[n:2 * n].copy_(xr)
        xA[2 * n:].copy_(xe)
        engine.roll_clips(xr, st, None, out=xB[:n])
        engine.roll_clips(xr, None, sf, out=xB[n:2 * n])
        engine.roll_clips(xsy, st, None, out=xB[2 * n:3 * n])
        engine.roll_clips(xsy, None, sf, out=xB[3 * n:])
        engine.roll_clips(xe, st, None, out=xC[:n])
        engine.roll_clips(xe, None, sf, out=xC[n:])
        sp, sbn, snbt = self.model.flat_tensors()
        tp, tbn, tnbt = self.ema_model.flat_tensors()
        S = lambda k: dict(params=sp, bn=sbn, nbt=snbt, n=k)
        Tg = lambda k: dict(params=tp, bn=tbn, nbt=tnbt, n=k)
        seed = self.dropout_seed
        encA, encB, encC = self.enc[:3 * n], self.enc[3 * n:7 * n], self.enc[7 * n:9 * n]
        self.planA.forward([S(n), S(n), Tg(n)], xA, train=True, save=True, seed=seed, step=3 * global_step, enc=encA)
        self.planB.forward([S(n), S(n), S(n), S(n)], xB, train=True, save=True, seed=seed, step=3 * global_step + 1,
                           enc=encB)
        self.planC.forward([Tg(n), Tg(n)], xC, train=True, save=False, seed=seed, step=3 * global_step + 2, enc=encC)
        pp, tpp = self.params[self.n_crnn:], self.ema_params[self.n_crnn:]
        s_logits, s_strong, s_weak = self.s_logits[:6 * n], self.s_strong[:6 * n], self.s_weak[:6 * n]
        t_logits, t_strong, t_weak = self.t_logits[:3 * n], self.t_strong[:3 * n], self.t_weak[:3 * n]
        sl = lambda t, a, b: t[a:b]
        outS = lambda a, b: (sl(s_logits, a, b), sl(s_strong, a, b), sl(s_weak, a, b))
        outT = lambda a, b: (sl(t_logits, a, b), sl(t_strong, a, b), sl(t_weak, a, b))
        self.planA.predictor_forward(pp, encA[:2 * n], out=outS(0, 2 * n))
        self.planA.predictor_forward(pp, encB, out=outS(2 * n, 6 * n))
        self.planA.predictor_forward(tpp, encA[2 * n:], out=outT(0, n))
        self.planA.predictor_forward(tpp, encC, out=outT(n, 3 * n))

        # student clips: syn [0,n) real [n,2n) real>>t [2n,3n) real>>f [3n,4n) syn>>t [4n,5n) syn>>f [5n,6n)
        # teacher clips: real [0,n) real>>t [n,2n) real>>f [2n,3n)
        from ._lib import LOSS_BCE_STRONG as BS, LOSS_BCE_WEAK as BW, LOSS_MSE_STRONG as MS, LOSS_MSE_WEAK as MW
        cc = float(max_consistency_cost * rampup_value)
        widx = n // 2
        ts = ts.float().contiguous()
        tw = target_weak.float().contiguous()
        Ss, Sw, Ts, Tw = s_strong, s_weak, t_strong, t_weak
        terms = [
            dict(kind=BS, pred_first=0, n=n, ref=ts, slot=0),                                            # :475
            dict(kind=BW, pred_first=0, n=n, ref=ts, ref_is_strong=True, slot=1),                        # :433-434
            dict(kind=BW, pred_first=n, n=n, ref=tw, slot=1),                                            # :437
            dict(kind=MS, pred_first=n, n=n, ref=Ts[:n], weight=cc, slot=2),                             # :489
            dict(kind=MW, pred_first=n, n=n, ref=Tw[:n], weight=cc, slot=3),                             # :494
            dict(kind=BW, pred_first=5 * n, n=n, ref=ts, ref_is_strong=True, slot=4),                    # :448
            dict(kind=BS, pred_first=4 * n, n=n, ref=ts, roll=pool_shift, slot=5),                       # :481
            dict(kind=BS, pred_first=5 * n, n=n, ref=ts, slot=6),                                        # :482
            dict(kind=MS, pred_first=4 * n, n=n, ref=Ss[:n], roll=pool_shift, weight=cc / 2, slot=7),    # :523
            dict(kind=MS, pred_first=2 * n, n=n, ref=Ss[n:2 * n], roll=pool_shift, weight=cc / 2, slot=7),
            dict(kind=MS, pred_first=2 * n, n=n, ref=Ts[n:2 * n], weight=cc, grad_weight=cc / 2, slot=8),   # :500, :529
            dict(kind=MS, pred_first=3 * n, n=n, ref=Ts[2 * n:], weight=cc, grad_weight=cc / 2, slot=9),    # :508, :529
        ]
        if widx > 0:
            terms.insert(6, dict(kind=BW, pred_first=3 * n, n=widx, ref=tw[:widx], slot=4))              # :448 (real half)
        if n - widx > 0:   # logged only (:504, :512)
            terms += [dict(kind=MW, pred_first=2 * n + widx, n=n - widx, ref=Tw[n + widx:2 * n], weight=cc, grad_weight=0.0, slot=10),
                      dict(kind=MW, pred_first=3 * n + widx, n=n - widx, ref=Tw[2 * n + widx:], weight=cc, grad_weight=0.0, slot=11)]
        losses, d_strong, d_weak = engine.loss_terms(Ss, Sw, terms, len(ISP_SLOTS))

        gp = self.grads[self.n_crnn:]
        dA, dB = self.d_enc[:3 * n], self.d_enc[3 * n:7 * n]
        self.planA.predictor_backward(pp, encA[:2 * n], s_logits[:2 * n], Ss[:2 * n], Sw[:2 * n], d_strong[:2 * n],
                                      d_weak[:2 * n], gp, accumulate=False, d_enc=dA[:2 * n])
        self.planA.predictor_backward(pp, encB, s_logits[2 * n:], Ss[2 * n:], Sw[2 * n:], d_strong[2 * n:],
                                      d_weak[2 * n:], gp, accumulate=True, d_enc=dB)
        gc = self.grads[:self.n_crnn]
        self.planA.backward(0b011, dA, gc, accumulate=False)
        self.planB.backward(0b1111, dB, gc, accumulate=True)
        self.opt_step += 1
        if self.dp is not None:    # reduce-scatter over peer memory + Adam + EMA + all-gather in one kernel
            self.dp.opt_ema_step(self.m, self.v, step=self.opt_step, ema_step=global_step + 1, kind="adam", lr=self.lr,
                                 betas=self.betas, eps=self.eps, weight_decay=self.weight_decay)
        else:
            grad_scale = shard.allreduce_gradients(self.grads, self.pg)
            engine.opt_ema_step(self.params, self.grads, self.m, self.v, self.ema_params, step=self.opt_step,
                                ema_step=global_step + 1, kind="adam", lr=self.lr, betas=self.betas, eps=self.eps,
                                weight_decay=self.weight_decay, grad_scale=grad_scale)
        engine.ema_buffers(sbn, tbn, snbt, tnbt, global_step + 1)
        self.last = dict(strong=Ss, weak=Sw, losses=losses)
        return losses

    @staticmethod
    def total(losses):
        """loss of src/main_baseline.py:517-529 from the 12 terms."""
        return losses[:8].sum() + 0.5 * (losses[8] + losses[9])


def _generic_step(model, predictor, ema_model, ema_predictor, optimizer, batch, syn_batch, global_step,
                  rampup_value):
    """The reference's statement order (src/main.py:250-254, 335-343, 376-477, 517-523) on the autograd
    wrappers; losses/gradients of the head come from the same CUDA loss kernel."""
    (x, x_ema, _target), (xs, _xs_ema, ts) = batch, syn_batch
    if ema_model is not None:
        enc_ema, _ = ema_model(x_ema)
        strong_ema, weak_ema = ema_predictor(enc_ema)
        strong_ema, weak_ema = strong_ema.detach(), weak_ema.detach()
    optimizer.zero_grad()
    enc_s, _ = model(xs)
    syn_strong, syn_weak = predictor(enc_s)
    enc, _ = model(x)
    strong, weak = predictor(enc)
    ns, nr = xs.shape[0], x.shape[0]
    cat_s = torch.cat([syn_strong, strong]).contiguous()
    cat_w = torch.cat([syn_weak, weak]).contiguous()
    cons_w = cfg.max_consistency_cost * rampup_value if ema_model is not None else 0.0
    losses, d_strong, d_weak = engine.mt_loss(cat_s.detach(), cat_w.detach(), 0, ns, ts.contiguous().float(),
                                              ns if ema_model is not None else 0, nr if ema_model is not None else 0,
                                              strong_ema if ema_model is not None else None,
                                              weak_ema if ema_model is not None else None, cons_w)
    torch.autograd.backward([cat_s, cat_w], [d_strong, d_weak])
    shard.allreduce_module_grads([model, predictor])     # data parallel (no-op on one rank)
    optimizer.step()
    if ema_model is not None:
        update_ema_variables(model, ema_model, 0.999, global_step + 1)
        update_ema_variables(predictor, ema_predictor, 0.999, global_step + 1)
    return losses


def adversarial_step(model, predictor, discriminator, optimizer_crnn, optimizer_d, batch_input, syn_batch_input):
    """src/main_scmt_ada_weak_seperate.py:314-335.  Returns the domain loss (device scalar)."""
    syn_encoded_x, syn_d_input = model(syn_batch_input)
    syn_strong_pred, _ = predictor(syn_encoded_x)
    encoded_x, d_input = model(batch_input)
    strong_pred, _ = predictor(encoded_x)
    optimizer_crnn.zero_grad()
    optimizer_d.zero_grad()
    domain_loss = discriminator(syn_strong_pred, syn_d_input, strong_pred, d_input)
    domain_loss.backward()
    shard.allreduce_module_grads([model, discriminator])     # data parallel: replicas take the same adversarial step
    optimizer_crnn.step()
    optimizer_d.step()
    return domain_loss.detach()


def train_mt(train_loader, syn_loader, model, optimizer, c_epoch, ema_model=None, ema_predictor=None, mask_weak=None,
             mask_strong=None, adjust_lr=False, discriminator=None, optimizer_d=None, predictor=None,
             optimizer_crnn=None, ISP=False):
    """One epoch of the mean-teacher model; same arguments as the reference (src/main.py:163).
    Loaders yield (((student_input, teacher_input), target), filename)."""
    if mask_weak is not None or mask_strong is not None:
        raise NotImplementedError("masked-real-label branches are outside this round's hot path")
    if ISP and (ema_model is None or not isinstance(optimizer, FusedAdam)):
        raise ValueError("ISP=True runs the fused shift-consistency step: it needs the teacher and a FusedAdam optimizer")
    if discriminator is not None and (optimizer_d is None or optimizer_crnn is None):
        raise ValueError("the adversarial update needs optimizer_d and optimizer_crnn (as in the reference)")
    if predictor is None:
        raise ValueError("train_mt needs the Predictor module (the reference passes it as `predictor`)")
    start = time.time()
    syn_iter = iter(syn_loader)
    n_syn_batches = len(syn_loader)
    losses = None
    fused = isinstance(optimizer, FusedAdam)
    if shard.world_size() > 1:
        # ranks enter an epoch together: rank-local work between epochs (validation, checkpoints on rank 0) must not eat
        # into the peer-arrival timeout of the fused data-parallel step
        torch.distributed.barrier()
    for i, data1 in enumerate(train_loader):
        try:
            data2 = next(syn_iter)
        except StopIteration:
            syn_iter = iter(syn_loader)
            data2 = next(syn_iter)
        ((batch_input, ema_batch_input), target), _ = data1
        ((syn_batch_input, syn_ema_batch_input), syn_target), _ = data2
        global_step = c_epoch * n_syn_batches + i
        rampup_len = cfg.n_epoch_rampup * n_syn_batches
        rampup_value = ramps.exp_rampup(global_step, rampup_len)
        if adjust_lr:
            # ISP=True is src/main_baseline.py's train_mt, whose adjust_learning_rate halves the rate after epoch 100
            adjust_learning_rate(optimizer, rampup_value, optimizer_d=optimizer_d, optimizer_crnn=optimizer_crnn,
                                 c_epoch=c_epoch, step_decay=bool(ISP))
        dev = model._flat.device
        x = batch_input.to(dev, non_blocking=True)
        x_ema = ema_batch_input.to(dev, non_blocking=True)
        xs = syn_batch_input.to(dev, non_blocking=True)
        ts = syn_target.to(dev, non_blocking=True)
        fused_ada = (discriminator is not None and isinstance(optimizer, FusedSGD) and isinstance(optimizer_d, FusedSGD)
                     and isinstance(optimizer_crnn, FusedSGD) and not ISP)
        if fused_ada:
            # the whole iteration of src/main_scmt_ada_weak_seperate.py (adversarial update + mean-teacher update with the
            # weak labels of the real batch) in libbsed.so kernels
            tr = optimizer._trainer
            if tr is None:
                g, ga = optimizer.param_groups[0], optimizer_crnn.param_groups[0]
                tr = AdaptationTrainer(model, predictor, ema_model, ema_predictor, discriminator, lr=g['lr'], lr_adv=ga['lr'],
                                       momentum=g['momentum'], weight_decay=g['weight_decay'], n_syn=xs.shape[0],
                                       n_real=x.shape[0], dropout_seed=_dropout_state["seed"])
                optimizer._trainer = optimizer_d._trainer = optimizer_crnn._trainer = tr
            tr.lr, tr.lr_adv = optimizer.param_groups[0]['lr'], optimizer_crnn.param_groups[0]['lr']
            tgt = target.to(dev, non_blocking=True).float()
            target_weak = tgt if tgt.dim() == 2 else tgt.max(-2)[0]
            losses, domain_loss = tr.step(x, x_ema, target_weak, xs, ts, global_step, rampup_len)
            continue
        if discriminator is not None:
            domain_loss = adversarial_step(model, predictor, discriminator, optimizer_crnn, optimizer_d, x, xs)
        if ISP:
            if xs.shape[0] != x.shape[0]:
                # the two loaders are cycled independently and have no drop_last: a short last batch on one side.  The
                # reference's loop applies one per-clip shift list to both batches (src/main_baseline.py:232-277) and
                # cannot run such a pair either; here the longer batch is cut to the shorter one
                k = min(xs.shape[0], x.shape[0])
                x, x_ema, target, xs, ts = x[:k], x_ema[:k], target[:k], xs[:k], ts[:k]
            tr = optimizer._trainer
            if tr is None:
                g = optimizer.param_groups[0]
                tr = ShiftConsistencyTrainer(model, predictor, ema_model, ema_predictor, lr=g['lr'], betas=g['betas'],
                                             eps=g['eps'], weight_decay=g['weight_decay'], n=x.shape[0],
                                             dropout_seed=_dropout_state["seed"])
                optimizer._attach(tr)
            tr.lr = optimizer.param_groups[0]['lr']
            # src/main_baseline.py:232-233: per-clip random shifts, +-64 pooled frames in time, +-4 mel bins
            shift_list = [random.randint(-64, 64) * cfg.pooling_time_ratio for _ in range(x.shape[0])]
            freq_shift_list = [random.randint(-4, 4) for _ in range(x.shape[0])]
            tgt = target.to(dev, non_blocking=True).float()
            target_weak = tgt if tgt.dim() == 2 else tgt.max(-2)[0]
            losses12 = tr.step(x, x_ema, target_weak, xs, ts, shift_list, freq_shift_list, global_step, rampup_value)
            losses = torch.cat([losses12[:8], 0.5 * losses12[8:10]])    # the terms that make up the loss (:517-529)
        elif fused:
            tr = optimizer._trainer
            if tr is None:
                g = optimizer.param_groups[0]
                tr = MeanTeacherTrainer(model, predictor, ema_model, ema_predictor, lr=g['lr'], betas=g['betas'],
                                        eps=g['eps'], weight_decay=g['weight_decay'], n_syn=xs.shape[0],
                                        n_real=x.shape[0], dropout_seed=_dropout_state["seed"])
                optimizer._attach(tr)
            tr.lr = optimizer.param_groups[0]['lr']
            losses = tr.step(x, x_ema, xs, ts, global_step, rampup_len)
        else:
            target_d = target.to(dev, non_blocking=True)
            losses = _generic_step(model, predictor, ema_model, ema_predictor, optimizer, (x, x_ema, target_d),
                                   (xs, None, ts), global_step, rampup_value)
    loss = losses.sum() if losses is not None else None
    if getattr(optimizer, "_trainer", None) is not None:
        optimizer._trainer.check_health()
    if losses is not None:
        lv = losses.tolist()   # the only host sync of the epoch
        log.info("Epoch: %d\t Time %.2f\t strong %.4f weak %.4f cons_strong %.4f cons_weak %.4f", c_epoch,
                 time.time() - start, *lv[:4])
        assert not (sum(lv) != sum(lv) or sum(lv) > 1e5), 'Loss explosion: {}'.format(sum(lv))
    return loss
