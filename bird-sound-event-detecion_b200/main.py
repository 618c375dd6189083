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
Out of scope here: the ISP/ICT shift-consistency branches.
"""
import logging
import time

import torch
from torch import nn

from . import engine
from .data import config as cfg
from .models.CRNN import CRNN, Predictor, _dropout_state
from .utilities import ramps, shard

log = logging.getLogger("bsed_b200.train")


def adjust_learning_rate(optimizer, rampup_value, rampdown_value=1, optimizer_d=None, optimizer_crnn=None,
                         c_epoch=None, rampup_value_adv=None):
    """lr = rampup * rampdown * max_lr; the d / crnn optimizers get 0.1 x (src/main.py:51-83)."""
    lr = rampup_value * rampdown_value * cfg.max_learning_rate
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

    def step(self, closure=None):
        raise RuntimeError("FusedAdam is driven by MeanTeacherTrainer.step / train_mt")


def _rehome(modules, device):
    """Put the flat parameter buffers of several modules back to back in one tensor."""
    sizes = [m._flat.numel() for m in modules]
    joint = torch.empty(sum(sizes), dtype=torch.float32, device=device)
    o = 0
    for m, n in zip(modules, sizes):
        m._reflatten(flat=joint[o:o + n])
        o += n
    return joint, sizes


class MeanTeacherTrainer:
    """One fused mean-teacher iteration (src/main.py:190-523, pretrain stage, -mt)."""

    def __init__(self, model, predictor, ema_model=None, ema_predictor=None, lr=cfg.default_learning_rate,
                 betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, n_syn=cfg.batch_size, n_real=cfg.batch_size,
                 ema_flavour="state_dict", dropout_seed=2023, process_group=None, precision=None):
        assert isinstance(model, CRNN) and isinstance(predictor, Predictor)
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
        B = n_syn + n_real + (n_real if self.has_teacher else 0)
        self.B = B
        self.plan = engine.Plan(engine.make_cfg(**model.cfg_kwargs), max_clips=B, device=dev,
                                precision=precision or model.precision)
        assert self.plan.n_params == self.n_crnn and self.plan.n_pred_params == self.n_pred
        self.x = torch.empty(B, 1, cfg.max_frames, cfg.n_mels, dtype=torch.float32, device=dev)
        self.enc = torch.empty(B, self.plan.t_out, 256, dtype=torch.float32, device=dev)
        self.d_enc = torch.zeros(B, self.plan.t_out, 256, dtype=torch.float32, device=dev)
        self.last = {}

    def step(self, x, x_ema, xs, ts, global_step, rampup_length, max_consistency_cost=cfg.max_consistency_cost):
        """x / x_ema: real batch (student / teacher inputs), xs / ts: synthetic batch and its strong
        targets; all CUDA tensors.  Returns the 4 loss terms as a device tensor
        [strong_bce, weak_bce, cons_strong, cons_weak] (no host sync)."""
        ns, nr = self.n_syn, self.n_real
        nst = ns + nr
        m, p = self.model, self.predictor
        self.x[:ns].copy_(xs.reshape(ns, 1, cfg.max_frames, cfg.n_mels))
        self.x[ns:nst].copy_(x.reshape(nr, 1, cfg.max_frames, cfg.n_mels))
        sp, sbn, snbt = m.flat_tensors()
        groups = [dict(params=sp, bn=sbn, nbt=snbt, n=ns), dict(params=sp, bn=sbn, nbt=snbt, n=nr)]
        if self.has_teacher:
            self.x[nst:].copy_(x_ema.reshape(nr, 1, cfg.max_frames, cfg.n_mels))
            tp, tbn, tnbt = self.ema_model.flat_tensors()
            groups.append(dict(params=tp, bn=tbn, nbt=tnbt, n=nr))
        self.plan.forward(groups, self.x, train=True, save=True, seed=self.dropout_seed, step=global_step, enc=self.enc)
        pp = self.params[self.n_crnn:]
        logits, strong, weak = self.plan.predictor_forward(pp, self.enc[:nst])
        if self.has_teacher:
            _, strong_ema, weak_ema = self.plan.predictor_forward(self.ema_params[self.n_crnn:], self.enc[nst:])
            cons_w = max_consistency_cost * ramps.exp_rampup(global_step, rampup_length)
            losses, d_strong, d_weak = engine.mt_loss(strong, weak, 0, ns, ts.contiguous().float(), ns, nr, strong_ema,
                                                      weak_ema, cons_w)
        else:
            losses, d_strong, d_weak = engine.mt_loss(strong, weak, 0, ns, ts.contiguous().float(), 0, 0, None, None, 0.0)
        self.plan.predictor_backward(pp, self.enc[:nst], logits, strong, weak, d_strong, d_weak,
                                     self.grads[self.n_crnn:], accumulate=False, d_enc=self.d_enc[:nst])
        self.plan.backward(0b011, self.d_enc, self.grads[:self.n_crnn], accumulate=False)
        grad_scale = shard.allreduce_gradients(self.grads, self.pg)   # NCCL sum over NVLink; 1/N folded below
        self.opt_step += 1
        engine.opt_ema_step(self.params, self.grads, self.m, self.v, self.ema_params if self.has_teacher else None,
                            step=self.opt_step, ema_step=global_step + 1, kind="adam", lr=self.lr, betas=self.betas,
                            eps=self.eps, weight_decay=self.weight_decay, grad_scale=grad_scale)
        if self.has_teacher and self.ema_flavour == "state_dict":
            engine.ema_buffers(sbn, tbn, snbt, tnbt, global_step + 1)
        self.last = dict(strong=strong, weak=weak, losses=losses)
        return losses


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
    optimizer_crnn.step()
    optimizer_d.step()
    return domain_loss.detach()


def train_mt(train_loader, syn_loader, model, optimizer, c_epoch, ema_model=None, ema_predictor=None, mask_weak=None,
             mask_strong=None, adjust_lr=False, discriminator=None, optimizer_d=None, predictor=None,
             optimizer_crnn=None, ISP=False):
    """One epoch of the mean-teacher model; same arguments as the reference (src/main.py:163).
    Loaders yield (((student_input, teacher_input), target), filename)."""
    if ISP or mask_weak is not None or mask_strong is not None:
        raise NotImplementedError("ISP / masked-real-label branches are outside this round's hot path")
    if discriminator is not None and (optimizer_d is None or optimizer_crnn is None):
        raise ValueError("the adversarial update needs optimizer_d and optimizer_crnn (as in the reference)")
    if predictor is None:
        raise ValueError("train_mt needs the Predictor module (the reference passes it as `predictor`)")
    start = time.time()
    syn_iter = iter(syn_loader)
    n_syn_batches = len(syn_loader)
    losses = None
    fused = isinstance(optimizer, FusedAdam)
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
            adjust_learning_rate(optimizer, rampup_value, optimizer_d=optimizer_d, optimizer_crnn=optimizer_crnn,
                                 c_epoch=c_epoch)
        dev = model._flat.device
        x = batch_input.to(dev, non_blocking=True)
        x_ema = ema_batch_input.to(dev, non_blocking=True)
        xs = syn_batch_input.to(dev, non_blocking=True)
        ts = syn_target.to(dev, non_blocking=True)
        if discriminator is not None:
            domain_loss = adversarial_step(model, predictor, discriminator, optimizer_crnn, optimizer_d, x, xs)
        if fused:
            tr = optimizer._trainer
            if tr is None:
                g = optimizer.param_groups[0]
                tr = MeanTeacherTrainer(model, predictor, ema_model, ema_predictor, lr=g['lr'], betas=g['betas'],
                                        eps=g['eps'], weight_decay=g['weight_decay'], n_syn=xs.shape[0],
                                        n_real=x.shape[0], dropout_seed=_dropout_state["seed"])
                optimizer._trainer = tr
            tr.lr = optimizer.param_groups[0]['lr']
            losses = tr.step(x, x_ema, xs, ts, global_step, rampup_len)
        else:
            target_d = target.to(dev, non_blocking=True)
            losses = _generic_step(model, predictor, ema_model, ema_predictor, optimizer, (x, x_ema, target_d),
                                   (xs, None, ts), global_step, rampup_value)
    loss = losses.sum() if losses is not None else None
    if losses is not None:
        lv = losses.tolist()   # the only host sync of the epoch
        log.info("Epoch: %d\t Time %.2f\t strong %.4f weak %.4f cons_strong %.4f cons_weak %.4f", c_epoch,
                 time.time() - start, *lv)
        assert not (sum(lv) != sum(lv) or sum(lv) > 1e5), 'Loss explosion: {}'.format(sum(lv))
    return loss
