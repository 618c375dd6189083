"""ResNet-18 weak-label tagger with the reference's entry points (src/audio_tagging_system_cnn.py):

    Net_resnet(pretrained=True)                                                             (:50-64)
    train_mt(train_unlabeled_loader, train_weak_loader, syn_loader, model, optimizer, c_epoch, ema_model=None, ...)   (:199-416)

One iteration = the three loaders' batches (synthetic with strong targets; weakly labelled; pseudo-labelled unlabeled, the
last two concatenated into the "real" batch as the reference does, :247-250), two model calls, loss = BCE(syn_weak,
max_t syn_target) + BCE(weak[:half], target_weak[:half]) (:345-352), backward, optimiser step, optional parameter EMA of a
teacher copy (:404-406).  With a `FusedAdam` optimizer (bsed_b200.main) the iteration is `TaggerTrainer.step` (forward,
loss, backward and Adam over the flat parameter buffer, all in libbsed.so); with any other torch optimizer the reference's
statement order runs through Net_resnet's autograd function.  The ISP branch of that script is not built here."""
import logging
import time

import torch

from . import engine
from .main import FusedAdam, update_ema_variables
from .models.ResNet import Net_resnet, TaggerTrainer  # noqa: F401  (Net_resnet is part of this module's reference API)

log = logging.getLogger("bsed_b200.tagging")


def train_mt(train_unlabeled_loader, train_weak_loader, syn_loader, model, optimizer, c_epoch, ema_model=None,
             ema_predictor=None, mask_weak=None, mask_strong=None, adjust_lr=False, discriminator=None, optimizer_d=None,
             predictor=None, optimizer_crnn=None, ISP=False):
    """One epoch of the tagger; same arguments as the reference (src/audio_tagging_system_cnn.py:199).
    Loaders yield (((student_input, teacher_input), target), filename)."""
    if ISP or discriminator is not None:
        raise NotImplementedError("the ISP and discriminator arguments of the tagger's train_mt are not built")
    start = time.time()
    it_u, it_w = iter(train_unlabeled_loader), iter(train_weak_loader)
    dev = next(model.parameters()).device
    fused = isinstance(optimizer, FusedAdam)
    loss = None
    global_step = c_epoch * len(syn_loader)
    for i, data_syn in enumerate(syn_loader):
        try:
            data_u = next(it_u)
        except StopIteration:
            it_u = iter(train_unlabeled_loader)
            data_u = next(it_u)
        try:
            data_w = next(it_w)
        except StopIteration:
            it_w = iter(train_weak_loader)
            data_w = next(it_w)
        ((u_in, _u_ema), target_pl), _ = data_u
        ((w_in, _w_ema), target), _ = data_w
        ((syn_in, _s_ema), syn_target), _ = data_syn
        half = syn_in.shape[0] // 2
        if u_in.shape[0] != half or w_in.shape[0] != half:          # :243-246
            continue
        target_weak = target.max(-2)[0] if target.dim() == 3 else target
        batch_input = torch.cat((w_in, u_in), dim=0).to(dev, non_blocking=True)
        target_weak = torch.cat((target_weak, target_pl), dim=0).to(dev, non_blocking=True).float()
        syn_in = syn_in.to(dev, non_blocking=True)
        syn_target = syn_target.to(dev, non_blocking=True).float()
        if fused:
            tr = optimizer._trainer
            if tr is None:
                g = optimizer.param_groups[0]
                tr = TaggerTrainer(model, lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"])
                optimizer._trainer = tr
            tr.lr = optimizer.param_groups[0]["lr"]
            loss = tr.step(syn_in, syn_target, batch_input, target_weak)
        else:
            model.train()
            syn_weak_pred = model(syn_in)
            weak_pred = model(batch_input)
            from ._lib import LOSS_BCE_WEAK
            weak = torch.cat([syn_weak_pred, weak_pred]).contiguous()
            terms = [dict(kind=LOSS_BCE_WEAK, pred_first=0, n=syn_in.shape[0], ref=syn_target.contiguous(),
                          ref_is_strong=syn_target.dim() == 3, slot=0),
                     dict(kind=LOSS_BCE_WEAK, pred_first=syn_in.shape[0], n=half, ref=target_weak[:half].contiguous(), slot=0)]
            T = syn_target.shape[1] if syn_target.dim() == 3 else 1
            dummy = torch.zeros(weak.shape[0], T, weak.shape[1], dtype=torch.float32, device=dev)
            losses, _, d_weak = engine.loss_terms(dummy, weak.detach(), terms, 1)
            optimizer.zero_grad()
            weak.backward(d_weak)
            optimizer.step()
            loss = losses[0]
        global_step += 1
        if ema_model is not None:
            update_ema_variables(model, ema_model, 0.999, global_step, flavour="params")
    if loss is not None:
        lv = float(loss)    # the only host sync of the epoch
        log.info("Epoch: %d\t Time %.2f\t weak_class_loss %.4f", c_epoch, time.time() - start, lv)
        assert not (lv != lv or lv > 1e5), 'Loss explosion: {}'.format(lv)
    return loss
