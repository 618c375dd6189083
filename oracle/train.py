"""Oracle (TEST INFRASTRUCTURE): restatement of one mean-teacher training step.

PINNED for model / loss / EMA arithmetic through the reference modules (tests/make_golden.py runs
the same step with the reference's CRNN/Predictor classes); the step assembly itself restates
src/main.py:train_mt because the script cannot be imported (tensorboardX / librosa / module-level
globals).

Follows (pretrain stage, `-mt`, no discriminator, mask_weak = mask_strong = None):
  * src/main.py:250-254   teacher forward (teacher stays in .train() mode, main.py:947-948), detach
  * src/main.py:335-343   student forward on the synthetic batch, then on the real batch (two calls:
                          BatchNorm statistics are per call)
  * src/main.py:376,405   weak  BCE(syn_weak_pred, syn_target.max(-2))
  * src/main.py:434       strong BCE(syn_strong_pred, syn_target)
  * src/main.py:439-449   consistency MSE(student real, teacher real) * max_consistency_cost * rampup
  * src/main.py:474-477   loss = strong + weak + cons_weak + cons_strong
  * src/main.py:517-518   backward, Adam step (lr 5e-4, betas (.9,.999), eps 1e-8, wd 0; main.py:823-828)
  * src/main.py:86-100,520-523   update_ema_variables(model, ema, .999, global_step + 1) on the whole
                          state_dict (BN buffers and num_batches_tracked included), and
  * src/main_origin.py:85-89     the parameters-only in-place flavour.
  * src/utilities/ramps.py:4-16  exp_rampup
"""
import numpy as np
import torch
from torch import nn


def exp_rampup(current, rampup_length):
    if rampup_length == 0:
        return 1.0
    current = float(np.clip(current, 0.0, rampup_length))
    phase = 1.0 - current / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def sigmoid_rampdown(current, rampup_length):
    if rampup_length == 0:
        return 1.0
    current = float(np.clip(current, 0.0, rampup_length))
    phase = 1.0 - current / rampup_length
    return float(np.exp(-12.5 * phase * phase))


def ema_alpha(alpha, global_step):
    return min(1 - 1 / (global_step + 1), alpha)


def update_ema_state_dict(model, ema_model, alpha, global_step):
    """main.py:86-100: every state_dict entry, integer counters cast back by load_state_dict."""
    a = ema_alpha(alpha, global_step)
    with torch.no_grad():
        msd = model.state_dict()
        esd = ema_model.state_dict()
        for k in esd.keys():
            esd[k] = esd[k].clone() * a + msd[k].clone() * (1.0 - a)
        ema_model.load_state_dict(esd)


def update_ema_params(model, ema_model, alpha, global_step):
    """main_origin.py:85-89: parameters only, in place."""
    a = ema_alpha(alpha, global_step)
    with torch.no_grad():
        for pe, p in zip(ema_model.parameters(), model.parameters()):
            pe.mul_(a).add_(p, alpha=1 - a)


def losses(syn_strong, syn_weak, syn_target, strong, weak, strong_ema, weak_ema, cons_w):
    bce = nn.BCELoss()
    mse = nn.MSELoss()
    syn_target_weak = syn_target.max(-2)[0]
    weak_class = bce(syn_weak, syn_target_weak)
    strong_class = bce(syn_strong, syn_target)
    cons_strong = cons_w * mse(strong, strong_ema)
    cons_weak = cons_w * mse(weak, weak_ema)
    total = strong_class + weak_class + (cons_weak + cons_strong)
    return total, dict(weak_class=weak_class, strong_class=strong_class,
                       cons_strong=cons_strong, cons_weak=cons_weak)


def mt_step(model, predictor, ema_model, ema_predictor, optimizer, x, x_ema, xs, ts,
            global_step, rampup_length, ema_flavour="state_dict", max_consistency_cost=1.0,
            dropout_hook=None):
    """One iteration of train_mt.  x / x_ema: real batch student / teacher inputs (B,1,1255,128);
    xs, ts: synthetic batch and its (B,313,20) targets.  `dropout_hook(tag)` is called before each
    forward with tag in {'teacher','syn','real'} so callers can key the hash-dropout masks.
    Returns (loss, parts, outputs)."""
    rampup = exp_rampup(global_step, rampup_length)
    hook = dropout_hook or (lambda tag: None)

    hook("teacher")
    enc_ema, _ = ema_model(x_ema)
    strong_ema, weak_ema = ema_predictor(enc_ema)
    strong_ema = strong_ema.detach()
    weak_ema = weak_ema.detach()

    optimizer.zero_grad()
    hook("syn")
    enc_s, _ = model(xs)
    syn_strong, syn_weak = predictor(enc_s)
    hook("real")
    enc, _ = model(x)
    strong, weak = predictor(enc)

    loss, parts = losses(syn_strong, syn_weak, ts, strong, weak, strong_ema, weak_ema,
                         max_consistency_cost * rampup)
    loss.backward()
    grads = {}
    for prefix, mod in (("crnn.", model), ("pred.", predictor)):
        for n, p in mod.named_parameters():
            grads[prefix + n] = p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)
    optimizer.step()

    gs = global_step + 1
    if ema_flavour == "none":
        pass
    elif ema_flavour == "state_dict":
        update_ema_state_dict(model, ema_model, 0.999, gs)
        update_ema_state_dict(predictor, ema_predictor, 0.999, gs)
    else:
        update_ema_params(model, ema_model, 0.999, gs)
        update_ema_params(predictor, ema_predictor, 0.999, gs)
    outs = dict(strong_ema=strong_ema, weak_ema=weak_ema, syn_strong=syn_strong.detach(),
                syn_weak=syn_weak.detach(), strong=strong.detach(), weak=weak.detach(), grads=grads)
    return loss.detach(), {k: v.detach() for k, v in parts.items()}, outs
