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


def isp_step(model, predictor, ema_model, ema_predictor, optimizer, x, x_ema, target_weak, xs, ts, shift_list,
             freq_shift_list, global_step, rampup_value, max_consistency_cost=1.0, pooling_time_ratio=4,
             dropout_hook=None):
    """One iteration of the shift-consistency (ISP / SCT) step of src/main_baseline.py:train_mt with ISP=True, a teacher
    and no discriminator update.  PARITY UNPINNED for the step assembly (the script imports librosa / tensorboardX and
    cannot be executed here); the modules it drives are the pinned ones.

      * :229-277   per clip k: inputs rolled by shift_list[k] frames (time) and by freq_shift_list[k] bins (frequency),
                   the same shifts for the real, teacher and synthetic batches
      * :337-341   student forward: synthetic, real
      * :352-370   teacher forward (detached): real, time-shifted real, frequency-shifted real
      * :375-408   rolled (detached) student predictions and rolled synthetic targets, roll = shift / pooling_time_ratio
      * :412-422   student forward: shifted real, freq-shifted real, shifted synthetic, freq-shifted synthetic
      * :433-452   weak BCE (synthetic + real weak targets), SCT weak term on the freq-shifted batches (real: first half)
      * :475-482   strong BCE; strong BCE of the shifted / freq-shifted synthetic batch
      * :485-512   teacher-student consistency (MSE) incl. the shifted variants (the weak shifted ones are only logged)
      * :517-529   loss assembly
      * :571-584   backward, optimizer step, state-dict EMA with global_step + 1
    x / x_ema / xs: (B,1,T,128); target_weak (B,20); ts (B,313,20); shift_list multiples of pooling_time_ratio.
    dropout_hook(tag) is called before each forward with tag in {'syn','real','teacher','teacher_shift',
    'teacher_fshift','real_shift','real_fshift','syn_shift','syn_fshift'}."""
    hook = dropout_hook or (lambda tag: None)
    B = x.shape[0]

    def roll_batch(t, shifts, dim):
        return torch.stack([torch.roll(t[k], int(shifts[k]), dims=dim) for k in range(t.shape[0])])

    x_shift, x_fshift = roll_batch(x, shift_list, 1), roll_batch(x, freq_shift_list, 2)
    xe_shift, xe_fshift = roll_batch(x_ema, shift_list, 1), roll_batch(x_ema, freq_shift_list, 2)
    xs_shift, xs_fshift = roll_batch(xs, shift_list, 1), roll_batch(xs, freq_shift_list, 2)

    def run(m, p, inp, tag):
        hook(tag)
        enc, _ = m(inp)
        return p(enc)

    syn_strong, syn_weak = run(model, predictor, xs, "syn")
    strong, weak = run(model, predictor, x, "real")
    strong_ema, weak_ema = [t.detach() for t in run(ema_model, ema_predictor, x_ema, "teacher")]
    strong_shift_ema, weak_shift_ema = [t.detach() for t in run(ema_model, ema_predictor, xe_shift, "teacher_shift")]
    strong_fshift_ema, weak_fshift_ema = [t.detach() for t in run(ema_model, ema_predictor, xe_fshift, "teacher_fshift")]

    pool_shift = [int(s / pooling_time_ratio) for s in shift_list]
    strong_pred_shift = roll_batch(strong, pool_shift, 0).detach()
    syn_strong_pred_shift = roll_batch(syn_strong, pool_shift, 0).detach()
    syn_target_shift = roll_batch(ts, pool_shift, 0)

    strong_shift, weak_shift = run(model, predictor, x_shift, "real_shift")
    strong_fshift, weak_fshift = run(model, predictor, x_fshift, "real_fshift")
    syn_strong_shift, syn_weak_shift = run(model, predictor, xs_shift, "syn_shift")
    syn_strong_fshift, syn_weak_fshift = run(model, predictor, xs_fshift, "syn_fshift")

    bce, mse = nn.BCELoss(), nn.MSELoss()
    syn_target_weak = ts.max(-2)[0]
    widx = target_weak.shape[0] // 2
    cc = max_consistency_cost * rampup_value
    parts = dict(
        strong_class=bce(syn_strong, ts),
        weak_class=bce(syn_weak, syn_target_weak) + bce(weak, target_weak),
        cons_strong=cc * mse(strong, strong_ema),
        cons_weak=cc * mse(weak, weak_ema),
        weak_freq_shift_class=bce(syn_weak_fshift, syn_target_weak) + bce(weak_fshift[:widx], target_weak[:widx]),
        strong_shift_class=bce(syn_strong_shift, syn_target_shift),
        strong_freq_shift_class=bce(syn_strong_fshift, ts),
        cons_shift=cc / 2 * (mse(syn_strong_shift, syn_strong_pred_shift) + mse(strong_shift, strong_pred_shift)),
        cons_strong_shift=cc * mse(strong_shift, strong_shift_ema),
        cons_strong_freq_shift=cc * mse(strong_fshift, strong_fshift_ema),
        cons_weak_shift=cc * mse(weak_shift[widx:], weak_shift_ema[widx:]),            # logged only
        cons_weak_freq_shift=cc * mse(weak_fshift[widx:], weak_fshift_ema[widx:]))     # logged only
    loss = parts["strong_class"] + parts["weak_class"]
    loss = loss + (parts["cons_weak"] + parts["cons_strong"])
    loss = loss + (parts["weak_freq_shift_class"] + parts["strong_shift_class"] + parts["strong_freq_shift_class"]
                   + parts["cons_shift"])
    loss = loss + 1 / 2 * (parts["cons_strong_shift"] + parts["cons_strong_freq_shift"])
    optimizer.zero_grad()
    loss.backward()
    grads = {}
    for prefix, mod in (("crnn.", model), ("pred.", predictor)):
        for n, p in mod.named_parameters():
            grads[prefix + n] = p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)
    optimizer.step()
    gs = global_step + 1
    update_ema_state_dict(model, ema_model, 0.999, gs)
    update_ema_state_dict(predictor, ema_predictor, 0.999, gs)
    outs = dict(strong=strong.detach(), weak=weak.detach(), syn_strong=syn_strong.detach(),
                strong_shift=strong_shift.detach(), syn_strong_fshift=syn_strong_fshift.detach(), grads=grads)
    return loss.detach(), {k: v.detach() for k, v in parts.items()}, outs


def ada_step(model, predictor, ema_model, ema_predictor, disc, optimizer, optimizer_crnn, optimizer_d, x, x_ema,
             target_weak, xs, ts, global_step, rampup_length, grl_iter, max_consistency_cost=1.0, dropout_hook=None):
    """One iteration of src/main_scmt_ada_weak_seperate.py:train_mt with a discriminator (no ISP):
      :314-335  adversarial update -- student forward of the synthetic and the real batch, ConditionalDomainAdversarialLoss
                (oracle/da.py: cdan_clip_loss, GRL coefficient of iteration `grl_iter`), backward, optimizer_crnn.step(),
                optimizer_d.step()
      :337-521  mean-teacher update -- strong / weak BCE on the synthetic batch, weak BCE of the real batch against
                `target_weak` (:437-445), MSE consistency with the teacher, optimizer.step(), EMA (parameters + buffers)
    `dropout_hook(tag)` with tag in {'adv_syn', 'adv_real', 'teacher', 'syn', 'real'} keys the hash-dropout masks.
    Returns (loss, parts incl. 'domain', outputs)."""
    from oracle import da as oda
    hook = dropout_hook or (lambda tag: None)
    rampup = exp_rampup(global_step, rampup_length)

    hook("adv_syn")
    _, syn_d_input = model(xs)
    hook("adv_real")
    _, d_input = model(x)
    optimizer_crnn.zero_grad()
    optimizer_d.zero_grad()
    domain_loss, _ = oda.cdan_clip_loss(disc, syn_d_input, d_input, grl_iter)
    domain_loss.backward()
    adv_grads = {"crnn." + n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    d_grads = {n: p.grad.detach().clone() for n, p in disc.named_parameters()}
    optimizer_crnn.step()
    optimizer_d.step()

    hook("teacher")
    enc_ema, _ = ema_model(x_ema)
    strong_ema, weak_ema = ema_predictor(enc_ema)
    strong_ema, weak_ema = strong_ema.detach(), weak_ema.detach()
    optimizer.zero_grad()
    hook("syn")
    syn_strong, syn_weak = predictor(model(xs)[0])
    hook("real")
    strong, weak = predictor(model(x)[0])
    bce, mse = nn.BCELoss(), nn.MSELoss()
    weak_class = bce(syn_weak, ts.max(-2)[0]) + bce(weak, target_weak)
    strong_class = bce(syn_strong, ts)
    cc = max_consistency_cost * rampup
    cons_strong, cons_weak = cc * mse(strong, strong_ema), cc * mse(weak, weak_ema)
    loss = strong_class + weak_class + (cons_weak + cons_strong)
    loss.backward()
    grads = {}
    for prefix, mod in (("crnn.", model), ("pred.", predictor)):
        for n, p in mod.named_parameters():
            grads[prefix + n] = p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)
    optimizer.step()
    update_ema_state_dict(model, ema_model, 0.999, global_step + 1)
    update_ema_state_dict(predictor, ema_predictor, 0.999, global_step + 1)
    parts = dict(strong_class=strong_class, weak_class=weak_class, cons_strong=cons_strong, cons_weak=cons_weak,
                 domain=domain_loss)
    outs = dict(strong=strong.detach(), weak=weak.detach(), syn_strong=syn_strong.detach(), grads=grads, adv_grads=adv_grads,
                d_grads=d_grads)
    return loss.detach(), {k: v.detach() for k, v in parts.items()}, outs
