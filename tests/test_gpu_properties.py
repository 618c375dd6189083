"""Size-independent properties at the full BASELINE sizes (24 ten-second clips, 1255 x 128 features, 36-clip step) that
hold bit for bit whatever the data: scaling by a power of two, shifts by whole hops, independence of the clips of a
batch, idempotence of the decoder.  They complement the oracle comparisons, which run at sizes the CPU finishes fast."""
import numpy as np
import pytest
import torch

from helpers import bsed_models, oracle_models
from bsed_b200.utilities import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clips24():
    return torch.from_numpy(synth.make_clips(24, seed=77)).cuda()


def test_melspec_is_homogeneous_under_power_of_two_gain(clips24):
    from bsed_b200 import engine
    a = engine.melspec(clips24)
    b = engine.melspec(clips24 * 4.0)
    assert torch.equal(b, a * 4.0)                    # every operation of the chain is linear or |.|: exact in fp32
    assert a.shape == (24, 1255, 128) and float(a.min()) >= 0.0


def test_melspec_clips_are_independent_of_their_batch(clips24):
    from bsed_b200 import engine
    full = engine.melspec(clips24)
    for k in (0, 7, 23):
        assert torch.equal(engine.melspec(clips24[k:k + 1])[0], full[k])


def test_melspec_shift_by_whole_hops_shifts_the_frames(clips24):
    from bsed_b200 import engine
    k = 5
    x = clips24[:4]
    shifted = torch.zeros_like(x)
    shifted[:, 255 * k:] = x[:, :-255 * k]
    a, b = engine.melspec(x), engine.melspec(shifted)
    # frames whose 2048-sample window lies inside the common, unpadded region see the same samples
    lo, hi = k + 5, 1255 - 5
    assert torch.equal(b[:, lo:hi], a[:, lo - k:hi - k])


def test_db_transform_gain_and_clamp(clips24):
    from bsed_b200 import engine
    mel = engine.melspec(clips24)
    a = engine.amp_to_db(mel, 1255)
    b = engine.amp_to_db(mel * 10.0, 1255)
    top = a.amax(dim=(1, 2), keepdim=True)
    assert bool((a >= top - 80.0 - 1e-4).all())                     # per-clip 80 dB floor (librosa top_db)
    live = mel > 1e-4                                               # above amin in both versions
    assert float((b - a - 20.0)[live].abs().max()) < 2e-4           # +20 dB of gain, away from the amin clamp
    # ragged input: rows past t_in are zero-padded after the log, longer inputs are truncated
    short = engine.amp_to_db(mel[:, :900].contiguous(), 1255)
    assert float(short[:, 900:].abs().max()) == 0.0
    long = engine.amp_to_db(mel, 1000)
    assert long.shape == (24, 1000, 128)


def test_crnn_eval_clips_are_independent_of_their_batch():
    oc, op = oracle_models(seed=5, linear_std=0.2)
    m, p = bsed_models(oc, op)
    m.eval(); p.eval()
    x = torch.from_numpy(synth.make_logmel_like(24, seed=31)).cuda()
    with torch.no_grad():
        enc, _ = m(x)
        strong, weak = p(enc)
        e1, _ = m(x[5:6])
        s1, w1 = p(e1)
    assert torch.equal(e1[0], enc[5]) and torch.equal(s1[0], strong[5]) and torch.equal(w1[0], weak[5])
    assert bool(torch.isfinite(enc).all()) and float(strong.min()) >= 0.0 and float(strong.max()) <= 1.0


def test_decoder_is_idempotent_on_its_own_output():
    from bsed_b200 import engine
    g = torch.Generator().manual_seed(5)
    p = torch.rand(24, 313, 20, generator=g).cuda()
    ev, n = engine.median_decode(p, 0.5, 14)
    # rebuild the filtered binary matrix from the events and decode it again with a 1-frame window: same events
    rebuilt = torch.zeros(24, 313, 20)
    evc, nc = ev.cpu().numpy(), n.cpu().numpy()
    for b in range(24):
        for c, on, off in evc[b, :nc[b]]:
            rebuilt[b, on:off, c] = 1.0
    ev2, n2 = engine.median_decode(rebuilt.cuda(), 0.5, 1)
    assert torch.equal(n, n2)
    for b in range(24):
        assert np.array_equal(evc[b, :nc[b]], ev2[b, :nc[b]].cpu().numpy())
    # sortedness: class-major, then time, non-overlapping runs inside a class
    for b in range(24):
        e = evc[b, :nc[b]]
        key = e[:, 0] * 1000 + e[:, 1]
        assert np.all(np.diff(key) > 0) and np.all(e[:, 2] > e[:, 1])


def test_full_size_train_step_replicas_and_counters():
    """Two trainers with the same seed take the same step bit for bit (no atomics-order dependence in the parameters), and
    the BatchNorm counters advance as the reference's would (2 student calls, 1 teacher call per iteration)."""
    from bsed_b200.main import MeanTeacherTrainer
    outs = []
    for rep in range(2):
        oc, op = oracle_models(seed=5, linear_std=0.2)
        tc, tp = oracle_models(seed=6, linear_std=0.2)
        m, p = bsed_models(oc, op, dropout=0.5)
        em, ep = bsed_models(tc, tp, dropout=0.5)
        for mod in (m, p, em, ep):
            mod.train()
        for prm in list(em.parameters()) + list(ep.parameters()):
            prm.detach_()
        x = torch.from_numpy(synth.make_logmel_like(12, seed=41)).cuda()
        xs = torch.from_numpy(synth.make_logmel_like(12, seed=42)).cuda()
        ts = torch.from_numpy(synth.make_targets(12, seed=43)).cuda()
        tr = MeanTeacherTrainer(m, p, em, ep, lr=5e-4, n_syn=12, n_real=12, dropout_seed=2023, precision="fp32")
        for it in range(2):
            losses = tr.step(x, x, xs, ts, it, 500)
        assert bool(torch.isfinite(losses).all())
        assert int(m.cnn.batchnorm0.num_batches_tracked) == 4 and int(em.cnn.batchnorm0.num_batches_tracked) == 2
        outs.append((tr.params.clone(), tr.ema_params.clone(), losses.clone()))
    # fp64 atomics accumulate the BatchNorm statistics and split-K partials in a run-dependent order: the replicas agree to
    # rounding, not bitwise
    assert float((outs[0][0] - outs[1][0]).abs().max()) < 1e-5
    assert float((outs[0][1] - outs[1][1]).abs().max()) < 1e-5     # EMA alpha is 0 and 1/2 in the first two steps
    assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-4, atol=1e-7)
