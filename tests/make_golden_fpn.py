"""Generate tests/golden/fpn_*.npz by EXECUTING THE REFERENCE's CRNN_fpn / CNN_FPN / Predictor (imported from
/root/reference/src) in the build container.  The fixtures pin oracle/crnn.py:OracleCRNNfpn and the CUDA CRNN_fpn.

    python tests/make_golden_fpn.py

Dropout: the reference's nn.Dropout modules are replaced by the hash dropout the kernels use (one module is called
several times per forward in CNN_FPN / CRNN_fpn, hence CycleHashDropout).  Weights and inputs are regenerated from
seeds by the tests.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from oracle import crnn as ocrnn  # noqa: E402
from oracle import train as otrain  # noqa: E402
from bsed_b200.utilities import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def reference_models(dropout, seed, linear_std=0.2):
    """The reference's own CRNN_fpn / Predictor carrying the seeded oracle weights (strict load: the key sets and shapes
    of the restatement and of the reference are identical)."""
    from models.CRNN import CRNN_fpn, Predictor
    kw = dict(ocrnn.CRNN_KWARGS)
    kw["dropout"] = dropout
    rc, rp = CRNN_fpn(**kw), Predictor(**ocrnn.PREDICTOR_KWARGS)
    oc = ocrnn.OracleCRNNfpn(**{**ocrnn.CRNN_KWARGS, "dropout": 0.0})
    op = ocrnn.OraclePredictor(**ocrnn.PREDICTOR_KWARGS)
    ocrnn.reference_style_init(oc, op, seed, linear_std)
    rc.load_state_dict(oc.state_dict(), strict=True)
    rp.load_state_dict(op.state_dict(), strict=True)
    # hash dropout in place of nn.Dropout
    for i in range(7):
        setattr(rc.cnn.cnn, f"dropout{i}", ocrnn.HashDropout(dropout, i))
    rc.cnn.dropout = ocrnn.CycleHashDropout(0.5, ocrnn.STREAM_FCN)                           # CNN_FPN.py:79
    rc.dropout = ocrnn.CycleHashDropout(dropout, ocrnn.STREAM_RNN_OUT_FPN, layout="BCT")     # CRNN.py:270
    return rc, rp


def set_keys(mod, seed, step, batch_offset):
    for m in mod.modules():
        if isinstance(m, ocrnn.HashDropout):
            m.key = ocrnn.mix_key(seed, step, m.stream)
            m.batch_offset = batch_offset
        elif isinstance(m, ocrnn.CycleHashDropout):
            m.set(seed, step, batch_offset)


def sample(v, n=4096):
    g = v.detach().numpy().reshape(-1)
    return g if g.size <= n else g[:: max(1, g.size // n)][:n]


PARAM_KEYS = ("cnn.cnn.conv0.weight", "cnn.cnn.glu4.linear.weight", "cnn.cnn_fcn.weight", "cnn.glu.linear.weight",
              "cnn.bn_fcn.weight", "cnn.conv1x1.weight", "rnn.rnn.weight_hh_l0", "rnn_2.rnn.weight_ih_l1_reverse",
              "rnn_4.rnn.bias_hh_l0", "conv1x1_2.weight", "conv1x1_4.bias", "cnn.bn_fcn.running_mean",
              "cnn.bn_fcn.running_var", "cnn.cnn.batchnorm5.running_mean")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)

    rc, rp = reference_models(0.5, seed=5)
    sd = rc.state_dict()
    np.savez(os.path.join(OUT, "fpn_state_dict_keys.npz"), keys=np.array(list(sd.keys())),
             shapes=np.array([str(tuple(v.shape)) for v in sd.values()]),
             param_keys=np.array([n for n, _ in rc.named_parameters()]),
             n_params=sum(p.numel() for p in rc.parameters()))

    # ---------------------------------------------------------------- eval forward
    x = torch.from_numpy(synth.make_logmel_like(2, seed=11))
    rc.eval(); rp.eval()
    with torch.no_grad():
        enc, d_in = rc(x)
        strong, weak = rp(enc)
    assert d_in is enc or torch.equal(d_in, enc)
    np.savez_compressed(os.path.join(OUT, "fpn_eval.npz"), x_sum=float(x.double().sum()), enc=enc.numpy()[:, ::8],
                        enc_sum=float(enc.double().sum()), strong=strong.numpy(), weak=weak.numpy())

    # ---------------------------------------------------------------- train-mode forward + backward, hash dropout
    rc, rp = reference_models(0.5, seed=5)
    rc.train(); rp.train()
    set_keys(rc, seed=2023, step=3, batch_offset=0)
    enc, _ = rc(x)
    strong, weak = rp(enc)
    w = torch.from_numpy(np.random.default_rng(7).standard_normal(strong.shape).astype(np.float32))
    ((strong * w).mean() + weak.mean()).backward()
    rec = dict(strong=strong.detach().numpy(), weak=weak.detach().numpy(), enc_sum=float(enc.detach().double().sum()),
               nbt_fcn=int(rc.cnn.bn_fcn.num_batches_tracked), nbt0=int(rc.cnn.cnn.batchnorm0.num_batches_tracked),
               rm_fcn=rc.cnn.bn_fcn.running_mean.numpy(), rv_fcn=rc.cnn.bn_fcn.running_var.numpy())
    for n, p in rc.named_parameters():
        if p.grad is None:
            rec["none_" + n] = 1
            continue
        rec["g_" + n] = sample(p.grad)
        rec["gn_" + n] = float(p.grad.double().norm())
    np.savez_compressed(os.path.join(OUT, "fpn_train.npz"), **rec)
    print("fpn_train: nbt_fcn", rec["nbt_fcn"], "unused:", [k for k in rec if k.startswith("none_")])

    # ---------------------------------------------------------------- mean-teacher steps, state-dict EMA (main.py:86-100)
    for name, p_drop in (("fpn_mt_step_drop", 0.5),):
        rc, rp = reference_models(p_drop, seed=5)
        tc, tp = reference_models(p_drop, seed=6)
        for m in (rc, rp, tc, tp):
            m.train()
        for prm in list(tc.parameters()) + list(tp.parameters()):
            prm.detach_()
        xs = torch.from_numpy(synth.make_logmel_like(2, seed=21))
        xr = torch.from_numpy(synth.make_logmel_like(2, seed=22))
        xr_ema = xr + 0.5 * torch.from_numpy(synth.make_logmel_like(2, seed=23)) * 0.1
        ts = torch.from_numpy(synth.make_targets(2, seed=24))
        opt = torch.optim.Adam(list(rc.parameters()) + list(rp.parameters()), lr=5e-4, betas=(0.9, 0.999))
        rec = {}
        for it in range(2):
            gstep = 100 + it

            def hook(tag, gstep=gstep):
                # device batch order: syn clips [0,2), real [2,4), teacher [4,6)
                if tag == "teacher":
                    set_keys(tc, 2023, gstep, 4)
                elif tag == "syn":
                    set_keys(rc, 2023, gstep, 0)
                else:
                    set_keys(rc, 2023, gstep, 2)

            # the reference's own state-dict EMA works on CRNN_fpn: oracle/train.py restates main.py:86-100
            loss, parts, outs = otrain.mt_step(rc, rp, tc, tp, opt, xr, xr_ema, xs, ts, gstep, rampup_length=50 * 10,
                                               ema_flavour="state_dict", dropout_hook=hook)
            rec[f"loss{it}"] = float(loss)
            for k, v in parts.items():
                rec[f"{k}{it}"] = float(v)
            if it == 0:
                for k, v in outs["grads"].items():
                    rec["g_" + k] = sample(v)
                    rec["gn_" + k] = float(v.double().norm())
                rec["strong0"] = outs["strong"].numpy()
                rec["strong_ema0"] = outs["strong_ema"].numpy()
                rec["weak0"] = outs["weak"].numpy()
        ssd, tsd = rc.state_dict(), tc.state_dict()
        for k in PARAM_KEYS:
            rec["s_" + k] = ssd[k].numpy().reshape(-1)[:2048]
            rec["t_" + k] = tsd[k].numpy().reshape(-1)[:2048]
        rec["t_nbt_fcn"] = int(tsd["cnn.bn_fcn.num_batches_tracked"])
        rec["s_nbt_fcn"] = int(ssd["cnn.bn_fcn.num_batches_tracked"])
        rec["t_nbt0"] = int(tsd["cnn.cnn.batchnorm0.num_batches_tracked"])
        rec["s_nbt0"] = int(ssd["cnn.cnn.batchnorm0.num_batches_tracked"])
        rec["s_dense_w"] = rp.dense.weight.detach().numpy().reshape(-1)
        rec["t_dense_w"] = tp.dense.weight.detach().numpy().reshape(-1)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, {k: v for k, v in rec.items() if isinstance(v, (float, int))})

    print("fpn golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
