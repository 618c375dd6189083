"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (imported from /root/reference/src) in the
build container.  The fixtures are what pins oracle/crnn.py and oracle/train.py; the GPU box has no
/root/reference, so it only ever sees these files.

    python tests/make_golden.py

Inputs and weights are regenerated from seeds by the tests (numpy PCG64 streams), so the fixtures
hold outputs plus small checksums of the regenerated inputs.
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


def reference_models(dropout):
    from models.CRNN import CRNN, Predictor   # the reference's own classes
    kw = dict(ocrnn.CRNN_KWARGS)
    kw["dropout"] = dropout
    return CRNN(**kw), Predictor(**ocrnn.PREDICTOR_KWARGS)


def load_oracle_weights_into_reference(ref_crnn, ref_pred, seed, linear_std):
    oc = ocrnn.OracleCRNN(**{**ocrnn.CRNN_KWARGS, "dropout": 0.0})
    op = ocrnn.OraclePredictor(**ocrnn.PREDICTOR_KWARGS)
    ocrnn.reference_style_init(oc, op, seed, linear_std)
    sd = oc.state_dict()
    # the reference CNN overrides state_dict()/load_state_dict() (models/CNN.py:71-75): CRNN.load_state_dict
    # does not route through it, so load per sub-module
    ref_crnn.cnn.load_state_dict({k[len("cnn."):]: v for k, v in sd.items() if k.startswith("cnn.")})
    ref_crnn.rnn.load_state_dict({k[len("rnn."):]: v for k, v in sd.items() if k.startswith("rnn.")})
    ref_pred.load_state_dict(op.state_dict())
    return oc, op


def inject_hash_dropout(ref_crnn, p):
    """Replace the reference's nn.Dropout modules by the hash dropout the CUDA kernels use."""
    for i in range(7):
        setattr(ref_crnn.cnn.cnn, f"dropout{i}", ocrnn.HashDropout(p, i))
    ref_crnn.dropout = ocrnn.HashDropout(p, ocrnn.STREAM_RNN_OUT)


def set_keys(mod, seed, step, batch_offset):
    for m in mod.modules():
        if isinstance(m, ocrnn.HashDropout):
            m.key = ocrnn.mix_key(seed, step, m.stream)
            m.batch_offset = batch_offset


def ref_state_dict(ref_crnn):
    sd = {"cnn." + k: v for k, v in ref_crnn.cnn.state_dict().items()}
    sd.update({"rnn." + k: v for k, v in ref_crnn.rnn.state_dict().items()})
    return sd


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)

    # ---------------------------------------------------------------- key list / shapes
    rc, rp = reference_models(0.5)
    keys = list(rc.state_dict().keys())
    shapes = [tuple(v.shape) for v in rc.state_dict().values()]
    np.savez(os.path.join(OUT, "state_dict_keys.npz"), keys=np.array(keys), shapes=np.array([str(s) for s in shapes]),
             pred_keys=np.array(list(rp.state_dict().keys())))

    # ---------------------------------------------------------------- eval forward (config 1 shapes, B=2)
    x = torch.from_numpy(synth.make_logmel_like(2, seed=11))
    rc, rp = reference_models(0.5)
    load_oracle_weights_into_reference(rc, rp, seed=5, linear_std=0.2)
    rc.eval(); rp.eval()
    with torch.no_grad():
        enc, d_in = rc(x)
        strong, weak = rp(enc)
    np.savez_compressed(os.path.join(OUT, "crnn_eval.npz"), x_sum=float(x.double().sum()), enc=enc.numpy()[:, ::8],
                        enc_sum=float(enc.double().sum()), strong=strong.numpy(), weak=weak.numpy())

    # ---------------------------------------------------------------- train-mode forward with hash dropout
    rc, rp = reference_models(0.5)
    load_oracle_weights_into_reference(rc, rp, seed=5, linear_std=0.2)
    inject_hash_dropout(rc, 0.5)
    rc.train(); rp.train()
    set_keys(rc, seed=2023, step=3, batch_offset=0)
    with torch.no_grad():
        enc, _ = rc(x)
        strong, weak = rp(enc)
    sd = ref_state_dict(rc)
    np.savez_compressed(os.path.join(OUT, "crnn_train_fwd.npz"), strong=strong.numpy(), weak=weak.numpy(),
                        enc_sum=float(enc.double().sum()),
                        rm0=sd["cnn.batchnorm0.running_mean"].numpy(), rv0=sd["cnn.batchnorm0.running_var"].numpy(),
                        rm6=sd["cnn.batchnorm6.running_mean"].numpy(), rv6=sd["cnn.batchnorm6.running_var"].numpy(),
                        nbt=int(sd["cnn.batchnorm3.num_batches_tracked"]))

    # ---------------------------------------------------------------- mean-teacher steps (2 syn + 2 real clips)
    for name, p_drop in (("mt_step_nodrop", 0.0), ("mt_step_drop", 0.5)):
        rc, rp = reference_models(p_drop)
        load_oracle_weights_into_reference(rc, rp, seed=5, linear_std=0.2)
        tc, tp = reference_models(p_drop)
        load_oracle_weights_into_reference(tc, tp, seed=6, linear_std=0.2)
        for m in (rc, tc):
            inject_hash_dropout(m, p_drop)
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

            # the state-dict EMA of the reference crashes on plain CRNN; use its intent:
            # blend every entry of cnn.cnn.* / rnn.* -- implemented on sub-modules
            loss, parts, outs = otrain.mt_step(rc, rp, tc, tp, opt, xr, xr_ema, xs, ts, gstep, rampup_length=50 * 10,
                                               ema_flavour="none", dropout_hook=hook)
            a = otrain.ema_alpha(0.999, gstep + 1)
            with torch.no_grad():
                for sm, tm in ((rc.cnn.cnn, tc.cnn.cnn), (rc.rnn, tc.rnn), (rp, tp)):
                    ssd, tsd = sm.state_dict(), tm.state_dict()
                    for k in tsd.keys():
                        tsd[k] = tsd[k].clone() * a + ssd[k].clone() * (1.0 - a)
                    tm.load_state_dict(tsd)
            rec[f"loss{it}"] = float(loss)
            for k, v in parts.items():
                rec[f"{k}{it}"] = float(v)
            if it == 0:
                for k, v in outs["grads"].items():
                    g = v.numpy().reshape(-1)
                    rec["g_" + k] = g if g.size <= 4096 else g[:: max(1, g.size // 4096)][:4096]
                    rec["gn_" + k] = float(np.sqrt((g.astype(np.float64) ** 2).sum()))
                rec["strong0"] = outs["strong"].numpy()
                rec["strong_ema0"] = outs["strong_ema"].numpy()
                rec["weak0"] = outs["weak"].numpy()
        ssd = ref_state_dict(rc)
        tsd = ref_state_dict(tc)
        for k in ("cnn.conv0.weight", "cnn.conv3.bias", "cnn.batchnorm2.weight", "cnn.glu4.linear.weight",
                  "rnn.rnn.weight_hh_l0", "rnn.rnn.bias_ih_l1_reverse", "cnn.batchnorm1.running_var",
                  "cnn.batchnorm5.running_mean"):
            rec["s_" + k] = ssd[k].numpy().reshape(-1)[:2048]
            rec["t_" + k] = tsd[k].numpy().reshape(-1)[:2048]
        rec["t_nbt"] = int(tsd["cnn.batchnorm0.num_batches_tracked"])
        rec["s_nbt"] = int(ssd["cnn.batchnorm0.num_batches_tracked"])
        rec["s_dense_w"] = rp.dense.weight.detach().numpy().reshape(-1)
        rec["t_dense_w"] = tp.dense.weight.detach().numpy().reshape(-1)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, {k: v for k, v in rec.items() if isinstance(v, float)})

    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
