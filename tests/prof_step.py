"""Minimal driver for ncu: builds the bench models, runs W warm-up + K mean-teacher steps (and optionally the
frontend) and nothing else.   python tests/prof_step.py [--steps K] [--warmup W] [--frontend] [--precision tf32]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bsed_b200 import engine  # noqa: E402
from bsed_b200.main import MeanTeacherTrainer  # noqa: E402
from bsed_b200.models import CRNN, CRNN_fpn, Predictor  # noqa: E402
from bsed_b200.utilities import synth  # noqa: E402
from bsed_b200.utilities.utils import weights_init  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frontend", action="store_true")
    ap.add_argument("--precision", default=None)
    ap.add_argument("--model", default="crnn", choices=["crnn", "crnn_fpn"])
    ap.add_argument("--ada", action="store_true", help="the SCMT + adversarial domain adaptation iteration (config 3)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(2023)

    def make():
        cls = CRNN_fpn if a.model == "crnn_fpn" else CRNN
        m, p = cls(**engine.REFERENCE_CRNN_KWARGS), Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
        weights_init(m)
        weights_init(p)
        return m.to(dev).train(), p.to(dev).train()

    if a.ada:
        from bsed_b200.DA.cdan_frame import ConditionalDomainAdversarialLoss
        from bsed_b200.main import AdaptationTrainer
        from bsed_b200.models.CRNN import Clip_Discriminator
        model, predictor = make()
        ema_model, ema_predictor = make()
        disc = Clip_Discriminator(256).to(dev).train()
        tr = AdaptationTrainer(model, predictor, ema_model, ema_predictor, ConditionalDomainAdversarialLoss(disc), lr=5e-4,
                               momentum=0.9, weight_decay=1e-4, n_syn=12, n_real=12, dropout_seed=2023)
        xs = torch.from_numpy(synth.make_logmel_like(12, seed=3)).to(dev)
        x = torch.from_numpy(synth.make_logmel_like(12, seed=40)).to(dev)
        ts = torch.from_numpy(synth.make_targets(12, seed=5)).to(dev)
        tw = torch.from_numpy(synth.make_targets(12, seed=8)).max(1)[0].to(dev)
        for i in range(a.warmup + a.steps):
            tr.step(x, x, tw, xs, ts, i, 5000)
        torch.cuda.synchronize()
        print("prof_step (ada) done: precision", tr.plan.precision)
        return
    model, predictor = make()
    ema_model, ema_predictor = make()
    tr = MeanTeacherTrainer(model, predictor, ema_model, ema_predictor, lr=5e-4, n_syn=12, n_real=12,
                            precision=a.precision)
    x = torch.from_numpy(synth.make_logmel_like(12, seed=1)).to(dev)
    xs = torch.from_numpy(synth.make_logmel_like(12, seed=2)).to(dev)
    ts = torch.from_numpy(synth.make_targets(12, seed=3)).to(dev)
    for i in range(a.warmup + a.steps):
        tr.step(x, x, xs, ts, i, 5000)
    torch.cuda.synchronize()
    if a.frontend:
        clips = torch.from_numpy(synth.make_clips(8, seed=5)).to(dev).repeat(32, 1)
        for _ in range(2):
            engine.amp_to_db(engine.melspec(clips), 1255)
        torch.cuda.synchronize()
    print("prof_step done: precision", tr.plan.precision)


if __name__ == "__main__":
    main()
