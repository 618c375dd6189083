"""Shared builders for the tests: seeded oracle models, inputs, comparisons."""
import os

import numpy as np
import torch

from oracle import crnn as ocrnn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def oracle_models(seed=5, linear_std=0.2, dropout=0.0, train=False):
    oc = ocrnn.OracleCRNN(**{**ocrnn.CRNN_KWARGS, "dropout": dropout})
    op = ocrnn.OraclePredictor(**ocrnn.PREDICTOR_KWARGS)
    ocrnn.reference_style_init(oc, op, seed, linear_std)
    oc.train(train)
    op.train(train)
    return oc, op


def oracle_fpn_models(seed=5, linear_std=0.2, dropout=0.0, train=False):
    oc = ocrnn.OracleCRNNfpn(**{**ocrnn.CRNN_KWARGS, "dropout": dropout})
    op = ocrnn.OraclePredictor(**ocrnn.PREDICTOR_KWARGS)
    ocrnn.reference_style_init(oc, op, seed, linear_std)
    oc.train(train)
    op.train(train)
    return oc, op


def bsed_fpn_models(oc, op, dropout=0.0, device="cuda", precision=None):
    """Our CRNN_fpn / Predictor carrying the oracle's weights (state-dict keys are identical to the reference's)."""
    from bsed_b200 import engine
    from bsed_b200.models import CRNN_fpn, Predictor
    kw = dict(engine.REFERENCE_CRNN_KWARGS)
    kw["dropout"] = dropout
    m = CRNN_fpn(**kw, precision=precision)
    p = Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
    m.load_state_dict(oc.state_dict())
    p.load_state_dict(op.state_dict())
    return m.to(device), p.to(device)


def bsed_models(oc, op, dropout=0.0, device="cuda"):
    """Our CRNN/Predictor carrying the oracle's weights (state-dict keys are identical)."""
    from bsed_b200 import engine
    from bsed_b200.models import CRNN, Predictor
    kw = dict(engine.REFERENCE_CRNN_KWARGS)
    kw["dropout"] = dropout
    m = CRNN(**kw)
    p = Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
    m.load_state_dict(oc.state_dict())
    p.load_state_dict(op.state_dict())
    return m.to(device), p.to(device)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_abs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


def logmel_close(a, b, rtol=1e-4):
    """north_star tolerance for log-mel: |a-b| <= 1e-4 * max(1, |b|)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    return float(err.max()), float((err > rtol).mean())
