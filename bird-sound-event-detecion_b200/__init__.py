"""bird-sound-event-detecion_b200 -- the sound-event-detection hot path of
fumchin/bird-sound-event-detecion rebuilt for NVIDIA B200 (sm_100a).

Layout mirrors the reference's `src/` tree for the path in scope:
  data/preprocess.py   data/Transforms.py   data/config.py        (log-mel frontend)
  models/CRNN.py  models/CNN.py  models/RNN.py                    (CRNN + Predictor)
  utilities/ramps.py  ManyHotEncoder.py  Scaler.py  utils.py      (host helpers)
  evaluation_measures.py                                          (get_predictions post-processing)
  main.py                                                         (train_mt, update_ema_variables)
All arithmetic runs in csrc/ (hand-written CUDA behind the C ABI of include/bsed.h).  Import this
directory as `bsed_b200` (the repo root carries a small alias package, since the directory name is
not a Python identifier).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "engine", "models", "data", "utilities", "evaluation_measures", "main", "pseudo_labeling"]
