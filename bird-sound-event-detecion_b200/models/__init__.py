from .CRNN import CRNN, Predictor, set_dropout_seed  # noqa: F401
