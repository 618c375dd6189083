from .CRNN import CRNN, Predictor  # noqa: F401
