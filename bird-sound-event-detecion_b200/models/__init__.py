from .CRNN import CRNN, CRNN_fpn, Clip_Discriminator, Predictor, set_dropout_seed  # noqa: F401
