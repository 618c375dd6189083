from .CRNN import CRNN, CRNN_fpn, Clip_Discriminator, Predictor, set_dropout_seed  # noqa: F401
from .ResNet import Net_resnet  # noqa: F401
