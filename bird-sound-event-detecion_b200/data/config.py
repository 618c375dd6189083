"""Constants of the hot path (reference: src/data/config.py:47-63,67-109).  Dataset paths, model
names and the 2.8 GB of module-level torch.randn (config.py:90-91) are out of scope."""
import math

# audio                                                     src/data/config.py:47-57
sr = 32000
seg_sec = 10
n_window = 2048
hop_size = 255
n_mels = 128
mel_f_min = 0.
mel_f_max = 16000.
max_len_seconds = 10.
max_frames = math.ceil(max_len_seconds * sr / hop_size)     # 1255
pooling_time_ratio = 4

noise_snr = 30
median_window_s = 0.45
out_nb_frames_1s = sr / hop_size / pooling_time_ratio

batch_size = 12
n_epoch = 300
n_epoch_rampup = 50
adjust_lr = False
max_learning_rate = 0.0005
default_learning_rate = 0.0005
max_consistency_cost = 1

bird_list = [
    "EATO", "WOTH", "BCCH", "BTNW", "TUTI",
    "NOCA", "REVI", "AMCR", "BLJA", "OVEN",
    "COYE", "BGGN", "SCTA", "AMRE", "KEWA",
    "BHCO", "BHVI", "HETH", "RBWO", "BAWW",
]
