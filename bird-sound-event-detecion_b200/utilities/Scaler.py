"""Dataset mean/std per mel bin (reference: src/utilities/Scaler.py:38-110): float64 running mean
of per-file means over the time axis, std from the mean of squares."""
import numpy as np
import torch


class Scaler:
    def __init__(self):
        self.mean_ = None
        self.mean_of_square_ = None
        self.std_ = None

    def means(self, dataset):
        counter = 0
        for sample in dataset:
            x = sample[0] if isinstance(sample, (tuple, list)) and len(sample) == 2 else sample
            if isinstance(x, (tuple, list)):
                x = x[0]
            x = x.numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
            m = np.mean(x, axis=-2, dtype=np.float64)
            m2 = np.mean(x.astype(np.float64) ** 2, axis=-2, dtype=np.float64)
            self.mean_ = m if self.mean_ is None else self.mean_ + m
            self.mean_of_square_ = m2 if self.mean_of_square_ is None else self.mean_of_square_ + m2
            counter += 1
        self.mean_ /= counter
        self.mean_of_square_ /= counter
        return self

    def calculate_scaler(self, dataset):
        self.means(dataset)
        self.std_ = np.sqrt(self.mean_of_square_ - self.mean_ ** 2)
        return self.mean_, self.std_

    def normalize(self, batch):
        if isinstance(batch, torch.Tensor):
            return torch.Tensor((batch.numpy() - self.mean_) / self.std_)
        return (batch - self.mean_) / self.std_

    def state_dict(self):
        return {"mean_": self.mean_, "mean_of_square_": self.mean_of_square_}

    def load_state_dict(self, state):
        self.mean_ = np.asarray(state["mean_"])
        self.mean_of_square_ = np.asarray(state["mean_of_square_"])
        self.std_ = np.sqrt(self.mean_of_square_ - self.mean_ ** 2)
