"""Frontend stage B with the reference's transform classes (src/data/Transforms.py).

get_transforms(frames, scaler=None, add_axis=0, noise_dict_params=None) returns a Compose of
  [AugmentGaussianNoise] -> ApplyLog -> PadOrTrunc -> ToTensor -> [Normalize]        (:304-322)
Calling it on a (features, label) sample runs the whole chain as ONE fused CUDA pass
(csrc/frontend.cu: amp_to_db), not transform by transform; the classes carry the configuration.
`Compose.batch` is the device-resident batched form used by the training loop.
"""
import numpy as np
import torch

from .. import engine


class Transform:
    def transform_data(self, data):
        raise NotImplementedError("transforms are fused; apply them through Compose")

    def transform_label(self, label):
        return label


class ApplyLog(Transform):
    """librosa.amplitude_to_db (ref=1, amin=1e-5, top_db=80)   (:74-86)"""


class PadOrTrunc(Transform):
    """(:112-139)"""

    def __init__(self, nb_frames, apply_to_label=False):
        self.nb_frames = nb_frames
        self.apply_to_label = apply_to_label


class AugmentGaussianNoise(Transform):
    """(:142-197) only the snr form is in scope: noisy = x + N(0, sqrt(mean_t(x^2) 10^(-snr/10)))"""

    def __init__(self, mean=0., std=None, snr=None):
        if snr is None:
            raise NotImplementedError("only AugmentGaussianNoise(snr=...) is supported")
        self.mean, self.std, self.snr = mean, std, snr


class ToTensor(Transform):
    """(:200-227)"""

    def __init__(self, unsqueeze_axis=None):
        self.unsqueeze_axis = unsqueeze_axis


class Normalize(Transform):
    """(:230-250)"""

    def __init__(self, scaler):
        self.scaler = scaler


def pad_trunc_seq(x, max_len):
    """(:89-109) host helper for labels."""
    if x.shape[-2] <= max_len:
        pad = [(0, 0)] * (x.ndim - 2) + [(0, max_len - x.shape[-2]), (0, 0)]
        return np.pad(x, pad, mode="constant")
    return x[..., :max_len, :]


class Compose(object):
    def __init__(self, transforms):
        self.transforms = list(transforms)
        self._plan()

    def _plan(self):
        kinds = [type(t) for t in self.transforms]
        core = [k for k in kinds if k in (ApplyLog, PadOrTrunc, ToTensor)]
        if core != [ApplyLog, PadOrTrunc, ToTensor]:
            raise NotImplementedError("supported chain: [AugmentGaussianNoise] ApplyLog PadOrTrunc ToTensor [Normalize]")
        self.noise = next((t for t in self.transforms if isinstance(t, AugmentGaussianNoise)), None)
        self.pad = next(t for t in self.transforms if isinstance(t, PadOrTrunc))
        self.tot = next(t for t in self.transforms if isinstance(t, ToTensor))
        self.norm = next((t for t in self.transforms if isinstance(t, Normalize)), None)

    def add_transform(self, transform):
        return Compose(self.transforms + [transform])

    def _scaler(self, device):
        if self.norm is None:
            return None, None
        sc = self.norm.scaler
        mean = torch.as_tensor(np.asarray(sc.mean_, dtype=np.float32).reshape(-1), device=device)
        std = torch.as_tensor(np.asarray(sc.std_, dtype=np.float32).reshape(-1), device=device)
        return mean.contiguous(), std.contiguous()

    def batch(self, mel, unit_noise=None):
        """mel (B, t_in, 128) CUDA amplitude-mel -> clean (B,1,frames,128), or (clean, noisy) when the
        chain has AugmentGaussianNoise (unit_noise: standard-normal draws of mel's shape; drawn with
        torch if omitted)."""
        mean, std = self._scaler(mel.device)
        frames = self.pad.nb_frames
        clean = engine.amp_to_db(mel, frames, None, 0.0, mean, std)
        ax = self.tot.unsqueeze_axis
        if ax is not None:
            clean = clean.unsqueeze(ax + 1)
        if self.noise is None:
            return clean
        if unit_noise is None:
            unit_noise = torch.randn_like(mel)
        noisy = engine.amp_to_db(mel, frames, unit_noise.contiguous(), float(self.noise.snr), mean, std)
        if ax is not None:
            noisy = noisy.unsqueeze(ax + 1)
        return clean, noisy

    def __call__(self, sample):
        """(features (t,128) ndarray, label ndarray) -> ((clean, noisy) | clean, label) CPU tensors,
        like the reference's per-sample pipeline (host in, host out)."""
        data, label = sample
        dev = torch.device("cuda", torch.cuda.current_device())
        mel = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).to(dev)[None]
        noise = None
        if self.noise is not None:
            # np.random.normal(0, std, shape) == std * standard_normal(shape) draw for draw
            noise = torch.from_numpy(np.random.standard_normal(data.shape).astype(np.float32)).to(dev)[None]
        out = self.batch(mel, noise)
        lab = torch.from_numpy(np.asarray(label)).float()
        if self.pad.apply_to_label:
            lab = torch.from_numpy(pad_trunc_seq(np.asarray(label), self.pad.nb_frames)).float()
        if isinstance(out, tuple):
            return (out[0][0].cpu(), out[1][0].cpu()), lab
        return out[0].cpu(), lab

    def __repr__(self):
        return "Compose(" + ", ".join(type(t).__name__ for t in self.transforms) + ")"


def get_transforms(frames, scaler=None, add_axis=0, noise_dict_params=None, combine_channels_args=None):
    if combine_channels_args is not None:
        raise NotImplementedError("CombineChannels (source separation) is outside the hot path")
    transf = []
    if noise_dict_params is not None:
        transf.append(AugmentGaussianNoise(**noise_dict_params))
    transf.extend([ApplyLog(), PadOrTrunc(nb_frames=frames), ToTensor(unsqueeze_axis=add_axis)])
    if scaler is not None:
        transf.append(Normalize(scaler=scaler))
    return Compose(transf)
