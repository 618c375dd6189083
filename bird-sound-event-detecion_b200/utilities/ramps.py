"""Ramp schedules (reference: src/utilities/ramps.py:4-31).  Host-side scalars feeding the
consistency weight and the learning rate."""
import math


def _clip(v, lo, hi):
    return max(lo, min(hi, v))


def exp_rampup(current, rampup_length):
    """exp(-5 (1 - t)^2), t = clip(current / rampup_length, 0, 1)."""
    if rampup_length == 0:
        return 1.0
    phase = 1.0 - _clip(float(current), 0.0, float(rampup_length)) / rampup_length
    return float(math.exp(-5.0 * phase * phase))


def cosine_rampdown(current, rampdown_length):
    assert 0 <= current <= rampdown_length
    return float(.5 * (math.cos(math.pi * current / rampdown_length) + 1))


def sigmoid_rampdown(current, rampup_length):
    """Named 'rampdown' in the reference but is exp(-12.5 (1 - t)^2) (ramps.py:24-31)."""
    if rampup_length == 0:
        return 1.0
    phase = 1.0 - _clip(float(current), 0.0, float(rampup_length)) / rampup_length
    return float(math.exp(-12.5 * phase * phase))
