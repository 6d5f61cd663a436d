"""Small helpers, mirror of Henbun/tf_wraps.py: eye (:26-30), clip (:33-39), log_sum_exp (:42-48).
The reference's dead native-op loader (tf_wraps.py:50-71) has no counterpart: the native boundary of
this package is libhenbun_b200.so (include/henbun_b200.h)."""
from __future__ import annotations

import torch

from ._settings import settings


def eye(N, device=None):
    return torch.eye(int(N), dtype=torch.float32, device=device if device is not None else 'cuda')


def clip(tensor):
    """Clip to [clip_value_min, clip_value_max] when settings.numerics.clip_by_value is on."""
    if settings.numerics.clip_by_value:
        return torch.clamp(tensor, settings.numerics.clip_value_min, settings.numerics.clip_value_max)
    return tensor


def log_sum_exp(tensor, axis=-1):
    maxtensor = torch.amax(tensor, dim=axis, keepdim=True)
    return maxtensor.squeeze(axis) + torch.log(torch.sum(torch.exp(tensor - maxtensor), dim=axis))
