"""Constrained-parameter transforms, mirror of Henbun/transforms.py:73-180,271.

Each transform has numpy ``forward``/``backward`` (host side, used by assignment and ``.value``) and
``tf_forward``/``tf_log_jacobian`` acting on device tensors inside ``tf_mode``: one CUDA kernel each, forward
and backward (csrc/transforms.cu, C ABI hb_transform_fwd / _bwd / _logjac / _logjac_bwd) -- the same code
serves a one-element hyper-parameter and the [S, n] sample tensor of a transformed variational
(variationals.py:110,129,204-208).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class Transform(object):
    def forward(self, x):
        raise NotImplementedError

    def backward(self, y):
        raise NotImplementedError

    def tf_forward(self, x):
        raise NotImplementedError

    def tf_log_jacobian(self, x):
        raise NotImplementedError

    def free_state_size(self, variable_shape):
        return np.prod(variable_shape)

    def __str__(self):
        raise NotImplementedError

    def __getstate__(self):
        return self.__dict__.copy()

    def __setstate__(self, d):
        self.__dict__ = d


class Identity(Transform):
    def tf_forward(self, x):
        return x

    def forward(self, x):
        return x

    def backward(self, y):
        return y

    def tf_log_jacobian(self, x):
        return torch.zeros((1,), dtype=x.dtype, device=x.device)

    def __str__(self):
        return '(none)'


class Exp(Transform):
    def __init__(self, lower=1e-6):
        self._lower = lower

    def tf_forward(self, x):
        return ops.transform_forward("exp", x, self._lower)

    def forward(self, x):
        return np.exp(x) + self._lower

    def backward(self, y):
        return np.log(y - self._lower)

    def tf_log_jacobian(self, x):
        return ops.transform_log_jacobian("exp", x, self._lower)

    def __str__(self):
        return '+ve'


class Log1pe(Transform):
    """y = log(1 + exp(x)) + lower  (softplus; Henbun/transforms.py:110-143)."""

    def __init__(self, lower=1e-6):
        self._lower = lower

    def forward(self, x):
        return np.logaddexp(0.0, x) + self._lower

    def tf_forward(self, x):
        return ops.transform_forward("log1pe", x, self._lower)

    def tf_log_jacobian(self, x):
        return ops.transform_log_jacobian("log1pe", x, self._lower)

    def backward(self, y):
        y = np.asarray(y)
        return np.log(np.expm1(y - self._lower)).astype(y.dtype if y.dtype.kind == 'f' else np.float64)

    def __str__(self):
        return '+ve'


class Logistic(Transform):
    def __init__(self, a=0., b=1.):
        assert b > a
        self.a, self.b = a, b

    def tf_forward(self, x):
        return ops.transform_forward("logistic", x, self.a, self.b)

    def forward(self, x):
        return self.a + (self.b - self.a) / (1. + np.exp(-x))

    def backward(self, y):
        return -np.log((self.b - self.a) / (y - self.a) - 1.)

    def tf_log_jacobian(self, x):
        return ops.transform_log_jacobian("logistic", x, self.a, self.b)

    def __str__(self):
        return '[' + str(self.a) + ', ' + str(self.b) + ']'


positive = Log1pe()
