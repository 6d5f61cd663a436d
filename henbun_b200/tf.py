"""A small ``tf``-shaped namespace so objectives written for the reference port by changing one
import (``import henbun_b200.tf as tf``).  Heavy ops (matmul, cholesky, triangular solve) dispatch to
this package's CUDA kernels; elementwise / reduction glue on already-reduced or O(n) tensors uses
torch's device ops."""
from __future__ import annotations

import numpy as np
import torch

from . import ops as _ops
from . import nn as _nn
from . import train  # noqa: F401  (tf.train.AdamOptimizer)
from . import trace as _trace

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32


def _t(x):
    if isinstance(x, _trace.Sym):
        raise _trace.TraceError("this tf op is not traced")
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float32)).cuda()


def constant(value, dtype=None):
    return _t(value)


convert_to_tensor = constant


def matmul(a, b, transpose_a=False, transpose_b=False):
    if _trace.is_sym(a, b):
        return _trace.Sym('matmul', a, b, ta=bool(transpose_a), tb=bool(transpose_b))
    return _ops.matmul(_t(a), _t(b), transpose_a, transpose_b)


def cholesky(a):
    return _ops.cholesky(_t(a))


def matrix_triangular_solve(matrix, rhs, lower=True, adjoint=False):
    """L^{-1} rhs (lower, not adjoint) through the right-sided blocked TRSM: (rhs^T L^{-T})^T."""
    if not lower or adjoint:
        raise NotImplementedError
    return _ops.trsm_right(matrix, _t(rhs).transpose(-1, -2).contiguous(), 1).transpose(-1, -2)


def reduce_sum(x, axis=None, keep_dims=False):
    if isinstance(x, _trace.Sym):
        return _trace.Sym('reduce_sum', x, axis=axis, keep_dims=keep_dims)
    x = _t(x)
    return torch.sum(x) if axis is None else torch.sum(x, dim=axis, keepdim=keep_dims)


def reduce_mean(x, axis=None, keep_dims=False):
    x = _t(x)
    return torch.mean(x) if axis is None else torch.mean(x, dim=axis, keepdim=keep_dims)


def reduce_max(x, axis=None, keep_dims=False):
    x = _t(x)
    return torch.max(x) if axis is None else torch.amax(x, dim=axis, keepdim=keep_dims)


def sqrt(x): return _trace.Sym('sqrt', x) if isinstance(x, _trace.Sym) else torch.sqrt(_t(x))
def square(x): return torch.square(_t(x))
def exp(x): return torch.exp(_t(x))
def log(x): return torch.log(_t(x))
def abs(x): return torch.abs(_t(x))
def negative(x): return -_t(x)
def lgamma(x): return torch.lgamma(_t(x))
def identity(x): return _t(x)
def add(a, b): return _t(a) + _t(b)
def multiply(a, b): return _t(a) * _t(b)
def expand_dims(x, axis): return _t(x).unsqueeze(axis)
def squeeze(x, axis=None): return _t(x).squeeze() if axis is None else _t(x).squeeze(axis[0] if isinstance(axis, (list, tuple)) else axis)
def reshape(x, shape): return _t(x).reshape(list(shape))
def shape(x): return _t(x).shape
def transpose(x, perm=None): return _t(x).t() if perm is None else _t(x).permute(*perm)
def ones(shape, dtype=None): return torch.ones(list(shape), device='cuda')
def zeros(shape, dtype=None): return torch.zeros(list(shape), device='cuda')
def ones_like(x, dtype=None): return torch.ones_like(_t(x))
def zeros_like(x, dtype=None): return torch.zeros_like(_t(x))
def cast(x, dtype): return _t(x).to(dtype)
def stack(values, axis=0): return torch.stack([_t(v) for v in values], dim=axis)
def tile(x, multiples): return _t(x).repeat(*multiples)
def clip_by_value(x, lo, hi): return torch.clamp(_t(x), lo, hi)
def matrix_band_part(x, lower, upper):
    if lower == -1 and upper == 0:
        return torch.tril(_t(x))
    if lower == 0 and upper == -1:
        return torch.triu(_t(x))
    raise NotImplementedError
def matrix_diag_part(x): return torch.diagonal(_t(x), dim1=-2, dim2=-1)
def diag_part(x): return torch.diagonal(_t(x))
def diag(x): return torch.diag(_t(x))


sigmoid = _nn.sigmoid
tanh = _nn.tanh


class nn:  # noqa: N801  (tf.nn)
    sigmoid = staticmethod(_nn.sigmoid)
    relu = staticmethod(_nn.relu)
    tanh = staticmethod(_nn.tanh)

    @staticmethod
    def softplus(x):
        return torch.nn.functional.softplus(_t(x), threshold=1e9)
