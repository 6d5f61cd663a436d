"""Log-densities, mirror of Henbun/densities.py.  Every elementwise density (:25-74, :95-103) is one CUDA
kernel forward and one backward from the same family (csrc/density_family.cu, C ABI hb_density_logpdf /
hb_density_logpdf_bwd); ``multivariate_normal`` (:75-91) runs on the blocked triangular solve."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _t(x, like):
    if isinstance(x, torch.Tensor):
        return x if x.dtype == torch.float32 else x.to(torch.float32)
    return torch.as_tensor(np.asarray(x, dtype=np.float32), device=like.device)


def _like(*xs):
    for x in xs:
        if isinstance(x, torch.Tensor):
            return x
    raise TypeError("at least one argument must be a device tensor")


def _call(name, *args):
    from . import trace as _trace
    if _trace.is_sym(*args):
        raise _trace.TraceError(f"densities.{name} is not traced")
    like = _like(*args)
    return ops.density(name, *[_t(a, like) for a in args])


def gaussian(x, mu, var):
    from . import trace as _trace
    if _trace.is_sym(x, mu, var):
        return _trace.Sym('gaussian', x, mu, var)
    return _call("gaussian", x, mu, var)


def lognormal(x, mu, var):
    return _call("lognormal", x, mu, var)


def bernoulli(p, y):
    return _call("bernoulli", p, y)


def poisson(lamb, y):
    return _call("poisson", lamb, y)


def exponential(lamb, y):
    return _call("exponential", lamb, y)


def gamma(shape, scale, x):
    return _call("gamma", shape, scale, x)


def student_t(x, mean, scale, deg_free):
    return _call("student_t", x, mean, scale, deg_free)


def beta(alpha, beta, y):
    return _call("beta", alpha, beta, y)


def laplace(mu, sigma, y):
    return _call("laplace", mu, sigma, y)


def multivariate_normal(x, mu, L):
    """L is the Cholesky factor of the covariance (densities.py:75-91); the triangular solve runs on
    the blocked TRSM of this package: alpha^T = d^T L^{-T}."""
    d = x - mu
    if d.dim() == 1:
        d = d[:, None]
    alpha_t = ops.trsm_right(L, d.transpose(0, 1).contiguous(), 1)       # [cols, n]
    num_col = d.shape[1]
    num_dims = d.shape[0]
    ret = - 0.5 * num_dims * num_col * float(np.log(2 * np.pi))
    ret = ret - num_col * torch.sum(torch.log(torch.diagonal(L)))
    ret = ret - 0.5 * torch.sum(torch.square(alpha_t))
    return ret


def bimixture(fraction, logp0, logp1):
    """log(fraction*exp(logp0) + (1-fraction)*exp(logp1)) (densities.py:95-103), one fused kernel instead of
    log / stack / reduce_max / exp / reduce_sum / log."""
    return _call("bimixture", fraction, logp0, logp1)
