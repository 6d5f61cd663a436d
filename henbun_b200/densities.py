"""Log-densities, mirror of Henbun/densities.py.  ``gaussian`` (:25-27) is the hot-path one and runs
this repo's CUDA kernels forward and backward; the others (:30-103) are elementwise epilogue
variants kept for API completeness and evaluated with torch elementwise ops."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .tf_wraps import log_sum_exp


def _t(x, like):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float32), device=like.device)


def _like(*xs):
    for x in xs:
        if isinstance(x, torch.Tensor):
            return x
    raise TypeError("at least one argument must be a device tensor")


def gaussian(x, mu, var):
    like = _like(x, mu, var)
    return ops.gaussian_logpdf(_t(x, like), _t(mu, like), _t(var, like))


def lognormal(x, mu, var):
    lnx = torch.log(x)
    return gaussian(lnx, mu, var) - lnx


def bernoulli(p, y):
    return torch.log(torch.where(y == 1, p, 1 - p))


def poisson(lamb, y):
    return y * torch.log(lamb) - lamb - torch.lgamma(y + 1.)


def exponential(lamb, y):
    return - y / lamb - torch.log(lamb)


def gamma(shape, scale, x):
    like = _like(shape, scale, x)
    shape, scale, x = _t(shape, like), _t(scale, like), _t(x, like)
    return -shape * torch.log(scale) - torch.lgamma(shape) + (shape - 1.) * torch.log(x) - x / scale


def student_t(x, mean, scale, deg_free):
    like = _like(x, mean, scale, deg_free)
    x, mean, scale, deg_free = _t(x, like), _t(mean, like), _t(scale, like), _t(deg_free, like)
    const = torch.lgamma((deg_free + 1.) * 0.5) - torch.lgamma(deg_free * 0.5) \
        - 0.5 * (torch.log(torch.square(scale)) + torch.log(deg_free) + float(np.log(np.pi)))
    return const - 0.5 * (deg_free + 1.) * torch.log(1. + (1. / deg_free) * (torch.square((x - mean) / scale)))


def beta(alpha, beta, y):
    y = torch.clamp(y, 1e-6, 1 - 1e-6)
    return (alpha - 1.) * torch.log(y) + (beta - 1.) * torch.log(1. - y) + torch.lgamma(alpha + beta) \
        - torch.lgamma(alpha) - torch.lgamma(beta)


def laplace(mu, sigma, y):
    like = _like(mu, sigma, y)
    mu, sigma, y = _t(mu, like), _t(sigma, like), _t(y, like)
    return - torch.abs(mu - y) / sigma - torch.log(2. * sigma)


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
    return log_sum_exp(torch.stack([logp0 + torch.log(fraction), logp1 + torch.log(1.0 - fraction)], dim=-1), axis=-1)
