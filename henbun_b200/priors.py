"""Priors, mirror of Henbun/priors.py.  ``Normal`` (:44-52) is the one on the hot path (its log-density
is folded into the sampler kernel's KL reduction for variationals.Normal); the rest (:55-116) are
elementwise variants."""
from __future__ import annotations

import numpy as np
import torch

from .param import Parameterized
from . import densities

np_float_type = np.float32


class Prior(Parameterized):
    def logp(self, x):
        raise NotImplementedError

    def __str__(self):
        raise NotImplementedError


class Normal(Prior):
    """Zero-mean unit-variance Gaussian prior."""

    def logp(self, x):
        return -0.5 * torch.sum(float(np.log(2 * np.pi)) + torch.square(x))

    def __str__(self):
        return "N(" + str(0) + "," + str(1) + ")"


class Gaussian(Prior):
    def __init__(self, mu, var):
        Prior.__init__(self)
        self.mu = np.atleast_1d(np.array(mu, np_float_type))
        self.var = np.atleast_1d(np.array(var, np_float_type))

    def logp(self, x):
        return torch.sum(densities.gaussian(x, self.mu, self.var))

    def __str__(self):
        return "N(" + str(self.mu) + "," + str(self.var) + ")"


class LogNormal(Prior):
    def __init__(self, mu, var):
        Prior.__init__(self)
        self.mu = np.atleast_1d(np.array(mu, np_float_type))
        self.var = np.atleast_1d(np.array(var, np_float_type))

    def logp(self, x):
        return torch.sum(densities.lognormal(x, self.mu, self.var))

    def __str__(self):
        return "logN(" + str(self.mu) + "," + str(self.var) + ")"


class Gamma(Prior):
    def __init__(self, shape, scale):
        Prior.__init__(self)
        self.shape = np.atleast_1d(np.array(shape, np_float_type))
        self.scale = np.atleast_1d(np.array(scale, np_float_type))

    def logp(self, x):
        return torch.sum(densities.gamma(self.shape, self.scale, x))

    def __str__(self):
        return "Ga(" + str(self.shape) + "," + str(self.scale) + ")"


class Laplace(Prior):
    def __init__(self, mu, sigma):
        Prior.__init__(self)
        self.mu = np.atleast_1d(np.array(mu, np_float_type))
        self.sigma = np.atleast_1d(np.array(sigma, np_float_type))

    def logp(self, x):
        return torch.sum(densities.laplace(self.mu, self.sigma, x))

    def __str__(self):
        return "Lap.(" + str(self.mu) + "," + str(self.sigma) + ")"


class Uniform(Prior):
    def __init__(self, lower=0, upper=1):
        Prior.__init__(self)
        self.log_height = - np.log(upper - lower)
        self.lower, self.upper = lower, upper

    def logp(self, x):
        return self.log_height * float(x.numel())

    def __str__(self):
        return "U(" + str(self.lower) + "," + str(self.upper) + ")"
