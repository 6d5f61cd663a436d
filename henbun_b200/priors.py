"""Priors over (transformed) parameters -- the public surface of Henbun/priors.py (Normal :44-52, Gaussian :55-65,
LogNormal :68-78, Gamma :81-91, Laplace :94-104, Uniform :107-116) on top of this package's density kernels.

Design: a prior is (a member of the densities.py kernel family, where the random variable sits among that member's
operands, its hyper-parameters).  ``logp`` is one elementwise kernel launch of that member plus the sum the reference
applies (priors are univariate, so an array argument means the sum of the element log-densities); the backward is the
member's backward kernel.  Hyper-parameters are kept as float32 host arrays (attributes named as in the reference,
``mu``/``var``/``shape``/``scale``/``sigma``) and travel to the device once per call.
"""
from __future__ import annotations

import numpy as np
import torch

from .param import Parameterized
from . import ops

np_float_type = np.float32
_LOG_2PI = float(np.log(2.0 * np.pi))


def _hyper(value):
    return np.atleast_1d(np.asarray(value, dtype=np_float_type))


class Prior(Parameterized):
    """Base class.  Subclasses either fill in the three class attributes below or override ``logp``.

    _family : name of the densities.py member evaluated by the kernel family (ops.DENSITY_KINDS)
    _fields : attribute names of the hyper-parameters, in the order given to ``__init__``
    _x_slot : index of the random variable among the member's operands (hyper-parameters fill the other slots in order)
    _tag    : prefix used by ``__str__``
    """
    _family = None
    _fields = ()
    _x_slot = 0
    _tag = "prior"

    def __init__(self, *hyper):
        Parameterized.__init__(self)
        if len(hyper) != len(self._fields):
            raise TypeError("%s takes %d hyper-parameter(s)" % (type(self).__name__, len(self._fields)))
        for name, value in zip(self._fields, hyper):
            setattr(self, name, _hyper(value))

    def _operands(self, x):
        hyper = [torch.as_tensor(getattr(self, name), device=x.device) for name in self._fields]
        hyper.insert(self._x_slot, x if x.dtype == torch.float32 else x.to(torch.float32))
        return hyper

    def logp(self, x):
        """Sum of the element log-densities at x."""
        if self._family is None:
            raise NotImplementedError
        return torch.sum(ops.density(self._family, *self._operands(x)))

    def __str__(self):
        if self._family is None:
            raise NotImplementedError
        return self._tag + "(" + ",".join(str(getattr(self, name)) for name in self._fields) + ")"


class Normal(Prior):
    """Standard normal prior; this is the prior of variationals.Normal / Gaussian, whose KL never calls it because the
    sampler kernel already reduces -1/2 sum(logdet + u^2 - z^2) (variationals.py:225-230).  Used by the generic
    Variational._KL only."""

    def __init__(self):
        Prior.__init__(self)

    def logp(self, x):
        return -0.5 * (_LOG_2PI * x.numel() + torch.sum(torch.square(x)))

    def __str__(self):
        return "N(0,1)"


class Gaussian(Prior):
    _family, _fields, _x_slot, _tag = "gaussian", ("mu", "var"), 0, "N"

    def __init__(self, mu, var):
        Prior.__init__(self, mu, var)


class LogNormal(Prior):
    _family, _fields, _x_slot, _tag = "lognormal", ("mu", "var"), 0, "logN"

    def __init__(self, mu, var):
        Prior.__init__(self, mu, var)


class Gamma(Prior):
    _family, _fields, _x_slot, _tag = "gamma", ("shape", "scale"), 2, "Ga"

    def __init__(self, shape, scale):
        Prior.__init__(self, shape, scale)


class Laplace(Prior):
    _family, _fields, _x_slot, _tag = "laplace", ("mu", "sigma"), 2, "Lap."

    def __init__(self, mu, sigma):
        Prior.__init__(self, mu, sigma)


class Uniform(Prior):
    """Flat density on [lower, upper]: logp is a constant per element, no kernel involved."""

    def __init__(self, lower=0, upper=1):
        Prior.__init__(self)
        self.lower, self.upper = lower, upper
        self.log_height = -float(np.log(upper - lower))

    def logp(self, x):
        return self.log_height * float(x.numel())

    def __str__(self):
        return "U(%s,%s)" % (self.lower, self.upper)
