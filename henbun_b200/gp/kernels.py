"""GP kernels, mirror of Henbun/gp/kernels.py (Kern :28, UnitStationary :36-101, UnitRBF :103-111,
UnitCsymRBF :113-131).  K is one CUDA kernel (distance + exp + optional jitter, csrc/gram.cu);
``Cholesky`` fuses Gram + jitter + blocked potrf in one buffer (ops.kern_cholesky)."""
from __future__ import annotations

import numpy as np
import torch

from .. import transforms, ops
from ..param import Variable, Parameterized, graph_key
from ..variationals import Variational
from .._settings import settings

np_float_type = np.float32


class Kern(Parameterized):
    def __init__(self):
        Parameterized.__init__(self)
        self.scoped_keys.extend(['K', 'Kdiag'])


class UnitStationary(Kern):
    _csym = False

    def __init__(self, lengthscales=np.ones(1), n_batch=None, collections=[graph_key.VARIABLES]):
        Kern.__init__(self)
        if isinstance(lengthscales, np.ndarray):
            self.lengthscales = Variable(lengthscales.shape, transform=transforms.positive, collections=collections)
            self.lengthscales = lengthscales          # set initial values
        elif isinstance(lengthscales, (Variable, Variational)):
            self.lengthscales = lengthscales
        else:
            raise TypeError
        self.scoped_keys.extend(['square_dist', 'euclid_dist', 'Cholesky'])

    def square_dist(self, X, X2=None):
        """-2 Xe Xe'^T + |Xe|^2 + |Xe'|^2 (gp/kernels.py:54-84); only used when a caller asks for the
        distance itself -- K() never materialises it."""
        Xeff = X / self.lengthscales
        Xs = torch.sum(torch.square(Xeff), -1)
        if X2 is None:
            return -2 * ops.matmul(Xeff, Xeff, transpose_b=True) + Xs.unsqueeze(-1) + Xs.unsqueeze(-2)
        X2eff = X2 / self.lengthscales
        X2s = torch.sum(torch.square(X2eff), -1)
        return -2 * ops.matmul(Xeff, X2eff, transpose_b=True) + Xs.unsqueeze(-1) + X2s.unsqueeze(-2)

    def euclid_dist(self, X, X2):
        return torch.sqrt(self.square_dist(X, X2) + 1e-12)

    def Kdiag(self, X):
        return torch.ones(X.shape[:-1], dtype=torch.float32, device=X.device)

    def K(self, X, X2=None):
        raise NotImplementedError

    def Cholesky(self, X):
        """Cholesky factor of K(X) + jitter*I; [n,d] -> [n,n], [N,n,d] -> [N,n,n] (gp/kernels.py:93-101)."""
        from .. import trace as _trace
        if isinstance(X, _trace.Sym):
            return _trace.Sym('cholesky', self, X)
        X = _dev(X)
        return ops.kern_cholesky(X, self.lengthscales, settings.numerics.jitter_level, self._csym)


def _dev(x):
    from .. import trace as _trace
    if isinstance(x, _trace.Sym):
        raise _trace.TraceError("this kernel method is not traced")
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np_float_type)).cuda()


class UnitRBF(UnitStationary):
    """K(x,x2) = exp(-(x-x2)^2 / (2 l^2))."""

    def K(self, X, X2=None):
        return ops.rbf_K(_dev(X), None if X2 is None else _dev(X2), self.lengthscales, False)


class UnitCsymRBF(UnitStationary):
    """exp(-(x-x2)^2/(2 l^2)) + exp(-(x+x2)^2/(2 l^2))."""
    _csym = True

    def K(self, X, X2=None):
        return ops.rbf_K(_dev(X), None if X2 is None else _dev(X2), self.lengthscales, True)

    def Kdiag(self, X):
        Xeff = X / self.lengthscales
        Xs = torch.sum(torch.square(Xeff), -1)
        return torch.ones_like(Xs) + torch.exp(-2 * Xs)


# names used by BASELINE.json's wording
RBF = UnitRBF
Stationary = UnitStationary
