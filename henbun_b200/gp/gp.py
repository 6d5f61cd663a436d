"""GP samplers, mirror of Henbun/gp/gp.py (GP :9-50, SparseGP :53-192)."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..param import Variable, Parameterized, graph_key
from .._settings import settings


class GP(Parameterized):
    """Samples from the GP posterior: u L^T with L = chol(K(x,x)) (gp/gp.py:37-50)."""

    def __init__(self, kern):
        Parameterized.__init__(self)
        self.kern = kern

    def samples(self, x, u):
        """x: [n,d]; u: variational samples [N,n] -> [N,n]."""
        L = self.kern.Cholesky(x)
        return ops.matmul(u, L, transpose_b=True)


class SparseGP(GP):
    """Sparse GP with inducing points z [m,d] (gp/gp.py:53-192).  Listed as a 'next' row in
    SURVEY.md 8f: built from the same kernels (Gram, potrf, TRSM, GEMM); the inducing points z are trainable
    (gradient through K(x, z) and chol(K(z, z)) via hb_rbf_gram_bwd_x2)."""

    def __init__(self, kern, z, collections=[graph_key.VARIABLES]):
        GP.__init__(self, kern)
        self.z = Variable(shape=z.shape, collections=collections)
        self.z = z
        self.m = len(z)

    def _z(self):
        return object.__getattribute__(self, 'z').tensor()

    def samples(self, x, u, q_shape='diagonal'):
        assert (q_shape in ['diagonal', 'neglected', 'fullrank'])
        jitter = settings.numerics.jitter_level
        N = u.shape[0]
        LnT = self._effective_LT(x)
        if x.dim() == 2:
            samples = ops.matmul(u, LnT)                                     # [N,n]
        else:
            samples = ops.matmul(u.unsqueeze(1), LnT).squeeze(1)             # [N,1,m] @ [N,m,n]
        if q_shape == 'neglected':
            return samples
        ctx = getattr(self.highest_parent, '_run_ctx', None)
        seed = ctx.seed if ctx is not None else 0
        if q_shape == 'diagonal':
            diag_cov = self._additional_cov(x, LnT, 'diagonal')
            off = ctx.take(diag_cov.numel()) if ctx is not None else 0
            return samples + torch.sqrt(torch.abs(diag_cov)) * ops.randn_philox(tuple(x.shape[:-1]), seed, off, x.device)
        n = x.shape[-2]
        cov = self._additional_cov(x, LnT, 'fullrank') + jitter * torch.eye(n, device=x.device)
        chol = ops.cholesky(cov)
        if x.dim() == 2:
            off = ctx.take(N * n) if ctx is not None else 0
            return samples + ops.matmul(ops.randn_philox((N, n), seed, off, x.device), chol, transpose_b=True)
        off = ctx.take(N * n) if ctx is not None else 0
        return samples + ops.matmul(ops.randn_philox((N, 1, n), seed, off, x.device), chol, transpose_b=True).squeeze(1)

    def _effective_LT(self, x):
        """Lm^{-1} K(z,x)  (gp/gp.py:146-174)."""
        z = self._z()
        Lm = self.kern.Cholesky(z)
        if x.dim() == 2:
            Kxz = self.kern.K(x, z)                                          # [n,m]
            return ops.trsm_right(Lm, Kxz, 1).transpose(0, 1)                # (K_xz Lm^{-T})^T = Lm^{-1} K_zx
        elif x.dim() == 3:
            N = x.shape[0]
            zt = z.unsqueeze(0).expand(N, -1, -1).contiguous()
            Kxz = self.kern.K(x, zt)                                         # [N,n,m]
            n = x.shape[1]
            sol = ops.trsm_right(Lm, Kxz.reshape(N * n, self.m), 1)          # rows are independent
            return sol.reshape(N, n, self.m).transpose(1, 2)
        raise ValueError('shape is not specified for tensor x')

    def _additional_cov(self, x, LnT, q_shape):
        """Knn - Knm Kmm^-1 Kmn (gp/gp.py:177-192)."""
        if q_shape == 'diagonal':
            return self.kern.Kdiag(x) - torch.sum(torch.square(LnT), -2)
        Knn = self.kern.K(x)
        return Knn - ops.matmul(LnT.transpose(-1, -2).contiguous(), LnT.contiguous())
