"""Whole-step entry points of the C ABI, for drivers that do not need the autograd tape in between.

``LinearOperatorStep`` = BASELINE config 5 (full-covariance ``variationals.Normal([n])`` through a dense forward
operator with ``densities.gaussian``): forward, backward and the TF-1 Adam update are two C calls
(``hb_linop_elbo_local`` / ``hb_linop_elbo_update``, csrc/linop.cu) with ONE all-reduce in between when the operator
is row-sharded over ranks.  Same numbers as the model written against the Python API
(tests/test_gpu_linop.py compares the two), without the n x n gradient ever being materialised.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, parallel
from ._lib import ptr, stream, check


class LinearOperatorStep(object):
    """params packing: [ q_sqrt (n*n) | q_mu (n) | var (1) ], all in free space (var = softplus(free) + 1e-6).

    A_rows / y_rows: this rank's rows of the operator and of the data (device tensors, resident).
    m_total: rows of the whole operator (defaults to this rank's count = single GPU)."""

    def __init__(self, A_rows: torch.Tensor, y_rows: torch.Tensor, n_samples: int, m_total: int = None, seed: int = 0,
                 lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999, epsilon: float = 1e-8):
        self.lib = _lib.load()
        self.A = _lib.f32(A_rows).contiguous()
        self.y = _lib.f32(y_rows).contiguous().reshape(-1)
        M, n = self.A.shape
        if self.y.numel() != M:
            raise ValueError("y must have one entry per row of A")
        self.n, self.S = int(n), int(n_samples)
        self.cfg = _lib.LinopConfig(int(M), int(m_total if m_total is not None else M), self.n, self.S, int(seed), 0)
        dev = self.A.device
        self.count = int(self.lib.hb_linop_param_count(C.byref(self.cfg)))
        self.params = torch.zeros(self.count, device=dev)
        self.m = torch.zeros(self.count, device=dev)
        self.v = torch.zeros(self.count, device=dev)
        self.zbar_stats = torch.empty(self.S * self.n + 4, device=dev)
        self.out4 = torch.zeros(4, device=dev)
        self.ws_bytes = int(self.lib.hb_linop_workspace_bytes(C.byref(self.cfg)))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.step_dev = torch.ones(1, dtype=torch.int32, device=dev)      # Adam's t (starts at 1)
        self.hyper = (float(lr), float(beta1), float(beta2), float(epsilon))
        self._stride = (self.S * self.n + 3) // 4 * 4

    # ---- parameter access (free space) ----
    @property
    def q_sqrt(self):
        return self.params[:self.n * self.n].view(self.n, self.n)

    @property
    def q_mu(self):
        return self.params[self.n * self.n:self.n * self.n + self.n]

    @property
    def var_free(self):
        return self.params[self.n * self.n + self.n:]

    def set_params(self, q_mu, q_sqrt, var_free):
        dev = self.params.device
        self.q_sqrt.copy_(torch.as_tensor(np.asarray(q_sqrt, np.float32), device=dev))
        self.q_mu.copy_(torch.as_tensor(np.asarray(q_mu, np.float32), device=dev).reshape(-1))
        self.var_free.copy_(torch.as_tensor(np.asarray(var_free, np.float32), device=dev).reshape(-1))

    # ---- the step ----
    def local(self, eps=None, step_index=0):
        """Sampler, F = Z A^T, log-lik, partial Zbar = R A on this rank's rows -> self.zbar_stats."""
        self.cfg.offset = int(step_index) * self._stride
        check(self.lib.hb_linop_elbo_local(C.byref(self.cfg), ptr(self.A), ptr(self.y), ptr(self.params),
                                           ptr(eps) if eps is not None else None, ptr(self.zbar_stats), ptr(self.ws),
                                           self.ws_bytes, stream()), "hb_linop_elbo_local")
        return self.zbar_stats

    def update(self, grads: torch.Tensor = None, apply_adam: bool = True):
        """mu-bar, var-bar and the fused (gradient of q_sqrt + Adam) pass; returns out4 = {ELBO, loglik, kl, 0}."""
        lr, b1, b2, e = self.hyper
        check(self.lib.hb_linop_elbo_update(C.byref(self.cfg), ptr(self.params), ptr(self.zbar_stats), ptr(grads),
                                            ptr(self.m) if apply_adam else None, ptr(self.v) if apply_adam else None,
                                            lr, b1, b2, e, ptr(self.step_dev), 0, ptr(self.out4), ptr(self.ws),
                                            self.ws_bytes, stream()), "hb_linop_elbo_update")
        if apply_adam:
            check(self.lib.hb_increment_i32(ptr(self.step_dev), stream()), "hb_increment_i32")
        return self.out4

    def step(self, eps=None, step_index=0):
        """One ELBO + gradient + Adam step; the only collective is the all-reduce of [S*n + 4] floats."""
        self.local(eps, step_index)
        parallel.allreduce_sum_(self.zbar_stats)
        return self.update()

    def value_and_grads(self, eps=None):
        """ELBO and d ELBO / d params without touching the parameters (parity tests)."""
        g = torch.zeros(self.count, device=self.params.device)
        self.local(eps)
        parallel.allreduce_sum_(self.zbar_stats)
        out = self.update(grads=g, apply_adam=False)
        return out, g
