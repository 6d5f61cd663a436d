"""Variational distributions: mirror of Henbun/variationals.py (Variational :31-209, Normal :213-230,
Gaussian :232-291, OffsetGaussian :293-314, Beta :316-381).

The reference builds the sampling ops once and draws ONE sample per ``session.run``
(variationals.py:107,127).  Here a sample is drawn at the start of every objective evaluation
(``_new_run``): S independent draws at once when the Optimizer was compiled with ``n_samples=S``
(leading sample axis; the objective's sums then run over the S samples and the driver divides by S).
The sampler and the one-sample KL are one fused CUDA kernel (csrc/sampler.cu); eps comes from
device Philox unless injected (``eps={variational: tensor}``, the reference's
``feed_dict={variational.u: eps}``).
"""
from __future__ import annotations

from functools import reduce

import numpy as np
import torch

from . import transforms, priors, densities, ops
from . import trace as _trace
from .tf_wraps import clip
from .param import Variable, graph_key, Parameterized, _is, _in_collection

np_float_type = np.float32
LOG2PI = float(np.log(2.0 * np.pi))


class RunContext(object):
    """Per-evaluation sampling context handed down the tree by Model._new_run."""

    def __init__(self, n_samples=1, seed=0, offset=0, eps=None):
        self.n_samples = int(n_samples)
        self.seed = int(seed)
        self.offset = int(offset)
        self.eps = eps or {}
        self.first_sample = 0                 # multi-GPU: this rank's first sample ...
        self.total_samples = self.n_samples   # ... of the samples drawn over all ranks

    def take(self, count):
        """Reserve `count` Philox stream positions (rounded up to a multiple of 4)."""
        off = self.offset
        self.offset += (int(count) + 3) // 4 * 4
        return off

    def take_sharded(self, per_sample):
        """Stream position of this rank's first draw of a [n_samples, per_sample] block: the block of ALL ranks'
        samples is reserved, this rank reads its contiguous window of it (parallel.rank_philox_offset)."""
        total = max(int(self.total_samples), int(self.n_samples))
        off = self.take(total * int(per_sample))
        return off + (int(self.first_sample) * int(per_sample)) // 4 * 4


_default_ctx = RunContext()


class Variational(Parameterized):
    """Multivariate-Gaussian variational posterior, 'diagonal' or 'fullrank' over the flattened
    `shape` axes (variationals.py:31-110)."""

    def __init__(self, shape, n_layers=[], n_batch=None, q_shape='diagonal', mean=0.0, stddev=1.0,
                 prior=None, transform=transforms.Identity(), collections=[graph_key.VARIABLES]):
        Parameterized.__init__(self)
        self._shape = list([shape]) if isinstance(shape, int) else list(shape)
        self.n_layers = list([n_layers]) if isinstance(n_layers, int) else list(n_layers)
        self.n_batch = n_batch
        self.size = int(reduce(np.multiply, self._shape))
        self.collections = collections
        assert (q_shape in ['diagonal', 'fullrank'])
        self.q_shape = q_shape
        self.q_mu = Variable(self.size, n_layers=n_layers, n_batch=self.n_batch, mean=mean, stddev=0.1 * stddev,
                             collections=collections)
        if self.q_shape == 'diagonal':
            # log(q_sqrt) is stored (variationals.py:87-92)
            self.q_sqrt = Variable(self.size, n_layers=n_layers, n_batch=self.n_batch, mean=np.log(stddev), stddev=0.1,
                                   collections=collections)
        else:
            self.q_sqrt = Variable([self.size, self.size], n_layers=n_layers, n_batch=self.n_batch, mean=stddev,
                                   stddev=0.1 * stddev, collections=collections)
        self.transform = transform
        self.prior = prior
        self.u = None
        self._tensor = None
        self.transformed_tensor = None
        self._kl_core = None
        self._ctx = _default_ctx
        self._S = 1

    # ---- run protocol -------------------------------------------------------------------------
    @property
    def is_local(self):
        return _is(self.collections, graph_key.LOCAL)

    def _new_run(self, ctx):
        Parameterized._new_run(self, ctx)
        self._ctx = ctx
        self._tensor = None
        self.transformed_tensor = None
        self._kl_core = None
        self.u = None

    def _params(self):
        object.__getattribute__(self, 'q_mu')
        q_mu = object.__getattribute__(self, 'q_mu').tensor()
        q_sqrt = object.__getattribute__(self, 'q_sqrt').tensor()
        return q_mu, q_sqrt

    def _draw(self):
        """Sample for the current run (non-LOCAL: lazily at first use; LOCAL: from feed())."""
        q_mu, q_sqrt = self._params()
        if q_mu is None:
            raise ValueError('local variable ' + self.long_name + ' is not fed.')
        ctx = self._ctx
        S = ctx.n_samples
        eps = ctx.eps.get(self, None)
        if eps is not None and not isinstance(eps, torch.Tensor):
            eps = torch.as_tensor(np.asarray(eps), dtype=torch.float32).to(q_mu.device)
        self._tensor = self._sample(eps, q_mu, q_sqrt, S)
        self.transformed_tensor = self.transform.tf_forward(self._tensor)
        self._S = S

    def _sample(self, u, q_mu=None, q_sqrt=None, S=None):
        """variationals.py:131-150.  u: injected i.i.d. draws shaped [S, *q_mu.shape] ([*q_mu.shape] when S
        is 1), or None for device Philox."""
        if q_mu is None:
            q_mu, q_sqrt = self._params()
        if S is None:
            S = 1 if (u is None or u.dim() == q_mu.dim()) else u.shape[0]
        lead = tuple(q_mu.shape)
        if self.q_shape == 'diagonal':
            ue = None if u is None else u.reshape((S,) + lead)
            off = self._ctx.take_sharded(q_mu.numel()) if u is None else 0
            z, kl = ops.sample_diag(q_mu, q_sqrt, ue, self._ctx.seed, off, S)
            self.u = ue
        else:
            n = self.size
            B = int(q_mu.numel() // n)
            if u is None:
                off = self._ctx.take_sharded(B * n)
                ue = ops.randn_philox((S, B, n), self._ctx.seed, off, q_mu.device)
            else:
                ue = u.reshape(S, B, n)
            eps_b = ue.permute(1, 0, 2).contiguous()                       # [B,S,n]
            zb, kl = ops.sample_tril(q_mu.reshape(B, n), q_sqrt.reshape(B, n, n), eps_b)
            z = zb.permute(1, 0, 2).reshape((S,) + lead)
            self.u = ue.reshape((S,) + lead)
        self._kl_core = kl
        return z if S > 1 else z.reshape(lead)

    def tensor(self):
        """In tf_mode this object is seen as a sample from the variational distribution
        (variationals.py:112-119)."""
        if _trace.active():
            return _trace.Sym('sample', self, getattr(self, '_sym_feed', None))
        if self._tensor is None:
            if self.is_local:
                return None
            self._draw()
        t = self.transformed_tensor
        lead = [self._S] if self._S > 1 else []
        if not self.is_local and self.n_batch is None:
            return clip(t.reshape(lead + self.n_layers + self._shape))
        # the batch axis is named, not inferred: an empty one (n_batch = 0, test_variationals.py:288-322) is legal upstream
        # and torch refuses to infer a -1 for a tensor without elements
        batch = int(t.shape[len(lead) + len(self.n_layers)])
        return clip(t.reshape(lead + self.n_layers + [batch] + self._shape))

    def feed(self, x):
        """LOCAL: route the encoder output into q_mu / q_sqrt and sample (variationals.py:121-129)."""
        if _trace.active():
            object.__setattr__(self, '_sym_feed', x)
            return
        Parameterized.feed(self, x)
        if self.is_local:
            self._draw()

    def _einsum_matmul(self):
        """Index string of the reference's (commented-out) einsum sampler (variationals.py:152-176)."""
        alphabet = 'abcdefghijklmnopqrstuvwxyz'
        n = len(self.n_layers)
        if not self.is_local and self.n_batch is None:
            return alphabet[:n + 2] + ',' + alphabet[:n] + alphabet[n + 1] + '->' + alphabet[:n + 1]
        return alphabet[:n + 3] + ',' + alphabet[:n + 1] + alphabet[n + 2] + '->' + alphabet[:n + 2]

    @property
    def logdet(self):
        """Log-determinant of the posterior (variationals.py:178-186)."""
        q_mu, q_sqrt = self._params()
        if self.q_shape == 'diagonal':
            return 2.0 * q_sqrt
        return torch.log(torch.square(torch.diagonal(q_sqrt, dim1=-2, dim2=-1)))

    def KL(self, collection=None):
        if _trace.active():
            return _trace.Sym('KL', self, collection)
        if collection is None or _in_collection(collection, self.collections):
            return self._KL()
        return np.zeros([], dtype=np_float_type)

    def _ensure_sampled(self):
        if self._tensor is None:
            if self.is_local:
                raise ValueError('local variable ' + self.long_name + ' is not fed.')
            self._draw()

    def _entropy_term(self):
        """-0.5*sum(log2pi + logdet + u^2) from the kernel's fused reduction:
        kl_core = -0.5*sum(logdet + u^2 - z^2)."""
        z = self._tensor
        return self._kl_core - 0.5 * torch.sum(torch.square(z)) - 0.5 * LOG2PI * z.numel()

    def _KL(self):
        """One-sample MC KL with an arbitrary prior / transform (variationals.py:198-209)."""
        self._ensure_sampled()
        kl = self._entropy_term()
        if self.prior is not None:
            kl = kl - torch.sum(self.prior.logp(self.transformed_tensor))
            kl = kl - torch.sum(self.transform.tf_log_jacobian(self._tensor))
        return kl


class Normal(Variational):
    """Normal prior, no transform; KL shortcut (variationals.py:213-230)."""

    def __init__(self, shape, n_layers=[], n_batch=None, q_shape='diagonal', mean=0.0, stddev=1.0,
                 collections=[graph_key.VARIABLES]):
        Variational.__init__(self, shape, q_shape=q_shape, n_layers=n_layers, n_batch=n_batch, mean=mean,
                             stddev=stddev, prior=priors.Normal(), transform=transforms.Identity(),
                             collections=collections)

    def _KL(self):
        # -0.5*sum(logdet + u^2 - z^2): exactly the reduction fused into the sampler kernel
        self._ensure_sampled()
        return self._kl_core


class Gaussian(Normal):
    """scale * Normal (variationals.py:232-291); the KL is that of the unscaled Normal."""

    def __init__(self, shape, n_layers=[], n_batch=None, q_shape='diagonal', mean=0.0, stddev=1.0,
                 collections=[graph_key.VARIABLES], scale_shape=None, scale_n_layers=None):
        if np.abs(mean) < stddev:
            scale_mean = stddev
            q_mean = mean / stddev
            q_std = 1.0
        else:
            scale_mean = np.abs(mean)
            q_mean = 1.0
            q_std = stddev / np.abs(mean)
        Variational.__init__(self, shape, q_shape=q_shape, n_layers=n_layers, n_batch=n_batch, mean=q_mean,
                             stddev=q_std, prior=priors.Normal(), transform=transforms.Identity(),
                             collections=collections)
        scale_shape = scale_shape or [1 for s in self._shape]
        scale_layer = scale_n_layers or [1 for s in self.n_layers]
        self.scale = Variable(scale_shape, n_layers=scale_layer, n_batch=n_batch, mean=scale_mean,
                              stddev=0.1 * scale_mean, transform=transforms.positive, collections=collections)

    def tensor(self):
        if _trace.active():
            return _trace.Sym('sample', self, getattr(self, '_sym_feed', None))
        t = Normal.tensor(self)
        if t is None:
            return None
        return object.__getattribute__(self, 'scale').tensor() * t


class OffsetGaussian(Gaussian):
    """Gaussian + offset (variationals.py:293-314)."""

    def __init__(self, shape, n_layers=[], n_batch=None, q_shape='diagonal', mean=0.0, stddev=1.0,
                 collections=[graph_key.VARIABLES], scale_shape=None, scale_n_layers=None):
        Gaussian.__init__(self, shape=shape, n_layers=n_layers, n_batch=n_batch, q_shape=q_shape, mean=0.0,
                          stddev=stddev, collections=collections, scale_shape=scale_shape,
                          scale_n_layers=scale_n_layers)
        offset_shape = scale_shape or [1 for s in self._shape]
        offset_layer = scale_n_layers or [1 for s in self.n_layers]
        self.offset = Variable(offset_shape, n_layers=offset_layer, n_batch=n_batch, mean=mean, stddev=0.1 * mean,
                               collections=collections)

    def tensor(self):
        if _trace.active():
            return _trace.Sym('sample', self, getattr(self, '_sym_feed', None))
        t = Gaussian.tensor(self)
        if t is None:
            return None
        return t + object.__getattribute__(self, 'offset').tensor()


class Beta(Variational):
    """Logistic-transformed Gaussian with a Beta prior whose alpha, beta are Variables
    (variationals.py:316-381)."""

    def __init__(self, shape, n_layers=[], n_batch=None, q_shape='diagonal', mean=0.0, stddev=1.0,
                 collections=[graph_key.VARIABLES], scale_shape=None, scale_n_layers=None):
        Variational.__init__(self, shape, q_shape=q_shape, n_layers=n_layers, n_batch=n_batch, mean=mean,
                             stddev=stddev, transform=transforms.Logistic(), collections=collections)
        scale_shape = scale_shape or [1 for s in self._shape]
        scale_layer = scale_n_layers or [1 for s in self.n_layers]
        self.alpha = Variable(scale_shape, n_layers=scale_layer, n_batch=n_batch, mean=1.0, stddev=0.1,
                              transform=transforms.positive, collections=collections)
        self.beta = Variable(scale_shape, n_layers=scale_layer, n_batch=n_batch, mean=1.0, stddev=0.1,
                             transform=transforms.positive, collections=collections)

    def _KL(self):
        self._ensure_sampled()
        kl = self._entropy_term()
        # NOTE the reference guards this block with `if self.prior is not None` although Beta never sets a
        # prior (variationals.py:376), so upstream the Beta density is silently dropped; we keep the
        # upstream behaviour for parity.
        if self.prior is not None:
            a = object.__getattribute__(self, 'alpha').tensor()
            b = object.__getattribute__(self, 'beta').tensor()
            t = self.transformed_tensor.reshape([-1] + self.n_layers + ([] if self.n_batch is None else [self.n_batch]) + self._shape)
            kl = kl - torch.sum(densities.beta(a, b, t))
            kl = kl - torch.sum(self.transform.tf_log_jacobian(self._tensor))
        return kl
