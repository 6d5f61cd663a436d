"""CPU oracle for the Henbun Monte-Carlo ELBO hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain restatement (torch-CPU / numpy, fp64 by default) of the
arithmetic the reference performs on the hot path named by BASELINE.json.  It is
the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``henbun_b200``) never imports anything from ``oracle/``.

Every function cites the reference file:line it follows (paths relative to the
upstream Henbun tree).  The arithmetic of the reference lives in TensorFlow 1.x
(un-vendored, un-pinned: ``setup.py:34-37`` ``tensorflow>=1.0``), which is not
installable here, so TF op semantics are restated from their documentation:

* ``tf.cholesky``            lower factor of a symmetric PD matrix
* ``tf.matrix_band_part(x,-1,0)``  lower triangle including the diagonal
* ``tf.matrix_triangular_solve``   lower=True, adjoint=False by default
* ``tf.nn.softplus``         log(1+exp(x))
* ``tf.train.AdamOptimizer`` lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA;
                             theta -= lr_t*m/(sqrt(v)+eps)   (eps outside sqrt)

Pinning status (see DESIGN.md "Oracle"): the sampler, logdet, K, Kdiag,
Cholesky reconstruction, NeuralNet forward, transforms and the LOCAL feed order
are pinned by the reference's own known-answer tests, re-created in
``tests/golden/`` (``make_golden.py`` executes the *unmodified* reference
modules on a numpy-backed TF-1 shim).  ``densities.gaussian``, the ELBO scalar,
every gradient value, the Cholesky gradient and the Adam trajectory have no
reference test: for those rows parity is UNPINNED beyond the shim run.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch

LOG2PI = math.log(2.0 * math.pi)


def _t(x, dtype=torch.float64):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


# --------------------------------------------------------------------------
# transforms  (Henbun/transforms.py:73-180, :271)
# --------------------------------------------------------------------------

def softplus(x):
    """tf.nn.softplus, overflow-safe: max(x,0)+log1p(exp(-|x|))."""
    return torch.clamp(x, min=0) + torch.log1p(torch.exp(-torch.abs(x)))


def log1pe_forward(x, lower=1e-6):
    """transforms.Log1pe.tf_forward (transforms.py:133-134): softplus(x)+lower."""
    return softplus(x) + lower


def log1pe_backward(y, lower=1e-6):
    """transforms.Log1pe.backward (transforms.py:139-140): log(exp(y-lower)-1)."""
    y = np.asarray(y, dtype=np.float64)
    return np.log(np.expm1(y - lower))


def log1pe_log_jacobian(x):
    """transforms.Log1pe.tf_log_jacobian (transforms.py:136-137)."""
    return -torch.sum(torch.log(1.0 + torch.exp(-x)))


def exp_forward(x, lower=1e-6):
    """transforms.Exp.tf_forward (transforms.py:94-95)."""
    return torch.exp(x) + lower


def exp_log_jacobian(x):
    """transforms.Exp.tf_log_jacobian (transforms.py:103-104)."""
    return torch.sum(x)


def logistic_forward(x, a=0.0, b=1.0):
    """transforms.Logistic.tf_forward (transforms.py:154-156)."""
    return a + (b - a) / (1.0 + torch.exp(-x))


def logistic_log_jacobian(x, a=0.0, b=1.0):
    """transforms.Logistic.tf_log_jacobian (transforms.py:165-166)."""
    return torch.sum(x - 2.0 * torch.log(torch.exp(x) + 1.0) + math.log(b - a))


def clip(x, enabled=False, lo=-50.0, hi=50.0):
    """tf_wraps.clip (tf_wraps.py:33-39); defaults from henbunrc:12-14."""
    return torch.clamp(x, lo, hi) if enabled else x


# --------------------------------------------------------------------------
# Variational sampler / logdet / KL   (Henbun/variationals.py)
# --------------------------------------------------------------------------

def sample_diag(q_mu, q_sqrt, u):
    """Variational._sample, 'diagonal' (variationals.py:138-142).
    q_sqrt stores log-std (variationals.py:87-92).  u broadcasts over leading
    (sample) axes."""
    return q_mu + torch.exp(q_sqrt) * u


def sample_fullrank(q_mu, q_sqrt, u):
    """Variational._sample, 'fullrank' (variationals.py:144-146):
    mu + tril(q_sqrt) @ u, batched over leading axes.
    q_sqrt [..., n, n]; q_mu, u [..., n]."""
    L = torch.tril(q_sqrt)
    return q_mu + torch.matmul(L, u.unsqueeze(-1)).squeeze(-1)


def logdet_diag(q_sqrt):
    """Variational.logdet 'diagonal' (variationals.py:183-184)."""
    return 2.0 * q_sqrt


def logdet_fullrank(q_sqrt):
    """Variational.logdet 'fullrank' (variationals.py:185-186)."""
    return torch.log(torch.square(torch.diagonal(q_sqrt, dim1=-2, dim2=-1)))


def kl_normal(logdet, u, z):
    """Normal._KL (variationals.py:225-230): one-sample MC KL,
    -0.5*sum(logdet + u^2 - z^2).  logdet broadcasts against u."""
    return -0.5 * torch.sum(logdet + torch.square(u) - torch.square(z))


def kl_generic(logdet, u, z, prior_logp=None, log_jacobian=None):
    """Variational._KL (variationals.py:198-209) with an arbitrary prior and
    transform.  prior_logp / log_jacobian are already-reduced scalars."""
    kl = -0.5 * torch.sum(LOG2PI + logdet + torch.square(u))
    if prior_logp is not None:
        kl = kl - prior_logp
        kl = kl - log_jacobian
    return kl


def prior_normal_logp(x):
    """priors.Normal.logp (priors.py:48-49)."""
    return -0.5 * torch.sum(LOG2PI + torch.square(x))


# --------------------------------------------------------------------------
# densities  (Henbun/densities.py)
# --------------------------------------------------------------------------

def gaussian(x, mu, var):
    """densities.gaussian (densities.py:25-27), elementwise, broadcasting."""
    var = _t(var, x.dtype) if not isinstance(var, torch.Tensor) else var
    return -0.5 * LOG2PI - 0.5 * torch.log(var) - 0.5 * torch.square(mu - x) / var


def student_t(x, mean, scale, deg_free):
    """densities.student_t (densities.py:52-59)."""
    deg_free = _t(deg_free, x.dtype)
    scale = _t(scale, x.dtype)
    const = (torch.lgamma((deg_free + 1.0) * 0.5) - torch.lgamma(deg_free * 0.5)
             - 0.5 * (torch.log(torch.square(scale)) + torch.log(deg_free) + math.log(math.pi)))
    return const - 0.5 * (deg_free + 1.0) * torch.log(
        1.0 + (1.0 / deg_free) * torch.square((x - mean) / scale))


def lognormal(x, mu, var):
    """densities.lognormal (densities.py:30-32)."""
    lnx = torch.log(x)
    return gaussian(lnx, mu, var) - lnx


def bernoulli(p, y):
    """densities.bernoulli (densities.py:35-36; tf.select = where)."""
    return torch.log(torch.where(y == 1, p, 1 - p))


def poisson(lamb, y):
    """densities.poisson (densities.py:39-40)."""
    return y * torch.log(lamb) - lamb - torch.lgamma(y + 1.0)


def exponential(lamb, y):
    """densities.exponential (densities.py:43-44)."""
    return -y / lamb - torch.log(lamb)


def gamma(shape, scale, x):
    """densities.gamma (densities.py:47-49)."""
    return -shape * torch.log(scale) - torch.lgamma(shape) + (shape - 1.0) * torch.log(x) - x / scale


def beta(alpha, beta_, y):
    """densities.beta (densities.py:62-68), y clipped to [1e-6, 1-1e-6]."""
    y = torch.clamp(y, 1e-6, 1 - 1e-6)
    return ((alpha - 1.0) * torch.log(y) + (beta_ - 1.0) * torch.log(1.0 - y)
            + torch.lgamma(alpha + beta_) - torch.lgamma(alpha) - torch.lgamma(beta_))


def laplace(mu, sigma, y):
    """densities.laplace (densities.py:71-72)."""
    return -torch.abs(mu - y) / sigma - torch.log(2.0 * sigma)


def bimixture(fraction, logp0, logp1):
    """densities.bimixture (densities.py:94-103) with tf_wraps.log_sum_exp (tf_wraps.py:42-48)."""
    t = torch.stack([logp0 + torch.log(fraction), logp1 + torch.log(1.0 - fraction)], dim=-1)
    m = torch.amax(t, dim=-1, keepdim=True)
    return m.squeeze(-1) + torch.log(torch.sum(torch.exp(t - m), dim=-1))


DENSITIES = {"gaussian": gaussian, "lognormal": lognormal, "bernoulli": bernoulli, "poisson": poisson,
             "exponential": exponential, "gamma": gamma, "student_t": student_t, "beta": beta, "laplace": laplace,
             "bimixture": bimixture}


def multivariate_normal(x, mu, L):
    """densities.multivariate_normal (densities.py:75-91)."""
    d = x - mu
    if d.dim() == 1:
        d = d[:, None]
    alpha = torch.linalg.solve_triangular(L, d, upper=False)
    num_col = d.shape[1]
    num_dims = d.shape[0]
    ret = -0.5 * num_dims * num_col * LOG2PI
    ret = ret - num_col * torch.sum(torch.log(torch.diagonal(L)))
    ret = ret - 0.5 * torch.sum(torch.square(alpha))
    return ret


# --------------------------------------------------------------------------
# GP kernels (Henbun/gp/kernels.py)
# --------------------------------------------------------------------------

def square_dist(X, lengthscales, X2=None):
    """UnitStationary.square_dist (gp/kernels.py:54-84).  Follows the reference
    formula -2 Xe Xe'^T + |Xe|^2 + |Xe'|^2 (no clamping at zero)."""
    Xe = X / lengthscales
    Xs = torch.sum(torch.square(Xe), -1)
    if X2 is None:
        return -2.0 * torch.matmul(Xe, Xe.transpose(-1, -2)) + Xs.unsqueeze(-1) + Xs.unsqueeze(-2)
    X2e = X2 / lengthscales
    X2s = torch.sum(torch.square(X2e), -1)
    return -2.0 * torch.matmul(Xe, X2e.transpose(-1, -2)) + Xs.unsqueeze(-1) + X2s.unsqueeze(-2)


def rbf_K(X, lengthscales, X2=None):
    """UnitRBF.K (gp/kernels.py:110-111)."""
    return torch.exp(-square_dist(X, lengthscales, X2) / 2.0)


def rbf_K_direct(X, lengthscales, X2=None):
    """Same function as rbf_K with r^2 summed from squared differences instead of the reference's
    -2 x.x' + |x|^2 + |x'|^2 expansion (gp/kernels.py:71-84).  Identical in exact arithmetic; in fp32 the
    expansion loses ~|x|^2 * 2^-24 absolute per entry, which at N=2000 on the 1-D notebook grid with l=0.2
    already breaks positive-definiteness at jitter 3e-4.  Only used as the fp32 comparator of the config-2
    full-size test (the CUDA kernel sums differences as well, csrc/gram.cu)."""
    Xe = X / lengthscales
    X2e = Xe if X2 is None else X2 / lengthscales
    d = Xe.unsqueeze(-2) - X2e.unsqueeze(-3)
    return torch.exp(-0.5 * torch.sum(torch.square(d), -1))


def csym_rbf_K(X, lengthscales, X2=None):
    """UnitCsymRBF.K (gp/kernels.py:122-126)."""
    if X2 is None:
        X2 = X
    return (torch.exp(-square_dist(X, lengthscales, X2) / 2.0)
            + torch.exp(-square_dist(X, lengthscales, -X2) / 2.0))


def Kdiag(X):
    """UnitStationary.Kdiag (gp/kernels.py:90-91)."""
    return torch.ones(X.shape[:-1], dtype=X.dtype, device=X.device)


def kern_cholesky(X, lengthscales, jitter=1e-5, K_fn=rbf_K):
    """UnitStationary.Cholesky (gp/kernels.py:93-101) + tf_wraps.eye (tf_wraps.py:26-30)."""
    n = X.shape[-2]
    K = K_fn(X, lengthscales)
    return torch.linalg.cholesky(K + jitter * torch.eye(n, dtype=X.dtype, device=X.device))


def gp_samples(L, u):
    """GP.samples (gp/gp.py:37-50): u [N,n] @ L^T -> [N,n]."""
    return torch.matmul(u, L.transpose(-1, -2))


# --------------------------------------------------------------------------
# SparseGP (gp/gp.py:53-192)  -- "next" row, restated for completeness
# --------------------------------------------------------------------------

def sparse_effective_LT(x, z, lengthscales, jitter=1e-5):
    """SparseGP._effective_LT (gp/gp.py:146-174): Lm^{-1} K(z,x)."""
    Lm = kern_cholesky(z, lengthscales, jitter)
    if x.dim() == 2:
        return torch.linalg.solve_triangular(Lm, rbf_K(z, lengthscales, x), upper=False)
    N = x.shape[0]
    Lminv = torch.linalg.solve_triangular(Lm, torch.eye(z.shape[0], dtype=x.dtype), upper=False)
    zt = z.unsqueeze(0).expand(N, -1, -1)
    return torch.matmul(Lminv.unsqueeze(0), rbf_K(zt, lengthscales, x))


def sparse_additional_cov(x, LnT, lengthscales, q_shape="diagonal"):
    """SparseGP._additional_cov (gp/gp.py:177-192)."""
    if q_shape == "diagonal":
        return Kdiag(x) - torch.sum(torch.square(LnT), -2)
    return rbf_K(x, lengthscales) - torch.matmul(LnT.transpose(-1, -2), LnT)


# --------------------------------------------------------------------------
# Neural net (Henbun/nn.py)
# --------------------------------------------------------------------------

_ACT = {
    "sigmoid": torch.sigmoid,
    "relu": torch.relu,
    "tanh": torch.tanh,
    "identity": lambda x: x,
}


def matbias(x, w, b, clip_enabled=False):
    """MatBias.__call__ (nn.py:31-32): clip(x@w + b); w [*n_layers,in,out], b [*n_layers,1,out]."""
    return clip(torch.matmul(x, w) + b, clip_enabled)


def neural_net(x, ws: Sequence, bs: Sequence, acts: Sequence[str], clip_enabled=False):
    """NeuralNet.__call__ (nn.py:73-84): act_i(matbias_i(y)) for all but the last
    layer, last layer linear."""
    y = x
    for i in range(len(ws) - 1):
        y = _ACT[acts[i]](matbias(y, ws[i], bs[i], clip_enabled))
    return matbias(y, ws[-1], bs[-1], clip_enabled)


def local_feed_split(x, sizes: Sequence[int]):
    """Parameterized.feed (param.py:516-537): split the last axis of the fed
    tensor into the children's feed sizes in name-sorted order (for a diagonal
    variational: q_mu then q_sqrt)."""
    out, begin = [], 0
    for s in sizes:
        out.append(x[..., begin:begin + s])
        begin += s
    return out


# --------------------------------------------------------------------------
# TF-1 Adam  (model.py:206,220 -> tf.train.AdamOptimizer defaults)
# --------------------------------------------------------------------------

def adam_tf1_step(theta, grad_of_loss, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """One tf.train.AdamOptimizer step on loss = -objective.  t is the 1-based
    step count.  Returns (theta, m, v) as new arrays (numpy, same dtype)."""
    lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    m = b1 * m + (1.0 - b1) * grad_of_loss
    v = b2 * v + (1.0 - b2) * grad_of_loss * grad_of_loss
    theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v


# --------------------------------------------------------------------------
# Model-level ELBOs (the five BASELINE configs' graphs)
# --------------------------------------------------------------------------

def gpr_elbo(p: Dict[str, torch.Tensor], X, Y, U, q_shape="diagonal", jitter=1e-5,
             return_parts=False):
    """S-sample mean of the reference one-sample ELBO of
    notebooks/GaussianProcess.ipynb:109-148 (``ELBO_gaussian``):

        y_fit = matmul(kern.Cholesky(X), q) * sqrt(k_var)
        ELBO  = reduce_sum(gaussian(Y, y_fit, var)) - KL()

    with q = variationals.Gaussian(shape=X.shape[:1]+[1], q_shape) (scale*Normal,
    variationals.py:290-291; KL of the unscaled Normal :225-230).

    p: free-space parameters: 'q_mu' [n], 'q_sqrt' [n] (log-std) or [n,n],
       'scale' [1] (free), 'lengthscales' [1 or D] (free), 'k_var' [1], 'var' [1].
    U: [S, n] standard-normal draws (one row per MC sample / per session.run).
    Y: [n] (the notebook's [n,1] column, flattened)."""
    ell = log1pe_forward(p["lengthscales"])
    kvar = log1pe_forward(p["k_var"])
    var = log1pe_forward(p["var"])
    scale = log1pe_forward(p["scale"])
    L = kern_cholesky(X, ell, jitter)
    if q_shape == "diagonal":
        Z = sample_diag(p["q_mu"], p["q_sqrt"], U)             # [S,n]
        logdet = logdet_diag(p["q_sqrt"])
    else:
        Z = p["q_mu"] + torch.matmul(U, torch.tril(p["q_sqrt"]).T)
        logdet = logdet_fullrank(p["q_sqrt"])
    W = scale * Z
    F = torch.matmul(W, L.T) * torch.sqrt(kvar)                # [S,n]
    ll = torch.sum(gaussian(Y, F, var))
    kl = kl_normal(logdet, U, Z)
    S = U.shape[0]
    elbo = (ll - kl) / S
    if return_parts:
        return elbo, dict(L=L, Z=Z, F=F, ll=ll, kl=kl)
    return elbo


def gpr_elbo_closed_form_grads(p: Dict[str, np.ndarray], X, Y, U, jitter=1e-5):
    """Hand-derived backward of ``gpr_elbo`` (mean-field q), numpy fp64.  This is
    the blueprint the CUDA backward follows; the tests require it
    to agree with torch autograd of ``gpr_elbo`` to 1e-9."""
    f8 = np.float64
    X = np.asarray(X, f8); Y = np.asarray(Y, f8); U = np.asarray(U, f8)
    sp = lambda x: np.logaddexp(0.0, x) + 1e-6
    sg = lambda x: 1.0 / (1.0 + np.exp(-x))
    ell = sp(np.asarray(p["lengthscales"], f8)); kvar = sp(np.asarray(p["k_var"], f8))[0]
    var = sp(np.asarray(p["var"], f8))[0]; scale = sp(np.asarray(p["scale"], f8))[0]
    mu = np.asarray(p["q_mu"], f8); om = np.asarray(p["q_sqrt"], f8)
    S, n = U.shape
    Xe = X / ell
    d2 = ((Xe[:, None, :] - Xe[None, :, :]) ** 2).sum(-1)
    K = np.exp(-0.5 * d2)
    L = np.linalg.cholesky(K + jitter * np.eye(n))
    Z = mu + np.exp(om) * U
    W = scale * Z
    F = np.sqrt(kvar) * W @ L.T
    E = F - Y
    ll = np.sum(-0.5 * LOG2PI - 0.5 * np.log(var) - 0.5 * E ** 2 / var)
    kl = -0.5 * np.sum(2.0 * om + U ** 2 - Z ** 2)
    elbo = (ll - kl) / S
    R = -E / (var * S)
    g_var = np.sum(-0.5 / var + 0.5 * E ** 2 / var ** 2) / S
    g_kvar = np.sum(R * F) / (2.0 * kvar)
    Wb = np.sqrt(kvar) * R @ L
    Lb = np.sqrt(kvar) * np.tril(R.T @ W)
    g_scale = np.sum(Wb * Z)
    Zb = scale * Wb - Z / S
    g_mu = Zb.sum(0)
    g_om = (Zb * np.exp(om) * U).sum(0) + 1.0
    G = chol_rev_recursive(L, Lb, nb=max(1, n // 4))
    Dm = (X[:, None, :] - X[None, :, :]) ** 2                       # [n,n,D]
    if ell.shape[0] == 1:
        g_ell = np.array([np.sum(G * K * Dm.sum(-1)) / ell[0] ** 3])
    else:
        g_ell = np.einsum("ij,ij,ijd->d", G, K, Dm) / ell ** 3
    grads = {
        "q_mu": g_mu, "q_sqrt": g_om,
        "scale": np.array([g_scale * sg(np.asarray(p["scale"], f8))[0]]),
        "lengthscales": g_ell * sg(np.asarray(p["lengthscales"], f8)),
        "k_var": np.array([g_kvar * sg(np.asarray(p["k_var"], f8))[0]]),
        "var": np.array([g_var * sg(np.asarray(p["var"], f8))[0]]),
    }
    return elbo, grads


# --------------------------------------------------------------------------
# Reverse-mode Cholesky -- the algorithm the CUDA host recursion follows.
# TF's _CholeskyGrad computes the same quantity (gradient w.r.t. the full
# symmetric input, symmetrised).
# --------------------------------------------------------------------------

def chol_rev_base(L, Lbar):
    """G = 0.5*(S+S^T), S = L^{-T} Phi(L^T tril(Lbar)) L^{-1}; Phi = tril with
    halved diagonal."""
    n = L.shape[0]
    P = np.tril(L.T @ np.tril(Lbar))
    P[np.diag_indices(n)] *= 0.5
    Linv = np.linalg.inv(L)
    S = Linv.T @ P @ Linv
    return 0.5 * (S + S.T)


def chol_rev_recursive(L, Lbar, nb=32):
    """Recursive blocked reverse-mode Cholesky (all level-3).  Input: lower
    factor L and dELBO/dL (lower part used).  Output: full symmetric G with
    d ELBO = sum_ij G_ij dK_ij for symmetric dK.

        G22 = rev(L22, Lb22)
        T   = (Lb21 - 2 G22 L21) L11^{-1};  G21 = T/2
        G11 = rev(L11, Lb11 - tril(T^T L21))
    """
    n = L.shape[0]
    if n <= nb:
        return chol_rev_base(L, Lbar)
    n1 = (n // 2)
    L11, L21, L22 = L[:n1, :n1], L[n1:, :n1], L[n1:, n1:]
    G = np.zeros_like(L)
    G22 = chol_rev_recursive(L22, Lbar[n1:, n1:], nb)
    T = (Lbar[n1:, :n1] - 2.0 * G22 @ L21) @ np.linalg.inv(L11)
    Lb11 = Lbar[:n1, :n1] - np.tril(T.T @ L21)
    G11 = chol_rev_recursive(L11, Lb11, nb)
    G[:n1, :n1] = G11
    G[n1:, n1:] = G22
    G[n1:, :n1] = 0.5 * T
    G[:n1, n1:] = 0.5 * T.T
    return G


# --------------------------------------------------------------------------
# Other BASELINE configs
# --------------------------------------------------------------------------

def expert_gpr_elbo(p, X, Y, U3, q_shapes=("fullrank", "fullrank", "fullrank"), jitter=3e-4, K_fn=rbf_K):
    """notebooks/Expert_GPR.ipynb:101-149 (``ELBO``), S-sample mean.
    p: 'q_{s,l,r}.q_mu' [n], 'q_{s,l,r}.q_sqrt', 'q_{s,l,r}.scale' [1],
       'kern_{s,l,r}.lengthscales' [1], 'k_var','k_var_r','var' [1] (all free).
    U3: dict name -> [S,n]."""
    S = U3["s"].shape[0]
    fs, kl = {}, 0.0
    for name, qs in zip(("s", "l", "r"), q_shapes):
        ell = log1pe_forward(p[f"kern_{name}.lengthscales"])
        L = kern_cholesky(X, ell, jitter, K_fn)
        mu, sq = p[f"q_{name}.q_mu"], p[f"q_{name}.q_sqrt"]
        U = U3[name]
        if qs == "diagonal":
            Z = sample_diag(mu, sq, U); ld = logdet_diag(sq)
        else:
            Z = mu + torch.matmul(U, torch.tril(sq).T); ld = logdet_fullrank(sq)
        kl = kl + kl_normal(ld, U, Z)
        Wq = log1pe_forward(p[f"q_{name}.scale"]) * Z
        fs[name] = torch.matmul(Wq, L.T)
    f_r = fs["r"] * torch.sqrt(log1pe_forward(p["k_var_r"]))
    frac = torch.sigmoid(f_r)
    f = (frac * fs["s"] + (1 - frac) * fs["l"]) * log1pe_forward(p["k_var"])
    ll = torch.sum(gaussian(Y, f, log1pe_forward(p["var"])))
    return (ll - kl) / S


def amortised_elbo(p, Xmb, U, enc_acts, dec_acts=None, n_total=None):
    """Amortised local-variable model (BASELINE config 4): encoder
    nn.NeuralNet (nn.py:34-87) feeds a LOCAL Normal([latent]) (param.py:386-392,
    516-537: q_mu | q_sqrt halves); z = mu + exp(omega)*u (variationals.py:140),
    KL variationals.py:225-230; likelihood densities.gaussian(x, dec(z), var).

    p: 'enc.w{i}','enc.b{i}', optional 'dec.w{i}','dec.b{i}', 'var' [1] (free).
    Xmb [B, Din]; U [S, B, latent].  If there is no decoder the "lite" model
    gaussian(x[:, :latent], z, var) is used.
    The sum is scaled by n_total/B if n_total is given (user-side scaling,
    SURVEY.md section 9)."""
    n_enc = len([k for k in p if k.startswith("enc.w")])
    h = neural_net(Xmb, [p[f"enc.w{i}"] for i in range(n_enc)],
                   [p[f"enc.b{i}"] for i in range(n_enc)], enc_acts)
    latent = U.shape[-1]
    mu, om = local_feed_split(h, [latent, latent])
    Z = sample_diag(mu, om, U)                                  # [S,B,latent]
    kl = kl_normal(logdet_diag(om), U, Z)
    var = log1pe_forward(p["var"])
    n_dec = len([k for k in p if k.startswith("dec.w")])
    if n_dec:
        xr = neural_net(Z, [p[f"dec.w{i}"] for i in range(n_dec)],
                        [p[f"dec.b{i}"] for i in range(n_dec)], dec_acts)
        ll = torch.sum(gaussian(Xmb, xr, var))
    else:
        ll = torch.sum(gaussian(Xmb[:, :latent], Z, var))
    S = U.shape[0]
    sc = 1.0 if n_total is None else float(n_total) / Xmb.shape[0]
    return sc * (ll - kl) / S


def linear_operator_elbo(p, A, y, U):
    """BASELINE config 5: full-covariance Normal([n]) latent through a dense
    forward operator: gaussian(y, A z, var) - KL  (same building blocks:
    variationals.py:144-146,185-186,225-230; densities.py:25-27).
    p: 'q_mu' [n], 'q_sqrt' [n,n], 'var' [1] (free).  U [S,n]."""
    Z = p["q_mu"] + torch.matmul(U, torch.tril(p["q_sqrt"]).T)  # [S,n]
    F = torch.matmul(Z, A.T)                                    # [S,M]
    ll = torch.sum(gaussian(y, F, log1pe_forward(p["var"])))
    kl = kl_normal(logdet_fullrank(p["q_sqrt"]), U, Z)
    return (ll - kl) / U.shape[0]


def value_and_grads(fn, p: Dict[str, np.ndarray], *args, dtype=torch.float64, device="cpu", **kw):
    """Evaluate fn(p, ...) and d fn / d p by torch autograd (fp64).  device="cuda" runs the SAME restatement through
    torch's fp64 library kernels on the GPU -- used by the parity tests at the named sizes (N >= 16384), where the CPU
    would need minutes; it is still only the checker."""
    tp = {k: torch.tensor(np.asarray(v), dtype=dtype, device=device, requires_grad=True) for k, v in p.items()}

    def conv(a):
        if isinstance(a, (np.ndarray, torch.Tensor)):
            return _t(a, dtype).to(device)
        if isinstance(a, dict):
            return {k: _t(v, dtype).to(device) for k, v in a.items()}
        return a
    targs = [conv(a) for a in args]
    val = fn(tp, *targs, **kw)
    val.backward()
    grads = {k: (v.grad.detach().cpu().numpy().copy() if v.grad is not None else
                 np.zeros_like(np.asarray(p[k], dtype=np.float64))) for k, v in tp.items()}
    return float(val.detach()), grads
