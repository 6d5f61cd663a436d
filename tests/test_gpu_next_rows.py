"""'Next' rows of SURVEY.md 8(f) on the GPU: SparseGP with trainable inducing points (Henbun/gp/gp.py:53-192) --
gradients through K(x, z) and chol(K(z, z)) w.r.t. the kernel inputs.  Oracle: oracle/henbun_oracle.py in torch fp64
with autograd.  Tolerances: values 1e-5 relative, gradients 1e-4 normwise (fp32 kernels, small well-conditioned cases)."""
import numpy as np
import pytest
import torch

from oracle import henbun_oracle as O

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("n,m,D,n_ell", [(37, 21, 2, 1), (200, 64, 3, 3), (70, 70, 8, 1), (33, 150, 12, 12)])
def test_rbf_gram_input_gradients(n, m, D, n_ell):
    from henbun_b200 import ops
    rng = np.random.RandomState(n + m)
    X = rng.randn(n, D); Z = rng.randn(m, D); ell = np.exp(0.3 * rng.randn(n_ell)); W = rng.randn(n, m)
    tX, tZ, tl = (torch.tensor(a, requires_grad=True) for a in (X, Z, ell))
    (O.rbf_K(tX, tl, tZ) * torch.tensor(W)).sum().backward()
    dX, dZ, dl = (torch.tensor(a, dtype=torch.float32, device="cuda", requires_grad=True) for a in (X, Z, ell))
    K = ops.rbf_K(dX, dZ, dl)
    (K * torch.tensor(W, dtype=torch.float32, device="cuda")).sum().backward()
    assert rel_err(dX.grad.cpu(), tX.grad) < 1e-4
    assert rel_err(dZ.grad.cpu(), tZ.grad) < 1e-4
    assert rel_err(dl.grad.cpu(), tl.grad) < 1e-4
    # K(X, X): both arguments are the same tensor
    W2 = rng.randn(n, n)
    tX2 = torch.tensor(X, requires_grad=True)
    (O.rbf_K(tX2, torch.tensor(ell)) * torch.tensor(W2)).sum().backward()
    dX2 = torch.tensor(X, dtype=torch.float32, device="cuda", requires_grad=True)
    (ops.rbf_K(dX2, None, torch.tensor(ell, dtype=torch.float32, device="cuda")) *
     torch.tensor(W2, dtype=torch.float32, device="cuda")).sum().backward()
    assert rel_err(dX2.grad.cpu(), tX2.grad) < 1e-4


@pytest.mark.parametrize("m,D", [(40, 2), (300, 4)])
def test_kernel_cholesky_input_gradient(m, D):
    from henbun_b200 import ops
    rng = np.random.RandomState(m)
    Z = rng.randn(m, D); ell = np.array([0.7]); W = np.tril(rng.randn(m, m))
    tZ = torch.tensor(Z, requires_grad=True)
    (O.kern_cholesky(tZ, torch.tensor(ell), 1e-3) * torch.tensor(W)).sum().backward()
    dZ = torch.tensor(Z, dtype=torch.float32, device="cuda", requires_grad=True)
    L = ops.kern_cholesky(dZ, torch.tensor(ell, dtype=torch.float32, device="cuda"), 1e-3)
    (L * torch.tensor(W, dtype=torch.float32, device="cuda")).sum().backward()
    assert rel_err(dZ.grad.cpu(), tZ.grad) < 2e-4


def test_sparse_gp_trains_inducing_points():
    """SparseGP.samples (q_shape='neglected' and 'diagonal' mean part) and its gradient w.r.t. z, u and the lengthscale."""
    import henbun_b200 as hb
    rng = np.random.RandomState(3)
    n, m, D, N = 150, 24, 2, 5
    x = rng.randn(n, D); z = rng.randn(m, D); u = rng.randn(N, m); w = rng.randn(N, n)
    model = hb.model.Model()
    model.gp = hb.gp.SparseGP(hb.gp.kernels.UnitRBF(np.ones(1) * 0.8), z)
    model.initialize()
    xd = torch.tensor(x, dtype=torch.float32, device="cuda")
    ud = torch.tensor(u, dtype=torch.float32, device="cuda", requires_grad=True)
    with model.tf_mode():
        s = model.gp.samples(xd, ud, q_shape='neglected')
    (s * torch.tensor(w, dtype=torch.float32, device="cuda")).sum().backward()
    # oracle
    tz = torch.tensor(z, requires_grad=True); tu = torch.tensor(u, requires_grad=True)
    ell = torch.tensor([0.8], dtype=torch.float64)
    LnT = O.sparse_effective_LT(torch.tensor(x), tz, ell, hb.settings.get_settings().numerics.jitter_level)
    so = tu @ LnT
    (so * torch.tensor(w)).sum().backward()
    assert rel_err(s.detach().cpu(), so.detach()) < 2e-5
    assert rel_err(ud.grad.cpu(), tu.grad) < 1e-4
    free = [p for p in model.get_tf_variables() if p.grad is not None and tuple(p.shape) == tuple(z.shape)]
    assert free, "no gradient reached the inducing points"
    assert rel_err(free[0].grad.cpu(), tz.grad) < 2e-4


@pytest.mark.parametrize("n,m,D", [(37, 21, 2), (64, 64, 1), (50, 90, 5)])
def test_csym_rbf_input_gradients(n, m, D):
    """UnitCsymRBF (gp/kernels.py:113-131): K(x, x2) + K(x, -x2).  Gradients w.r.t. both arguments, w.r.t. X when both
    arguments are X, and through the kernel Cholesky -- all vs the oracle's csym_rbf_K under autograd."""
    from henbun_b200 import ops
    rng = np.random.RandomState(7 * n + m)
    X = 0.6 * rng.randn(n, D); Z = 0.6 * rng.randn(m, D); ell = np.array([0.9]); W = rng.randn(n, m)
    tX, tZ = (torch.tensor(a, requires_grad=True) for a in (X, Z))
    (O.csym_rbf_K(tX, torch.tensor(ell), tZ) * torch.tensor(W)).sum().backward()
    c = lambda a, g=False: torch.tensor(a, dtype=torch.float32, device="cuda", requires_grad=g)
    dX, dZ = c(X, True), c(Z, True)
    (ops.rbf_K(dX, dZ, c(ell), True) * c(W)).sum().backward()
    assert rel_err(dX.grad.cpu(), tX.grad) < 1e-4
    assert rel_err(dZ.grad.cpu(), tZ.grad) < 1e-4
    W2 = rng.randn(n, n)
    tX2 = torch.tensor(X, requires_grad=True)
    (O.csym_rbf_K(tX2, torch.tensor(ell)) * torch.tensor(W2)).sum().backward()
    dX2 = c(X, True)
    (ops.rbf_K(dX2, None, c(ell), True) * c(W2)).sum().backward()
    assert rel_err(dX2.grad.cpu(), tX2.grad) < 1e-4
    # through the fused kernel-Cholesky (K-bar symmetric, lower storage)
    Wl = np.tril(rng.randn(n, n))
    tX3 = torch.tensor(X, requires_grad=True)
    (O.kern_cholesky(tX3, torch.tensor(ell), 1e-2, K_fn=O.csym_rbf_K) * torch.tensor(Wl)).sum().backward()
    dX3 = c(X, True)
    (ops.kern_cholesky(dX3, c(ell), 1e-2, True) * c(Wl)).sum().backward()
    assert rel_err(dX3.grad.cpu(), tX3.grad) < 5e-4
