"""transforms.py as kernels (csrc/transforms.cu): forward, log-Jacobian and both backward passes against the vectors the
unmodified reference produced (tests/golden/densities_transforms.npz), the fp64 oracle (with autograd gradients), and
through a transformed variational's generic one-sample KL (variationals.py:198-209).
fp32 vs fp64: 2e-6 relative per element on the forward, 1e-5 norm-wise on gradients, 1e-5 relative on the summed
log-Jacobian."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import henbun_b200 as hb
from henbun_b200 import _lib
from oracle import henbun_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def dev(a, grad=False):
    return torch.tensor(np.asarray(a, np.float32), device="cuda", requires_grad=grad)


def test_transforms_vs_reference():
    g = np.load(os.path.join(G, "densities_transforms.npz"))
    xs = dev(g["xs"])
    t = hb.transforms.Log1pe()
    assert np.allclose(t.tf_forward(xs).cpu().numpy(), g["log1pe_fwd"], rtol=2e-6, atol=1e-7)
    assert abs(float(t.tf_log_jacobian(xs)) - float(g["log1pe_logjac"])) <= 1e-5 * abs(float(g["log1pe_logjac"]))
    lg = hb.transforms.Logistic(7.3, 19.4)
    assert np.allclose(lg.tf_forward(xs).cpu().numpy(), g["logistic_fwd"], rtol=2e-6)
    assert abs(float(lg.tf_log_jacobian(xs)) - float(g["logistic_logjac"])) <= 1e-5 * abs(float(g["logistic_logjac"]))
    ex = hb.transforms.Exp()
    assert np.allclose(ex.tf_forward(xs).cpu().numpy(), g["exp_fwd"], rtol=2e-6)


CASES = [("exp", hb.transforms.Exp(), O.exp_forward, lambda x: O.exp_log_jacobian(x)),
         ("log1pe", hb.transforms.Log1pe(), O.log1pe_forward, lambda x: O.log1pe_log_jacobian(x)),
         ("logistic", hb.transforms.Logistic(-1.5, 4.0), lambda x: O.logistic_forward(x, -1.5, 4.0),
          lambda x: O.logistic_log_jacobian(x, -1.5, 4.0))]


@pytest.mark.parametrize("name,tr,fwd64,lj64", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("shape", [(1,), (10,), (3, 4097), (64, 65536)])
def test_forward_logjac_and_gradients_vs_oracle(name, tr, fwd64, lj64, shape):
    rng = np.random.RandomState(11)
    xv = (2.5 * rng.randn(*shape)).astype(np.float32)
    w = rng.randn(*shape)
    x = dev(xv, grad=True)
    y = tr.tf_forward(x)
    lj = tr.tf_log_jacobian(x)
    ((y * dev(w)).sum() + 0.7 * lj).backward()
    x64 = torch.tensor(xv, dtype=torch.float64, requires_grad=True)
    y64 = fwd64(x64); lj64v = lj64(x64)
    ((y64 * torch.tensor(w)).sum() + 0.7 * lj64v).backward()
    assert y.shape == tuple(shape)
    assert np.allclose(y.detach().cpu().numpy(), y64.detach().numpy(), rtol=2e-6, atol=1e-6)
    assert abs(float(lj) - float(lj64v)) <= 1e-5 * max(abs(float(lj64v)), 1.0)
    assert rel_err(x.grad.cpu().numpy(), x64.grad.numpy()) < 1e-5


def test_extreme_arguments_do_not_overflow():
    x = dev([-100.0, -30.0, 0.0, 30.0, 80.0], grad=True)
    t = hb.transforms.Log1pe()
    y = t.tf_forward(x); lj = t.tf_log_jacobian(x)
    (y.sum() + lj).backward()
    assert torch.isfinite(y).all() and torch.isfinite(lj) and torch.isfinite(x.grad).all()
    assert np.allclose(y.detach().cpu().numpy(), [1e-6, 1e-6 + np.exp(-30.0), np.log(2.0) + 1e-6, 30.0, 80.0], rtol=1e-6)


def test_cabi_errors():
    lib = _lib.load()
    x = torch.zeros(8, device="cuda"); y = torch.empty(8, device="cuda")
    st = _lib.stream()
    assert lib.hb_transform_fwd(0, _lib.ptr(x), 8, 0.0, 1.0, _lib.ptr(y), st) == _lib.HB_ERR_ARG      # identity has no kernel
    assert lib.hb_transform_fwd(4, _lib.ptr(x), 8, 0.0, 1.0, _lib.ptr(y), st) == _lib.HB_ERR_ARG
    assert lib.hb_transform_fwd(3, _lib.ptr(x), 8, 2.0, 1.0, _lib.ptr(y), st) == _lib.HB_ERR_ARG      # Logistic needs b > a
    assert lib.hb_transform_fwd(2, _lib.ptr(x), 0, 0.0, 1.0, None, st) == 0                            # empty
    assert lib.hb_transform_logjac(2, _lib.ptr(x), 8, 0.0, 1.0, _lib.ptr(y), None, 0, st) == _lib.HB_ERR_WORKSPACE


def test_transformed_variational_generic_kl():
    """Variational(transform=Exp, prior=Gamma): z -> exp(z) + 1e-6, KL = -entropy - prior.logp(T(z)) - log|J|
    (variationals.py:198-209), S samples, against the oracle's kl_generic; gradients w.r.t. q_mu / q_sqrt too."""
    rng = np.random.RandomState(5)
    n, S = 300, 6
    m = hb.model.Model()
    m.q = hb.variationals.Variational([n], q_shape='diagonal', prior=hb.priors.Gamma(2.0, 1.5), transform=hb.transforms.Exp())
    m.q.q_mu = 0.3 * rng.randn(n)
    m.q.q_sqrt = 0.2 * rng.randn(n) - 1.0
    m.initialize()
    q = object.__getattribute__(m, "q")
    U = rng.randn(S, n).astype(np.float32)
    m._begin_run(S, {q: U})
    with m.tf_mode():
        sample = m.q
        kl = m.KL()
    kl.backward()
    mu = torch.tensor(q.q_mu._free_numpy().astype(np.float64), requires_grad=True)
    om = torch.tensor(q.q_sqrt._free_numpy().astype(np.float64), requires_grad=True)
    U64 = torch.tensor(U, dtype=torch.float64)
    Z = O.sample_diag(mu, om, U64)
    T = O.exp_forward(Z)
    prior = torch.sum(O.gamma(torch.tensor(2.0, dtype=torch.float64), torch.tensor(1.5, dtype=torch.float64), T))
    ref = O.kl_generic(O.logdet_diag(om), U64, Z, prior, O.exp_log_jacobian(Z))
    ref.backward()
    assert np.allclose(sample.detach().cpu().numpy(), T.detach().numpy().reshape(sample.shape), rtol=3e-6)
    assert abs(float(kl) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_err(q.q_mu._tensor.grad.cpu().numpy().ravel(), mu.grad.numpy()) < 1e-5
    assert rel_err(q.q_sqrt._tensor.grad.cpu().numpy().ravel(), om.grad.numpy()) < 1e-5
