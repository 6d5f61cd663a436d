"""Ports of the reference's remaining hot-path tests onto the CUDA path:
* testing/test_gp.py (GP.samples, SparseGP._effective_LT / _additional_cov / samples in all three q_shape modes, batched
  and non-batched x, "works with very small jitter"),
* testing/test_densities.py (bimixture with a broadcast fraction, student_t with scalar / tensor parameters),
* testing/test_tf_wraps.py (clip_by_value on NeuralNet outputs and variational samples),
each with the reference's own tolerance, plus oracle comparisons (fp64) where the reference only checks shapes."""
import numpy as np
import pytest
import torch
from scipy.special import loggamma

import henbun_b200 as hb
import henbun_b200.tf as tf
from oracle import henbun_oracle as O

pytestmark = pytest.mark.gpu
f32 = np.float32


def dev(a):
    return torch.tensor(np.asarray(a, f32), device="cuda")


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


# ------------------------------------------------------------------------------------------ test_gp.py
def test_sparse_gp_works_with_very_small_jitter():          # test_gp.py:10-29
    rng = np.random.RandomState(0)
    m = hb.model.Model()
    m.sparse_gp = hb.gp.SparseGP(z=np.random.RandomState(1).randn(600, 1),
                                 kern=hb.gp.kernels.UnitRBF(lengthscales=np.ones(1, f32)))
    m.u = hb.variationals.Normal(shape=[1, 600])
    x = dev(rng.randn(400, 1))
    m.initialize()
    with m.tf_mode():
        s = m.run(m.sparse_gp.samples(x, m.u, 'neglected'))
        assert s.shape == (1, 400) and not np.any(np.isnan(s))
        s = m.run(m.sparse_gp.samples(x, m.u, 'diagonal'))
        assert s.shape == (1, 400) and not np.any(np.isnan(s))
    hb.ops.check_numerics()                                    # no failed pivot was flagged


def test_gp_dense_samples_shape_gradient_and_value():      # test_gp.py:32-55
    rng = np.random.RandomState(0)
    m = hb.model.Model()
    m.gp = hb.gp.GP(kern=hb.gp.kernels.UnitRBF(lengthscales=np.ones(1, f32)))
    m.u = hb.variationals.Normal(shape=[20, 30])
    xv = rng.randn(30, 2)
    x = dev(xv)
    m.initialize()
    U = rng.randn(600).astype(f32)
    q = object.__getattribute__(m, "u")
    m._begin_run(1, {q: U})
    with m.tf_mode():
        u = m.u
        samples = m.gp.samples(x, u)
    assert tuple(samples.shape) == (20, 30)
    torch.sum(samples * samples).backward()
    grads = [p.grad for p in m.get_tf_variables() if p.grad is not None]
    assert len(grads) > 0 and all(torch.isfinite(g).all() for g in grads)
    # value: u L^T with L = chol(K + jitter I) (gp/gp.py:37-50)
    L = O.kern_cholesky(torch.tensor(xv), torch.tensor([1.0], dtype=torch.float64), hb.settings.numerics.jitter_level)
    ref = u.detach().double().cpu() @ L.T
    # cond(K + 1e-5 I) = 1.2e6 on these 30 points: LAPACK in fp32 is 6.6e-5 away from fp64
    assert rel_err(samples.detach().cpu().numpy(), ref.numpy()) < 5e-4


@pytest.fixture()
def sparse_model():                                          # test_gp.py:58-66
    m = hb.model.Model()
    m.sparse_gp = hb.gp.SparseGP(z=np.linspace(-2.0, 2.0, 60).reshape(-1, 2),
                                 kern=hb.gp.kernels.UnitRBF(lengthscales=np.ones(1, f32) * 0.5))
    m.u = hb.variationals.Normal(shape=[20, 30])
    m.initialize()
    return m


def test_effective_LT(sparse_model):                        # test_gp.py:68-91
    m = sparse_model
    zx = np.linspace(-2.0, 2.0, 60).reshape(-1, 2)
    with m.tf_mode():
        LT_eff = m.run(m.sparse_gp._effective_LT(dev(zx)))
        L = m.run(tf.transpose(m.sparse_gp.kern.Cholesky(dev(zx))))
    assert np.allclose(LT_eff, L, atol=0.005)                  # x == z: effective L^T is the Cholesky factor
    with m.tf_mode():
        LT1 = m.run(m.sparse_gp._effective_LT(dev(zx.reshape(1, -1, 2))))
        L1 = m.run(tf.transpose(m.sparse_gp.kern.Cholesky(dev(zx.reshape(1, -1, 2))), [0, 2, 1]))
    assert np.allclose(LT1, LT_eff, atol=0.005) and np.allclose(L1, L, atol=0.005)
    xb = np.array([zx for _ in range(20)])
    with m.tf_mode():
        LTb = m.run(m.sparse_gp._effective_LT(dev(xb)))
        Lb = m.run(tf.transpose(m.sparse_gp.kern.Cholesky(dev(xb)), [0, 2, 1]))
    assert LTb.shape == (20, 30, 30) and np.allclose(LTb, Lb, atol=0.005)
    # and against the fp64 oracle on a generic x
    x = np.random.RandomState(2).randn(40, 2)
    with m.tf_mode():
        got = m.run(m.sparse_gp._effective_LT(dev(x)))
    ref = O.sparse_effective_LT(torch.tensor(x), torch.tensor(zx), torch.tensor([0.5], dtype=torch.float64),
                                hb.settings.numerics.jitter_level).numpy()
    assert np.allclose(got, ref, atol=1e-2)                    # cond(Kmm + 1e-5 I) = 6e5: fp32 LAPACK is 2.5e-3 off (max abs)


def test_additional_cov(sparse_model):                      # test_gp.py:95-131 (both methods named test_additional_cov1)
    m = sparse_model
    zx = np.linspace(-2.0, 2.0, 60).reshape(-1, 2)
    for x in (zx, np.array([zx for _ in range(20)])):          # x == z: the additional covariance vanishes
        for q_shape in ('diagonal', 'fullrank'):
            with m.tf_mode():
                LT = m.sparse_gp._effective_LT(dev(x))
                cov = m.run(m.sparse_gp._additional_cov(dev(x), LT, q_shape))
            assert np.allclose(cov, 0.0, atol=0.005), q_shape
    rng = np.random.RandomState(0)
    x = rng.randn(20, 2)
    with m.tf_mode():
        LT = m.sparse_gp._effective_LT(dev(x))
        cov = m.run(m.sparse_gp._additional_cov(dev(x), LT, 'fullrank'))
        cov_diag = m.run(m.sparse_gp._additional_cov(dev(x), LT, 'diagonal'))
    assert np.allclose(np.diagonal(cov), cov_diag, atol=1e-4)
    ref = O.sparse_additional_cov(torch.tensor(x), O.sparse_effective_LT(torch.tensor(x), torch.tensor(zx), torch.tensor(
        [0.5], dtype=torch.float64), hb.settings.numerics.jitter_level), torch.tensor([0.5], dtype=torch.float64), "fullrank").numpy()
    assert np.allclose(cov, ref, atol=1e-2)
    xb = rng.randn(21, 20, 2)
    with m.tf_mode():
        LT = m.sparse_gp._effective_LT(dev(xb))
        cov = m.run(m.sparse_gp._additional_cov(dev(xb), LT, 'fullrank'))
        cov_diag = m.run(m.sparse_gp._additional_cov(dev(xb), LT, 'diagonal'))
    assert cov.shape == (21, 20, 20) and cov_diag.shape == (21, 20)
    for i in range(len(cov)):
        assert np.allclose(np.diagonal(cov[i]), cov_diag[i], atol=1e-4)


@pytest.mark.parametrize("batched", [False, True])
def test_sparse_samples_all_modes(sparse_model, batched):   # test_gp.py:133-176
    m = sparse_model
    rng = np.random.RandomState(0)
    x = dev(rng.randn(20, 40, 2) if batched else rng.randn(40, 2))
    m._begin_run(1, None)
    with m.tf_mode():
        samples = m.sparse_gp.samples(x, m.u)
    assert tuple(samples.shape) == (20, 40)
    torch.sum(samples * samples).backward()
    grads = [p.grad for p in m.get_tf_variables() if p.grad is not None]
    assert len(grads) > 0 and all(torch.isfinite(g).all() for g in grads)
    for q_shape in ('neglected', 'fullrank'):
        m._begin_run(1, None)
        with m.tf_mode():
            s = m.sparse_gp.samples(x, m.u, q_shape=q_shape)
        assert tuple(s.shape) == (20, 40) and torch.isfinite(s).all()


# ------------------------------------------------------------------------------------------ test_densities.py
def test_bimixture_with_broadcast_fraction():               # test_densities.py:11-24
    rng = np.random.RandomState(0)
    a = rng.randn(2, 3, 4); b = rng.randn(2, 3, 4); frac = rng.uniform(size=(2, 1, 1))
    logp0 = hb.densities.gaussian(dev(a), 0.0, 2.0)
    logp1 = hb.densities.student_t(dev(b), 0.0, 2.0, 3.0)
    mix = hb.densities.bimixture(dev(frac), logp0, logp1).cpu().numpy()
    ref = np.log(frac * np.exp(logp0.cpu().numpy()) + (1 - frac) * np.exp(logp1.cpu().numpy()))
    assert mix.dtype == np.float32 and np.allclose(mix, ref)


def _student_t_ref(x, mu, scale, nu):                        # test_densities.py:26-32
    const = loggamma(0.5 * (nu + 1.0)) - loggamma(0.5 * nu) - 0.5 * (np.log(scale * scale) + np.log(nu) + np.log(np.pi))
    return const - 0.5 * (nu + 1.) * np.log(1.0 + (1.0 / nu) * ((x - mu) / scale) ** 2.0)


def test_student_t_scalar_and_tensor_parameters():          # test_densities.py:34-70
    rng = np.random.RandomState(0)
    x = rng.randn(2, 3, 4).astype(f32); mu = rng.randn(2, 3, 4).astype(f32)
    scale = np.exp(rng.randn(2, 3, 4).astype(f32)); nu = np.exp(rng.randn(2, 3, 4).astype(f32))
    ref = _student_t_ref(x, mu, scale, 3.0)
    got = hb.densities.student_t(dev(x), dev(mu), dev(scale), 3.0).cpu().numpy()
    assert got.dtype == np.float32 and np.allclose(got, ref, atol=1e-5)
    got = hb.densities.student_t(dev(x), dev(mu), dev(scale), dev(nu)).cpu().numpy()
    assert np.allclose(got, _student_t_ref(x.astype(np.float64), mu, scale, nu.astype(np.float64)), atol=1e-5)
    got = hb.densities.student_t(dev(x), 0.5, 2.0, dev(nu)).cpu().numpy()          # scalar mean / scale, tensor deg_free
    assert np.allclose(got, _student_t_ref(x.astype(np.float64), 0.5, 2.0, nu.astype(np.float64)), atol=1e-5)


# ------------------------------------------------------------------------------------------ test_tf_wraps.py
def test_clip_by_value_on_nn_and_variational():             # test_tf_wraps.py:10-47
    rng = np.random.RandomState(0)
    x = dev(rng.randn(101, 100))
    cfg = hb.settings.get_settings()
    out = {}
    for clip in (False, True):
        cfg.numerics.clip_by_value = clip
        with hb.settings.temp_settings(cfg):
            m = hb.model.Model()
            m.nn = hb.nn.NeuralNet([100, 99, 98], neuron_types=tf.nn.relu, stddev=1.0)
            m.v = hb.variationals.Gaussian([100], stddev=100.0, mean=100.0)
            m.initialize()
            with m.tf_mode():
                y = m.run(m.nn(x))
                v = m.run(m.v)
        out[clip] = (y, v)
    assert np.max(out[False][0]) > 90 and np.max(out[False][1]) > 90
    # clip on: the NN output is clipped to [-50, 50]; the Gaussian's sample is clipped BEFORE its scale is applied
    # (variationals.py:112-119, 290-291), so it still exceeds 90 -- exactly what the reference asserts
    assert np.max(out[True][0]) < 90 and np.max(out[True][0]) <= hb.settings.numerics.clip_value_max
    assert np.max(out[True][1]) > 90


def test_log_sum_exp():                                      # test_tf_wraps.py:46-60
    rng = np.random.RandomState(0)
    a = rng.randn(2, 3, 4); b = rng.randn(2, 3, 4); c = rng.randn(2, 3, 4)
    value = hb.tf_wraps.log_sum_exp(tf.stack([dev(a), dev(b), dev(c)], axis=1), axis=1).cpu().numpy()
    assert np.allclose(value, np.log(np.exp(a) + np.exp(b) + np.exp(c)), atol=1e-6)
