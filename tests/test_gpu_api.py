"""The Henbun-shaped Python surface on the GPU: ports of the reference's own tests for the hot path
(testing/test_variationals.py, test_kernels.py, test_nn.py, test_model.py, test_gp.py) plus ELBO /
gradient / Adam-trajectory known-answer tests against the fp64 oracle for the BASELINE configs at
reduced size."""
import numpy as np
import pytest
import torch

import henbun_b200 as hb
import henbun_b200.tf as tf
from oracle import henbun_oracle as O

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


# ------------------------------------------------------------------ variationals (test_variationals.py)
@pytest.fixture()
def var_models():
    rng = np.random.RandomState(0)
    sqrts = {'fullrank': rng.randn(3, 10, 10) * 0.5, 'diagonal': rng.randn(3, 10) * 0.5 - 0.5}
    for i in range(3):
        for j in range(10):
            sqrts['fullrank'][i, j, j] = np.exp(sqrts['fullrank'][i, j, j])
            for k in range(j + 1, 10):
                sqrts['fullrank'][i, j, k] = 0.
    x = rng.randn(3, 10) * 0.3
    ms = {}
    for shape in ('fullrank', 'diagonal'):
        m = hb.model.Model()
        m.m = hb.variationals.Normal(x.shape[-1], n_layers=[3], q_shape=shape)
        m.m.q_mu = x
        m.m.q_sqrt = sqrts[shape]
        m.initialize()
        ms[shape] = m
    samples_iid = rng.randn(3, 10).astype(np.float32)
    return ms, sqrts, x, samples_iid


def test_logdet(var_models):                       # test_variationals.py:69-83
    ms, sqrts, x, _ = var_models
    ref = {'fullrank': 2.0 * np.log(np.stack([np.diag(s) for s in sqrts['fullrank']])), 'diagonal': 2.0 * sqrts['diagonal']}
    for shape, m in ms.items():
        v = object.__getattribute__(m, 'm')
        with m.tf_mode():
            logdet = v.logdet.detach().cpu().numpy()
        assert np.allclose(logdet, ref[shape], atol=1e-6)


def test_project_samples(var_models):              # test_variationals.py:85-106
    ms, sqrts, x, u = var_models
    post = {'fullrank': np.stack([x[i] + sqrts['fullrank'][i] @ u[i] for i in range(3)]),
            'diagonal': x + np.exp(sqrts['diagonal']) * u}
    for shape, m in ms.items():
        v = object.__getattribute__(m, 'm')
        with m.tf_mode():
            out = v._sample(torch.as_tensor(u).cuda()).detach().cpu().numpy()
        assert np.allclose(out, post[shape], rtol=1e-5, atol=1e-6)
        assert isinstance(v._kl_core, torch.Tensor)


def gaussian_KL(mu, L, q_shape):                   # analytic KL[N(mu, LL^T) || N(0, I)]
    KL = 0.0
    for i in range(mu.shape[0]):
        n = mu.shape[1]
        if q_shape == 'diagonal':
            logdet = 2.0 * np.sum(L[i]); trace = np.sum(np.exp(2.0 * L[i]))
        else:
            logdet = np.sum(np.log(np.square(np.diagonal(L[i])))); trace = np.sum(np.square(L[i]))
        KL += -logdet - n + trace + mu[i] @ mu[i]
    return 0.5 * KL


def test_KL_statistical(var_models):               # test_variationals.py:108-122 (rtol 0.1, 100 draws)
    ms, sqrts, x, _ = var_models
    for shape, m in ms.items():
        KL = 0.0
        with m.tf_mode():
            for _ in range(100):
                KL += m.run(m.KL())
        KL /= 100
        ana = gaussian_KL(x, sqrts[shape], shape)
        assert np.allclose(KL, ana, rtol=0.1), (shape, KL, ana)
    # and with S samples at once the same estimator, summed over the sample axis
    m = ms['diagonal']
    v = object.__getattribute__(m, 'm')
    kl = m.run(lambda: m.KL(), n_samples=400) / 400.0
    assert np.allclose(kl, gaussian_KL(x, sqrts['diagonal'], 'diagonal'), rtol=0.05)
    assert v.tensor().shape == (400, 3, 10)


def test_local_feed_order_and_shapes():           # test_variationals.py:166-203, test_param.py:117-124
    rng = np.random.RandomState(0)
    m = hb.model.Model()
    m.v = hb.variationals.Normal([2, 3], n_layers=[4], collections=hb.param.graph_key.LOCAL)
    assert m.v.feed_size == 12
    x = rng.randn(4, 7, 12).astype(np.float32)       # [n_layers, minibatch, q_mu(6) | q_sqrt(6)]
    with m.tf_mode():
        m.v = torch.as_tensor(x).cuda()
        s = m.v
        v = object.__getattribute__(m, 'v')
        logdet = v.logdet.detach().cpu().numpy()
        u = (v._tensor - torch.as_tensor(x[..., :6]).cuda()) / torch.exp(torch.as_tensor(x[..., 6:]).cuda())
    assert tuple(s.shape) == (4, 7, 2, 3)
    assert np.allclose(logdet, 2.0 * x[..., 6:], atol=1e-6)
    assert abs(float(u.mean())) < 0.3 and 0.7 < float(u.std()) < 1.3


def test_fullrank_init_diag_positive():            # test_variationals.py:279-286
    m = hb.model.Model()
    m.v = hb.variationals.Normal([5], q_shape='fullrank', stddev=0.5)
    assert np.all(np.diagonal(m.v.q_sqrt.value) > 0)


# ------------------------------------------------------------------ kernels (test_kernels.py)
def ref_sqdist(X, X2, l):
    return (((X[..., :, None, :] - X2[..., None, :, :]) / l) ** 2).sum(-1)


@pytest.fixture()
def kern_model():
    rng = np.random.RandomState(0)
    m = hb.model.Model()
    l1 = np.exp(rng.randn(1)); l2 = np.exp(rng.randn(2))
    m.k1 = hb.gp.kernels.UnitRBF(lengthscales=l1)
    m.k2 = hb.gp.kernels.UnitRBF(lengthscales=l2)
    m.k3 = hb.gp.kernels.UnitCsymRBF(lengthscales=l1)
    X = rng.randn(5, 2); X2 = rng.randn(6, 2); Xb = rng.randn(10, 5, 2); X2b = rng.randn(10, 6, 2)
    m.initialize()
    return m, l1, l2, X, X2, Xb, X2b


def test_K_all_variants(kern_model):               # test_kernels.py:90-182 (atol 1e-4)
    m, l1, l2, X, X2, Xb, X2b = kern_model
    with m.tf_mode():
        for A, B in ((X, None), (X, X2), (Xb, None), (Xb, X2b)):
            Bn = A if B is None else B
            K1 = m.k1.K(A, B).detach().cpu().numpy(); K2 = m.k2.K(A, B).detach().cpu().numpy(); K3 = m.k3.K(A, B).detach().cpu().numpy()
            assert np.allclose(K1, np.exp(-0.5 * ref_sqdist(A, Bn, l1)), atol=1e-5)
            assert np.allclose(K2, np.exp(-0.5 * ref_sqdist(A, Bn, l2)), atol=1e-5)
            assert np.allclose(K3, np.exp(-0.5 * ref_sqdist(A, Bn, l1)) + np.exp(-0.5 * ref_sqdist(A, -Bn, l1)), atol=1e-5)
        assert np.allclose(m.k1.Kdiag(torch.as_tensor(X).cuda()).detach().cpu().numpy(), np.ones(5))
        assert np.allclose(m.k1.K(X).detach().cpu().numpy(), m.k1.K(X.reshape(1, -1, 2)).detach().cpu().numpy()[0])   # :110-123
        # gradients exist (test_kernels.py:134-139)
        loss = tf.reduce_sum(m.k2.K(X, X2))
        g = torch.autograd.grad(loss, m.get_tf_variables(), allow_unused=True)
        assert sum(x is not None for x in g) > 0


def test_cholesky(kern_model):                     # test_kernels.py:184-226
    m, l1, l2, X, X2, Xb, X2b = kern_model
    j = hb.settings.numerics.jitter_level
    with m.tf_mode():
        for k in (m.k1, m.k2, m.k3):
            K = k.K(X).detach().cpu().numpy(); L = k.Cholesky(X).detach().cpu().numpy()
            assert L.shape == (5, 5) and np.allclose(np.triu(L, 1), 0)
            assert np.allclose(K + j * np.eye(5), L @ L.T, atol=1e-5)
            Kb = k.K(Xb).detach().cpu().numpy(); Lb = k.Cholesky(Xb).detach().cpu().numpy()
            assert Lb.shape == (10, 5, 5)
            for i in range(10):
                assert np.allclose(Kb[i] + j * np.eye(5), Lb[i] @ Lb[i].T, atol=1e-5)
        loss = tf.reduce_sum(m.k1.Cholesky(X))
        g = torch.autograd.grad(loss, m.get_tf_variables(), allow_unused=True)
        # value check of that gradient against the oracle (the reference only checks existence, :199-203)
        t = torch.tensor(np.log(np.expm1(l1 - 1e-6)), dtype=torch.float64, requires_grad=True)
        Lr = O.kern_cholesky(torch.tensor(X), O.log1pe_forward(t), j)
        Lr.sum().backward()
        got = [x for x in g if x is not None][0].detach().cpu().numpy()
        assert rel_err(got, t.grad.numpy()) < 1e-3      # 5 nearly collinear 2-D points: cond-limited in fp32


# ------------------------------------------------------------------ nn (test_nn.py)
def test_nn_matches_manual_chain():
    rng = np.random.RandomState(0)
    m = hb.model.Model()
    m.nn = hb.nn.NeuralNet([3, 2, 4], n_layers=[5], neuron_types=tf.sigmoid)
    m.nn2 = hb.nn.NeuralNet([3, 2, 4, 5], n_layers=[6, 5], neuron_types=[tf.nn.sigmoid, tf.nn.relu])
    m.initialize()
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))
    x = rng.randn(5, 6, 3).astype(np.float32)
    w1, b1, w2, b2 = (m.nn.matbias0.w.value, m.nn.matbias0.b.value, m.nn.matbias1.w.value, m.nn.matbias1.b.value)
    with m.tf_mode():
        y = m.nn(torch.as_tensor(x).cuda()).cpu().detach().numpy()
    assert np.allclose(y, sig(x @ w1 + b1) @ w2 + b2, atol=1e-4)         # test_nn.py:11-29
    x = rng.randn(6, 5, 6, 3).astype(np.float32)
    W = [m.nn2[i].w.value for i in range(3)]; Bv = [m.nn2[i].b.value for i in range(3)]
    with m.tf_mode():
        y = m.nn2(torch.as_tensor(x).cuda()).cpu().detach().numpy()
    ref = np.maximum(sig(x @ W[0] + Bv[0]) @ W[1] + Bv[1], 0) @ W[2] + Bv[2]
    assert np.allclose(y, ref, atol=1e-4)                                  # test_nn.py:31-52


def test_nn_gradients_match_oracle():
    rng = np.random.RandomState(1)
    m = hb.model.Model()
    m.nn = hb.nn.NeuralNet([7, 5, 3], neuron_types=tf.tanh, stddev=0.5)
    m.initialize()
    x = rng.randn(9, 7).astype(np.float32)
    with m.tf_mode():
        y = m.nn(torch.as_tensor(x).cuda())
        loss = tf.reduce_sum(tf.square(y))
    vs = m.get_tf_variables()
    g = torch.autograd.grad(loss, vs)
    tp = [torch.tensor(v.detach().cpu().numpy(), dtype=torch.float64, requires_grad=True) for v in vs]
    names = [v.long_name for v in m.get_variables()]
    d = dict(zip(names, tp))
    yr = O.neural_net(torch.tensor(x, dtype=torch.float64), [d['model.nn.matbias0.w'], d['model.nn.matbias1.w']],
                      [d['model.nn.matbias0.b'], d['model.nn.matbias1.b']], ['tanh'])
    (yr ** 2).sum().backward()
    for a, b in zip(g, tp):
        assert rel_err(a.detach().cpu().numpy(), b.grad.numpy()) < 1e-5


# ------------------------------------------------------------------ model / Adam (test_model.py)
class SquareModel(hb.model.Model):
    def setUp(self):
        self.p = hb.param.Variable([2, 3])
        self.q = hb.param.Variable([2, 3], collections=['another'])

    @hb.model.AutoOptimize()
    def likelihood(self):
        return -tf.reduce_sum(tf.square(self.p))

    @hb.model.AutoOptimize()
    def likelihood_q(self):
        return -tf.reduce_sum(tf.square(self.q))


def test_adam_converges_and_collections(tmp_path):
    m = SquareModel()
    m.likelihood().compile(optimizer=tf.train.AdamOptimizer(0.01), verbose=False)
    q0 = m.q.value.copy()
    m.likelihood().optimize(maxiter=1500)                                   # test_model.py:21-29
    assert np.allclose(m.p.value, 0.0, atol=1e-4)
    assert np.allclose(m.q.value, q0)                                       # other collection untouched (:61-74)
    m.likelihood_q().compile(optimizer=tf.train.AdamOptimizer(0.01), collection='another', verbose=False)
    m.likelihood_q().optimize(maxiter=1500)
    assert np.allclose(m.q.value, 0.0, atol=1e-4)
    # save / restore round trip (test_model.py:76-105)
    m.p = np.ones((2, 3)); m.initialize()
    path = m.save(str(tmp_path / "sq.ckpt"))
    m.p = np.zeros((2, 3)); m.initialize()
    m.restore(path)
    assert np.allclose(m.p.value, 1.0)


def test_adam_trajectory_matches_tf1_rule():
    m = SquareModel()
    p0 = m.p.value.astype(np.float64).copy()
    m.likelihood().compile(optimizer=tf.train.AdamOptimizer(0.05), verbose=False)
    th, mm, vv = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for t in range(1, 21):
        m.likelihood().optimize(maxiter=1)
        th, mm, vv = O.adam_tf1_step(th, 2.0 * th, mm, vv, t, lr=0.05)      # grad of loss = +2p
    assert np.allclose(m.p.value, th, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ BASELINE configs at reduced size
class GPR(hb.model.Model):                        # notebooks/GaussianProcess.ipynb:109-148
    def setUp(self, X=None, Y=None, q_shape='fullrank'):
        self.X = hb.param.Data(X)
        self.Y = hb.param.Data(Y)
        self.q = hb.variationals.Gaussian(shape=X.shape[:1] + (1,), q_shape=q_shape)
        self.kern = hb.gp.kernels.UnitRBF()
        self.k_var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)
        self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)

    @hb.model.AutoOptimize()
    def ELBO_gaussian(self):
        y_fit = tf.matmul(self.kern.Cholesky(self.X), self.q) * tf.sqrt(self.k_var)
        return tf.reduce_sum(hb.densities.gaussian(self.Y, y_fit, self.var)) - self.KL()


def _gpr_free_params(m, q_shape):
    g = lambda v: v._free_numpy().astype(np.float64)
    n = m.q.size
    return dict(q_mu=g(m.q.q_mu).reshape(n), q_sqrt=g(m.q.q_sqrt).reshape((n, n) if q_shape == 'fullrank' else (n,)),
                scale=g(m.q.scale).reshape(1), lengthscales=g(m.kern.lengthscales).reshape(-1),
                k_var=g(m.k_var).reshape(1), var=g(m.var).reshape(1))


@pytest.mark.parametrize("n,D,S,q_shape", [(100, 1, 10, 'fullrank'), (300, 8, 16, 'diagonal')])
def test_gpr_elbo_and_gradients_through_api(n, D, S, q_shape):
    rng = np.random.RandomState(0)
    if D == 1:
        X = np.linspace(0, 6, n).reshape(-1, 1); Y = np.sin(X) + 0.3 * rng.randn(n, 1); jitter = 1e-3
    else:
        X = rng.randn(n, D); Y = np.sin(X.sum(1, keepdims=True) / np.sqrt(D)) + 0.1 * rng.randn(n, 1); jitter = 1e-5
    cfg = hb.settings.get_settings(); cfg.numerics.jitter_level = jitter
    m = GPR(X=X, Y=Y, q_shape=q_shape)
    with hb.settings.temp_settings(cfg):
        m.ELBO_gaussian().compile(n_samples=S, verbose=False)
    p = _gpr_free_params(m, q_shape)
    U = rng.randn(S, n, 1).astype(np.float32)
    q = object.__getattribute__(m, 'q')
    val = m.ELBO_gaussian().run(eps={q: U})
    ref, gref = O.value_and_grads(lambda pp, *a: O.gpr_elbo(pp, *a, q_shape=q_shape, jitter=jitter), p, X, Y[:, 0], U[:, :, 0])
    K = O.rbf_K(torch.tensor(X), O.log1pe_forward(torch.tensor(p['lengthscales']))).numpy() + jitter * np.eye(n)
    tol = max(1e-5, 10 * np.linalg.cond(K) * 2.0 ** -24)
    assert abs(float(val) - ref) <= tol * abs(ref)
    opt = m.ELBO_gaussian()
    opt._flat_grad.zero_()
    opt._evaluate(opt.feed_dict(None), eps={q: U}, grad=True).backward()
    for name, var in (('q_mu', m.q.q_mu), ('q_sqrt', m.q.q_sqrt), ('scale', m.q.scale), ('lengthscales', m.kern.lengthscales),
                      ('k_var', m.k_var), ('var', m.var)):
        got = var._tensor.grad.detach().cpu().numpy().reshape(gref[name].shape)
        assert rel_err(got, gref[name]) <= 3 * tol, name


def test_gpr_ten_adam_steps_match_oracle():
    """Adam trajectory KAT: 10 steps with injected eps vs the fp64 oracle + TF-1 Adam rule."""
    rng = np.random.RandomState(3)
    n, D, S = 200, 8, 8
    X = rng.randn(n, D); Y = np.sin(X.sum(1, keepdims=True) / np.sqrt(D)) + 0.1 * rng.randn(n, 1)
    m = GPR(X=X, Y=Y, q_shape='diagonal')
    m.ELBO_gaussian().compile(optimizer=tf.train.AdamOptimizer(0.01), n_samples=S, verbose=False)
    p = _gpr_free_params(m, 'diagonal')
    mom = {k: np.zeros_like(v) for k, v in p.items()}; vel = {k: np.zeros_like(v) for k, v in p.items()}
    q = object.__getattribute__(m, 'q')
    for t in range(1, 11):
        U = rng.randn(S, n, 1).astype(np.float32)
        m.ELBO_gaussian().optimize(maxiter=1, eps={q: U})
        _, g = O.value_and_grads(O.gpr_elbo, p, X, Y[:, 0], U[:, :, 0])
        for k in p:
            p[k], mom[k], vel[k] = O.adam_tf1_step(p[k], -g[k], mom[k], vel[k], t, lr=0.01)
    got = _gpr_free_params(m, 'diagonal')
    for k in p:
        assert np.allclose(got[k], p[k], rtol=2e-4, atol=2e-5), k


class Amortised(hb.model.Model):                  # BASELINE config 4 at reduced size
    def setUp(self, X=None, latent=4, hidden=16):
        self.X = hb.param.MinibatchData(X)
        self.enc = hb.nn.NeuralNet([X.shape[1], hidden, 2 * latent], stddev=0.3)
        self.dec = hb.nn.NeuralNet([latent, hidden, X.shape[1]], stddev=0.3)
        self.q_local = hb.variationals.Normal([latent], collections=hb.param.graph_key.LOCAL)
        self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)

    @hb.model.AutoOptimize()
    def ELBO(self):
        self.q_local = self.enc(self.X)
        x_rec = self.dec(self.q_local)
        return tf.reduce_sum(hb.densities.gaussian(self.X, x_rec, self.var)) - self.KL(hb.param.graph_key.LOCAL)


def test_amortised_model_matches_oracle():
    rng = np.random.RandomState(0)
    Xall = rng.randn(500, 12).astype(np.float32)
    m = Amortised(X=Xall)
    S, B = 5, 32
    m.ELBO().compile(n_samples=S, verbose=False)
    idx = np.arange(B) * 3
    fd = {object.__getattribute__(m, 'X'): idx}
    U = rng.randn(S, B, 4).astype(np.float32)
    ql = object.__getattribute__(m, 'q_local')
    opt = m.ELBO()
    opt._flat_grad.zero_()
    val = opt._evaluate(fd, eps={ql: U}, grad=True)
    val.backward()
    g = lambda v: v._free_numpy().astype(np.float64)
    p = {'var': g(m.var)}
    for i in range(2):
        p[f'enc.w{i}'] = g(m.enc[i].w); p[f'enc.b{i}'] = g(m.enc[i].b)
        p[f'dec.w{i}'] = g(m.dec[i].w); p[f'dec.b{i}'] = g(m.dec[i].b)
    ref, gref = O.value_and_grads(lambda pp, X_, U_: O.amortised_elbo(pp, X_, U_, ['sigmoid'], ['sigmoid']), p,
                                  Xall[idx].astype(np.float64), U.astype(np.float64))
    assert abs(float(val) - ref) <= 1e-5 * abs(ref)
    pairs = [('var', m.var)] + [(f'enc.w{i}', m.enc[i].w) for i in range(2)] + [(f'enc.b{i}', m.enc[i].b) for i in range(2)] \
        + [(f'dec.w{i}', m.dec[i].w) for i in range(2)] + [(f'dec.b{i}', m.dec[i].b) for i in range(2)]
    for name, var in pairs:
        assert rel_err(var._tensor.grad.detach().cpu().numpy(), gref[name]) < 2e-5, name
    m.ELBO().optimize(maxiter=3, minibatch_size=B)      # the minibatch path runs end to end


class LinearOperator(hb.model.Model):             # BASELINE config 5 at reduced size
    def setUp(self, A=None, y=None):
        self.A = hb.param.Data(A)
        self.y = hb.param.Data(y)
        self.q = hb.variationals.Normal([A.shape[1]], q_shape='fullrank', stddev=0.1)
        self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)

    @hb.model.AutoOptimize()
    def ELBO(self):
        f = tf.matmul(self.q, self.A, transpose_b=True)          # [S,n] @ A^T -> [S,M]
        return tf.reduce_sum(hb.densities.gaussian(self.y, f, self.var)) - self.KL()


def test_linear_operator_model_matches_oracle():
    rng = np.random.RandomState(0)
    M, n, S = 150, 40, 6
    A = (rng.randn(M, n) / np.sqrt(n)).astype(np.float32); y = (A @ rng.randn(n) + 0.1 * rng.randn(M)).astype(np.float32)
    m = LinearOperator(A=A, y=y)
    m.q.q_sqrt = 0.1 * np.eye(n) + 1e-3 * np.tril(rng.randn(n, n))
    m.ELBO().compile(n_samples=S, verbose=False)
    U = rng.randn(S, n).astype(np.float32)
    q = object.__getattribute__(m, 'q')
    opt = m.ELBO()
    opt._flat_grad.zero_()
    val = opt._evaluate(opt.feed_dict(None), eps={q: U}, grad=True)
    val.backward()
    g = lambda v: v._free_numpy().astype(np.float64)
    p = dict(q_mu=g(m.q.q_mu), q_sqrt=g(m.q.q_sqrt), var=g(m.var))
    ref, gref = O.value_and_grads(O.linear_operator_elbo, p, A.astype(np.float64), y.astype(np.float64), U.astype(np.float64))
    assert abs(float(val) - ref) <= 1e-5 * abs(ref)
    for name, var in (('q_mu', m.q.q_mu), ('q_sqrt', m.q.q_sqrt), ('var', m.var)):
        assert rel_err(var._tensor.grad.detach().cpu().numpy(), gref[name]) < 2e-5, name


# ------------------------------------------------------------------ compile() binds recognised graphs to the fused entry points
@pytest.mark.parametrize("q_shape,n", [('diagonal', 260), ('fullrank', 130)])
def test_fused_binding_matches_the_eager_tape(q_shape, n):
    """The same notebook objective compiled with and without the whole-step binding: identical Philox windows, so five
    Adam steps must land on the same parameters (and on the oracle's, through the injected-eps trajectory test above)."""
    rng = np.random.RandomState(5)
    D, S = 4, 6
    X = rng.randn(n, D); Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.randn(n, 1)
    ms = []
    for fused in (True, False):
        np.random.seed(77)
        m = GPR(X=X, Y=Y, q_shape=q_shape)
        m.ELBO_gaussian().compile(optimizer=tf.train.AdamOptimizer(0.01), n_samples=S, seed=11, verbose=False, fused=fused)
        assert (m.ELBO_gaussian().fused_entry == 'GpElboBinding') == fused
        objs = [float(m.ELBO_gaussian().optimize(maxiter=1)) for _ in range(5)]
        ms.append((_gpr_free_params(m, q_shape), objs))
    (pf, of), (pt, ot) = ms
    assert np.allclose(of, ot, rtol=2e-5)
    for k in pf:
        assert np.allclose(pf[k], pt[k], rtol=1e-4, atol=2e-5), k      # two fp32 summation orders through 5 Adam steps


def test_fused_linear_operator_binding_matches_the_eager_tape():
    rng = np.random.RandomState(2)
    M, n, S = 300, 64, 8
    A = (rng.randn(M, n) / np.sqrt(n)).astype(np.float32); y = (A @ rng.randn(n) + 0.1 * rng.randn(M)).astype(np.float32)
    out = []
    for fused in (True, False):
        np.random.seed(3)
        m = LinearOperator(A=A, y=y)
        m.q.q_sqrt = 0.1 * np.eye(n) + 1e-3 * np.tril(rng.randn(n, n)) if False else 0.1 * np.eye(n)
        m.ELBO().compile(optimizer=tf.train.AdamOptimizer(0.01), n_samples=S, seed=5, verbose=False, fused=fused)
        assert (m.ELBO().fused_entry == 'LinopElboBinding') == fused
        objs = [float(m.ELBO().optimize(maxiter=1)) for _ in range(4)]
        g = lambda v: v._free_numpy().astype(np.float64)
        out.append((np.tril(g(m.q.q_sqrt)), g(m.q.q_mu), g(m.var), objs))
    for a, b in zip(out[0][:3], out[1][:3]):
        assert np.allclose(a, b, rtol=1e-4, atol=2e-6)
    assert np.allclose(out[0][3], out[1][3], rtol=2e-5)


class TwoObjectives(hb.model.Model):              # testing/test_gp.py compiles two objectives over the same variables
    def setUp(self):
        self.p = hb.param.Variable([4])

    @hb.model.AutoOptimize()
    def first(self):
        return -tf.reduce_sum(tf.square(self.p - 1.0))

    @hb.model.AutoOptimize()
    def second(self):
        return -tf.reduce_sum(tf.square(self.p + 1.0))


def test_two_optimizers_over_the_same_variables_alternate():
    m = TwoObjectives()
    m.first().compile(optimizer=tf.train.AdamOptimizer(0.05), verbose=False)
    m.second().compile(optimizer=tf.train.AdamOptimizer(0.05), verbose=False)     # re-binds p to the second optimizer's buffers
    p0 = m.p.value.copy()
    m.first().optimize(maxiter=200)                                               # ... the first must take it back and train
    assert np.all(np.abs(m.p.value - 1.0) < np.abs(p0 - 1.0))
    assert np.allclose(m.p.value, 1.0, atol=0.05)
    m.second().optimize(maxiter=400)
    assert np.allclose(m.p.value, -1.0, atol=0.05)
    m.first().optimize(maxiter=400)
    assert np.allclose(m.p.value, 1.0, atol=0.05)


def test_data_can_be_replaced_between_runs():
    """testing/test_data.py test_replacement: assigning a new array to a Data object takes effect at the next run."""
    class M(hb.model.Model):
        def setUp(self):
            self.d = hb.param.Data(np.ones((3, 2)))
            self.p = hb.param.Variable([1])

        @hb.model.AutoOptimize()
        def obj(self):
            return -tf.reduce_sum(tf.square(self.d * self.p))
    m = M()
    m.obj().compile(verbose=False)
    m.p = np.ones(1) * 2.0
    v1 = float(m.obj().run())
    m.d = np.full((3, 2), 3.0)
    v2 = float(m.obj().run())
    m.obj().optimize(maxiter=2)
    m.d = np.full((3, 2), 1.0)
    v3 = float(m.obj().run())
    assert np.isclose(v1, -24.0) and np.isclose(v2, -216.0)
    assert abs(v3) < 24.0 + 1e-3
    with pytest.raises(ValueError):
        m.d = np.ones((4, 2))


def test_device_minibatch_index_matches_its_definition():
    """hb_random_index: out[i] = pool[(w_i * pool_size) >> 64], w_i = two consecutive Philox words -- checked against the
    raw generator; draws stay inside the train split and are reproducible per (seed, offset)."""
    import ctypes as C
    from henbun_b200 import _lib
    lib = _lib.load()
    ix = hb.model.Indexer()
    ix.setUp(1000)
    a = ix.device_index(257, seed=5, offset=8).cpu().numpy()
    b = ix.device_index(257, seed=5, offset=8).cpu().numpy()
    c = ix.device_index(257, seed=6, offset=8).cpu().numpy()
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert set(a.tolist()) <= set(ix._train_index.tolist())
    raw = torch.zeros(4 * 129, dtype=torch.int32, device="cuda")
    ctr = (C.c_uint32 * 4)(2, 0, 0, 0); key = (C.c_uint32 * 2)(5, 0)
    assert lib.hb_philox4x32_10(_lib.ptr(raw), 129, ctr, key, _lib.stream()) == 0
    w = raw.cpu().numpy().view(np.uint32).astype(np.uint64)
    w64 = (w[1::2] << np.uint64(32)) | w[0::2]
    j = np.array([(int(x) * ix.train_size) >> 64 for x in w64[:257]])
    assert np.array_equal(a, ix._train_index[j])
    t = ix.device_index(64, seed=1, offset=0, training=False).cpu().numpy()
    assert set(t.tolist()) <= set(ix._test_index.tolist())


def test_minibatch_training_runs_on_device_indices():
    rng = np.random.RandomState(0)
    Xall = rng.randn(400, 12).astype(np.float32)
    m = Amortised(X=Xall)
    m.ELBO().compile(n_samples=3, verbose=False)                      # device_index=True is the default
    v = [float(m.ELBO().optimize(maxiter=1, minibatch_size=32)) for _ in range(3)]
    assert all(np.isfinite(v))
    m2 = Amortised(X=Xall)
    m2.ELBO().compile(n_samples=3, verbose=False, device_index=False)  # numpy's global RNG, as upstream
    assert np.isfinite(float(m2.ELBO().optimize(maxiter=2, minibatch_size=32)))
