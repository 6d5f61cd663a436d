"""hb_linop_elbo_local / hb_linop_elbo_update (csrc/linop.cu, BASELINE config 5) against the fp64 oracle
(oracle.linear_operator_elbo + the TF-1 Adam rule), the Python-API model of the same graph, and -- at the named size
n=16384, M=65536, S=64 -- the oracle's graph evaluated in fp64 plus size-independent properties (row-sharded ==
unsharded, untouched upper triangle, Philox regeneration).

Tolerance: 1e-5 relative on the ELBO and norm-wise on every gradient block (north star, fp32 vs fp64)."""
import ctypes as C

import numpy as np
import pytest
import torch

import henbun_b200 as hb
import henbun_b200.tf as tf
from henbun_b200 import _lib, ops
from henbun_b200.fused import LinearOperatorStep
from oracle import henbun_oracle as O

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def make_problem(M, n, S, seed=0):
    rng = np.random.RandomState(seed)
    A = (rng.randn(M, n) / np.sqrt(n)).astype(np.float32)
    zstar = rng.randn(n)
    y = (A @ zstar + 0.1 * rng.randn(M)).astype(np.float32)
    p = dict(q_mu=0.1 * rng.randn(n), q_sqrt=0.1 * np.eye(n) + 1e-2 * np.tril(rng.randn(n, n)) + 0.3 * np.triu(rng.randn(n, n), 1),
             var=np.array([0.3]))
    p = {k: v.astype(np.float32).astype(np.float64) for k, v in p.items()}
    U = rng.randn(S, n).astype(np.float32)
    return A, y, p, U


def build(A, y, p, S, m_total=None, **kw):
    st = LinearOperatorStep(torch.tensor(A, device="cuda"), torch.tensor(y, device="cuda"), S, m_total=m_total, **kw)
    st.set_params(p["q_mu"], p["q_sqrt"], p["var"])
    return st


@pytest.mark.parametrize("M,n,S", [(300, 70, 5), (512, 128, 64), (1000, 257, 3), (64, 64, 1), (2048, 512, 96), (4096, 1024, 64)])
def test_value_and_gradients_vs_oracle(M, n, S):
    A, y, p, U = make_problem(M, n, S)
    st = build(A, y, p, S)
    out, g = st.value_and_grads(torch.tensor(U, device="cuda"))
    ref, gref = O.value_and_grads(O.linear_operator_elbo, p, A.astype(np.float64), y.astype(np.float64), U.astype(np.float64))
    out = out.cpu().numpy(); g = g.cpu().numpy()
    assert abs(out[0] - ref) <= 1e-5 * abs(ref)
    gL = g[:n * n].reshape(n, n)
    assert np.all(np.triu(gL, 1) == 0) and np.all(np.triu(gref["q_sqrt"], 1) == 0)     # dead upper triangle
    assert rel_err(gL, gref["q_sqrt"]) < 1e-5
    assert rel_err(g[n * n:n * n + n], gref["q_mu"]) < 1e-5
    assert rel_err(g[n * n + n:], gref["var"]) < 1e-5
    # nothing moved
    assert np.array_equal(st.q_sqrt.cpu().numpy(), p["q_sqrt"].astype(np.float32))


@pytest.mark.parametrize("M,n,S", [(2048, 512, 64), (4096, 1024, 64), (1024, 256, 8), (2048, 512, 128), (520, 264, 16)])
def test_presplit_route_vs_oracle(M, n, S):
    """Both passes over A on the pre-split fp16 hi/lo engine (hb_linop_prepare + cfg.presplit: 256 x 64 pair tiles, the
    operator as the M-side operand): same bars as the fp32-operand route."""
    A, y, p, U = make_problem(M, n, S, seed=3)
    st = build(A, y, p, S, presplit=True)
    assert st.cfg.presplit == 1
    out, g = st.value_and_grads(torch.tensor(U, device="cuda"))
    ref, gref = O.value_and_grads(O.linear_operator_elbo, p, A.astype(np.float64), y.astype(np.float64), U.astype(np.float64))
    out = out.cpu().numpy(); g = g.cpu().numpy()
    assert abs(out[0] - ref) <= 1e-5 * abs(ref)
    assert rel_err(g[:n * n].reshape(n, n), gref["q_sqrt"]) < 1e-5
    assert rel_err(g[n * n:n * n + n], gref["q_mu"]) < 1e-5
    assert rel_err(g[n * n + n:], gref["var"]) < 1e-5
    # and against the fp32-operand route of the same library
    st0 = build(A, y, p, S, presplit=False)
    out0, g0 = st0.value_and_grads(torch.tensor(U, device="cuda"))
    assert rel_err(g, g0.cpu().numpy()) < 5e-6


def test_adam_trajectory_vs_oracle():
    """5 fused steps == 5 x (fp64 oracle gradient + TF-1 Adam rule); the upper triangle never moves."""
    M, n, S = 400, 96, 8
    A, y, p, _ = make_problem(M, n, S, seed=2)
    st = build(A, y, p, S, lr=0.01)
    rng = np.random.RandomState(5)
    mom = {k: np.zeros_like(v) for k, v in p.items()}; vel = {k: np.zeros_like(v) for k, v in p.items()}
    upper0 = np.triu(p["q_sqrt"], 1).astype(np.float32)
    for t in range(1, 6):
        U = rng.randn(S, n).astype(np.float32)
        st.step(torch.tensor(U, device="cuda"))
        _, g = O.value_and_grads(O.linear_operator_elbo, p, A.astype(np.float64), y.astype(np.float64), U.astype(np.float64))
        for k in p:
            p[k], mom[k], vel[k] = O.adam_tf1_step(p[k], -g[k], mom[k], vel[k], t, lr=0.01)
    assert np.allclose(st.q_mu.cpu().numpy(), p["q_mu"], rtol=2e-4, atol=2e-5)
    assert np.allclose(np.tril(st.q_sqrt.cpu().numpy()), np.tril(p["q_sqrt"]), rtol=2e-4, atol=2e-5)
    assert np.allclose(st.var_free.cpu().numpy(), p["var"], rtol=2e-4, atol=2e-5)
    assert np.array_equal(np.triu(st.q_sqrt.cpu().numpy(), 1), upper0)
    assert int(st.step_dev.item()) == 6


def test_row_sharded_equals_unsharded():
    """rank g holds M/G rows; the partial Zbar and {loglik, sum E^2} add up.  Two 'ranks' emulated on
    one GPU (sum instead of the all-reduce), unequal shard sizes, one empty shard."""
    M, n, S = 700, 128, 16
    A, y, p, U = make_problem(M, n, S, seed=4)
    Ud = torch.tensor(U, device="cuda")
    whole = build(A, y, p, S)
    out_w, g_w = whole.value_and_grads(Ud)
    for cuts in ([0, 300, 700], [0, 0, 700], [0, 100, 350, 700]):
        parts = [build(A[a:b], y[a:b], p, S, m_total=M) for a, b in zip(cuts[:-1], cuts[1:])]
        total = None
        for st in parts:
            z = st.local(Ud).clone()
            total = z if total is None else total + z
        last = parts[-1]
        last.zbar_stats.copy_(total)
        g = torch.zeros(last.count, device="cuda")
        out = last.update(grads=g, apply_adam=False)
        assert abs(float(out[0]) - float(out_w[0])) <= 2e-6 * abs(float(out_w[0]))
        assert rel_err(g.cpu().numpy(), g_w.cpu().numpy()) < 2e-6


def test_philox_path_is_reproducible():
    M, n, S = 256, 64, 8
    A, y, p, _ = make_problem(M, n, S, seed=6)
    st = build(A, y, p, S, seed=11)
    o1, g1 = st.value_and_grads(None)
    o1 = o1.clone()
    o2, g2 = st.value_and_grads(None)
    assert torch.equal(o1, o2) and torch.equal(g1, g2)
    eps = ops.randn_philox((S, n), 11, 0, torch.device("cuda"))
    o3, g3 = st.value_and_grads(eps)
    assert torch.equal(o1, o3) and torch.equal(g1, g3)
    st.cfg.offset = 2          # not a multiple of 4
    rc = st.lib.hb_linop_elbo_local(C.byref(st.cfg), _lib.ptr(st.A), _lib.ptr(st.y), _lib.ptr(st.params), None,
                                    _lib.ptr(st.zbar_stats), _lib.ptr(st.ws), st.ws_bytes, _lib.stream())
    assert rc == _lib.HB_ERR_ARG
    st.cfg.offset = 0
    rc = st.lib.hb_linop_elbo_local(C.byref(st.cfg), _lib.ptr(st.A), _lib.ptr(st.y), _lib.ptr(st.params), None,
                                    _lib.ptr(st.zbar_stats), _lib.ptr(st.ws), 1024, _lib.stream())
    assert rc == _lib.HB_ERR_WORKSPACE


def test_matches_python_api_model():
    """The same graph written against the Henbun API (autograd tape over the same kernels) gives the same numbers."""
    M, n, S = 384, 128, 8
    A, y, p, U = make_problem(M, n, S, seed=8)

    class LinearOperator(hb.model.Model):
        def setUp(self):
            self.A = hb.param.Data(A); self.y = hb.param.Data(y)
            self.q = hb.variationals.Normal([n], q_shape='fullrank', stddev=0.1)
            self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)

        @hb.model.AutoOptimize()
        def ELBO(self):
            f = tf.matmul(self.q, self.A, transpose_b=True)
            return tf.reduce_sum(hb.densities.gaussian(self.y, f, self.var)) - self.KL()

    m = LinearOperator()
    vals = {"model.q.q_mu": p["q_mu"], "model.q.q_sqrt": p["q_sqrt"], "model.var": p["var"]}
    for v in m.get_variables():
        if v.is_parameter and v.long_name in vals:
            v._pending = vals[v.long_name].astype(np.float32).reshape(v._host.shape); v._assigned = True
    m.ELBO().compile(n_samples=S, verbose=False)
    opt = m.ELBO()
    q = object.__getattribute__(m, "q")
    opt._flat_grad.zero_()
    val = opt._evaluate(opt.feed_dict(None), eps={q: U}, grad=True)
    val.backward()
    st = build(A, y, p, S)
    out, g = st.value_and_grads(torch.tensor(U, device="cuda"))
    assert abs(float(val) - float(out[0])) <= 2e-6 * abs(float(out[0]))
    grads = {v.long_name: v._tensor.grad.detach().cpu().numpy() for v in m.get_variables() if v.is_parameter}
    g = g.cpu().numpy()
    assert rel_err(np.tril(grads["model.q.q_sqrt"].reshape(n, n)), g[:n * n].reshape(n, n)) < 3e-6
    assert rel_err(grads["model.q.q_mu"].ravel(), g[n * n:n * n + n]) < 3e-6
    assert rel_err(grads["model.var"].ravel(), g[n * n + n:]) < 3e-6


def test_full_size_config5():
    """BASELINE config 5 at its named size (n=16384 full-covariance q, A 65536x16384 fp32 = 4.3 GB, S=64).
    (1) ELBO, mu-bar, var-bar and a 512-row band of the q_sqrt gradient against the oracle's graph evaluated in fp64 on
    the device tensors; (2) two row shards == the whole operator; (3) a real Adam step leaves the strict upper triangle
    bit-identical and moves every lower-triangle entry by at most lr (|m/(sqrt(v)+eps)| <= 1/sqrt(1-b2) * sqrt(1-b2^1)...
    = 1 at t=1)."""
    M, n, S = 65536, 16384, 64
    gen = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(M, n, device="cuda", generator=gen) / np.sqrt(n)
    y = A[:, :256] @ torch.randn(256, device="cuda", generator=gen) + 0.1 * torch.randn(M, device="cuda", generator=gen)
    st = LinearOperatorStep(A, y, S, lr=1e-3)
    st.q_sqrt.copy_(0.1 * torch.eye(n, device="cuda") + 1e-3 * torch.randn(n, n, device="cuda", generator=gen))
    st.q_mu.copy_(0.1 * torch.randn(n, device="cuda", generator=gen))
    st.var_free.fill_(0.3)
    U = torch.randn(S, n, device="cuda", generator=gen)
    g = torch.zeros(st.count, device="cuda")
    st.local(U)
    out = st.update(grads=g, apply_adam=False).clone()

    # (1) fp64 evaluation of the oracle's graph on the same device tensors
    p64 = {"q_mu": st.q_mu.double().requires_grad_(True), "q_sqrt": st.q_sqrt.double().requires_grad_(True),
           "var": st.var_free.double().requires_grad_(True)}
    ref = O.linear_operator_elbo(p64, A.double(), y.double(), U.double())
    ref.backward()
    assert abs(float(out[0]) - float(ref)) <= 1e-5 * abs(float(ref))
    nn_ = n * n
    assert rel_err(g[nn_:nn_ + n].cpu().numpy(), p64["q_mu"].grad.cpu().numpy()) < 1e-5
    assert rel_err(g[nn_ + n:].cpu().numpy(), p64["var"].grad.cpu().numpy()) < 1e-5
    band = slice(n - 512, n)
    assert rel_err(g[:nn_].view(n, n)[band].cpu().numpy(), p64["q_sqrt"].grad[band].cpu().numpy()) < 1e-5
    assert rel_err(g[:nn_].view(n, n)[:512].cpu().numpy(), p64["q_sqrt"].grad[:512].cpu().numpy()) < 1e-5
    del p64, ref

    # (2) row-sharded: 2 shards of M/2 rows, partial results summed
    half = M // 2
    total = None
    for a in (0, half):
        sh = LinearOperatorStep(A[a:a + half], y[a:a + half], S, m_total=M)
        sh.params.copy_(st.params)
        z = sh.local(U).clone()
        total = z if total is None else total + z
        del sh
    st.zbar_stats.copy_(total)
    g2 = torch.zeros(st.count, device="cuda")
    out2 = st.update(grads=g2, apply_adam=False)
    assert abs(float(out2[0]) - float(out[0])) <= 2e-6 * abs(float(out[0]))
    assert float((g2 - g).norm() / g.norm()) < 2e-6
    del g2, total

    # (3) one real step
    before = st.q_sqrt.clone()
    st.step(U)
    after = st.q_sqrt
    iu = torch.triu(torch.ones(n, n, dtype=torch.bool, device="cuda"), 1)
    assert torch.equal(after[iu], before[iu])
    d = (after - before).abs()
    assert float(d.max()) <= 1e-3 * 1.0001
    moved = (d[~iu] > 0).float().mean()
    assert float(moved) > 0.99
