"""The densities.py kernel family (csrc/density_family.cu) through the C ABI and through hb.densities:
* against the values and tf.gradients the unmodified reference produced (tests/golden/densities_all.npz),
* against the fp64 oracle on larger, ragged and broadcast shapes (vector path, scalar path, in-kernel reduction of
  scalar-operand gradients, expanded operands),
* size-independent properties at 2^24 elements: permutation invariance of the reduced gradient, linearity in g.

Tolerance (fp32 kernels vs fp64 numbers): |err| <= 2e-6 * (1 + |ref|) * scale per element, where scale is the size of the
largest term that cancels inside the density (stated per case), and 1e-5 relative norm-wise on gradients.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import henbun_b200 as hb
from henbun_b200 import _lib, ops
from oracle import henbun_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ["lognormal", "bernoulli", "poisson", "exponential", "gamma", "student_t", "beta", "laplace", "bimixture"]


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def dev(a, grad=False):
    return torch.tensor(np.asarray(a, np.float32), device="cuda", requires_grad=grad)


@pytest.mark.parametrize("name", NAMES)
def test_density_vs_reference(name):
    d = np.load(os.path.join(G, "densities_all.npz"))
    n_args = len([k for k in d.files if k.startswith(name + "/arg")])
    args = [dev(d[f"{name}/arg{i}"], grad=True) for i in range(n_args)]
    out = getattr(hb.densities, name)(*args)
    ref = d[name + "/out"]
    assert out.shape == ref.shape
    assert np.allclose(out.detach().cpu().numpy(), ref, rtol=3e-6, atol=3e-6), name
    (out * dev(d[name + "/w"])).sum().backward()
    for i, a in enumerate(args):
        gref = d[f"{name}/grad{i}"]
        if not np.any(gref):
            assert a.grad is None or not torch.any(a.grad), (name, i)
            continue
        assert a.grad.shape == gref.shape
        assert rel_err(a.grad.cpu().numpy(), gref) < 1e-5, (name, i)


def _inputs(name, shape, rng, suffix_shapes):
    """Random valid operands: `suffix_shapes` gives the shape of every operand (broadcast against `shape`)."""
    pos = lambda s: np.exp(0.4 * rng.randn(*s))
    sh = suffix_shapes
    if name == "gaussian":
        return [rng.randn(*sh[0]), rng.randn(*sh[1]), pos(sh[2])]
    if name == "lognormal":
        return [pos(sh[0]), rng.randn(*sh[1]), pos(sh[2])]
    if name == "bernoulli":
        return [rng.uniform(0.05, 0.95, sh[0]), (rng.rand(*sh[1]) < 0.5).astype(np.float64)]
    if name == "poisson":
        return [pos(sh[0]), rng.poisson(3.0, sh[1]).astype(np.float64)]
    if name == "exponential":
        return [pos(sh[0]), pos(sh[1])]
    if name == "gamma":
        return [pos(sh[0]) + 0.5, pos(sh[1]), pos(sh[2])]
    if name == "student_t":
        return [rng.randn(*sh[0]), rng.randn(*sh[1]), pos(sh[2]), pos(sh[3]) * 4.0]
    if name == "beta":
        return [pos(sh[0]) + 0.5, pos(sh[1]) + 0.5, rng.uniform(0.02, 0.98, sh[2])]
    if name == "laplace":
        return [rng.randn(*sh[0]), pos(sh[1]), rng.randn(*sh[2])]
    return [rng.uniform(0.05, 0.95, sh[0]), 3 * rng.randn(*sh[1]), 3 * rng.randn(*sh[2])]


NARGS = {"gaussian": 3, "lognormal": 3, "bernoulli": 2, "poisson": 2, "exponential": 2, "gamma": 3, "student_t": 4, "beta": 3,
         "laplace": 3, "bimixture": 3}


@pytest.mark.parametrize("name", ["gaussian"] + NAMES)
@pytest.mark.parametrize("layout", ["full_vec", "ragged", "scalars", "suffix", "expanded"])
def test_density_vs_oracle_layouts(name, layout):
    """full_vec: every operand full-size, total % 4 == 0 (float4 path); ragged: odd total (scalar path);
    scalars: all but the first operand are [1] (in-kernel gradient reduction); suffix: mixed suffix shapes;
    expanded: an operand that broadcasts along the LAST axis (host expands it)."""
    rng = np.random.RandomState(7)
    na = NARGS[name]
    if layout == "full_vec":
        shape = (6, 33, 8); shapes = [shape] * na
    elif layout == "ragged":
        shape = (5, 7, 3); shapes = [shape] * na
    elif layout == "scalars":
        shape = (4, 129, 4); shapes = [shape] + [(1,)] * (na - 1)
    elif layout == "suffix":
        shape = (3, 10, 12); shapes = [shape, (10, 12), (12,), (1, 12)][:na]
        if na == 2:
            shapes = [(10, 12), shape]
    else:
        shape = (3, 10, 12); shapes = [shape] + [(10, 1)] + [(1,)] * (na - 2)
    if name in ("bernoulli", "poisson") and layout in ("scalars", "expanded"):
        shapes = [shapes[1], shapes[0]]          # keep the data operand (y) full-size
    vals = _inputs(name, shape, rng, shapes)
    w = rng.randn(*shape)
    args = [dev(v, grad=True) for v in vals]
    out = getattr(hb.densities, name)(*args)
    (out * dev(w)).sum().backward()
    targs = [torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in vals]
    ref = O.DENSITIES[name](*targs)
    (ref * torch.tensor(w)).sum().backward()
    assert tuple(out.shape) == tuple(ref.shape) == shape
    assert np.allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=3e-6, atol=5e-6), (name, layout)
    for i, (a, t) in enumerate(zip(args, targs)):
        if t.grad is None or not torch.any(t.grad):
            continue
        assert tuple(a.grad.shape) == tuple(t.grad.shape), (name, layout, i)
        assert rel_err(a.grad.cpu().numpy(), t.grad.numpy()) < 1e-5, (name, layout, i)


def test_reduce_sum_backward_takes_the_scalar_path():
    """tf.reduce_sum(densities.gaussian(x, f, var)) -- the shape of every objective in the notebooks: the incoming
    gradient is a broadcast scalar (g_period 1) and d/dvar is reduced in-kernel.  Checked against the closed form."""
    rng = np.random.RandomState(3)
    S, B, Dd = 8, 512, 784
    x = dev(rng.randn(B, Dd)); f = dev(rng.randn(S, B, Dd), grad=True); var = dev([0.7], grad=True)
    ll = torch.sum(hb.densities.gaussian(x, f, var))
    ll.backward()
    e = (f.detach().double() - x.double())
    v = 0.7 + 0.0
    v = float(np.float32(0.7))
    ref = float((-0.5 * np.log(2 * np.pi) - 0.5 * np.log(v)) * e.numel() - 0.5 * float((e * e).sum()) / v)
    assert abs(float(ll) - ref) <= 1e-5 * abs(ref)
    assert rel_err(f.grad.cpu().numpy(), (-(e) / v).cpu().numpy()) < 1e-6
    gv = float(-0.5 * e.numel() / v + 0.5 * float((e * e).sum()) / v ** 2)
    assert abs(float(var.grad) - gv) <= 1e-5 * abs(gv)


def test_cabi_errors_and_empty():
    lib = _lib.load()
    assert lib.hb_density_nargs(6) == 4 and lib.hb_density_nargs(10) == -1 and lib.hb_density_nargs(-1) == -1
    x = torch.ones(8, device="cuda")
    arr = (C.c_void_p * 4)(x.data_ptr(), x.data_ptr(), x.data_ptr(), None)
    per = (C.c_longlong * 4)(8, 8, 8, 1)
    out = torch.empty(8, device="cuda")
    st = _lib.stream()
    assert lib.hb_density_logpdf(0, arr, per, 0, _lib.ptr(out), st) == 0                 # empty input: no launch
    assert lib.hb_density_logpdf(42, arr, per, 8, _lib.ptr(out), st) == _lib.HB_ERR_ARG   # unknown kind
    assert lib.hb_density_logpdf(0, arr, (C.c_longlong * 4)(8, 0, 8, 1), 8, _lib.ptr(out), st) == _lib.HB_ERR_ARG
    assert lib.hb_density_logpdf(0, arr, (C.c_longlong * 4)(8, 16, 8, 1), 8, _lib.ptr(out), st) == _lib.HB_ERR_ARG
    assert lib.hb_density_logpdf(6, arr, per, 8, _lib.ptr(out), st) == _lib.HB_ERR_ARG    # 4th operand missing
    # scalar-operand gradient without a workspace
    d = (C.c_void_p * 4)(None, None, out.data_ptr(), None)
    assert lib.hb_density_logpdf_bwd(0, arr, (C.c_longlong * 4)(8, 8, 1, 1), 8, _lib.ptr(x), 8, d, None, 0, st) == _lib.HB_ERR_WORKSPACE
    e = hb.densities.gaussian(torch.empty(0, 3, device="cuda"), torch.zeros(3, device="cuda"), torch.ones(1, device="cuda"))
    assert e.shape == (0, 3)


def test_full_size_properties():
    """2^24 elements (student_t, the heaviest member): the in-kernel reduced gradient of the scalar operands is
    invariant under a permutation of the data to fp32 round-off of the SUM (deterministic run to run: bitwise), the
    backward is linear in g, and forward == sum of per-chunk forwards."""
    n = 1 << 24
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, device="cuda", generator=gen)
    mean = torch.zeros(1, device="cuda", requires_grad=True)
    scale = torch.full((1,), 1.3, device="cuda", requires_grad=True)
    nu = torch.full((1,), 4.0, device="cuda", requires_grad=True)

    def grads(xx, gscale=1.0):
        for p in (mean, scale, nu):
            p.grad = None
        (gscale * torch.sum(hb.densities.student_t(xx, mean, scale, nu))).backward()
        return [float(p.grad) for p in (mean, scale, nu)]
    g1 = grads(x); g1b = grads(x)
    assert g1 == g1b                                                     # bitwise deterministic
    perm = torch.randperm(n, device="cuda", generator=gen)
    g2 = grads(x[perm])
    g3 = grads(x, 2.0)
    # d/dmean is a sum of n zero-mean terms of size ~0.5 (sum ~ sqrt(n)): compare on the scale of sqrt(n)
    assert abs(g1[0] - g2[0]) <= 1e-5 * np.sqrt(n)
    for a, b in zip(g1[1:], g2[1:]):
        assert abs(a - b) <= 1e-6 * abs(a) + 1e-5 * np.sqrt(n)
    for a, b in zip(g1, g3):
        assert abs(2 * a - b) <= 1e-6 * abs(b) + 1e-6
    full = hb.densities.student_t(x, mean, scale, nu).detach()
    parts = torch.cat([hb.densities.student_t(c, mean, scale, nu).detach() for c in x.split(n // 8 + 4)])
    assert torch.equal(full, parts)
    ref = O.student_t(x[:4096].double().cpu(), torch.zeros(1, dtype=torch.float64), 1.3, 4.0)
    assert np.allclose(full[:4096].cpu().numpy(), ref.numpy(), rtol=3e-6, atol=3e-6)
