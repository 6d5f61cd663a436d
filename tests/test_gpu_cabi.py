"""Parity of every C-ABI entry point against the CPU oracle (fp64) on seeded inputs.

All tests call the CUDA library through ctypes (henbun_b200._lib) -- exactly the boundary a
reference-side binding would use.  Tolerances are written next to each comparison; fp32 kernels are
compared with the fp64 oracle (rel 1e-5 on well-conditioned inputs).
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import henbun_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from henbun_b200 import _lib
    return _lib.load()


_KEEP = []      # device tensors must outlive the asynchronous kernel that reads them


@pytest.fixture(autouse=True)
def _release_device_tensors():
    yield
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    _KEEP.clear()


def dev(a, dtype=torch.float32):
    t = torch.as_tensor(np.asarray(a), dtype=dtype).cuda().contiguous()
    _KEEP.append(t)
    return t


def P(t):
    from henbun_b200._lib import ptr
    return ptr(t)


def ST():
    from henbun_b200._lib import stream
    return stream()


def WS():
    from henbun_b200._lib import reduce_ws
    return reduce_ws()


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def call_gemm(lib, A, B, C, M, N, K, transA=0, transB=0, a_tri=0, b_tri=0, c_tri=0, alpha=1.0, beta=0.0,
              bias=None, act=0, clip=0, batch=1, sA=0, sB=0, sC=0, sBias=0, lda=None, ldb=None, ldc=None):
    lda = lda if lda is not None else (M if transA else K)
    ldb = ldb if ldb is not None else (K if transB else N)
    ldc = ldc if ldc is not None else N
    rc = lib.hb_gemm(P(A), lda, sA, transA, a_tri, P(B), ldb, sB, transB, b_tri, P(C), ldc, sC, c_tri, M, N, K, batch,
                     alpha, beta, P(bias), sBias, act, clip, -50.0, 50.0, ST())
    assert rc == 0, rc
    torch.cuda.synchronize()


def mask_a(mode, M, K):
    m = np.arange(M)[:, None]; k = np.arange(K)[None, :]
    return {0: np.ones((M, K), bool), 1: k <= m, 2: k >= m, 3: k > m, 4: k < m}[mode]


def mask_b(mode, K, N):
    k = np.arange(K)[:, None]; n = np.arange(N)[None, :]
    return {0: np.ones((K, N), bool), 1: n <= k, 2: n >= k, 3: n > k, 4: n < k}[mode]


@pytest.mark.parametrize("engine", [1])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 7, 3), (64, 128, 8), (100, 100, 100), (129, 257, 65),
                                   (300, 200, 513), (64, 600, 600), (512, 384, 256)])
@pytest.mark.parametrize("transA,transB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_layouts(lib, engine, M, N, K, transA, transB):
    lib.hb_set_gemm_engine(engine)
    rng = np.random.RandomState(M * 131 + N * 17 + K)
    A = rng.randn(M, K); B = rng.randn(K, N); C0 = rng.randn(M, N)
    Ad = dev(A.T if transA else A); Bd = dev(B.T if transB else B); Cd = dev(C0)
    call_gemm(lib, Ad, Bd, Cd, M, N, K, transA, transB, alpha=0.7, beta=-0.3)
    ref = 0.7 * A @ B - 0.3 * C0
    assert rel_err(Cd.cpu().numpy(), ref) < 2e-6          # fp32 SIMT accumulate
    lib.hb_set_gemm_engine(0)


@pytest.mark.parametrize("a_tri,b_tri,c_tri", [(1, 0, 0), (2, 0, 0), (3, 0, 0), (4, 0, 0), (0, 1, 0), (0, 2, 0),
                                               (0, 3, 0), (0, 4, 0), (0, 0, 1), (1, 2, 1)])
@pytest.mark.parametrize("transA,transB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_triangular_masks(lib, a_tri, b_tri, c_tri, transA, transB):
    lib.hb_set_gemm_engine(1)
    n = 200
    rng = np.random.RandomState(7)
    A = rng.randn(n, n); B = rng.randn(n, n); C0 = rng.randn(n, n)
    Am = A * mask_a(a_tri, n, n); Bm = B * mask_b(b_tri, n, n)
    Ad = dev(A.T if transA else A); Bd = dev(B.T if transB else B); Cd = dev(C0)
    call_gemm(lib, Ad, Bd, Cd, n, n, n, transA, transB, a_tri, b_tri, c_tri, alpha=1.0, beta=1.0)
    ref = Am @ Bm + C0
    out = Cd.cpu().numpy()
    if c_tri:
        iu = np.triu_indices(n, 1)
        assert np.array_equal(out[iu], C0.astype(np.float32)[iu])     # strict upper untouched
        out = np.tril(out); ref = np.tril(ref)
    assert rel_err(out, ref) < 2e-6
    lib.hb_set_gemm_engine(0)


def test_gemm_batched_bias_act(lib):
    lib.hb_set_gemm_engine(1)
    rng = np.random.RandomState(3)
    nb, M, K, N = 5, 6, 3, 2
    x = rng.randn(nb, M, K); w = rng.randn(nb, K, N); b = rng.randn(nb, 1, N)
    out = torch.empty(nb, M, N, device="cuda")
    call_gemm(lib, dev(x), dev(w), out, M, N, K, batch=nb, sA=M * K, sB=K * N, sC=M * N, bias=dev(b), sBias=N, act=1)
    ref = 1.0 / (1.0 + np.exp(-(x @ w + b)))
    assert np.allclose(out.cpu().numpy(), ref, atol=1e-6)
    lib.hb_set_gemm_engine(0)


def test_philox_stream_consistency(lib):
    n = 4096 * 5 + 4
    a = torch.empty(n, device="cuda"); b = torch.empty(n - 8, device="cuda")
    assert lib.hb_randn_philox(P(a), n, 1234, 0, ST()) == 0
    assert lib.hb_randn_philox(P(b), n - 8, 1234, 8, ST()) == 0
    torch.cuda.synchronize()
    assert torch.equal(a[8:], b)                       # counter based: offset shifts the stream
    x = a.double().cpu().numpy()
    assert abs(x.mean()) < 0.03 and abs(x.std() - 1.0) < 0.03
    assert lib.hb_randn_philox(P(a), n, 1234, 3, ST()) == 1      # offset must be a multiple of 4


@pytest.mark.parametrize("S,rows,cols,strided", [(1, 3, 10, False), (7, 1, 1000, False), (4, 6, 10, True),
                                                 (64, 1, 4096, False), (3, 1, 5, False)])
def test_sample_diag_fwd_bwd(lib, S, rows, cols, strided):
    rng = np.random.RandomState(0)
    mu = rng.randn(rows, cols) * 0.3; om = rng.randn(rows, cols) * 0.5 - 0.5
    eps = rng.randn(S, rows, cols); zbar = rng.randn(S, rows, cols)
    c = 1.0 / S
    if strided:      # mu | omega halves of one [rows, 2*cols] encoder output (param.py:529-537)
        h = dev(np.concatenate([mu, om], 1))
        mu_d, om_d, ld = h, h[:, cols:], 2 * cols
        gh = torch.zeros_like(h); gmu_d, gom_d, ldg = gh, gh[:, cols:], 2 * cols
    else:
        mu_d, om_d, ld = dev(mu), dev(om), cols
        gmu_d, gom_d, ldg = torch.zeros(rows, cols, device="cuda"), torch.zeros(rows, cols, device="cuda"), cols
    z = torch.empty(S, rows, cols, device="cuda"); kl = torch.zeros(1, device="cuda")
    ws = WS()
    rc = lib.hb_sample_diag_fwd(P(mu_d), ld, P(om_d), ld, rows, cols, P(dev(eps)), 0, 0, S, P(z), P(kl), P(ws),
                                ws.numel(), ST())
    assert rc == 0
    tm, to, te = (torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in (mu, om, eps))
    z_ref = O.sample_diag(tm, to, te)
    kl_ref = O.kl_normal(O.logdet_diag(to), te, z_ref)
    torch.cuda.synchronize()
    # test_variationals.py:85-106 uses default allclose on 30 values; fp32 round-off of O(1) operands needs atol ~1e-6
    assert np.allclose(z.cpu().numpy(), z_ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    assert abs(kl.item() - kl_ref.item()) <= 1e-5 * max(1.0, abs(kl_ref.item()))
    obj = torch.sum(z_ref * torch.tensor(zbar)) - c * kl_ref
    obj.backward()
    rc = lib.hb_sample_diag_bwd(P(mu_d), ld, P(om_d), ld, rows, cols, P(dev(eps)), 0, 0, S, P(dev(zbar)), None, c, None,
                                P(gmu_d), ldg, P(gom_d), ldg, 0.0, ST())
    assert rc == 0
    torch.cuda.synchronize()
    g1 = (gmu_d[:, :cols] if strided else gmu_d).cpu().numpy()
    g2 = (gom_d[:, :cols] if strided else gom_d).cpu().numpy()
    assert rel_err(g1, tm.grad.numpy()) < 1e-5
    assert rel_err(g2, to.grad.numpy()) < 1e-5


def test_sample_diag_philox_matches_materialised_eps(lib):
    S, n = 8, 1000
    rng = np.random.RandomState(1)
    mu = dev(rng.randn(1, n)); om = dev(rng.randn(1, n) * 0.1)
    eps = torch.empty(S, n, device="cuda")
    assert lib.hb_randn_philox(P(eps), S * n, 42, 16, ST()) == 0
    ws = WS()
    z1 = torch.empty(S, n, device="cuda"); z2 = torch.empty(S, n, device="cuda")
    k1 = torch.zeros(1, device="cuda"); k2 = torch.zeros(1, device="cuda")
    assert lib.hb_sample_diag_fwd(P(mu), n, P(om), n, 1, n, P(eps), 0, 0, S, P(z1), P(k1), P(ws), ws.numel(), ST()) == 0
    assert lib.hb_sample_diag_fwd(P(mu), n, P(om), n, 1, n, None, 42, 16, S, P(z2), P(k2), P(ws), ws.numel(), ST()) == 0
    zb = dev(rng.randn(S, n))
    g = [torch.zeros(1, n, device="cuda") for _ in range(4)]
    assert lib.hb_sample_diag_bwd(P(mu), n, P(om), n, 1, n, P(eps), 0, 0, S, P(zb), None, 0.125, None, P(g[0]), n, P(g[1]), n, 0.0, ST()) == 0
    assert lib.hb_sample_diag_bwd(P(mu), n, P(om), n, 1, n, None, 42, 16, S, P(zb), None, 0.125, None, P(g[2]), n, P(g[3]), n, 0.0, ST()) == 0
    torch.cuda.synchronize()
    assert torch.equal(z1, z2) and torch.equal(k1, k2)
    assert torch.equal(g[0], g[2]) and torch.equal(g[1], g[3])


@pytest.mark.parametrize("batch,n,S", [(3, 10, 1), (1, 100, 10), (2, 130, 5)])
def test_sample_tril_fwd_bwd(lib, batch, n, S):
    rng = np.random.RandomState(0)
    Lq = rng.randn(batch, n, n) * 0.5
    for b in range(batch):
        Lq[b][np.diag_indices(n)] = np.exp(Lq[b][np.diag_indices(n)])
    mu = rng.randn(batch, n) * 0.3; eps = rng.randn(batch, S, n); zbar = rng.randn(batch, S, n)
    c = 1.0 / S
    ws = WS()
    z = torch.empty(batch, S, n, device="cuda"); kl = torch.zeros(1, device="cuda")
    Lq_d, mu_d, eps_d = dev(Lq), dev(mu), dev(eps)
    assert lib.hb_sample_tril_fwd(P(mu_d), P(Lq_d), n, batch, P(eps_d), S, P(z), P(kl), P(ws), ws.numel(), ST()) == 0
    tL, tm = (torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in (Lq, mu))
    te = torch.tensor(eps)
    z_ref = O.sample_fullrank(tm[:, None, :], tL[:, None, :, :], te)          # [batch,S,n]
    kl_ref = sum(O.kl_normal(O.logdet_fullrank(tL[b]), te[b], z_ref[b]) for b in range(batch))
    torch.cuda.synchronize()
    assert np.allclose(z.cpu().numpy(), z_ref.detach().numpy(), rtol=1e-5, atol=1e-5)
    assert abs(kl.item() - kl_ref.item()) <= 2e-5 * max(1.0, abs(kl_ref.item()))
    (torch.sum(z_ref * torch.tensor(zbar)) - c * kl_ref).backward()
    gmu = torch.zeros(batch, n, device="cuda"); gL = torch.full((batch, n, n), 7.0, device="cuda")
    scratch = torch.empty(batch, S, n, device="cuda")
    assert lib.hb_sample_tril_bwd(P(Lq_d), n, batch, P(eps_d), P(z), S, P(dev(zbar)), c, P(gmu), P(gL), P(scratch), ST()) == 0
    torch.cuda.synchronize()
    assert rel_err(gmu.cpu().numpy(), tm.grad.numpy()) < 1e-5
    assert rel_err(gL.cpu().numpy(), tL.grad.numpy()) < 1e-5          # includes zeros above the diagonal


def test_gaussian_density(lib):
    rng = np.random.RandomState(0)
    S, n = 5, 333
    f = rng.randn(S, n); y = rng.randn(n); var = np.array([0.37]); a = np.array([1.3])
    out = torch.empty(S, n, device="cuda")
    assert lib.hb_gaussian_logpdf(P(dev(y)), n, P(dev(f)), S * n, P(dev(var)), 1, S * n, P(out), ST()) == 0
    ref = O.gaussian(torch.tensor(y), torch.tensor(f), torch.tensor(var)).numpy()
    torch.cuda.synchronize()
    assert np.allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    ws = WS()
    resid = torch.empty(S, n, device="cuda"); out3 = torch.zeros(3, device="cuda")
    assert lib.hb_gauss_loglik_fwd(P(dev(f)), P(dev(a)), P(dev(y)), S * n, n, P(dev(var)), 1.0 / S, P(resid), P(out3),
                                   P(ws), ws.numel(), ST()) == 0
    tf = torch.tensor(f, requires_grad=True)
    ll = torch.sum(O.gaussian(torch.tensor(y), a[0] * tf, torch.tensor(var)))
    (ll / S).backward()
    torch.cuda.synchronize()
    o = out3.cpu().numpy()
    assert abs(o[0] - ll.item()) < 1e-5 * abs(ll.item())
    assert rel_err(resid.cpu().numpy() * a[0], tf.grad.numpy()) < 1e-5


@pytest.mark.parametrize("n,n2,D,n_ell,batch", [(5, 5, 2, 1, 1), (5, 6, 2, 2, 1), (5, 6, 2, 2, 10), (200, 200, 8, 8, 1),
                                                (130, 70, 40, 1, 2), (257, 257, 1, 1, 1)])
@pytest.mark.parametrize("csym", [0, 1])
def test_rbf_gram_fwd_bwd(lib, n, n2, D, n_ell, batch, csym):
    rng = np.random.RandomState(0)
    ell = np.exp(rng.randn(n_ell) * 0.3)
    X = rng.randn(batch, n, D); same = (n == n2 and n_ell != 2)
    X2 = None if same else rng.randn(batch, n2, D)
    K = torch.full((batch, n, n2), -3.0, device="cuda")
    Xd = dev(X); X2d = None if same else dev(X2); elld = dev(ell)
    assert lib.hb_rbf_gram_fwd(P(Xd), P(X2d), n, n2, D, batch, P(elld), n_ell, P(K), n2, n * n2, 0.0, 0, csym, ST()) == 0
    tell = torch.tensor(ell, requires_grad=True)
    fn = O.csym_rbf_K if csym else O.rbf_K
    Kref = fn(torch.tensor(X), tell, None if same else torch.tensor(X2))
    torch.cuda.synchronize()
    assert np.allclose(K.cpu().numpy(), Kref.detach().numpy(), atol=2e-6)       # reference tests: atol 1e-4
    if D > 32:
        return
    G = rng.randn(batch, n, n2)
    torch.sum(Kref * torch.tensor(G)).backward()
    ws = WS()
    g = torch.zeros(n_ell, device="cuda")
    assert lib.hb_rbf_gram_bwd(P(dev(G)), n2, n * n2, P(Xd), P(X2d), n, n2, D, batch, P(elld), n_ell, 0, csym, None, P(g),
                               P(ws), ws.numel(), ST()) == 0
    torch.cuda.synchronize()
    assert rel_err(g.cpu().numpy(), tell.grad.numpy()) < 2e-5


def test_rbf_gram_lower_only_jitter_and_sym_bwd(lib):
    rng = np.random.RandomState(2)
    n, D = 300, 3
    X = rng.randn(n, D); ell = np.array([0.8, 1.1, 1.7])
    K = torch.full((n, n), -3.0, device="cuda")
    assert lib.hb_rbf_gram_fwd(P(dev(X)), None, n, n, D, 1, P(dev(ell)), D, P(K), n, 0, 1e-3, 1, 0, ST()) == 0
    Kref = O.rbf_K(torch.tensor(X), torch.tensor(ell)).numpy() + 1e-3 * np.eye(n)
    torch.cuda.synchronize()
    assert np.allclose(np.tril(K.cpu().numpy()), np.tril(Kref), atol=2e-6)
    Gs = rng.randn(n, n); Gs = Gs + Gs.T
    tell = torch.tensor(ell, requires_grad=True)
    torch.sum(O.rbf_K(torch.tensor(X), tell) * torch.tensor(Gs)).backward()
    ws = WS(); g = torch.zeros(D, device="cuda"); sc = dev(np.array([0.5]))
    assert lib.hb_rbf_gram_bwd(P(dev(np.tril(Gs))), n, 0, P(dev(X)), None, n, n, D, 1, P(dev(ell)), D, 1, 0, P(sc), P(g),
                               P(ws), ws.numel(), ST()) == 0
    torch.cuda.synchronize()
    assert rel_err(g.cpu().numpy(), 0.5 * tell.grad.numpy()) < 2e-5


def spd(rng, n, D=3):
    X = rng.randn(n, D)
    K = O.rbf_K(torch.tensor(X), torch.tensor([1.0])).numpy() + 1e-2 * np.eye(n)
    return K


@pytest.mark.parametrize("n", [1, 5, 100, 128, 129, 300, 640, 1000])
def test_potrf_and_bwd(lib, n):
    lib.hb_set_gemm_engine(1)
    rng = np.random.RandomState(n)
    K = spd(rng, n)
    A = dev(K)
    wsb = lib.hb_potrf_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.hb_potrf_lower(P(A), n, 0, n, 1, 1, P(ws), wsb, P(err), ST()) == 0
    torch.cuda.synchronize()
    assert err.item() == 0
    L = A.cpu().numpy().astype(np.float64)
    assert np.array_equal(np.triu(L, 1), np.zeros_like(L))
    Lref = np.linalg.cholesky(K)
    cond = np.linalg.cond(K)
    # fp32 factorisation error grows with cond(K); LAPACK spotrf on the same matrices is ~2e-2*cond*2^-24
    assert rel_err(L, Lref) < max(5e-6, 0.05 * cond * 2.0 ** -24)
    assert np.allclose(L @ L.T, K, atol=1e-5)           # reference test_kernels.py:184-198 uses atol 9e-4
    # reverse mode vs torch autograd (fp64)
    Lb = np.tril(rng.randn(n, n))
    tK = torch.tensor(K, requires_grad=True)
    torch.sum(torch.linalg.cholesky(tK) * torch.tensor(Lb)).backward()
    Gref = 0.5 * (tK.grad + tK.grad.T).numpy()
    G = dev(Lb + np.triu(rng.randn(n, n), 1))            # garbage above the diagonal must be ignored
    assert lib.hb_potrf_lower_bwd(P(A), n, 0, P(G), n, 0, n, 1, P(ws), wsb, ST()) == 0
    torch.cuda.synchronize()
    assert rel_err(np.tril(G.cpu().numpy()), np.tril(Gref)) < max(3e-5, 0.5 * cond * 2.0 ** -24)
    lib.hb_set_gemm_engine(0)


def test_potrf_batched_and_nonpd_flag(lib):
    rng = np.random.RandomState(0)
    Ks = np.stack([spd(rng, 5, 2) for _ in range(10)])
    A = dev(Ks)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.hb_potrf_lower(P(A), 5, 25, 5, 10, 1, None, 0, P(err), ST()) == 0
    torch.cuda.synchronize()
    for b in range(10):
        Lb = A[b].cpu().numpy().astype(np.float64)
        assert np.allclose(Lb @ Lb.T, Ks[b], atol=1e-5)     # test_kernels.py:205-226 (atol 1e-4)
    bad = np.eye(6); bad[3, 3] = -1.0
    B = dev(bad)
    assert lib.hb_potrf_lower(P(B), 6, 0, 6, 1, 0, None, 0, P(err), ST()) == 0
    torch.cuda.synchronize()
    assert err.item() == 4          # 1 + index of the failing pivot


@pytest.mark.parametrize("m,n,trans", [(7, 5, 1), (64, 300, 1), (64, 300, 0), (500, 129, 0)])
def test_trsm_right_lower(lib, m, n, trans):
    lib.hb_set_gemm_engine(1)
    rng = np.random.RandomState(0)
    L = np.linalg.cholesky(spd(rng, n)); Xm = rng.randn(m, n)
    wsb = lib.hb_trsm_workspace_bytes(m, n)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    Xd = dev(Xm)
    assert lib.hb_trsm_right_lower(P(dev(L)), n, P(Xd), n, m, n, trans, P(ws), wsb, ST()) == 0
    ref = Xm @ np.linalg.inv(L).T if trans else Xm @ np.linalg.inv(L)
    torch.cuda.synchronize()
    assert rel_err(Xd.cpu().numpy(), ref) < 2e-5
    lib.hb_set_gemm_engine(0)


def test_adam_tf1(lib):
    rng = np.random.RandomState(0)
    n = 1000
    th = rng.randn(n).astype(np.float32); m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    thd, md, vd = dev(th), dev(m), dev(v)
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    th64, m64, v64 = th.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    for t in range(1, 11):
        g = rng.randn(n).astype(np.float32)               # gradient of the objective; loss = -objective
        assert lib.hb_increment_i32(P(step), ST()) == 0
        assert lib.hb_adam_tf1(P(thd), P(dev(g)), P(md), P(vd), n, -1.0, 1e-3, 0.9, 0.999, 1e-8, P(step), 0, ST()) == 0
        th64, m64, v64 = O.adam_tf1_step(th64, -g.astype(np.float64), m64, v64, t)
    torch.cuda.synchronize()
    assert np.allclose(thd.cpu().numpy(), th64, rtol=1e-5, atol=1e-6)


def pack_gp(p, order=("q_mu", "q_sqrt", "scale", "lengthscales", "k_var", "var")):
    return np.concatenate([np.asarray(p[k], np.float64).ravel() for k in order])


@pytest.mark.parametrize("n,D,S,n_ell,full", [(100, 1, 10, 1, 1), (100, 1, 10, 1, 0), (333, 8, 16, 8, 0),
                                              (600, 8, 64, 1, 0), (260, 3, 4, 3, 1)])
def test_gp_elbo_step_matches_oracle(lib, n, D, S, n_ell, full):
    """ELBO and every gradient of the GaussianProcess.ipynb graph vs the fp64 oracle (autograd).
    Well-conditioned inputs for D>1; for the 1-D notebook-style case the tolerance is scaled by
    cond(K+jI) as SURVEY.md 7 (hard part 2) prescribes."""
    from henbun_b200._lib import GpConfig
    lib.hb_set_gemm_engine(1)
    rng = np.random.RandomState(0)
    if D == 1:
        X = np.linspace(0, 6, n)[:, None]; Y = np.sin(X[:, 0]) + 0.3 * rng.randn(n); jitter = 1e-3
    else:
        X = rng.randn(n, D); Y = np.sin(X.sum(1) / math.sqrt(D)) + 0.1 * rng.randn(n); jitter = 1e-5
    p = dict(q_mu=0.1 * rng.randn(n),
             q_sqrt=(0.1 * np.eye(n) + 1e-2 * np.tril(rng.randn(n, n))) if full else (-1.0 + 0.1 * rng.randn(n)),
             scale=np.array([0.54]), lengthscales=np.full(n_ell, 0.54 if D != 3 else -1.0), k_var=np.array([0.54]),
             var=np.array([-0.5]))
    U = rng.randn(S, n)
    qs = "fullrank" if full else "diagonal"
    val, g = O.value_and_grads(lambda pp, *a: O.gpr_elbo(pp, *a, q_shape=qs, jitter=jitter), p, X, Y, U)
    cfg = GpConfig(n, D, S, n_ell, full, jitter, 0, 0)
    npar = lib.hb_gp_param_count(C.byref(cfg))
    params = dev(pack_gp(p)); assert params.numel() == npar
    grads = torch.zeros(npar, device="cuda"); out4 = torch.zeros(4, device="cuda")
    wsb = lib.hb_gp_elbo_workspace_bytes(C.byref(cfg))
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = lib.hb_gp_elbo_step(C.byref(cfg), P(dev(X)), P(dev(Y)), P(params), P(dev(U)), P(grads), P(out4), P(ws), wsb,
                             P(err), ST())
    assert rc == 0
    torch.cuda.synchronize()
    assert err.item() == 0
    ell = float(O.log1pe_forward(torch.tensor(p["lengthscales"]))[0])
    K = O.rbf_K(torch.tensor(X), torch.tensor([ell])).numpy() + jitter * np.eye(n)
    tol = max(1e-5, 10 * np.linalg.cond(K) * 2.0 ** -24)     # = 1e-5 for the well-conditioned D>1 cases
    assert abs(out4[0].item() - val) <= tol * abs(val), (out4.cpu().numpy(), val)
    gref = pack_gp(g)
    got = grads.cpu().numpy().astype(np.float64)
    off = 0
    for name in ("q_mu", "q_sqrt", "scale", "lengthscales", "k_var", "var"):
        sz = np.asarray(p[name]).size
        e = rel_err(got[off:off + sz], gref[off:off + sz])
        assert e <= (tol if name not in ("lengthscales",) else 3 * tol), (name, e, got[off:off + 3], gref[off:off + 3])
        off += sz
    lib.hb_set_gemm_engine(0)


def test_options_travel_with_the_call(lib):
    """hb_options is a per-call argument (the library keeps no configuration): the same product with two different option
    structs back to back takes two different engines, and forcing the tensor-core engine on a shape it cannot take is an
    argument error for THAT call only."""
    import ctypes as C
    from henbun_b200 import _lib
    g = torch.Generator("cuda").manual_seed(0)
    M = N = K = 512
    A = torch.randn(M, K, device="cuda", generator=g); B = torch.randn(K, N, device="cuda", generator=g)
    ref = (A.double() @ B.double())
    outs = {}
    for name, engine in (("simt", 1), ("tc", 2), ("auto", 0)):
        o = _lib.Options(); lib.hb_options_init(C.byref(o)); o.gemm_engine = engine
        Cm = torch.zeros(M, N, device="cuda")
        rc = lib.hb_gemm_ws(P(A), K, 0, 0, 0, P(B), N, 0, 0, 0, P(Cm), N, 0, 0, M, N, K, 1, 1.0, 0.0, None, 0, 0, 0, -50.0, 50.0, None, 0,
                            ST(), C.byref(o))
        assert rc == 0
        outs[name] = Cm
        assert (torch.linalg.norm(Cm.double() - ref) / torch.linalg.norm(ref)).item() < 3e-6
    assert not torch.equal(outs["simt"], outs["tc"])              # different arithmetic paths really ran
    o = _lib.Options(); lib.hb_options_init(C.byref(o)); o.gemm_engine = 2
    A3 = torch.randn(8, 7, device="cuda", generator=g); B3 = torch.randn(7, 5, device="cuda", generator=g); C3 = torch.zeros(8, 5, device="cuda")
    bad = lib.hb_gemm_ws(P(A3), 7, 0, 0, 0, P(B3), 5, 0, 0, 0, P(C3), 5, 0, 0, 8, 5, 7, 1, 1.0, 0.0, None, 0, 0, 0, -50.0, 50.0, None, 0, ST(),
                         C.byref(o))
    ok = lib.hb_gemm_ws(P(A3), 7, 0, 0, 0, P(B3), 5, 0, 0, 0, P(C3), 5, 0, 0, 8, 5, 7, 1, 1.0, 0.0, None, 0, 0, 0, -50.0, 50.0, None, 0, ST(),
                        None)
    assert bad == 1 and ok == 0
