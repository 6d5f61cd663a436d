"""GPU parity tests of the tcgen05 3xTF32 engine (generation 2: single-CTA, CTA-pair and split-K paths), the short-K
SIMT kernel and the column-recursive factorisations that run on them.  Oracle: numpy fp64 on the same inputs.
Tolerance: 3e-6 relative (Frobenius) per product -- fp32-grade; the k-looped fp32 SIMT kernel measures 2e-7 .. 1.2e-6
on the same shapes (accumulation order), the tensor-core engine 9e-7 .. 1.1e-6."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from henbun_b200 import _lib
    return _lib.load()


def P(t):
    from henbun_b200._lib import ptr
    return ptr(t)


def ST():
    from henbun_b200._lib import stream
    return stream()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def mask(mat, mode, kind):
    if mode == 0:
        return mat
    r, c = np.indices(mat.shape)
    if kind == "A":
        keep = [None, c <= r, c >= r, c > r, c < r][mode]
    else:   # op(B)[k, n]: 1 n<=k, 2 n>=k, 3 n>k, 4 n<k
        keep = [None, c <= r, c >= r, c > r, c < r][mode]
    return np.where(keep, mat, 0.0)


def run(lib, M, N, K, tA, tB, a_tri=0, b_tri=0, c_tri=0, alpha=1.0, beta=0.0, engine=2, with_ws=False, seed=0):
    rng = np.random.RandomState(seed + M + 3 * N + 7 * K)
    A = rng.randn(*((K, M) if tA else (M, K))).astype(np.float32)
    B = rng.randn(*((N, K) if tB else (K, N))).astype(np.float32)
    C0 = rng.randn(M, N).astype(np.float32)
    opA = mask((A.T if tA else A).astype(np.float64), a_tri, "A")
    opB = mask((B.T if tB else B).astype(np.float64), b_tri, "B")
    ref = alpha * opA @ opB + beta * C0
    Ad, Bd, Cd = dev(A), dev(B), dev(C0)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda") if with_ws else None
    lib.hb_set_gemm_engine(engine)
    try:
        rc = lib.hb_gemm_ws(P(Ad), A.shape[1], 0, tA, a_tri, P(Bd), B.shape[1], 0, tB, b_tri, P(Cd), N, 0, c_tri, M, N, K, 1,
                            alpha, beta, None, 0, 0, 0, -50.0, 50.0, P(ws) if with_ws else None, (64 << 20) if with_ws else 0,
                            ST())
        torch.cuda.synchronize()
    finally:
        lib.hb_set_gemm_engine(0)
    assert rc == 0, rc
    out = Cd.cpu().numpy().astype(np.float64)
    if c_tri:
        iu = np.triu_indices(M, 1, N)
        assert np.array_equal(out[iu], C0.astype(np.float64)[iu]), "entries above the diagonal were written"
        out = np.tril(out); ref = np.tril(ref)
    return rel_err(out, ref)


LAYOUTS = [(0, 1), (0, 0), (1, 0), (1, 1)]


@pytest.mark.parametrize("tA,tB", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(128 * 300, 128, 128), (128 * 149 + 4, 100, 272), (40000, 64, 1000), (128 * 200, 128, 16)])
def test_tc_persistent_tile_walk(lib, M, N, K, tA, tB):
    """More than 148 row tiles of a 128- (or 64-) wide short-K product: one persistent CTA per SM walks over the tiles
    (ring, barrier phases and TMEM buffers run through; separate epilogue staging)."""
    assert run(lib, M, N, K, tA, tB, alpha=-1.0, beta=1.0) < 3e-6
    assert run(lib, M, N, K, tA, tB, alpha=0.5, beta=0.0, seed=3) < 3e-6


@pytest.mark.parametrize("tri", [1, 2, 3, 4])
def test_tc_persistent_tile_walk_masks(lib, tri):
    assert run(lib, 128 * 170, 128, 256, 0, 0, b_tri=tri, alpha=-2.0, beta=1.0) < 3e-6
    assert run(lib, 128 * 170, 128, 384, 0, 1, c_tri=1, alpha=-1.0, beta=1.0) < 3e-6


@pytest.mark.parametrize("tA,tB", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 384, 200), (300, 500, 100), (1000, 128, 336), (64, 1024, 512)])
def test_tc_single_cta_layouts(lib, M, N, K, tA, tB):
    assert run(lib, M, N, K, tA, tB, alpha=-1.0, beta=1.0) < 3e-6


@pytest.mark.parametrize("tA,tB", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(4096, 64, 512), (300, 64, 100), (1000, 48, 336), (128, 16, 64)])
def test_tc_64_wide_tiles(lib, M, N, K, tA, tB):
    """N <= 64 runs the 128 x 64 tile instantiation (config 5 with the operator as the M operand)."""
    assert run(lib, M, N, K, tA, tB, alpha=-1.0, beta=1.0) < 3e-6


def test_tc_64_wide_tiles_split_k_and_masks(lib):
    assert run(lib, 256, 64, 8192, 1, 0, alpha=0.5, beta=0.0, with_ws=True) < 3e-6      # long K, split through the scratch
    assert run(lib, 16384, 64, 4096, 1, 0, with_ws=True) < 3e-6                         # the Zbar^T = A^T R^T shape class
    assert run(lib, 512, 64, 512, 0, 0, a_tri=1) < 3e-6
    assert run(lib, 512, 64, 512, 0, 1, b_tri=2) < 3e-6


@pytest.mark.parametrize("tA,tB", LAYOUTS)
def test_tc_cta_pair_layouts(lib, tA, tB):
    # >= 64 tiles of 256 x 256 -> cta_group::2 kernel; ragged edges in M, N and K
    assert run(lib, 2100, 2180, 520, tA, tB, alpha=-1.0, beta=1.0) < 3e-6


@pytest.mark.parametrize("size", [512, 2048])
@pytest.mark.parametrize("tri", [1, 2, 3, 4])
def test_tc_triangular_operands(lib, size, tri):
    n = size
    assert run(lib, n, n - 128, n, 0, 0, a_tri=tri, alpha=-2.0, beta=1.0) < 3e-6
    assert run(lib, n, n - 128, n, 1, 0, a_tri=tri) < 3e-6
    assert run(lib, n - 128, n, n, 0, 1, b_tri=tri) < 3e-6
    assert run(lib, n - 128, n, n, 0, 0, b_tri=tri) < 3e-6


@pytest.mark.parametrize("M,K", [(640, 384), (3000, 700)])
def test_tc_lower_triangular_output(lib, M, K):
    assert run(lib, M, M, K, 1, 0, c_tri=1, alpha=-1.0, beta=1.0) < 3e-6
    assert run(lib, M, M, K, 0, 1, c_tri=1, alpha=-1.0, beta=1.0) < 3e-6


@pytest.mark.parametrize("M,N,K", [(128, 128, 4096), (256, 128, 9000), (512, 512, 2048)])
def test_tc_split_k(lib, M, N, K):
    # long-K reduction into a small tile: split over K with a deterministic second pass
    assert run(lib, M, N, K, 1, 0, alpha=-2.0, beta=1.0, with_ws=True) < 3e-6
    assert run(lib, M, N, K, 1, 0, c_tri=1 if M == N else 0, alpha=-2.0, beta=1.0, with_ws=True) < 3e-6
    a = run(lib, M, N, K, 0, 1, with_ws=True, seed=5)
    b = run(lib, M, N, K, 0, 1, with_ws=True, seed=5)
    assert a == b                                           # deterministic


@pytest.mark.parametrize("engine", [0, 1, 2])
@pytest.mark.parametrize("m,k,trans", [(128, 128, 1), (1000, 128, 0), (5000, 96, 1), (20000, 128, 0), (60000, 128, 1), (33333, 64, 0)])
def test_in_place_panel_solve(lib, engine, m, k, trans):
    """X <- X * D (or D^T) with C aliasing A: the leaf step of the blocked triangular solves."""
    if engine == 2 and (k % 4 or (k + 8) % 4):
        pytest.skip("TMA needs 16-byte rows")
    rng = np.random.RandomState(m + k)
    X = rng.randn(m, k + 8).astype(np.float32)
    D = np.tril(rng.randn(128, 128)).astype(np.float32)
    Xd, Dd = dev(X), dev(D)
    lib.hb_set_gemm_engine(engine)
    try:
        rc = lib.hb_gemm_ws(P(Xd), k + 8, 0, 0, 0, P(Dd), 128, 0, 1 if trans else 0, 0, P(Xd), k + 8, 0, 0, m, k, k, 1,
                            1.0, 0.0, None, 0, 0, 0, -50.0, 50.0, None, 0, ST())
        torch.cuda.synchronize()
    finally:
        lib.hb_set_gemm_engine(0)
    assert rc == 0
    Dk = D[:k, :k].astype(np.float64)
    ref = X[:, :k].astype(np.float64) @ (Dk.T if trans else Dk)
    out = Xd.cpu().numpy()
    assert rel_err(out[:, :k], ref) < 3e-6
    assert np.array_equal(out[:, k:], X[:, k:])             # padding columns untouched


@pytest.mark.parametrize("tA,tB", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(1, 1, 4), (200, 300, 128), (128, 128, 256), (2000, 128, 100), (77, 130, 36)])
def test_short_k_kernel_layouts(lib, M, N, K, tA, tB):
    assert run(lib, M, N, K, tA, tB, alpha=0.7, beta=-0.3, engine=1) < 2e-6


@pytest.mark.parametrize("tri", [1, 2, 3, 4])
def test_short_k_kernel_masks(lib, tri):
    assert run(lib, 256, 256, 256, 0, 0, a_tri=tri, engine=1) < 2e-6
    assert run(lib, 256, 256, 256, 1, 1, b_tri=tri, engine=1) < 2e-6
    assert run(lib, 256, 256, 200, 1, 0, c_tri=1, alpha=-1.0, beta=1.0, engine=1) < 2e-6


@pytest.mark.parametrize("n", [1000, 2176, 3000, 4224, 8200, 12288])
def test_potrf_and_reverse_mode_on_tensor_cores(lib, n):
    """Column-recursive Cholesky + reverse mode with the automatic engine choice (tensor cores for the large
    products; from n = 4096 the big ones on the pre-split fp16 hi/lo engine, 8200 = ragged last block) against fp64
    LAPACK / torch autograd on the GPU."""
    g = torch.Generator("cuda").manual_seed(n)
    X = torch.randn(n, 8, device="cuda", generator=g, dtype=torch.float64)
    K64 = torch.exp(-0.5 * torch.cdist(X, X) ** 2 / 0.25) + 1e-3 * torch.eye(n, device="cuda", dtype=torch.float64)
    Lbar = torch.tril(torch.randn(n, n, device="cuda", generator=g, dtype=torch.float64))
    Kr = K64.clone().requires_grad_(True)
    Lref = torch.linalg.cholesky(Kr)
    (Lref * Lbar).sum().backward()
    Gref = torch.tril(0.5 * (Kr.grad + Kr.grad.T))
    A = K64.float().contiguous()
    G = (Lbar + torch.triu(torch.randn(n, n, device="cuda", generator=g, dtype=torch.float64), 1)).float().contiguous()
    wsb = lib.hb_potrf_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.hb_potrf_lower(P(A), n, 0, n, 1, 1, P(ws), wsb, P(err), ST()) == 0
    assert lib.hb_potrf_lower_bwd(P(A), n, 0, P(G), n, 0, n, 1, P(ws), wsb, ST()) == 0
    torch.cuda.synchronize()
    assert err.item() == 0
    eL = (torch.linalg.norm(A.double() - Lref.detach()) / torch.linalg.norm(Lref.detach())).item()
    eG = (torch.linalg.norm(torch.tril(G.double()) - Gref) / torch.linalg.norm(Gref)).item()
    assert torch.count_nonzero(torch.triu(A, 1)).item() == 0
    assert eL < 2e-6, eL                                    # measured 5e-8 .. 1.3e-7
    assert eG < 5e-6, eG                                    # measured 2e-7 .. 1e-6


@pytest.mark.parametrize("mode", [0, 1, 3])
@pytest.mark.parametrize("n", [300, 1000, 2500])
def test_panel_solve_modes(lib, n, mode):
    """hb_set_panel_refinement: explicit-inverse panels (0), refined (1) and the substitution kernel (3) all reproduce the
    fp64 factor, its reverse mode and the right-sided triangular solves (n not a multiple of the 128-wide leaf)."""
    g = torch.Generator("cuda").manual_seed(n + mode)
    X = torch.randn(n, 8, device="cuda", generator=g, dtype=torch.float64)
    K64 = torch.exp(-0.5 * torch.cdist(X, X) ** 2 / 0.25) + 1e-3 * torch.eye(n, device="cuda", dtype=torch.float64)
    Lbar = torch.tril(torch.randn(n, n, device="cuda", generator=g, dtype=torch.float64))
    Kr = K64.clone().requires_grad_(True)
    Lref = torch.linalg.cholesky(Kr)
    (Lref * Lbar).sum().backward()
    Gref = torch.tril(0.5 * (Kr.grad + Kr.grad.T))
    A = K64.float().contiguous(); G = Lbar.float().contiguous()
    wsb = lib.hb_potrf_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    try:
        assert lib.hb_set_panel_refinement(mode) == mode
        assert lib.hb_potrf_lower(P(A), n, 0, n, 1, 1, P(ws), wsb, P(err), ST()) == 0
        assert lib.hb_potrf_lower_bwd(P(A), n, 0, P(G), n, 0, n, 1, P(ws), wsb, ST()) == 0
        m = 77
        B = torch.randn(m, n, device="cuda", generator=g, dtype=torch.float64)
        tws = lib.hb_trsm_workspace_bytes(m, n)
        tw = torch.empty(tws, dtype=torch.uint8, device="cuda")
        sols = []
        for trans in (1, 0):
            Xs = B.float().contiguous()
            assert lib.hb_trsm_right_lower(P(A), n, P(Xs), n, m, n, trans, P(tw), tws, ST()) == 0
            sols.append(Xs)
        torch.cuda.synchronize()
    finally:
        lib.hb_set_panel_refinement(2)
    assert err.item() == 0
    L = Lref.detach()
    assert (torch.linalg.norm(A.double() - L) / torch.linalg.norm(L)).item() < 2e-6
    assert (torch.linalg.norm(torch.tril(G.double()) - Gref) / torch.linalg.norm(Gref)).item() < 5e-6
    ref_t = torch.linalg.solve_triangular(L, B.T, upper=False).T            # B L^{-T}
    ref_n = torch.linalg.solve_triangular(L.T, B.T, upper=True).T           # B L^{-1}
    assert (torch.linalg.norm(sols[0].double() - ref_t) / torch.linalg.norm(ref_t)).item() < 5e-6
    assert (torch.linalg.norm(sols[1].double() - ref_n) / torch.linalg.norm(ref_n)).item() < 5e-6


def test_gp_step_tensor_cores_match_simt(lib):
    """Fused ELBO + gradient step at a size where the tensor-core engine takes the large products, against the same
    step on the fp32 SIMT kernels (same eps stream): ELBO to 1e-5 relative (north-star tolerance), gradients normwise."""
    import ctypes as C
    from henbun_b200 import _lib
    from oracle import cpu_baseline as cb
    n, D, S = 2560, 8, 16
    X, Y, p = cb.make_gp_problem(n, D, S, seed=0)
    order = ("q_mu", "q_sqrt", "scale", "lengthscales", "k_var", "var")
    params = torch.from_numpy(np.concatenate([np.asarray(p[k], np.float32).ravel() for k in order])).cuda()
    cfg = _lib.GpConfig(n, D, S, 1, 0, 1e-5, 1234, 0)
    npar = lib.hb_gp_param_count(C.byref(cfg))
    wsb = lib.hb_gp_elbo_workspace_bytes(C.byref(cfg))
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    res = {}
    for engine in (1, 0):
        lib.hb_set_gemm_engine(engine)
        grads = torch.zeros(npar, device="cuda"); out4 = torch.zeros(4, device="cuda")
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        rc = lib.hb_gp_elbo_step(C.byref(cfg), P(Xd), P(Yd), P(params), None, P(grads), P(out4), P(ws), wsb, P(err), ST())
        torch.cuda.synchronize()
        lib.hb_set_gemm_engine(0)
        assert rc == 0 and err.item() == 0
        res[engine] = (out4[0].item(), grads.double().cpu().numpy())
    e1, g1 = res[1]; e0, g0 = res[0]
    assert abs(e0 - e1) <= 1e-5 * abs(e1)
    assert np.linalg.norm(g0 - g1) <= 2e-4 * np.linalg.norm(g1)


def test_full_size_factorisation_properties(lib):
    """BASELINE config-3 size (N = 65536, D = 8, lengthscale 0.5, jitter 1e-5): size-independent properties of the
    factorisation and its reverse mode, since no fp64 oracle finishes at this size.
      * round trip: L (L^T V) == K V for 64 random probe vectors (relative, fp32-grade)
      * reverse mode is deterministic (bitwise, split-K included) and exactly homogeneous under a power-of-two scaling
      * the lower triangle of the result is finite and the strict upper triangle of L is untouched (zeroed on request)."""
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    n, D, S = 65536, 8, 64
    if free < 120 * (1 << 30):
        pytest.skip("needs ~100 GB of free HBM")
    g = torch.Generator("cuda").manual_seed(0)
    X = torch.randn(n, D, device="cuda", generator=g)
    ell = torch.tensor([0.5], device="cuda")
    K = torch.empty(n, n, device="cuda")
    assert lib.hb_rbf_gram_fwd(P(X), None, n, n, D, 1, P(ell), 1, P(K), n, 0, 1e-5, 0, 0, ST()) == 0
    V = torch.randn(S, n, device="cuda", generator=g)            # probes, sample-major like GP.samples
    KV = torch.empty(S, n, device="cuda")

    def gemm(A, B, Cm, M, N, Kd, tB, b_tri):
        rc = lib.hb_gemm_ws(P(A), Kd, 0, 0, 0, P(B), n, 0, tB, b_tri, P(Cm), N, 0, 0, M, N, Kd, 1, 1.0, 0.0, None, 0, 0, 0,
                            -50.0, 50.0, None, 0, ST())
        assert rc == 0
    gemm(V, K, KV, S, n, n, 1, 0)                                  # V K^T = (K V^T)^T, K symmetric (full matrix built)
    wsb = lib.hb_potrf_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.hb_potrf_lower(P(K), n, 0, n, 1, 1, P(ws), wsb, P(err), ST()) == 0      # K <- L, upper zeroed
    torch.cuda.synchronize()
    assert err.item() == 0
    T = torch.empty(S, n, device="cuda"); LLV = torch.empty(S, n, device="cuda")
    gemm(V, K, T, S, n, n, 0, 1)                                   # T = V L   (op(B)[k][j] = L[k][j], keep j <= k)
    gemm(T, K, LLV, S, n, n, 1, 2)                                 # LLV = T L^T (op(B)[k][j] = L[j][k], keep k <= j)
    torch.cuda.synchronize()
    rel = (torch.linalg.norm(LLV.double() - KV.double()) / torch.linalg.norm(KV.double())).item()
    assert rel < 2e-5, rel
    assert torch.count_nonzero(K[:4096, 4096:8192]).item() == 0   # a strictly-upper block stays zero

    del V, KV, T, LLV
    G1 = torch.randn(n, n, device="cuda", generator=g).tril_()
    G2 = G1 * 4.0
    assert lib.hb_potrf_lower_bwd(P(K), n, 0, P(G1), n, 0, n, 1, P(ws), wsb, ST()) == 0
    assert lib.hb_potrf_lower_bwd(P(K), n, 0, P(G2), n, 0, n, 1, P(ws), wsb, ST()) == 0
    torch.cuda.synchronize()
    for r0, c1 in ((60000, 60000), (1024, 1024), (33000, 20000)):                       # rows r0.., columns < c1 <= r0: lower part
        a = G1[r0:r0 + 1024, :c1]; b = G2[r0:r0 + 1024, :c1]
        assert torch.isfinite(a).all()
        assert torch.equal(b * 0.25, a)                                                 # exact homogeneity
    # determinism (split-K reductions included): same input, same bits
    G2.copy_(G1)            # reuse the buffers: G2 <- K-bar(G1) is not an L-bar, so rebuild the input instead
    del G2
    g = torch.Generator("cuda").manual_seed(0)
    Xr = torch.randn(n, D, device="cuda", generator=g); Vr = torch.randn(S, n, device="cuda", generator=g); del Xr, Vr
    G3 = torch.randn(n, n, device="cuda", generator=g).tril_()                       # the same draw as G1's input
    assert lib.hb_potrf_lower_bwd(P(K), n, 0, P(G3), n, 0, n, 1, P(ws), wsb, ST()) == 0
    torch.cuda.synchronize()
    assert torch.equal(G3[60000:61024, :60000], G1[60000:61024, :60000])
