"""Probe: tcgen05 engine on every operand form the blocked Cholesky / reverse mode uses (transposes, triangular
masks, c_tri), accuracy vs fp64 and vs the SIMT engine; then potrf + reverse at a few sizes with timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from henbun_b200 import _lib
lib = _lib.load()
P, ST = _lib.ptr, _lib.stream


def mask(mat, mode, kind):
    # kind 'A': op(A) [M,K] (m,k): 1 k<=m, 2 k>=m, 3 k>m, 4 k<m ; kind 'B': op(B) [K,N] (k,n): 1 n<=k 2 n>=k 3 n>k 4 n<k
    if mode == 0: return mat
    r, c = np.indices(mat.shape)
    if kind == 'A': m, k = r, c; keep = [None, k <= m, k >= m, k > m, k < m][mode]
    else: k, n = r, c; keep = [None, n <= k, n >= k, n > k, n < k][mode]
    return np.where(keep, mat, 0.0)


def run(M, N, K, tA, tB, a_tri=0, b_tri=0, c_tri=0, alpha=1.0, beta=0.0, eng=2):
    rng = np.random.RandomState(M + 3 * N + 7 * K + tA + 2 * tB)
    A = rng.randn(*((K, M) if tA else (M, K))).astype(np.float32)
    B = rng.randn(*((N, K) if tB else (K, N))).astype(np.float32)
    C0 = rng.randn(M, N).astype(np.float32)
    opA = mask((A.T if tA else A).astype(np.float64), a_tri, 'A')
    opB = mask((B.T if tB else B).astype(np.float64), b_tri, 'B')
    ref = alpha * opA @ opB + beta * C0
    outs = []
    for engine in (eng, 1):
        Ad, Bd, Cd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(C0).cuda()
        wsb = lib.hb_gemm_tc_workspace_bytes(M, N, K)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        lib.hb_set_gemm_engine(engine); lib.hb_set_tc_option(int(os.environ.get('HB_TC_OPT', '0')))
        rc = lib.hb_gemm_ws(P(Ad), A.shape[1], 0, tA, a_tri, P(Bd), B.shape[1], 0, tB, b_tri, P(Cd), N, 0, c_tri, M, N, K, 1,
                            alpha, beta, None, 0, 0, 0, -50.0, 50.0, P(ws), wsb, ST())
        torch.cuda.synchronize()
        lib.hb_set_gemm_engine(0)
        out = Cd.cpu().numpy().astype(np.float64)
        r = ref
        if c_tri:
            iu = np.triu_indices(M, 1, N)
            assert np.array_equal(out[iu], C0.astype(np.float64)[iu]), "upper touched"
            out = np.tril(out); r = np.tril(ref)
        outs.append((rc, np.linalg.norm(out - r) / np.linalg.norm(r)))
    print(f"M={M} N={N} K={K} tA={tA} tB={tB} a_tri={a_tri} b_tri={b_tri} c_tri={c_tri} a={alpha} b={beta}: "
          f"tc rc={outs[0][0]} err={outs[0][1]:.2e} | simt err={outs[1][1]:.2e}", flush=True)
    return outs[0][1]


def chol(n, reps=1):
    rng = np.random.RandomState(n)
    X = rng.randn(n, 8)
    d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) if n <= 4096 else None
    if d2 is None:
        Xt = torch.from_numpy(X).cuda()
        d2t = torch.cdist(Xt, Xt) ** 2
        Kd64 = torch.exp(-0.5 * d2t / 0.25) + 1e-3 * torch.eye(n, device="cuda", dtype=torch.float64)
    else:
        Kd64 = torch.from_numpy(np.exp(-0.5 * d2 / 0.25) + 1e-3 * np.eye(n)).cuda()
    Lref = torch.linalg.cholesky(Kd64)
    Gbar64 = torch.tril(torch.from_numpy(rng.randn(n, n)).cuda())
    # reference reverse mode via autograd in fp64
    Kr = Kd64.clone().requires_grad_(True)
    Lr = torch.linalg.cholesky(Kr)
    (Lr * Gbar64).sum().backward()
    Kbar_ref = torch.tril(Kr.grad + Kr.grad.T) - torch.diag(torch.diagonal(Kr.grad))  # sym-full convention, lower
    res = {}
    for engine in (0, 1):
        lib.hb_set_gemm_engine(engine)
        wsb = lib.hb_potrf_workspace_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        err = torch.zeros(4, dtype=torch.int32, device="cuda")
        times = []
        for rep in range(reps + 1):
            A = Kd64.float().contiguous()
            G = Gbar64.float().contiguous()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            rc1 = lib.hb_potrf_lower(P(A), n, 0, n, 1, 1, P(ws), wsb, P(err), ST())
            e[1].record()
            rc2 = lib.hb_potrf_lower_bwd(P(A), n, 0, P(G), n, 0, n, 1, P(ws), wsb, ST())
            e[2].record()
            torch.cuda.synchronize()
            times.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
        eL = (torch.linalg.norm(A.double() - Lref) / torch.linalg.norm(Lref)).item()
        Gl = torch.tril(G.double())
        # our convention: dObj/dK full-symmetric, lower triangle valid
        Kg = Kr.grad
        ref_sym = torch.tril(0.5 * (Kg + Kg.T))
        eG = (torch.linalg.norm(Gl - ref_sym) / torch.linalg.norm(ref_sym)).item()
        res[engine] = (rc1, rc2, eL, eG, times[-1])
    lib.hb_set_gemm_engine(0)
    for engine, (rc1, rc2, eL, eG, t) in res.items():
        fl_f, fl_b = n ** 3 / 3.0, 2 * n ** 3 / 3.0
        print(f"chol n={n} engine={'auto' if engine == 0 else 'simt'} rc={rc1},{rc2} errL={eL:.2e} errKbar={eG:.2e} "
              f"fwd {t[0]:.2f} ms ({fl_f / t[0] / 1e9:.1f} TF/s) bwd {t[1]:.2f} ms ({fl_b / t[1] / 1e9:.1f} TF/s)", flush=True)


def inplace(m, k, trans):
    rng = np.random.RandomState(m + k)
    X = rng.randn(m, k + 8).astype(np.float32)
    D = np.tril(rng.randn(128, 128)).astype(np.float32)
    Xd, Dd = torch.from_numpy(X).cuda(), torch.from_numpy(D).cuda()
    lib.hb_set_gemm_engine(1)
    rc = lib.hb_gemm_ws(P(Xd), k + 8, 0, 0, 0, P(Dd), 128, 0, 1 if trans else 0, 0, P(Xd), k + 8, 0, 0, m, k, k, 1,
                        1.0, 0.0, None, 0, 0, 0, -50.0, 50.0, None, 0, ST())
    torch.cuda.synchronize(); lib.hb_set_gemm_engine(0)
    Dk = D[:k, :k].astype(np.float64)
    ref = X[:, :k].astype(np.float64) @ (Dk.T if trans else Dk)
    out = Xd.cpu().numpy()
    err = np.linalg.norm(out[:, :k] - ref) / np.linalg.norm(ref)
    pad_ok = np.array_equal(out[:, k:], X[:, k:])
    print(f"inplace m={m} k={k} trans={trans} rc={rc} err={err:.2e} pad_untouched={pad_ok}", flush=True)


if __name__ == "__main__":
    worst = 0.0
    for (m, k, t) in ((128, 128, 1), (1000, 128, 0), (5000, 96, 1), (64, 128, 0)):
        inplace(m, k, t)
    # short-K kernel (engine 1) against the k-looped kernel is covered by run(..., eng=1) vs fp64
    for (tA, tB) in ((0, 1), (0, 0), (1, 0), (1, 1)):
        run(200, 300, 128, tA, tB, eng=1); run(128, 128, 256, tA, tB, alpha=-1.0, beta=1.0, eng=1)
        run(2000, 128, 100, tA, tB, a_tri=0, eng=1)
    for tri in (1, 2, 3, 4):
        run(256, 256, 256, 0, 0, a_tri=tri, eng=1); run(256, 256, 256, 1, 1, b_tri=tri, eng=1)
    run(256, 256, 200, 1, 0, c_tri=1, alpha=-1.0, beta=1.0, eng=1)
    for (tA, tB) in ((0, 1), (0, 0), (1, 0), (1, 1)):
        worst = max(worst, run(256, 384, 200, tA, tB))
        worst = max(worst, run(300, 500, 100, tA, tB, alpha=-1.0, beta=1.0))
    for a_tri in (1, 2, 3, 4):
        worst = max(worst, run(512, 384, 512, 0, 0, a_tri=a_tri, alpha=-2.0, beta=1.0))
        worst = max(worst, run(512, 384, 512, 1, 0, a_tri=a_tri, alpha=-2.0, beta=1.0))
    for b_tri in (1, 2, 3, 4):
        worst = max(worst, run(384, 512, 512, 0, 1, b_tri=b_tri))
        worst = max(worst, run(384, 512, 512, 0, 0, b_tri=b_tri))
    worst = max(worst, run(640, 640, 384, 1, 0, c_tri=1, alpha=-1.0, beta=1.0))
    worst = max(worst, run(640, 640, 384, 0, 1, c_tri=1, alpha=-1.0, beta=1.0))
    worst = max(worst, run(1000, 1000, 1000, 1, 0, a_tri=3))
    print("worst tc err", worst, flush=True)
    for n in (128, 1024, 4096, 8192, 16384):
        chol(n)
