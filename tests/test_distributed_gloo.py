"""world_size-2 gloo test (CPU) of the multi-rank plumbing used by bench.py / Optimizer: sample-axis
sharding, disjoint Philox windows, and the single all-reduce of the packed gradient."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from henbun_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = parallel.shard_samples(7, world, rank)
    g = torch.full((10,), float(rank + 1)) * count           # "gradient summed over my samples"
    parallel.allreduce_sum_(g)
    w, r = parallel.world()
    out[rank] = (first, count, g.numpy().copy(), w, r)
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    world = 2
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (f0, c0, g0, w0, r0), (f1, c1, g1, w1, r1) = out[0], out[1]
    assert (f0, c0, f1, c1) == (0, 4, 4, 3) and (w0, r0, w1, r1) == (2, 0, 2, 1)
    assert np.allclose(g0, 1 * 4 + 2 * 3) and np.allclose(g0, g1)      # every rank holds the same summed gradient


def test_philox_windows_are_disjoint_and_cover_the_stream():
    S, per = 64, 1000
    for world in (1, 2, 4, 8):
        spans = []
        for r in range(world):
            first, cnt = parallel.shard_samples(S, world, r)
            off = parallel.rank_philox_offset(3, first, per, S)
            assert off % 4 == 0
            spans.append((off, off + cnt * per))
        spans.sort()
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 <= b0
        assert spans[0][0] == 3 * S * per and spans[-1][1] == 4 * S * per


# ---- config 5 (linear operator): row-sharded A, ONE all-reduce of [S*n + 4] floats -------------------
def _linop_partial(A, y, p, U):
    """What hb_linop_elbo_local computes on one rank's rows (oracle arithmetic, fp64): the partial
    Zbar = R A with R = (1/S) d loglik / d F, and {loglik, sum E^2}."""
    from oracle import henbun_oracle as O
    S = U.shape[0]
    Z = p["q_mu"] + U @ torch.tril(p["q_sqrt"]).T
    F = Z @ A.T
    var = O.log1pe_forward(p["var"])
    E = F - y
    ll = torch.sum(O.gaussian(y, F, var))
    R = -(E / var) / S
    return torch.cat([(R @ A).reshape(-1), torch.stack([ll, torch.sum(E * E), torch.zeros(()), torch.zeros(())]).to(A.dtype)])


def _linop_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(0)                       # every rank builds the same problem, keeps its rows
    M, n, S = 37, 6, 3
    A = torch.tensor(rng.randn(M, n)); y = torch.tensor(rng.randn(M))
    p = {"q_mu": torch.tensor(rng.randn(n)), "q_sqrt": torch.tensor(np.eye(n) + 0.1 * rng.randn(n, n)), "var": torch.tensor([0.2], dtype=torch.float64)}
    U = torch.tensor(rng.randn(S, n))
    first, cnt = parallel.shard_rows(M, world, rank)
    buf = _linop_partial(A[first:first + cnt], y[first:first + cnt], p, U)
    parallel.allreduce_sum_(buf)
    out[rank] = (first, cnt, buf.numpy().copy())
    dist.destroy_process_group()


def test_linop_row_shards_sum_to_the_whole_operator():
    from oracle import henbun_oracle as O
    world = 2
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_linop_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (f0, c0, b0), (f1, c1, b1) = out[0], out[1]
    assert (f0, c0, f1, c1) == (0, 19, 19, 18)
    assert np.array_equal(b0, b1)
    rng = np.random.RandomState(0)
    M, n, S = 37, 6, 3
    A = rng.randn(M, n); y = rng.randn(M)
    p = {"q_mu": rng.randn(n), "q_sqrt": np.eye(n) + 0.1 * rng.randn(n, n), "var": np.array([0.2])}
    U = rng.randn(S, n)
    whole = _linop_partial(torch.tensor(A), torch.tensor(y), {k: torch.tensor(v) for k, v in p.items()}, torch.tensor(U)).numpy()
    assert np.allclose(b0, whole, rtol=1e-12, atol=1e-12)
    # and the update stage's algebra: mu-bar = colsum(Zbar - Z/S) reproduces the oracle's autograd gradient
    ref, g = O.value_and_grads(O.linear_operator_elbo, p, A, y, U)
    Z = p["q_mu"] + U @ np.tril(p["q_sqrt"]).T
    Zt = b0[:S * n].reshape(S, n) - Z / S
    assert np.allclose(Zt.sum(0), g["q_mu"], rtol=1e-10)
    gL = np.tril(Zt.T @ U) + np.diag(1.0 / np.diag(p["q_sqrt"]))
    assert np.allclose(gL, g["q_sqrt"], rtol=1e-10, atol=1e-12)


# ---- column-block-cyclic Cholesky + reverse mode: the algebra of the multi-rank path, on CPU -------------------------------
def _bc_problem(n):
    g = torch.Generator().manual_seed(n)
    X = torch.randn(n, 3, dtype=torch.float64, generator=g)
    K = torch.exp(-0.5 * torch.cdist(X, X) ** 2) + 1e-2 * torch.eye(n, dtype=torch.float64)
    Lbar = torch.tril(torch.randn(n, n, dtype=torch.float64, generator=g))
    return K, Lbar


def _bc_worker(rank, world, port, out, n, block, turn):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.block_cyclic import potrf_block_cyclic, chol_rev_block_cyclic
    K, Lbar = _bc_problem(n)
    P = (n + block - 1) // block
    mine = [b for b in range(P) if (b // turn) % world == rank]
    # a rank starts from its OWN columns only: poison everything else, the broadcasts must fill it in
    A = torch.full_like(K, float("nan")); G = torch.full_like(K, float("nan"))
    for b in mine:
        A[:, b * block:(b + 1) * block] = K[:, b * block:(b + 1) * block]
        G[:, b * block:(b + 1) * block] = Lbar[:, b * block:(b + 1) * block]
    L = torch.tril(potrf_block_cyclic(A, block, rank, world, turn))
    Kbar = torch.tril(chol_rev_block_cyclic(L, G, block, rank, world, turn))
    out[rank] = (L.numpy().copy(), Kbar.numpy().copy())
    dist.destroy_process_group()


def test_block_cyclic_cholesky_and_reverse_mode_world2():
    """oracle/block_cyclic.py (the restatement of csrc/linalg.cu's potrf_flat / chol_rev_flat / rev_block) on two gloo ranks:
    each starts from its own column blocks only and both end with the complete factor and the complete gradient, equal to
    LAPACK's Cholesky and torch autograd's gradient of it, for plain and `turn`-wise ownership and a ragged last block."""
    for n, block, turn in ((96, 16, 1), (100, 16, 2), (70, 32, 1)):
        K, Lbar = _bc_problem(n)
        Kr = K.clone().requires_grad_(True)
        Lref = torch.linalg.cholesky(Kr)
        (Lref * Lbar).sum().backward()
        Gref = torch.tril(0.5 * (Kr.grad + Kr.grad.T)).numpy()
        mgr = mp.Manager(); out = mgr.dict()
        mp.spawn(_bc_worker, args=(2, _free_port(), out, n, block, turn), nprocs=2, join=True)
        for r in (0, 1):
            L, Kbar = out[r]
            assert np.isfinite(L).all() and np.isfinite(Kbar).all()
            assert np.allclose(L, Lref.detach().numpy(), rtol=1e-10, atol=1e-12), (n, block, turn, r)
            assert np.allclose(Kbar, Gref, rtol=1e-8, atol=1e-10), (n, block, turn, r)
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])     # bit-identical on both ranks
