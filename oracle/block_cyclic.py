"""TEST INFRASTRUCTURE (only tests/ may import this): a CPU restatement, in plain torch fp64, of the column-block-cyclic
right-looking Cholesky and of its reverse mode as henbun_b200/csrc/linalg.cu schedules them over ranks (potrf_flat /
chol_rev_flat / rev_block / rev_update), with torch.distributed broadcasts for the panels.  It states the ALGEBRA of the
multi-rank path -- who owns which block, what a finished panel carries, which three products take a K-bar panel into the
columns to its left, the full-symmetric convention of the result -- so that it can be checked against LAPACK + autograd on
CPU (gloo, world 2) where no GPU exists.  The op it distributes is tf.cholesky and its gradient
(reference: Henbun/gp/kernels.py:100-101, reached through Optimizer.compile, Henbun/model.py:220).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .henbun_oracle import chol_rev_recursive


def _owner(b, turn, world):
    return (b // turn) % world


def _bcast(panel, src):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        panel = panel.contiguous()
        dist.broadcast(panel, src=src)
    return panel


def potrf_block_cyclic(A: torch.Tensor, block: int, rank: int, world: int, turn: int = 1) -> torch.Tensor:
    """In place on this rank's copy of the matrix (lower triangle; only the columns this rank owns have to be valid on
    entry).  On exit every rank holds the complete factor."""
    n = A.shape[0]
    P = (n + block - 1) // block
    for p in range(P):
        c0, c1 = p * block, min(n, (p + 1) * block)
        if _owner(p, turn, world) == rank:
            D = torch.linalg.cholesky(A[c0:c1, c0:c1])
            A[c0:c1, c0:c1] = D
            if c1 < n:     # panel <- panel D^{-T}
                A[c1:, c0:c1] = torch.linalg.solve_triangular(D, A[c1:, c0:c1].T, upper=False).T
        A[c0:, c0:c1] = _bcast(A[c0:, c0:c1].clone(), _owner(p, turn, world))
        for j in range(p + 1, P):                         # U(p, j): the blocks to the right that this rank owns
            if _owner(j, turn, world) != rank:
                continue
            j0, j1 = j * block, min(n, (j + 1) * block)
            A[j0:, j0:j1] -= A[j0:, c0:c1] @ A[j0:j1, c0:c1].T
    return A


def chol_rev_block_cyclic(L: torch.Tensor, G: torch.Tensor, block: int, rank: int, world: int, turn: int = 1) -> torch.Tensor:
    """G (lower triangle) holds dObj/dL in the columns this rank owns; on exit every rank holds dObj/dK in the
    full-symmetric convention (an off-diagonal entry holds half of the lower-triangle gradient)."""
    n = L.shape[0]
    P = (n + block - 1) // block
    for p in range(P - 1, -1, -1):
        c0, c1 = p * block, min(n, (p + 1) * block)
        if _owner(p, turn, world) == rank:                # rev_block: rows below the block first, the diagonal block last
            LD = L[c0:c1, c0:c1]
            if c1 < n:
                Y = 0.5 * torch.linalg.solve_triangular(LD, G[c1:, c0:c1], upper=False, left=False)   # G_R L_DD^{-1} / 2
                G[c0:c1, c0:c1] -= 2.0 * torch.tril(Y.T @ L[c1:, c0:c1])
                G[c1:, c0:c1] = Y
            # the diagonal block alone: full symmetric gradient block (the library's leaves store diagonal blocks full, too)
            G[c0:c1, c0:c1] = torch.from_numpy(chol_rev_recursive(torch.tril(LD).numpy(), torch.tril(G[c0:c1, c0:c1]).numpy()))
        G[c0:, c0:c1] = _bcast(G[c0:, c0:c1].clone(), _owner(p, turn, world))
        T = slice(c0, c1)
        symT = torch.tril(G[T, T]) + torch.tril(G[T, T], -1).T
        for j in range(p):                                # V(p, j): the blocks to the left that this rank owns
            if _owner(j, turn, world) != rank:
                continue
            J = slice(j * block, (j + 1) * block)
            if c1 < n:
                G[c1:, J] -= 2.0 * G[c1:, T] @ L[T, J]
                G[T, J] -= 2.0 * G[c1:, T].T @ L[c1:, J]
            G[T, J] -= 2.0 * symT @ L[T, J]
    return G
