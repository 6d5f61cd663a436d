"""Autograd operators over the C ABI: what a user objective written against the Henbun API is
made of when it runs on henbun_b200.  Every heavy op (matmul, Cholesky and its reverse mode, the RBF
Gram matrix, the reparameterised sampler + KL, densities.gaussian, MatBias) launches this repo's CUDA
kernels; torch only owns memory, the autograd tape and O(n) scalar glue.  No CPU path exists.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import ptr, stream, check, reduce_ws, ACT


def _L():
    return _lib.load()


def _c(t: torch.Tensor) -> torch.Tensor:
    t = _lib.f32(t)
    return t if t.is_contiguous() else t.contiguous()


def mark_lower(t: torch.Tensor) -> torch.Tensor:
    """Tag a tensor as lower-triangular so matmul can skip the zero half."""
    t._hb_lower = True
    return t


def is_lower(t) -> bool:
    return bool(getattr(t, "_hb_lower", False))


# --------------------------------------------------------------------------------------------
# raw GEMM call on 2-D / 3-D contiguous tensors
# --------------------------------------------------------------------------------------------
def gemm_raw(A, B, C_out, M, N, K, transA=0, transB=0, a_tri=0, b_tri=0, c_tri=0, alpha=1.0, beta=0.0, batch=1,
             sA=0, sB=0, sC=0, lda=None, ldb=None, ldc=None, bias=None, sBias=0, act=0, clip=0, lo=-50.0, hi=50.0):
    lda = lda if lda is not None else (M if transA else K)
    ldb = ldb if ldb is not None else (K if transB else N)
    ldc = ldc if ldc is not None else N
    ws, wsb = _gemm_scratch(C_out.device, M, N, K, batch)
    check(_L().hb_gemm_ws(ptr(A), lda, sA, transA, a_tri, ptr(B), ldb, sB, transB, b_tri, ptr(C_out), ldc, sC, c_tri,
                          M, N, K, batch, float(alpha), float(beta), ptr(bias), sBias, act, clip, float(lo), float(hi),
                          ptr(ws), wsb, stream()), "hb_gemm_ws")
    return C_out


_SCRATCH = {}


def _gemm_scratch(device, M, N, K, batch):
    """Split-K scratch of the tensor-core engine (long-K products with a small output, e.g. dW = x^T dz of MatBias).
    One 64 MiB buffer per device, only handed over for shapes that can use it; stream-ordered reuse is safe because
    every product runs on torch's current stream."""
    if batch != 1 or K < 512 or M * N > 74 * 128 * 256:
        return None, 0
    buf = _scratch64(device)
    return buf, buf.numel()


def _scratch64(device):
    key = (device.type, device.index)
    buf = _SCRATCH.get(key)
    if buf is None:
        buf = torch.empty(64 << 20, dtype=torch.uint8, device=device)
        _SCRATCH[key] = buf
    return buf


class _MatMul2D(torch.autograd.Function):
    """C = op(a) op(b) for 2-D (or row-flattened) operands with transpose flags handled by the GEMM
    itself (no transposed copies).  Triangular flags refer to the *stored* matrices."""

    @staticmethod
    def forward(ctx, a, b, ta, tb, a_lower, b_lower):
        a = _c(a); b = _c(b)
        lead = tuple(a.shape[:-1]) if (a.dim() > 2 and not ta) else None
        a2 = a.reshape(-1, a.shape[-1]) if lead is not None else a
        M, K = (a2.shape[1], a2.shape[0]) if ta else a2.shape
        N = b.shape[0] if tb else b.shape[1]
        # op(A)(m,k): stored lower -> ta=0: keep k<=m (1); ta=1: keep k>=m (2)
        a_tri = 0 if not a_lower else (2 if ta else 1)
        # op(B)(k,n): stored lower -> tb=0: B[k][n], keep n<=k (1); tb=1: B[n][k], keep k<=n (2)
        b_tri = 0 if not b_lower else (2 if tb else 1)
        out = torch.empty(M, N, device=a.device)
        gemm_raw(a2, b, out, M, N, K, transA=int(ta), transB=int(tb), a_tri=a_tri, b_tri=b_tri)
        ctx.save_for_backward(a2, b)
        ctx.meta = (ta, tb, a_lower, b_lower, M, N, K, lead, tuple(a.shape))
        return out.reshape(lead + (N,)) if lead is not None else out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ta, tb, al, bl, M, N, K, lead, ashape = ctx.meta
        g = _c(g).reshape(M, N)
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = torch.zeros_like(a) if al else torch.empty_like(a)
            ct = 1 if al else 0
            if not ta:      # dA[M,K] = G op(B)^T
                bt = 0 if not bl else (1 if tb else 2)      # op'(B)(n,k): tb=1 uses B[n][k] (lower: k<=n -> mode 1 in (k=n_,n=k_) space)
                gemm_raw(g, b, ga, M, K, N, transB=0 if tb else 1, b_tri=bt, c_tri=ct)
            else:           # dA stored [K,M] = op(B) G^T
                at = 0 if not bl else (2 if tb else 1)
                gemm_raw(b, g, ga, K, M, N, transA=1 if tb else 0, transB=1, a_tri=at, c_tri=ct)
            ga = ga.reshape(ashape)
        if ctx.needs_input_grad[1]:
            gb = torch.zeros_like(b) if bl else torch.empty_like(b)
            ct = 1 if bl else 0
            if not tb:      # dB[K,N] = op(A)^T G
                at = 0 if not al else (1 if ta else 2)
                gemm_raw(a, g, gb, K, N, M, transA=0 if ta else 1, a_tri=at, c_tri=ct)
            else:           # dB stored [N,K] = G^T op(A)
                bt = 0 if not al else (2 if ta else 1)
                gemm_raw(g, a, gb, N, K, M, transA=1, transB=1 if ta else 0, b_tri=bt, c_tri=ct)
        return ga, gb, None, None, None, None


class _MatMulBatched(torch.autograd.Function):
    """a [M,K] (optionally lower-triangular) @ b [S,K,1] / [S,K,N], and equal-batch a @ b."""

    @staticmethod
    def forward(ctx, a, b, a_lower):
        a = _c(a); b = _c(b)
        ctx.a_lower = a_lower
        ctx.save_for_backward(a, b)
        if a.dim() == 2 and b.dim() == 3 and b.shape[-1] == 1:
            # fold the sample axis into the GEMM: out[S,M] = b[S,K] a^T
            M, K = a.shape; S = b.shape[0]
            out = torch.empty(S, M, 1, device=a.device)
            gemm_raw(b, a, out, S, M, K, transB=1, b_tri=2 if a_lower else 0)
            ctx.mode = "fold"
            return out
        if a.dim() == 2 and b.dim() == 3:
            M, K = a.shape; S, _, N = b.shape
            out = torch.empty(S, M, N, device=a.device)
            gemm_raw(a, b, out, M, N, K, a_tri=1 if a_lower else 0, batch=S, sA=0, sB=K * N, sC=M * N)
            ctx.mode = "bcast_a"
            return out
        if a.dim() == b.dim() and a.dim() >= 3 and a.shape[:-2] == b.shape[:-2]:
            bs = a.shape[:-2]; nb = int(torch.Size(bs).numel())
            M, K = a.shape[-2:]; N = b.shape[-1]
            out = torch.empty(*bs, M, N, device=a.device)
            gemm_raw(a, b, out, M, N, K, batch=nb, sA=M * K, sB=K * N, sC=M * N)
            ctx.mode = "batched"
            return out
        raise ValueError(f"matmul: unsupported shapes {tuple(a.shape)} @ {tuple(b.shape)}")

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = _c(g)
        ga = gb = None
        lo = ctx.a_lower
        if ctx.mode == "fold":
            M, K = a.shape; S = b.shape[0]
            if ctx.needs_input_grad[0]:      # ga[M,K] = g[S,M]^T b[S,K]
                ga = torch.zeros(M, K, device=a.device) if lo else torch.empty(M, K, device=a.device)
                gemm_raw(g, b, ga, M, K, S, transA=1, c_tri=1 if lo else 0)
            if ctx.needs_input_grad[1]:      # gb[S,K] = g[S,M] a[M,K]
                gb = torch.empty(S, K, 1, device=a.device)
                gemm_raw(g, a, gb, S, K, M, b_tri=1 if lo else 0)
        elif ctx.mode == "bcast_a":
            M, K = a.shape; S, _, N = b.shape
            if ctx.needs_input_grad[0]:
                ga = torch.zeros(M, K, device=a.device)
                for s in range(S):
                    gemm_raw(g[s], b[s], ga, M, K, N, transB=1, beta=1.0, c_tri=1 if lo else 0)
            if ctx.needs_input_grad[1]:
                gb = torch.empty(S, K, N, device=a.device)
                gemm_raw(a, g, gb, K, N, M, transA=1, a_tri=2 if lo else 0, batch=S, sA=0, sB=M * N, sC=K * N)
        else:  # batched
            bs = a.shape[:-2]; nb = int(torch.Size(bs).numel())
            M, K = a.shape[-2:]; N = b.shape[-1]
            if ctx.needs_input_grad[0]:
                ga = torch.empty_like(a)
                gemm_raw(g, b, ga, M, K, N, transB=1, batch=nb, sA=M * N, sB=K * N, sC=M * K)
            if ctx.needs_input_grad[1]:
                gb = torch.empty_like(b)
                gemm_raw(a, g, gb, K, N, M, transA=1, batch=nb, sA=M * K, sB=M * N, sC=K * N)
        return ga, gb, None


def matmul(a, b, transpose_a=False, transpose_b=False):
    """tf.matmul.  2-D (and row-flattened) products take the transpose flags natively."""
    if b.dim() == 2 and (a.dim() == 2 or not transpose_a):
        return _MatMul2D.apply(a, b, bool(transpose_a), bool(transpose_b), is_lower(a), is_lower(b))
    lower = is_lower(a) and not transpose_a
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return _MatMulBatched.apply(a, b, lower)


# --------------------------------------------------------------------------------------------
# Cholesky (generic) and the fused kernel-Cholesky
# --------------------------------------------------------------------------------------------
def _potrf_ws(n, device):
    nbytes = _L().hb_potrf_workspace_bytes(int(n))
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


_err_flags = {}


def err_flag(device) -> torch.Tensor:
    key = str(device)
    if key not in _err_flags:
        _err_flags[key] = torch.zeros(1, dtype=torch.int32, device=device)
    return _err_flags[key]


def check_numerics(device=None):
    """Raise if a Cholesky reported a non-positive pivot (the reference surfaces this as a TF
    InvalidArgumentError from session.run).  Synchronises; call once per step or less."""
    for key, f in list(_err_flags.items()):
        v = int(f.item())
        if v != 0:
            f.zero_()
            raise FloatingPointError(f"Cholesky decomposition was not successful: non-positive pivot at row {v - 1}")


_INPLACE_BYTES = 256 << 20


def _own_grad(g):
    """Buffer the in-place reverse-mode factorisation may overwrite.  Small gradients are cloned; from 256 MB on
    the incoming gradient itself is used (it is the fresh output of the consumer's backward -- a second n x n copy
    costs 17 GB at N = 65536)."""
    g = g if g.is_contiguous() else g.contiguous()
    if g.numel() * 4 < _INPLACE_BYTES or g._base is not None:
        return g.clone()
    return g


def _mirror_lower_(G):
    """G <- tril(G) + tril(G, -1)^T block by block (bounded temporaries)."""
    n = G.shape[-1]
    step = 4096
    for i0 in range(0, n, step):
        i1 = min(n, i0 + step)
        for j0 in range(0, i0 + 1, step):
            j1 = min(n, j0 + step)
            if j0 == i0:
                blk = G[..., i0:i1, j0:j1]
                blk.copy_(blk.tril() + blk.tril(-1).transpose(-1, -2))
            else:
                G[..., j0:j1, i0:i1].copy_(G[..., i0:i1, j0:j1].transpose(-1, -2))
    return G


class _Cholesky(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A):
        A = _lib.f32(A)
        n = A.shape[-1]
        batch = int(A.numel() // (n * n)) if n else 0
        Lw = A.contiguous().clone()
        ws, nb = _potrf_ws(n, A.device)
        check(_L().hb_potrf_lower(ptr(Lw), n, n * n, n, batch, 1, ptr(ws), nb, ptr(err_flag(A.device)), stream()),
              "hb_potrf_lower")
        ctx.save_for_backward(Lw)
        return mark_lower(Lw)

    @staticmethod
    def backward(ctx, g):
        (Lw,) = ctx.saved_tensors
        n = Lw.shape[-1]
        batch = int(Lw.numel() // (n * n)) if n else 0
        G = _own_grad(g)
        ws, nb = _potrf_ws(n, Lw.device)
        check(_L().hb_potrf_lower_bwd(ptr(Lw), n, n * n, ptr(G), n, n * n, n, batch, ptr(ws), nb, stream()),
              "hb_potrf_lower_bwd")
        # full symmetric gradient, like TF: mirror the lower triangle in place (no n x n temporaries)
        G.tril_()
        G.add_(G.transpose(-1, -2).triu(1)) if G.numel() * 4 < _INPLACE_BYTES else _mirror_lower_(G)
        return G


def cholesky(A):
    return _Cholesky.apply(A)


def _kern_shapes(X, X2):
    if X.dim() == 2:
        batch, n, D = 1, X.shape[0], X.shape[1]
    elif X.dim() == 3:
        batch, n, D = X.shape
    else:
        raise ValueError("X must be [n,d] or [N,n,d]")
    n2 = n if X2 is None else X2.shape[-2]
    return batch, n, n2, D


def _check_grad_dims(ctx, D):
    """The Gram backward kernels hold one input row in registers (D <= 32); say so at forward time instead of
    failing in backward with a bare 'invalid argument'."""
    if D > 32 and any(ctx.needs_input_grad):
        raise ValueError(f"stationary kernels with more than 32 input dimensions (got {D}) can be evaluated but not "
                         "differentiated by this library")


class _RbfK(torch.autograd.Function):
    """UnitRBF.K / UnitCsymRBF.K (gp/kernels.py:110-111,122-126)."""

    @staticmethod
    def forward(ctx, X, X2, ell, csym):
        X = _c(X); X2c = None if X2 is None else _c(X2); ell = _c(ell).reshape(-1)
        batch, n, n2, D = _kern_shapes(X, X2c)
        _check_grad_dims(ctx, D)
        K = torch.empty((batch, n, n2) if X.dim() == 3 else (n, n2), device=X.device)
        check(_L().hb_rbf_gram_fwd(ptr(X), ptr(X2c), n, n2, D, batch, ptr(ell), ell.numel(), ptr(K), n2, n * n2, 0.0,
                                   0, int(csym), stream()), "hb_rbf_gram_fwd")
        ctx.save_for_backward(X, X2c if X2c is not None else X, ell)
        ctx.has_x2 = X2 is not None
        ctx.csym = int(csym)
        return K

    @staticmethod
    def backward(ctx, g):
        X, X2, ell = ctx.saved_tensors
        batch, n, n2, D = _kern_shapes(X, X2 if ctx.has_x2 else None)
        g = _c(g)
        gX = gX2 = gl = None
        lib = _L()
        want_x = ctx.needs_input_grad[0]
        want_x2 = ctx.has_x2 and ctx.needs_input_grad[1]
        if want_x or want_x2:
            gT = g.transpose(-1, -2).contiguous()

            def second_arg(Xa, Xb, sign):       # sign * sum_i G_ij K_rbf(xa_i, xb_j) (xa_i - xb_j)/ell^2
                d = torch.empty_like(Xb)
                check(lib.hb_rbf_gram_bwd_x2(ptr(g), n2, n * n2, ptr(Xa), ptr(Xb), n, n2, D, batch, ptr(ell), ell.numel(), 0,
                                             float(sign), ptr(d), stream()), "hb_rbf_gram_bwd_x2")
                return d

            def first_arg(Xa, Xb):              # the same contraction on G^T with swapped roles
                d = torch.empty_like(Xa)
                check(lib.hb_rbf_gram_bwd_x2(ptr(gT), n, n * n2, ptr(Xb), ptr(Xa), n2, n, D, batch, ptr(ell), ell.numel(), 0,
                                             1.0, ptr(d), stream()), "hb_rbf_gram_bwd_x2")
                return d
            d2 = second_arg(X, X2, 1.0) if (want_x2 or not ctx.has_x2) else None
            d1 = first_arg(X, X2) if want_x else None
            if ctx.csym:
                # UnitCsymRBF (gp/kernels.py:122-126) adds K_rbf(x, -x2): its first-argument gradient is that of K_rbf at
                # (x, -x2), its second-argument gradient the NEGATIVE of K_rbf's second-argument gradient at (x, -x2)
                nX2 = (-X2).contiguous()
                if d2 is not None:
                    d2 = d2 + second_arg(X, nX2, -1.0)
                if d1 is not None:
                    d1 = d1 + first_arg(X, nX2)
            if ctx.has_x2:
                gX, gX2 = d1, (d2 if want_x2 else None)
            elif want_x:
                gX = d1 + d2                    # K(X, X): both arguments are X
        if ctx.needs_input_grad[2]:
            ws = reduce_ws(X.device)
            gl = torch.empty(ell.numel(), device=X.device)
            check(lib.hb_rbf_gram_bwd(ptr(g), n2, n * n2, ptr(X), ptr(X2) if ctx.has_x2 else None, n, n2, D, batch,
                                      ptr(ell), ell.numel(), 0, ctx.csym, None, ptr(gl), ptr(ws), ws.numel(), stream()),
                  "hb_rbf_gram_bwd")
        return gX, gX2, gl, None


def rbf_K(X, X2, ell, csym=False):
    return _RbfK.apply(X, X2, ell, csym)


class _KernCholesky(torch.autograd.Function):
    """UnitStationary.Cholesky (gp/kernels.py:93-101) fused: Gram (lower tiles) + jitter + blocked
    potrf in one buffer; backward = reverse-mode Cholesky in place + lengthscale contraction, K is
    recomputed from X instead of being stored."""

    @staticmethod
    def forward(ctx, X, ell, jitter, csym):
        X = _c(X); ell = _c(ell).reshape(-1)
        batch, n, _, D = _kern_shapes(X, None)
        _check_grad_dims(ctx, D)
        Lw = torch.empty((batch, n, n) if X.dim() == 3 else (n, n), device=X.device)
        lib = _L()
        check(lib.hb_rbf_gram_fwd(ptr(X), None, n, n, D, batch, ptr(ell), ell.numel(), ptr(Lw), n, n * n, float(jitter),
                                  1, int(csym), stream()), "hb_rbf_gram_fwd")
        ws, nb = _potrf_ws(n, X.device)
        check(lib.hb_potrf_lower(ptr(Lw), n, n * n, n, batch, 1, ptr(ws), nb, ptr(err_flag(X.device)), stream()),
              "hb_potrf_lower")
        ctx.save_for_backward(X, ell, Lw)
        ctx.csym = int(csym)
        return mark_lower(Lw)

    @staticmethod
    def backward(ctx, g):
        X, ell, Lw = ctx.saved_tensors
        batch, n, _, D = _kern_shapes(X, None)
        G = _own_grad(g)
        lib = _L()
        ws, nb = _potrf_ws(n, X.device)
        check(lib.hb_potrf_lower_bwd(ptr(Lw), n, n * n, ptr(G), n, n * n, n, batch, ptr(ws), nb, stream()),
              "hb_potrf_lower_bwd")
        gl = None
        if ctx.needs_input_grad[1]:
            rws = reduce_ws(X.device)
            gl = torch.empty(ell.numel(), device=X.device)
            check(lib.hb_rbf_gram_bwd(ptr(G), n, n * n, ptr(X), None, n, n, D, batch, ptr(ell), ell.numel(), 1, ctx.csym,
                                      None, ptr(gl), ptr(rws), rws.numel(), stream()), "hb_rbf_gram_bwd")
        gX = None
        if ctx.needs_input_grad[0]:
            # K-bar is symmetric (lower triangle stored): both kernel arguments are X -> twice the second-argument part
            gX = torch.empty_like(X)
            check(lib.hb_rbf_gram_bwd_x2(ptr(G), n, n * n, ptr(X), None, n, n, D, batch, ptr(ell), ell.numel(), 1, 2.0,
                                         ptr(gX), stream()), "hb_rbf_gram_bwd_x2")
            if ctx.csym:    # + the mirrored term K_rbf(x, -x'), symmetric as well: -2 x its second-argument gradient at (X, -X)
                nX = (-X).contiguous()
                gc = torch.empty_like(X)
                check(lib.hb_rbf_gram_bwd_x2(ptr(G), n, n * n, ptr(X), ptr(nX), n, n, D, batch, ptr(ell), ell.numel(), 1, -2.0,
                                             ptr(gc), stream()), "hb_rbf_gram_bwd_x2")
                gX = gX + gc
        return gX, gl, None, None


def kern_cholesky(X, ell, jitter, csym=False):
    return _KernCholesky.apply(X, ell, jitter, csym)


class _TrsmRight(torch.autograd.Function):
    """X L^{-T} (trans=1) or X L^{-1} (trans=0); forward only is needed by the 'next' rows."""

    @staticmethod
    def forward(ctx, Lw, Xm, trans):
        Lw = _c(Lw); out = _c(Xm).clone()
        m, n = out.shape
        nb = _L().hb_trsm_workspace_bytes(m, n)
        ws = torch.empty(nb, dtype=torch.uint8, device=out.device)
        check(_L().hb_trsm_right_lower(ptr(Lw), n, ptr(out), n, m, n, int(trans), ptr(ws), nb, stream()),
              "hb_trsm_right_lower")
        ctx.save_for_backward(Lw, out)
        ctx.trans = int(trans)
        return out

    @staticmethod
    def backward(ctx, g):
        Lw, out = ctx.saved_tensors
        # Y = X op(L)^{-1}:  dX = g op(L)^{-T} ; dL = -tril(...) -- computed with the same primitives
        gx = _TrsmRight.apply(Lw, g, 1 - ctx.trans) if ctx.needs_input_grad[1] else None
        gl = None
        if ctx.needs_input_grad[0]:
            gxx = gx if gx is not None else _TrsmRight.apply(Lw, g, 1 - ctx.trans)
            # trans=1: Y = X L^{-T} -> dL = -tril(gX^T Y)... ; trans=0: Y = X L^{-1} -> dL = -tril(Y^T gX)
            A, B = (gxx, out) if ctx.trans == 1 else (out, gxx)
            n = Lw.shape[0]
            gl = torch.zeros(n, n, device=Lw.device)
            gemm_raw(A, B, gl, n, n, A.shape[0], transA=1, alpha=-1.0, c_tri=1)
        return gl, gx, None


def trsm_right(Lw, Xm, trans):
    return _TrsmRight.apply(Lw, Xm, trans)


# --------------------------------------------------------------------------------------------
# Reparameterised sampler + one-sample KL
# --------------------------------------------------------------------------------------------
def _rows_cols(t: torch.Tensor):
    """View t as [rows, cols] with a uniform row stride; returns (tensor, rows, cols, ld)."""
    if t.dim() == 0:
        t = t.reshape(1)
    cols = t.shape[-1]
    rows = int(t.numel() // cols) if cols else 0
    if t.is_contiguous():
        return t, rows, cols, cols
    if t.dim() >= 2 and t.stride(-1) == 1:
        # last-axis slice of a contiguous parent: uniform row stride if leading dims are packed
        ld = t.stride(-2)
        ok = True
        exp = ld * t.shape[-2]
        for d in range(t.dim() - 3, -1, -1):
            if t.shape[d] != 1 and t.stride(d) != exp:
                ok = False
                break
            exp *= t.shape[d]
        if ok:
            return t, rows, cols, ld
    t = t.contiguous()
    return t, rows, cols, cols


class _SampleDiag(torch.autograd.Function):
    """Variational._sample 'diagonal' + Normal-form KL pieces (variationals.py:138-142,183-184,225-230).
    Returns z [S, *mu.shape] and kl = -0.5*sum(2*omega + eps^2 - z^2)."""

    @staticmethod
    def forward(ctx, mu, omega, eps, seed, offset, S):
        mu = _lib.f32(mu); omega = _lib.f32(omega)
        mu2, rows, cols, ld_mu = _rows_cols(mu)
        om2, _, _, ld_om = _rows_cols(omega)
        epsc = None if eps is None else _c(eps)
        z = torch.empty((S,) + tuple(mu.shape), device=mu.device)
        kl = torch.empty(1, device=mu.device)
        ws = reduce_ws(mu.device)
        check(_L().hb_sample_diag_fwd(ptr(mu2), ld_mu, ptr(om2), ld_om, rows, cols, ptr(epsc), seed, offset, S, ptr(z),
                                      ptr(kl), ptr(ws), ws.numel(), stream()), "hb_sample_diag_fwd")
        ctx.save_for_backward(mu2, om2, epsc if epsc is not None else torch.empty(0, device=mu.device))
        ctx.meta = (rows, cols, ld_mu, ld_om, seed, offset, S, eps is not None, tuple(mu.shape))
        return z, kl.reshape(())

    @staticmethod
    def backward(ctx, gz, gkl):
        mu2, om2, epsc = ctx.saved_tensors
        rows, cols, ld_mu, ld_om, seed, offset, S, has_eps, shape = ctx.meta
        gmu = torch.empty(shape, device=mu2.device); gom = torch.empty(shape, device=mu2.device)
        gzc = None if gz is None else _c(gz)
        if gkl is None:
            coef, coef_dev = 0.0, None
        else:
            coef, coef_dev = -1.0, _c(gkl.reshape(1))
        check(_L().hb_sample_diag_bwd(ptr(mu2), ld_mu, ptr(om2), ld_om, rows, cols, ptr(epsc) if has_eps else None, seed,
                                      offset, S, ptr(gzc), None, coef, ptr(coef_dev), ptr(gmu), cols, ptr(gom), cols,
                                      0.0, stream()), "hb_sample_diag_bwd")
        return gmu, gom, None, None, None, None


def sample_diag(mu, omega, eps=None, seed=0, offset=0, S=1):
    return _SampleDiag.apply(mu, omega, eps, int(seed), int(offset), int(S))


class _SampleTril(torch.autograd.Function):
    """Variational._sample 'fullrank' (variationals.py:144-146) + KL with logdet=log(diag^2) (:185-186).
    mu [B,n], Lq [B,n,n], eps [B,S,n] -> z [B,S,n], kl scalar."""

    @staticmethod
    def forward(ctx, mu, Lq, eps):
        mu = _c(mu); Lq = _c(Lq); eps = _c(eps)
        B, S, n = eps.shape
        z = torch.empty_like(eps)
        kl = torch.empty(1, device=mu.device)
        ws = reduce_ws(mu.device)
        check(_L().hb_sample_tril_fwd(ptr(mu), ptr(Lq), n, B, ptr(eps), S, ptr(z), ptr(kl), ptr(ws), ws.numel(), stream()),
              "hb_sample_tril_fwd")
        ctx.save_for_backward(Lq, eps, z)
        return z, kl.reshape(())

    @staticmethod
    def backward(ctx, gz, gkl):
        Lq, eps, z = ctx.saved_tensors
        B, S, n = eps.shape
        c = torch.zeros((), device=z.device) if gkl is None else -gkl       # obj = g(z) - c*KL
        zt = (-c) * z if gz is None else gz - c * z                        # O(S n) glue
        zt = zt.contiguous()
        gmu = zt.sum(1)
        gL = torch.zeros_like(Lq)
        gemm_raw(zt, eps, gL, n, n, S, transA=1, c_tri=1, batch=B, sA=S * n, sB=S * n, sC=n * n)
        d = torch.diagonal(gL, dim1=-2, dim2=-1)
        d += c * float(S) / torch.diagonal(Lq, dim1=-2, dim2=-1)
        return gmu, gL, None


def sample_tril(mu, Lq, eps):
    return _SampleTril.apply(mu, Lq, eps)


def randn_philox(shape, seed, offset, device):
    out = torch.empty(shape, device=device)
    check(_L().hb_randn_philox(ptr(out), out.numel(), int(seed), int(offset), stream()), "hb_randn_philox")
    return out


# --------------------------------------------------------------------------------------------
# densities.gaussian
# --------------------------------------------------------------------------------------------
def _period(t: torch.Tensor, out_shape) -> Optional[int]:
    """numel if t broadcasts to out_shape 'modularly' (its shape is a suffix of out_shape), else None."""
    ts = list(t.shape)
    while ts and ts[0] == 1:
        ts = ts[1:]
    os_ = list(out_shape)
    if len(ts) <= len(os_) and ts == os_[len(os_) - len(ts):]:
        return max(1, int(t.numel()))
    return None


DENSITY_KINDS = {"gaussian": 0, "lognormal": 1, "bernoulli": 2, "poisson": 3, "exponential": 4, "gamma": 5,
                 "student_t": 6, "beta": 7, "laplace": 8, "bimixture": 9}


def _ptr_array(tensors):
    return (C.c_void_p * 4)(*[(t.data_ptr() if t is not None else None) for t in tensors] + [None] * (4 - len(tensors)))


class _Density(torch.autograd.Function):
    """One densities.py log-density (densities.py:25-103) = one kernel forward, one backward
    (csrc/density_family.cu).  Operands in the reference's argument order; broadcasting is modular
    (suffix-shaped operands / scalars are read in place, anything else is expanded first)."""

    @staticmethod
    def forward(ctx, kind, *args):
        shape = torch.broadcast_shapes(*[a.shape for a in args])
        total = int(torch.Size(shape).numel())
        ops_, periods = [], []
        for t in args:
            t = _lib.f32(t)
            p = _period(t, shape)
            if p is None:
                t = t.expand(shape)
                p = total
            ops_.append(_c(t)); periods.append(min(p, max(total, 1)))
        out = torch.empty(shape, device=args[0].device)
        if total:
            check(_L().hb_density_logpdf(kind, _ptr_array(ops_), (C.c_longlong * 4)(*periods + [1] * (4 - len(periods))),
                                         total, ptr(out), stream()), "hb_density_logpdf")
        ctx.save_for_backward(*ops_)
        ctx.meta = (kind, periods, total, tuple(shape), [tuple(a.shape) for a in args])
        return out

    @staticmethod
    def backward(ctx, g):
        ops_ = ctx.saved_tensors
        kind, periods, total, shape, arg_shapes = ctx.meta
        n = len(ops_)
        if total == 0:
            return (None,) + tuple(torch.zeros(s, device=g.device) if ctx.needs_input_grad[i + 1] else None
                                   for i, s in enumerate(arg_shapes))
        if all(st == 0 for st in g.stride()) and g.numel() > 0:      # the broadcast scalar a reduce_sum sends back
            gc, gp = g.reshape(-1)[:1].contiguous(), 1
        else:
            gc, gp = _c(g.expand(shape)), total
        outs = []
        for i in range(n):
            if not ctx.needs_input_grad[i + 1]:
                outs.append(None)
            elif periods[i] == 1 and total > 1:
                outs.append(torch.empty(1, device=g.device))            # reduced in-kernel
            else:
                outs.append(torch.empty(shape, device=g.device))
        ws = _lib.reduce_ws(g.device)
        check(_L().hb_density_logpdf_bwd(kind, _ptr_array(ops_), (C.c_longlong * 4)(*periods + [1] * (4 - n)), total,
                                         ptr(gc), gp, _ptr_array(outs), ptr(ws), ws.numel(), stream()),
              "hb_density_logpdf_bwd")
        grads = []
        for i in range(n):
            o = outs[i]
            if o is None:
                grads.append(None)
            elif o.numel() == 1 and total > 1:
                grads.append(o.reshape(arg_shapes[i]) if int(torch.Size(arg_shapes[i]).numel()) == 1 else o.expand(arg_shapes[i]))
            else:
                grads.append(o.sum_to_size(arg_shapes[i]))
        return (None,) + tuple(grads)


def density(name, *args):
    return _Density.apply(DENSITY_KINDS[name], *args)


def gaussian_logpdf(x, mu, var):
    return _Density.apply(0, x, mu, var)


# --------------------------------------------------------------------------------------------
# transforms.py: y = T(x) and sum log|T'(x)|
# --------------------------------------------------------------------------------------------
TRANSFORM_KINDS = {"exp": 1, "log1pe": 2, "logistic": 3}


class _TransformFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind, p0, p1):
        xc = _c(_lib.f32(x))
        y = torch.empty_like(xc)
        check(_L().hb_transform_fwd(kind, ptr(xc), xc.numel(), p0, p1, ptr(y), stream()), "hb_transform_fwd")
        ctx.save_for_backward(xc)
        ctx.meta = (kind, p0, p1)
        return y

    @staticmethod
    def backward(ctx, gy):
        (xc,) = ctx.saved_tensors
        kind, p0, p1 = ctx.meta
        gx = torch.empty_like(xc)
        check(_L().hb_transform_bwd(kind, ptr(xc), xc.numel(), p0, p1, ptr(_c(gy)), ptr(gx), stream()), "hb_transform_bwd")
        return gx, None, None, None


class _TransformLogJac(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind, p0, p1):
        xc = _c(_lib.f32(x))
        out = torch.empty(1, device=xc.device)
        ws = reduce_ws(xc.device)
        check(_L().hb_transform_logjac(kind, ptr(xc), xc.numel(), p0, p1, ptr(out), ptr(ws), ws.numel(), stream()),
              "hb_transform_logjac")
        ctx.save_for_backward(xc)
        ctx.meta = (kind, p0, p1)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        kind, p0, p1 = ctx.meta
        gx = torch.empty_like(xc)
        g1 = _c(g.reshape(1))
        check(_L().hb_transform_logjac_bwd(kind, ptr(xc), xc.numel(), p0, p1, ptr(g1), ptr(gx), stream()),
              "hb_transform_logjac_bwd")
        return gx, None, None, None


def transform_forward(name, x, p0=0.0, p1=1.0):
    return _TransformFwd.apply(x, TRANSFORM_KINDS[name], float(p0), float(p1))


def transform_log_jacobian(name, x, p0=0.0, p1=1.0):
    return _TransformLogJac.apply(x, TRANSFORM_KINDS[name], float(p0), float(p1))


# --------------------------------------------------------------------------------------------
# MatBias: act(clip(x w + b))   (nn.py:31-32, 80-84)
# --------------------------------------------------------------------------------------------
class _MatBias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act, clip, lo, hi):
        x = _c(x); w = _c(w); b = _c(b)
        K, N = w.shape[-2:]
        nb = int(w.numel() // (K * N))
        if w.dim() == 2:
            rows = int(x.numel() // K)
            out = torch.empty(*x.shape[:-1], N, device=x.device)
            gemm_raw(x, w, out, rows, N, K, bias=b.reshape(-1), act=act, clip=clip, lo=lo, hi=hi)
        else:
            if x.shape[:-2] != w.shape[:-2]:
                raise ValueError(f"MatBias: leading axes of x {tuple(x.shape)} must equal n_layers {tuple(w.shape[:-2])}")
            rows = x.shape[-2]
            out = torch.empty(*x.shape[:-1], N, device=x.device)
            gemm_raw(x, w, out, rows, N, K, batch=nb, sA=rows * K, sB=K * N, sC=rows * N, bias=b.reshape(-1), sBias=N,
                     act=act, clip=clip, lo=lo, hi=hi)
        ctx.save_for_backward(x, w, out)
        ctx.meta = (act, clip, lo, hi, nb, rows, K, N, tuple(b.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, out = ctx.saved_tensors
        act, clip, lo, hi, nb, rows, K, N, bshape = ctx.meta
        g = _c(g)
        lib = _L()
        dz = torch.empty_like(out)
        db = torch.empty(bshape, device=g.device)
        ws = _scratch64(g.device)
        for i in range(nb):
            off = i * rows * N
            check(lib.hb_act_bwd_colsum_ws(C.c_void_p(g.data_ptr() + 4 * off), C.c_void_p(out.data_ptr() + 4 * off),
                                           C.c_void_p(dz.data_ptr() + 4 * off), rows, N, N, act, clip, lo, hi,
                                           C.c_void_p(db.data_ptr() + 4 * i * N), ptr(ws), ws.numel(), stream()),
                  "hb_act_bwd_colsum_ws")
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            gemm_raw(dz, w, gx, rows, K, N, transB=1, batch=nb, sA=rows * N, sB=K * N, sC=rows * K)
        if ctx.needs_input_grad[1]:
            gw = torch.empty_like(w)
            gemm_raw(x, dz, gw, K, N, rows, transA=1, batch=nb, sA=rows * K, sB=rows * N, sC=K * N)
        return gx, gw, db, None, None, None, None


def matbias(x, w, b, act="none", clip=False, lo=-50.0, hi=50.0):
    a = ACT[act]
    if clip and a != 0:
        # clip precedes the activation in the reference; keep its gradient mask exact by splitting
        y = _MatBias.apply(x, w, b, 0, 1, float(lo), float(hi))
        return {1: torch.sigmoid, 2: torch.relu, 3: torch.tanh}[a](y)
    return _MatBias.apply(x, w, b, a, int(bool(clip)), float(lo), float(hi))


def gather_rows(src: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """dst[i] = src[index[i]] on device (MinibatchData.get_feed_dict, param.py:733-739)."""
    src = _c(src)
    idx = index.to(device=src.device, dtype=torch.int64).contiguous()
    row = int(src.numel() // src.shape[0]) if src.shape[0] else 0
    out = torch.empty((idx.numel(),) + tuple(src.shape[1:]), device=src.device)
    check(_L().hb_gather_rows(ptr(out), ptr(src), ptr(idx), idx.numel(), row, stream()), "hb_gather_rows")
    return out


def adam_tf1_(theta, grad, m, v, step_dev, lr, b1, b2, eps, grad_scale=-1.0):
    check(_L().hb_adam_tf1(ptr(theta), ptr(grad), ptr(m), ptr(v), theta.numel(), float(grad_scale), float(lr), float(b1),
                           float(b2), float(eps), ptr(step_dev), 0, stream()), "hb_adam_tf1")
