"""CPU baseline = the oracle restatement timed in torch-CPU fp32.  TEST/BENCH INFRASTRUCTURE ONLY.

TensorFlow (the reference's arithmetic engine, un-vendored: setup.py:34-37) cannot be installed in
this image, so the "reference TF CPU path" is stood in for by the oracle's restatement of the
GaussianProcess.ipynb:109-148 graph executed by torch-CPU (MKL GEMM, LAPACK potrf, autograd backward,
TF-1 Adam rule).  It batches the S samples and has no per-op session overhead, so it is FASTER than
real TF-1 Henbun: speed-ups quoted against it are conservative.
Only bench.py's cpu_baseline / --impl reference legs may call this.
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch

from . import henbun_oracle as O


def make_gp_problem(n, D, S, seed=0, dtype=np.float32):
    """Synthetic config-3 inputs: X~N(0,I_D), Y=sin(sum x/sqrt(D))+0.1 eps,
    UnitRBF lengthscale 0.5, Gaussian([n,1],'diagonal') with mu~0.1 randn, omega=-1, k_var=var=1.
    Lengthscale 0.5 instead of SURVEY's 1.0: at N=65536, D=8, ell=1 the Gram matrix has
    lambda_max ~ N/81 ~ 800 and a numerically zero lambda_min, so K + 1e-5 I is not positive definite
    in fp32 (cond ~ 8e7 > 2^24) -- the reference's own fp32 tf.cholesky would raise
    InvalidArgumentError there (our kernel reports the failing pivot through err_flag).  ell=0.5
    keeps the full-size problem well posed in the reference's default float_type (henbunrc:7)."""
    rng = np.random.RandomState(seed)
    X = rng.randn(n, D).astype(dtype)
    Y = (np.sin(X.sum(1) / math.sqrt(D)) + 0.1 * rng.randn(n)).astype(dtype)
    one = float(O.log1pe_backward(1.0))
    half = float(O.log1pe_backward(0.5))
    p = dict(q_mu=(0.1 * rng.randn(n)).astype(dtype), q_sqrt=np.full(n, -1.0, dtype),
             scale=np.array([one], dtype), lengthscales=np.array([half], dtype),
             k_var=np.array([one], dtype), var=np.array([one], dtype))
    return X, Y, p


def time_gpr_steps(n, D, S, steps=2, warmup=1, threads=None, seed=0):
    """Median seconds per (ELBO + gradient + TF-1 Adam) step of the oracle in torch-CPU fp32."""
    if threads:
        torch.set_num_threads(int(threads))
    X, Y, p = make_gp_problem(n, D, S, seed)
    tX, tY = torch.tensor(X), torch.tensor(Y)
    tp = {k: torch.tensor(v, requires_grad=True) for k, v in p.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v2 = {k: np.zeros_like(v) for k, v in p.items()}
    gen = torch.Generator().manual_seed(seed + 1)
    times, last = [], None
    for it in range(warmup + steps):
        U = torch.randn(S, n, generator=gen, dtype=torch.float32)
        t0 = time.perf_counter()
        for t in tp.values():
            t.grad = None
        elbo = O.gpr_elbo(tp, tX, tY, U)
        elbo.backward()
        with torch.no_grad():
            for k, t in tp.items():
                th, m[k], v2[k] = O.adam_tf1_step(t.numpy(), -t.grad.numpy(), m[k], v2[k], it + 1)
                t.copy_(torch.from_numpy(np.asarray(th, dtype=np.float32)))
        dt = time.perf_counter() - t0
        last = float(elbo.detach())
        if it >= warmup:
            times.append(dt)
    return float(np.median(times)), last, torch.get_num_threads()
