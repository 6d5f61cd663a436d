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


from henbun_b200.synthetic import make_gp_problem   # noqa: F401  (the generator lives with the product; re-exported)


def time_gpr_steps(n, D, S, steps=2, warmup=1, threads=None, seed=0, reduce="median"):
    """Median (or mean) seconds per (ELBO + gradient + TF-1 Adam) step of the oracle in torch-CPU fp32."""
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
    t = float(np.mean(times)) if reduce == "mean" else float(np.median(times))
    return t, last, torch.get_num_threads()
