"""Synthetic inputs of the BASELINE configs -- plain numpy, shared by bench.py, the tests and the CPU
baseline so that every leg sees the same problem."""
from __future__ import annotations

import math

import numpy as np


def positive_free(y, lower=1e-6):
    """Free-space value whose transforms.positive image is y (Henbun/transforms.py:139-143: log(exp(y - lower) - 1))."""
    y = np.asarray(y, dtype=np.float64) - lower
    return y + np.log(-np.expm1(-y))


def make_gp_problem(n, D, S, seed=0, dtype=np.float32, lengthscale=0.5):
    """Config 3: X~N(0,I_D), Y=sin(sum x/sqrt(D))+0.1 eps, UnitRBF, Gaussian([n,1],'diagonal') with mu~0.1 randn,
    omega=-1, k_var=var=1.
    Lengthscale 0.5 instead of SURVEY's 1.0: at N=65536, D=8, ell=1 the Gram matrix has lambda_max ~ N/81 ~ 800 and a
    numerically zero lambda_min, so K + 1e-5 I is not positive definite in fp32 (cond ~ 8e7 > 2^24) -- the reference's
    own fp32 tf.cholesky would raise InvalidArgumentError there (our kernel reports the failing pivot through
    err_flag).  ell=0.5 keeps the full-size problem well posed in the reference's default float_type (henbunrc:7)."""
    rng = np.random.RandomState(seed)
    X = rng.randn(n, D).astype(dtype)
    Y = (np.sin(X.sum(1) / math.sqrt(D)) + 0.1 * rng.randn(n)).astype(dtype)
    one = float(positive_free(1.0))
    ell = float(positive_free(lengthscale))
    p = dict(q_mu=(0.1 * rng.randn(n)).astype(dtype), q_sqrt=np.full(n, -1.0, dtype),
             scale=np.array([one], dtype), lengthscales=np.array([ell], dtype),
             k_var=np.array([one], dtype), var=np.array([one], dtype))
    return X, Y, p


GP_PARAM_ORDER = ("q_mu", "q_sqrt", "scale", "lengthscales", "k_var", "var")


def pack_gp_params(p, dtype=np.float32):
    """Packing of hb_gp_elbo_step's params / grads vectors (include/henbun_b200.h)."""
    return np.concatenate([np.asarray(p[k], dtype).ravel() for k in GP_PARAM_ORDER])
