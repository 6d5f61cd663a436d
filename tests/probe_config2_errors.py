"""Config 2 (Expert GPR, N=2000, jitter 3e-4, cond ~3e6) gradient error vs the fp64 oracle under each GEMM engine
setting, next to the fp32-CPU restatement's error.  Evidence for the accuracy discussion in DESIGN.md 4.2.
Lives under tests/ (not collected by pytest) because it uses the oracle as its checker.
    python tests/probe_config2_errors.py [n] [jitter]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import henbun_b200 as hb
from henbun_b200 import _lib
from oracle import henbun_oracle as O
import test_gpu_configs as T

n, S = int(sys.argv[1]) if len(sys.argv) > 1 else 2000, 4
jitter = float(sys.argv[2]) if len(sys.argv) > 2 else 3e-4
rng = np.random.RandomState(0)
X = np.linspace(0, 6, n).reshape(-1, 1)
Y = np.sin(0.1 * X * X * X) + rng.randn(*X.shape) * 0.1
q_shapes = ('fullrank', 'fullrank', 'diagonal')
m = T.ExpertGPR(X=X, Y=Y, q_shapes=q_shapes)
for nm in ("q_s", "q_l"):
    getattr(m, nm).q_sqrt = 0.3 * np.eye(n) + (0.2 / np.sqrt(n)) * np.tril(rng.randn(n, n))
opt = T._compile(m, jitter, S)
p = T._free(m, q_shapes)
U = {nm: rng.randn(S, n).astype(np.float32) for nm in ("s", "l", "r")}
eps = {object.__getattribute__(m, "q_" + nm): U[nm].reshape(S, n, 1) for nm in U}
fn = lambda pp, X_, Y_, U_: O.expert_gpr_elbo(pp, X_, Y_, U_, q_shapes=q_shapes, jitter=jitter)
fn32 = lambda pp, X_, Y_, U_: O.expert_gpr_elbo(pp, X_, Y_, U_, q_shapes=q_shapes, jitter=jitter, K_fn=O.rbf_K_direct)
ref, gref = O.value_and_grads(fn, p, X, Y[:, 0], U)
ref32, gref32 = O.value_and_grads(fn32, p, X, Y[:, 0], U, dtype=torch.float32)
keys = list(gref)
print("fp32-CPU   ELBO rel %.1e  grads %s" % (abs(ref32 - ref) / abs(ref), " ".join("%.0e" % T.rel_err(gref32[k], gref[k]) for k in keys)))
lib = _lib.load()
# n <= 2048: the factorisation's products are exact fp32 whatever the engine (linalg.cu kExactBelow); larger n (argv[1])
# shows the tensor-core engine.  refine: hb_set_panel_refinement mode (2 = default: refined for n <= 8192; 3 = substitution kernel).
for label, eng, opt_bits, refine in (("default", 0, 0, 2), ("inverse", 0, 0, 0), ("subst", 0, 0, 3), ("simt+subst", 1, 0, 3)):
    lib.hb_set_gemm_engine(eng); lib.hb_set_tc_option(opt_bits); lib.hb_set_panel_refinement(refine)
    val, grads = T._value_and_grads(m, opt, eps)
    errs = [T.rel_err(grads["model." + k].reshape(gref[k].shape), gref[k]) for k in keys]
    print("%-10s ELBO rel %.1e  grads %s" % (label, abs(val - ref) / abs(ref), " ".join("%.0e" % e for e in errs)))
print(keys)
import time
lib.hb_set_gemm_engine(0); lib.hb_set_tc_option(0)
for refine in (0, 1, 3):
    lib.hb_set_panel_refinement(refine)
    for _ in range(3):
        T._value_and_grads(m, opt, eps)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        T._value_and_grads(m, opt, eps)
    torch.cuda.synchronize()
    print("refine", refine, "ms per ELBO+grad", (time.perf_counter() - t0) * 100)
lib.hb_set_panel_refinement(2)
