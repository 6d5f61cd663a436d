"""Right-looking two-stream schedule of the blocked Cholesky and of its reverse mode (csrc/linalg.cu: potrf_flat /
chol_rev_flat; reference ops: tf.cholesky, Henbun/gp/kernels.py:100-101, and its gradient through minimize(), model.py:220),
on one GPU and column-block-cyclic over two: against fp64 LAPACK / autograd on the GPU (cuSOLVER as the checker), against the
column recursion, and the fused GP step with a shared factorisation against the single-GPU step."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from henbun_b200 import _lib
    return _lib.load()


def _problem(n, seed):
    g = torch.Generator("cuda").manual_seed(seed)
    X = torch.randn(n, 8, device="cuda", generator=g, dtype=torch.float64)
    K64 = torch.exp(-0.5 * torch.cdist(X, X) ** 2 / 0.25) + 1e-3 * torch.eye(n, device="cuda", dtype=torch.float64)
    Lbar = torch.tril(torch.randn(n, n, device="cuda", generator=g, dtype=torch.float64))
    Kr = K64.clone().requires_grad_(True)
    Lref = torch.linalg.cholesky(Kr)
    (Lref * Lbar).sum().backward()
    Gref = torch.tril(0.5 * (Kr.grad + Kr.grad.T))
    return K64.float().contiguous(), Lbar.float().contiguous(), Lref.detach(), Gref


@pytest.mark.parametrize("n,block,batch", [(1000, 256, 1), (2176, 128, 3), (4224, 512, 1), (4224, 256, 4), (8200, 1024, 2),
                                           (8200, 2048, 1), (12288, 1536, 2)])
def test_flat_schedule_vs_fp64(lib, n, block, batch):
    """hb_potrf_lower_dist / hb_potrf_lower_bwd_dist with world = 1: every block width / batching (ragged last blocks,
    blocks wider than the matrix's remainder, both engines) reproduces fp64 to the bars of the column recursion."""
    from henbun_b200 import _lib
    P, ST = _lib.ptr, _lib.stream
    A, G, Lref, Gref = _problem(n, n + block)
    env = _lib.Dist(None, 0, 1, block, 0, batch, 0, 1)
    wsb = lib.hb_potrf_dist_workspace_bytes(n, C.byref(env))
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.hb_potrf_lower_dist(P(A), n, n, C.byref(env), P(ws), wsb, P(err), ST()) == 0
    assert lib.hb_potrf_lower_bwd_dist(P(A), n, P(G), n, n, C.byref(env), P(ws), wsb, ST()) == 0
    torch.cuda.synchronize()
    assert err.item() == 0
    eL = (torch.linalg.norm(torch.tril(A).double() - Lref) / torch.linalg.norm(Lref)).item()
    eG = (torch.linalg.norm(torch.tril(G).double() - Gref) / torch.linalg.norm(Gref)).item()
    assert eL < 2e-6, eL
    assert eG < 5e-6, eG


def test_schedule_option_routes_the_standard_entry_points(lib):
    """hb_options.schedule: 1 = column recursion, >= 128 = right-looking with that block width, both through
    hb_potrf_lower / hb_potrf_lower_bwd; the two-stream schedule is deterministic (bitwise equal from run to run: no race
    between the chain and the trailing updates); lookahead = 0 issues the same products on one stream."""
    from henbun_b200 import _lib
    P, ST = _lib.ptr, _lib.stream
    n = 8320
    A0, G0, Lref, Gref = _problem(n, 7)
    wsb = lib.hb_potrf_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = {}
    try:
        for name, sched, look in (("recursion", 1, 1), ("flat", 1024, 1), ("flat_again", 1024, 1), ("flat_serial", 1024, 0)):
            lib.hb_set_schedule(sched)
            _lib.OPTIONS.lookahead = look
            A, G = A0.clone(), G0.clone()
            assert lib.hb_potrf_lower(P(A), n, 0, n, 1, 1, P(ws), wsb, P(err), ST()) == 0
            assert lib.hb_potrf_lower_bwd(P(A), n, 0, P(G), n, 0, n, 1, P(ws), wsb, ST()) == 0
            torch.cuda.synchronize()
            assert err.item() == 0
            assert torch.count_nonzero(torch.triu(A, 1)).item() == 0
            out[name] = (A, torch.tril(G))
    finally:
        lib.hb_set_schedule(0)
        _lib.OPTIONS.lookahead = 1
    for name, (A, G) in out.items():
        assert (torch.linalg.norm(A.double() - Lref) / torch.linalg.norm(Lref)).item() < 2e-6, name
        assert (torch.linalg.norm(G.double() - Gref) / torch.linalg.norm(Gref)).item() < 5e-6, name
    assert torch.equal(out["flat"][0], out["flat_again"][0]) and torch.equal(out["flat"][1], out["flat_again"][1])
    for i in (0, 1):     # without the leaf look-ahead a few products are not split: same values up to summation order
        assert (torch.linalg.norm(out["flat"][i] - out["flat_serial"][i]) / torch.linalg.norm(out["flat"][i])).item() < 1e-6
    assert not torch.equal(out["flat"][0], out["recursion"][0])       # a different summation order: the route was taken


def test_non_positive_pivot_is_flagged_by_the_flat_schedule(lib):
    from henbun_b200 import _lib
    P, ST = _lib.ptr, _lib.stream
    n = 1024
    A = -torch.eye(n, device="cuda")
    env = _lib.Dist(None, 0, 1, 256, 0, 1, 0, 1)
    wsb = lib.hb_potrf_dist_workspace_bytes(n, C.byref(env))
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.hb_potrf_lower_dist(P(A), n, n, C.byref(env), P(ws), wsb, P(err), ST()) == 0
    torch.cuda.synchronize()
    assert err.item() == 1
    bad = _lib.Dist(None, 0, 1, 200, 0, 1, 0, 1)                              # block must be a multiple of 128
    assert lib.hb_potrf_lower_dist(P(A), n, n, C.byref(bad), P(ws), wsb, P(err), ST()) == _lib.HB_ERR_ARG
    two = _lib.Dist(None, 0, 2, 256, 0, 1, 0, 1)                              # more than one rank needs a communicator
    assert lib.hb_potrf_lower_dist(P(A), n, n, C.byref(two), P(ws), 1 << 40, P(err), ST()) == _lib.HB_ERR_ARG


# ---- two GPUs ---------------------------------------------------------------------------------------------------------
WORKER = r'''
import os, sys, json, ctypes as C
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from henbun_b200 import _lib, parallel
import henbun_b200 as hb, henbun_b200.tf as tf
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
lib = _lib.load(); P, ST = _lib.ptr, _lib.stream
out = {}
# (a) the factorisation and its reverse mode, column-block-cyclic, against fp64
n = 4224
g = torch.Generator("cuda").manual_seed(3)
X = torch.randn(n, 8, device="cuda", generator=g, dtype=torch.float64)
K64 = torch.exp(-0.5 * torch.cdist(X, X) ** 2 / 0.25) + 1e-3 * torch.eye(n, device="cuda", dtype=torch.float64)
Lbar = torch.tril(torch.randn(n, n, device="cuda", generator=g, dtype=torch.float64))
Kr = K64.clone().requires_grad_(True); Lref = torch.linalg.cholesky(Kr); (Lref * Lbar).sum().backward()
Gref = torch.tril(0.5 * (Kr.grad + Kr.grad.T)); Lref = Lref.detach()
for block, batch, turn in ((512, 1, 1), (256, 3, 2), (1024, 1, 1), (128, 2, 3)):
    env = parallel.block_cyclic_env(block, batch, turn=turn)
    A = K64.float().contiguous(); G = Lbar.float().contiguous()
    wsb = lib.hb_potrf_dist_workspace_bytes(n, C.byref(env)); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.hb_potrf_lower_dist(P(A), n, n, C.byref(env), P(ws), wsb, P(err), ST()) == 0
    assert lib.hb_potrf_lower_bwd_dist(P(A), n, P(G), n, n, C.byref(env), P(ws), wsb, ST()) == 0
    torch.cuda.synchronize()
    eL = (torch.linalg.norm(torch.tril(A).double() - Lref) / torch.linalg.norm(Lref)).item()
    eG = (torch.linalg.norm(torch.tril(G).double() - Gref) / torch.linalg.norm(Gref)).item()
    # every rank ends with the complete factor / gradient, bit-identical
    same = True
    if world > 1:
        for T in (torch.tril(A), torch.tril(G)):
            ref = T.clone(); dist.broadcast(ref, src=0)
            flag = torch.tensor([float(torch.equal(ref, T))], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            same = same and bool(flag.item())
    out[f"chol_{block}x{batch}x{turn}"] = [eL, eG, int(err.item()), same]
del K64, Lbar, Kr, Lref, Gref, A, G, ws
# (b) the API step: samples sharded, ONE factorisation shared by the ranks (n >= GpElboBinding.SHARED_MIN_N)
rng = np.random.RandomState(0)
n, D, S = 4224, 8, 8
Xh = rng.randn(n, D); Yh = np.sin(Xh.sum(1, keepdims=True)) + 0.1 * rng.randn(n, 1)
class GPR(hb.model.Model):
    def setUp(self):
        self.X = hb.param.Data(Xh); self.Y = hb.param.Data(Yh)
        self.q = hb.variationals.Gaussian(shape=[n, 1], q_shape='diagonal')
        self.kern = hb.gp.kernels.UnitRBF()
        self.k_var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)
        self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)
    @hb.model.AutoOptimize()
    def ELBO(self):
        y_fit = tf.matmul(self.kern.Cholesky(self.X), self.q) * tf.sqrt(self.k_var)
        return tf.reduce_sum(hb.densities.gaussian(self.Y, y_fit, self.var)) - self.KL()
for shared in (True, False):
    np.random.seed(5)
    m = GPR()
    m.kern.lengthscales = np.array([0.5])      # the bench's conditioning: K + 1e-5 I positive definite in fp32
    m.ELBO().compile(optimizer=tf.train.AdamOptimizer(0.01), n_samples=S, seed=3, shard='samples', verbose=False,
                     shared_factorisation=shared)
    objs = [float(m.ELBO().optimize(maxiter=1)) for _ in range(3)]
    out[f"api_{shared}"] = [float(x) for x in np.concatenate([m.q.q_mu._free_numpy().ravel()[:24], m.q.q_sqrt._free_numpy().ravel()[:8],
                                                            m.kern.lengthscales._free_numpy().ravel(), m.k_var._free_numpy().ravel(),
                                                            m.var._free_numpy().ravel()])]
    out[f"entry_{shared}"] = m.ELBO().fused_entry
if rank == 0:
    print("RESULT " + json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
'''


def _run(world):
    code = WORKER % {"root": ROOT}
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    if world == 1:
        cmd = [sys.executable, "-c", code]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", "29741", "--no-python", sys.executable, "-c", code]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_block_cyclic_factorisation_and_shared_step_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    one, two = _run(1), _run(2)
    for k, v in two.items():
        if k.startswith("chol_"):
            eL, eG, err, same = v
            assert err == 0 and same, (k, v)
            assert eL < 2e-6 and eG < 5e-6, (k, v)
    assert two["entry_True"] == "GpElboBinding"
    # sample-sharded over two ranks, shared or replicated factorisation, equals the single-GPU step with the same S in total
    for k in ("api_True", "api_False"):
        assert np.allclose(one[k], two[k], rtol=2e-4, atol=2e-6), k
    assert np.allclose(two["api_True"], two["api_False"], rtol=2e-4, atol=2e-6)


def test_gp_step_on_the_right_looking_schedule_equals_the_recursive_step(lib):
    """hb_gp_elbo_step_dist with world = 1 (the whole fused step over the right-looking schedule, any block width) against
    hb_gp_elbo_step on the column recursion and against the fp64 oracle: ELBO and every gradient to 1e-5."""
    from henbun_b200 import _lib
    from henbun_b200.synthetic import make_gp_problem, pack_gp_params
    from oracle import henbun_oracle as O
    P, ST = _lib.ptr, _lib.stream
    n, D, S = 4224, 8, 8
    X, Y, p = make_gp_problem(n, D, S, seed=3, lengthscale=0.5)
    U = np.random.RandomState(5).randn(S, n)
    val, g = O.value_and_grads(O.gpr_elbo, {k: np.asarray(v, np.float64) for k, v in p.items()}, X.astype(np.float64),
                               Y.astype(np.float64), U)
    order = ("q_mu", "q_sqrt", "scale", "lengthscales", "k_var", "var")
    ref = np.concatenate([np.asarray(g[k], np.float64).ravel() for k in order])
    dev = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda()
    Xd, Yd, Ud, params = dev(X), dev(Y), dev(U), dev(pack_gp_params(p))
    cfg = _lib.GpConfig(n, D, S, 1, 0, 1e-5, 0, 0)
    npar = lib.hb_gp_param_count(C.byref(cfg))
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    outs = {}
    try:
        lib.hb_set_schedule(1)                       # the reference run: column recursion
        for name, env in (("recursion", None), ("flat512", _lib.Dist(None, 0, 1, 512, 0, 1, 0, 1)), ("flat1024x2_bwd256", _lib.Dist(None, 0, 1, 1024, 0, 2, 256, 1))):
            grads, out4 = torch.zeros(npar, device="cuda"), torch.zeros(4, device="cuda")
            if env is None:
                wsb = lib.hb_gp_elbo_workspace_bytes(C.byref(cfg)); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
                rc = lib.hb_gp_elbo_step(C.byref(cfg), P(Xd), P(Yd), P(params), P(Ud), P(grads), P(out4), P(ws), wsb, P(err), ST())
            else:
                wsb = lib.hb_gp_elbo_dist_workspace_bytes(C.byref(cfg), C.byref(env)); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
                rc = lib.hb_gp_elbo_step_dist(C.byref(cfg), C.byref(env), P(Xd), P(Yd), P(params), P(Ud), P(grads), P(out4), P(ws), wsb,
                                              P(err), ST())
            assert rc == 0
            torch.cuda.synchronize()
            assert err.item() == 0
            outs[name] = (out4[0].item(), grads.cpu().numpy().astype(np.float64))
            del ws
    finally:
        lib.hb_set_schedule(0)
    for name, (elbo, gr) in outs.items():
        assert abs(elbo - float(val)) <= 1e-5 * abs(float(val)), name
        assert np.linalg.norm(gr - ref) <= 1e-5 * np.linalg.norm(ref), (name, np.linalg.norm(gr - ref) / np.linalg.norm(ref))
    assert not np.array_equal(outs["recursion"][1], outs["flat512"][1])
