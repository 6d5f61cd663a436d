"""Right-looking ("flat") schedule of the factorisation and its reverse mode vs the column recursion: parity against fp64
at a moderate size, timing at the given size.  One GPU:  python tools/flat_probe.py 65536 2048
Several GPUs (column-block-cyclic):  torchrun --nproc-per-node 2 tools/flat_probe.py 65536 2048"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from henbun_b200 import _lib, parallel
lib = _lib.load()
P, ST = _lib.ptr, _lib.stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
blocks = [tuple(int(x) for x in b.split("x")) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [(2048, 1)]   # WxBATCH
check_n = int(sys.argv[3]) if len(sys.argv) > 3 else 4224
world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl")


def problem(n, seed=0):
    g = torch.Generator("cuda").manual_seed(seed)
    X = torch.randn(n, 8, device="cuda", generator=g)
    K0 = torch.cdist(X, X)
    K0.pow_(2).mul_(-2.0).exp_()
    K0.diagonal().add_(1e-3)
    G0 = torch.randn(n, n, device="cuda", generator=g)
    G0.tril_()
    return K0, G0


TRACE = os.environ.get("HB_FLAT_TRACE")


def dump_trace(label):
    cap = 4096
    buf = (C.c_double * (3 * cap))()
    k = lib.hb_flat_trace_end(buf, cap)
    rows = [(int(buf[3 * i]), int(buf[3 * i + 1]), buf[3 * i + 2]) for i in range(min(k, cap))]
    t0 = {}
    print(f"--- trace {label} rank {rank}: tag0/1 chain F(p) start/done, 2 panel available, 3 chain update done, 4/5 main bulk start/done")
    for tag, p_, ms in rows:
        t0.setdefault(p_, {})[tag] = ms
    for p_ in sorted(t0, reverse=rows[0][1] > rows[-1][1]):
        d = t0[p_]
        f = lambda a: f"{d[a]:8.2f}" if a in d else "       -"
        print(f"  p={p_:3d}  F {f(0)} ..{f(1)}  avail {f(2)}  chainU {f(3)}  bulk {f(4)} ..{f(5)}")


def run(n, K0, G0, env, reps=1):
    err = torch.zeros(4, dtype=torch.int32, device="cuda")
    if env is None:
        wsb = lib.hb_potrf_workspace_bytes(n)
    else:
        wsb = lib.hb_potrf_dist_workspace_bytes(n, C.byref(env))
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    best = None
    for rep in range(reps):
        A = K0.clone(); G = G0.clone()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        tr = TRACE and env is not None and rep == reps - 1 and n > 8192
        e[0].record()
        if tr:
            lib.hb_flat_trace_begin()
        if env is None:
            _lib.check(lib.hb_potrf_lower(P(A), n, 0, n, 1, 0, P(ws), wsb, P(err), ST()), "potrf")
        else:
            _lib.check(lib.hb_potrf_lower_dist(P(A), n, n, C.byref(env), P(ws), wsb, P(err), ST()), "potrf_dist")
        e[1].record()
        if tr:
            dump_trace(f"forward W={env.block}x{env.batch}")
            lib.hb_flat_trace_begin()
        if env is None:
            _lib.check(lib.hb_potrf_lower_bwd(P(A), n, 0, P(G), n, 0, n, 1, P(ws), wsb, ST()), "bwd")
        else:
            _lib.check(lib.hb_potrf_lower_bwd_dist(P(A), n, P(G), n, n, C.byref(env), P(ws), wsb, ST()), "bwd_dist")
        e[2].record()
        if tr:
            dump_trace(f"reverse W={env.block}x{env.batch}")
        torch.cuda.synchronize()
        t = (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]))
        best = t if best is None or sum(t) < sum(best) else best
    return A, G, best, int(err[0].item())


# ---- parity at check_n against fp64 (cuSOLVER + autograd as the checker) ----
if check_n > 0:
    K0, G0 = problem(check_n, 1)
    Kd = K0.double().requires_grad_(True)
    Ld = torch.linalg.cholesky(Kd)
    (Ld * G0.double()).sum().backward()
    Kbar = Kd.grad
    Kbar = torch.tril(Kbar + Kbar.T) - torch.diag(torch.diag(Kbar))       # lower triangle of the symmetric gradient, off-diagonals doubled
    ref_sym = torch.tril(0.5 * (Kd.grad + Kd.grad.T))                      # full-symmetric convention (what the library leaves)
    for env in [None] + [parallel.block_cyclic_env(b, k, turn=1, block_bwd=0) for b, k in ((256, 1), (512, 3), (256, 4), (1024, 2))]:
        A, G, t, err = run(check_n, K0, G0, env)
        eL = (torch.tril(A).double() - Ld.detach()).norm() / Ld.detach().norm()
        eG = (torch.tril(G).double() - ref_sym).norm() / ref_sym.norm()
        if rank == 0:
            print(f"check n={check_n} {'recursive' if env is None else f'flat W={env.block}x{env.batch} world={world}'}: |L-L64|/|L64| = {eL:.2e}  "
                  f"|Kbar-ref|/|ref| = {eG:.2e}  err={err}", flush=True)
    del K0, G0, Kd, Ld, Kbar, ref_sym, A, G
    torch.cuda.empty_cache()

# ---- timing at n ----
K0, G0 = problem(n, 0)
if world == 1:
    A, G, t, err = run(n, K0, G0, None, reps=2)
    print(f"n={n} recursive: fwd {t[0]:.1f} ms  bwd {t[1]:.1f} ms  err={err}", flush=True)
    Ar, Gr = (A, G) if n <= 32768 else (None, None)
    del A, G
for bb in blocks:
    b, k, tn = (tuple(bb) + (1, 1))[:3]                   # WxBATCHxTURN
    env = parallel.block_cyclic_env(b, k, turn=tn, block_bwd=0)
    A, G, t, err = run(n, K0, G0, env, reps=2)
    msg = ""
    if world == 1 and Ar is not None:
        dl = (torch.tril(A) - torch.tril(Ar)).norm() / torch.tril(Ar).norm()
        dg = (torch.tril(G) - torch.tril(Gr)).norm() / torch.tril(Gr).norm()
        msg = f"  vs recursive: dL {dl:.2e} dG {dg:.2e}"
    if rank == 0:
        print(f"n={n} flat W={b}x{k}x{tn} world={world}: fwd {t[0]:.1f} ms  bwd {t[1]:.1f} ms  err={err}{msg}", flush=True)
if world > 1:
    dist.destroy_process_group()
