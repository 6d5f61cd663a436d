"""The whole config-3 step (hb_gp_elbo_step + hb_adam_tf1: ~9500 launches incl. the side-stream look-ahead) captured in ONE
CUDA graph and replayed, against the same step launched eagerly.  The Philox offset is baked into the captured kernel
parameters, so every replay draws the same eps: a timing experiment (launch-gap saving), not a training loop."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from henbun_b200 import _lib
from henbun_b200.synthetic import make_gp_problem, pack_gp_params
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
D, S = 8, 64
X, Y, p = make_gp_problem(n, D, S, seed=0)
cfg = _lib.GpConfig(n, D, S, 1, 0, 1e-5, 1000, 0)
npar = lib.hb_gp_param_count(C.byref(cfg))
dev = torch.device("cuda")
params = torch.from_numpy(pack_gp_params(p)).to(dev); grads = torch.zeros(npar, device=dev)
am = torch.zeros(npar, device=dev); av = torch.zeros(npar, device=dev)
ctr = torch.zeros(1, dtype=torch.int32, device=dev); out4 = torch.zeros(4, device=dev); err = torch.zeros(1, dtype=torch.int32, device=dev)
wsb = lib.hb_gp_elbo_workspace_bytes(C.byref(cfg)); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
Xd, Yd = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
P, ST = _lib.ptr, _lib.stream


def step(it):
    cfg.offset = C.c_ulonglong(it * ((S * n + 3) // 4 * 4))
    _lib.check(lib.hb_gp_elbo_step(C.byref(cfg), P(Xd), P(Yd), P(params), None, P(grads), P(out4), P(ws), wsb, P(err), ST()), "step")
    _lib.check(lib.hb_increment_i32(P(ctr), ST()), "inc")
    _lib.check(lib.hb_adam_tf1(P(params), P(grads), P(am), P(av), npar, -1.0, 1e-3, 0.9, 0.999, 1e-8, P(ctr), 0, ST()), "adam")


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(reps):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for i in range(3):
    step(i)
l0 = lib.hb_launch_count(); step(3); launches = lib.hb_launch_count() - l0
reps = 5 if n >= 32768 else 20
eager = timed(lambda i: step(10 + i), reps)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step(100)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step(101)
torch.cuda.current_stream().wait_stream(side)
g.replay(); torch.cuda.synchronize()
graph = timed(lambda i: g.replay(), reps)
print(f"N={n}: {launches} launches/step; eager {eager:.2f} ms/step, CUDA-graph replay {graph:.2f} ms/step "
      f"({100 * (eager - graph) / eager:.1f} % saved); ELBO {float(out4[0]):.2f}, err {int(err.item())}", flush=True)
