"""Config 1 (N=100, 1-D, full-covariance q, S=10): the persistent single-CTA step against the multi-kernel path, same process,
alternating, CUDA events over 300 steps each; also as one launch with Adam inside, and under a CUDA-graph replay."""
import sys, os, ctypes as C, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from henbun_b200 import _lib
lib = _lib.load()
rng = np.random.RandomState(0)
n, D, S, jitter = 100, 1, 10, 1e-3
full = int(sys.argv[1]) if len(sys.argv) > 1 else 1
X = np.linspace(0, 6, n).reshape(-1, 1).astype(np.float32); Y = (np.sin(X[:, 0]) + 0.3 * rng.randn(n)).astype(np.float32)
cfg = _lib.GpConfig(n, D, S, 1, full, jitter, 0, 0)
npar = lib.hb_gp_param_count(C.byref(cfg))
q_sqrt = (0.3 * np.eye(n) + 0.02 * np.tril(rng.randn(n, n))).astype(np.float32) if full else np.full(n, -1.0, np.float32)
p0 = torch.tensor(np.concatenate([0.1 * rng.randn(n), q_sqrt.ravel(), [0.54], [0.54], [0.54], [-0.5]]).astype(np.float32), device="cuda")
grads = torch.zeros(npar, device="cuda"); out4 = torch.zeros(4, device="cuda"); err = torch.zeros(1, dtype=torch.int32, device="cuda")
wsb = lib.hb_gp_elbo_workspace_bytes(C.byref(cfg)); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
Xd, Yd = torch.tensor(X, device="cuda"), torch.tensor(Y, device="cuda")
P, ST = _lib.ptr, _lib.stream


def make(mode):
    params = p0.clone(); am = torch.zeros(npar, device="cuda"); av = torch.zeros(npar, device="cuda")
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    adam = _lib.AdamConfig(1e-3, 0.9, 0.999, 1e-8, -1.0, C.c_void_p(ctr.data_ptr()), 0)

    def step(it):
        cfg.offset = C.c_ulonglong(it * ((S * n + 3) // 4 * 4))
        if mode == "one":
            lib.hb_increment_i32(P(ctr), ST())
            lib.hb_gp_small_step(C.byref(cfg), P(Xd), P(Yd), P(params), None, P(grads), P(out4), P(am), P(av), C.byref(adam), P(ws), wsb,
                                 P(err), ST())
        else:
            lib.hb_set_small_gp_kernel(1 if mode == "small" else 0)
            lib.hb_gp_elbo_step(C.byref(cfg), P(Xd), P(Yd), P(params), None, P(grads), P(out4), P(ws), wsb, P(err), ST())
            lib.hb_increment_i32(P(ctr), ST())
            lib.hb_adam_tf1(P(params), P(grads), P(am), P(av), npar, -1.0, 1e-3, 0.9, 0.999, 1e-8, P(ctr), 0, ST())
    return step


def clock():
    try:
        return subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"]).decode().strip()
    except Exception:
        return "?"


def timed(step, reps=300):
    for i in range(30):
        step(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(reps):
        step(30 + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for rnd in range(2):
    for mode in ("multi", "small", "one"):
        st = make(mode)
        l0 = lib.hb_launch_count(); st(0); k = lib.hb_launch_count() - l0
        us = timed(st)
        print(f"round {rnd} {mode:6s}: {us:7.1f} us/step ({k} launches), SM clock now {clock()} MHz, ELBO {float(out4[0]):.3f}", flush=True)
# graph replay of the one-launch step (same Philox offset every replay: timing only)
st = make("one")
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    st(300)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        st(301)
torch.cuda.current_stream().wait_stream(side)
for _ in range(30):
    g.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(500):
    g.replay()
e1.record(); torch.cuda.synchronize()
print(f"one-launch step under CUDA-graph replay: {e0.elapsed_time(e1) / 500 * 1e3:.1f} us/step, SM clock {clock()} MHz", flush=True)
lib.hb_set_small_gp_kernel(1)
