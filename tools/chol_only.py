"""Run potrf_lower + potrf_lower_bwd once (after one warm-up) at size n -- the target of ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from henbun_b200 import _lib
lib = _lib.load()
P, ST = _lib.ptr, _lib.stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
X = torch.randn(n, 8, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
K0 = torch.exp(-0.5 * torch.cdist(X, X) ** 2 / 0.25) + 1e-3 * torch.eye(n, device="cuda")
G0 = torch.tril(torch.randn(n, n, device="cuda"))
wsb = lib.hb_potrf_workspace_bytes(n)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
err = torch.zeros(4, dtype=torch.int32, device="cuda")
for rep in range(reps):
    A = K0.clone(); G = G0.clone()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    lib.hb_potrf_lower(P(A), n, 0, n, 1, 0, P(ws), wsb, P(err), ST())
    e[1].record()
    lib.hb_potrf_lower_bwd(P(A), n, 0, P(G), n, 0, n, 1, P(ws), wsb, ST())
    e[2].record()
    torch.cuda.synchronize()
    print(f"n={n} rep={rep} fwd {e[0].elapsed_time(e[1]):.3f} ms  bwd {e[1].elapsed_time(e[2]):.3f} ms  err={err[0].item()}", flush=True)
