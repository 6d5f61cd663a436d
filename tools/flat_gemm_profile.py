"""Per-launch GEMM events of the reverse mode (or forward) under the flat schedule issued on ONE stream (lookahead = 0):
which products the block chain consists of.  python tools/flat_gemm_profile.py 65536 2048 bwd out.csv"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from henbun_b200 import _lib, parallel
lib = _lib.load(); P, ST = _lib.ptr, _lib.stream
n = int(sys.argv[1]); W = int(sys.argv[2]); which = sys.argv[3]; out = sys.argv[4]
g = torch.Generator("cuda").manual_seed(0)
X = torch.randn(n, 8, device="cuda", generator=g)
K0 = torch.cdist(X, X); K0.pow_(2).mul_(-2.0).exp_(); K0.diagonal().add_(1e-3)
G0 = torch.randn(n, n, device="cuda", generator=g); G0.tril_()
env = parallel.block_cyclic_env(W, 1)
wsb = lib.hb_potrf_dist_workspace_bytes(n, C.byref(env)); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
err = torch.zeros(4, dtype=torch.int32, device="cuda")
_lib.OPTIONS.lookahead = 0
for rep in range(2):
    A = K0.clone(); G = G0.clone()
    if rep == 1 and which == "fwd":
        lib.hb_profile_begin(200000)
    lib.hb_potrf_lower_dist(P(A), n, n, C.byref(env), P(ws), wsb, P(err), ST())
    if rep == 1 and which == "fwd":
        buf = (C.c_double * 8)(); lib.hb_profile_end_ex(buf); lib.hb_profile_dump_csv(out.encode())
    if rep == 1 and which == "bwd":
        lib.hb_profile_begin(200000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.hb_potrf_lower_bwd_dist(P(A), n, P(G), n, n, C.byref(env), P(ws), wsb, ST())
    e1.record()
    if rep == 1 and which == "bwd":
        buf = (C.c_double * 8)(); lib.hb_profile_end_ex(buf); lib.hb_profile_dump_csv(out.encode())
    torch.cuda.synchronize()
    print(f"rep {rep}: bwd serial {e0.elapsed_time(e1):.1f} ms; gemm launches {buf[0] if rep else 0}, gemm ms {buf[1] if rep else 0}")
