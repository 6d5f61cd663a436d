"""One product on the pre-split fp16 hi/lo engine (target for ncu): python tools/h2_one.py M N K tA tB [reps]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from henbun_b200 import _lib
lib = _lib.load(); P, ST = _lib.ptr, _lib.stream
M, N, K, tA, tB = [int(x) for x in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
g = torch.Generator("cuda").manual_seed(1)
A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g)
B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g)
C = torch.zeros(M, N, device="cuda")
wsb = lib.hb_gemm_presplit_workspace_bytes(M, N, K, tA, tB)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for r in range(reps):
    e0.record()
    rc = lib.hb_gemm_presplit(P(A), A.shape[1], tA, P(B), B.shape[1], tB, P(C), N, 0, M, N, K, 1.0, 0.0, 0, 0, 1 if r else 0,
                              P(ws), wsb, ST())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rc={rc} {ms:.3f} ms {2.0 * M * N * K / ms / 1e9:.1f} TF/s{'' if r else '  (includes the split passes)'}", flush=True)
