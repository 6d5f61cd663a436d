"""One GEMM shape on the tensor-core engine, a few repetitions (target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from henbun_b200 import _lib
lib = _lib.load(); P, ST = _lib.ptr, _lib.stream
M, N, K, tA, tB = [int(x) for x in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
c_tri = int(sys.argv[7]) if len(sys.argv) > 7 else 0
g = torch.Generator("cuda").manual_seed(1)
A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g)
B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g)
C = torch.zeros(M, N, device="cuda")
lib.hb_set_gemm_engine(2)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for r in range(reps):
    e0.record()
    rc = lib.hb_gemm_ws(P(A), A.shape[1], 0, tA, 0, P(B), B.shape[1], 0, tB, 0, P(C), N, 0, c_tri, M, N, K, 1, 1.0, 0.0, None, 0, 0, 0,
                        -50.0, 50.0, None, 0, ST())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rc={rc} {ms:.3f} ms {2.0 * M * N * K * (0.5 if c_tri else 1) / ms / 1e9:.1f} TF/s", flush=True)
