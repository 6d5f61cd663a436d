"""One launch each of the Gram forward / backward kernels at size n (ncu target; also prints their times)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from henbun_b200 import _lib
lib = _lib.load(); P, ST = _lib.ptr, _lib.stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
D = 8
g = torch.Generator("cuda").manual_seed(0)
X = torch.randn(n, D, device="cuda", generator=g)
ell = torch.full((1,), 0.5, device="cuda")
K = torch.empty(n, n, device="cuda")
G = torch.empty(n, n, device="cuda"); G.normal_(generator=g)
ws = torch.empty(lib.hb_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
one = torch.ones(1, device="cuda"); gl = torch.zeros(1, device="cuda")
for rep in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    assert lib.hb_rbf_gram_fwd(P(X), None, n, n, D, 1, P(ell), 1, P(K), n, 0, 1e-5, 1, 0, ST()) == 0
    e[1].record()
    assert lib.hb_rbf_gram_bwd(P(G), n, 0, P(X), None, n, n, D, 1, P(ell), 1, 1, 0, P(one), P(gl), P(ws), ws.numel(), ST()) == 0
    e[2].record()
    torch.cuda.synchronize()
    print(f"n={n} gram fwd {e[0].elapsed_time(e[1]):.3f} ms ({2.0 * n * n / 1e6 / e[0].elapsed_time(e[1]):.0f} GB/s of n^2/2 x 4 B written), "
          f"bwd {e[1].elapsed_time(e[2]):.3f} ms ({2.0 * n * n / 1e6 / e[1].elapsed_time(e[2]):.0f} GB/s read), g_ell {gl.item():.6e}", flush=True)
