"""Phase cycle counts of the one-CTA fp32 GP step (gp_leaf_step_kernel writes them behind the workspace's 4 S n elements)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from henbun_b200 import _lib
lib = _lib.load(); P, ST = _lib.ptr, _lib.stream
n, D, S, full = 100, 1, 10, 1
rng = np.random.RandomState(0)
X = np.linspace(0, 6, n).reshape(-1, 1).astype(np.float32); Y = (np.sin(X[:, 0]) + 0.3 * rng.randn(n)).astype(np.float32)
cfg = _lib.GpConfig(n, D, S, 1, full, 1e-3, 0, 0)
npar = lib.hb_gp_param_count(C.byref(cfg))
q_sqrt = (0.3 * np.eye(n) + 0.02 * np.tril(rng.randn(n, n))).astype(np.float32)
params = torch.tensor(np.concatenate([0.1 * rng.randn(n), q_sqrt.ravel(), [0.54], [0.54], [0.54], [-0.5]]).astype(np.float32), device="cuda")
grads = torch.zeros(npar, device="cuda"); out4 = torch.zeros(4, device="cuda"); err = torch.zeros(1, dtype=torch.int32, device="cuda")
wsb = lib.hb_gp_small_workspace_bytes(C.byref(cfg), 0); ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
Xd, Yd = torch.tensor(X, device="cuda"), torch.tensor(Y, device="cuda")
for _ in range(5):
    _lib.check(lib.hb_gp_small_step(C.byref(cfg), P(Xd), P(Yd), P(params), None, P(grads), P(out4), None, None, None, P(ws), wsb, P(err), ST()), "step")
torch.cuda.synchronize()
base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
st = ws[base:].view(torch.float32)[4 * S * n: 4 * S * n + 8].cpu().numpy()
names = ["hyper + Gram", "potrf", "sampler .. sampler bwd", "trinv", "L-bar", "3 products", "Gram bwd + scalars"]
prev = 0.0
for nm, v in zip(names, st):
    print(f"{nm:28s} {v - prev:10.0f} cycles"); prev = v
print("total", st[6], "cycles; ELBO", out4[0].item())
