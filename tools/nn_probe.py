"""BASELINE config 4 encoder (784-512-512-128, sigmoid) through ops.matbias: parity vs torch fp64 and timing per
engine (fwd + bwd of the three MatBias layers at B*S rows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from henbun_b200 import _lib, ops
lib = _lib.load()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator("cuda").manual_seed(0)
dims = [784, 512, 512, 128]
x = torch.randn(rows, dims[0], device="cuda", generator=g)
Ws = [(torch.randn(dims[i], dims[i + 1], device="cuda", generator=g) / dims[i] ** 0.5).requires_grad_(True) for i in range(3)]
bs = [torch.zeros(1, dims[i + 1], device="cuda").requires_grad_(True) for i in range(3)]
gout = torch.randn(rows, dims[-1], device="cuda", generator=g)

def fwd_bwd():
    y = x
    for i in range(3):
        y = ops.matbias(y, Ws[i], bs[i], act="sigmoid" if i < 2 else "none")
    y.backward(gout)
    return y

def ref():
    y = x.double()
    W64 = [w.detach().double().requires_grad_(True) for w in Ws]; b64 = [b.detach().double().requires_grad_(True) for b in bs]
    for i in range(3):
        y = y @ W64[i] + b64[i]
        if i < 2: y = torch.sigmoid(y)
    y.backward(gout.double())
    return y, W64, b64

yr, W64, b64 = ref()
flops = 3 * 2.0 * rows * sum(dims[i] * dims[i + 1] for i in range(3)) - 2.0 * rows * dims[0] * dims[1]   # no dx of layer 0
for engine, name in ((1, "simt"), (0, "auto")):
    lib.hb_set_gemm_engine(engine)
    for w in Ws + bs: w.grad = None
    y = fwd_bwd(); torch.cuda.synchronize()
    ey = (torch.linalg.norm(y.double() - yr) / torch.linalg.norm(yr)).item()
    ew = max((torch.linalg.norm(Ws[i].grad.double() - W64[i].grad) / torch.linalg.norm(W64[i].grad)).item() for i in range(3))
    eb = max((torch.linalg.norm(bs[i].grad.double() - b64[i].grad) / torch.linalg.norm(b64[i].grad)).item() for i in range(3))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        for w in Ws + bs: w.grad = None
        fwd_bwd()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"rows={rows} engine={name}: err y {ey:.2e} dW {ew:.2e} db {eb:.2e}  {ms:.3f} ms  {flops / ms / 1e9:.1f} TF/s", flush=True)
lib.hb_set_gemm_engine(0)
