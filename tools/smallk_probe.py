"""The launch classes of the factorisations' tail (profiles/r2_gemm_classes_n65536.txt) on each engine: hb_gemm_ws with
hb_options.gemm_engine = 1 (fp32 SIMT: short-K whole-K kernel / k-looped kernel) vs 2 (tcgen05 in-kernel split)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from henbun_b200 import _lib
lib = _lib.load(); P, ST = _lib.ptr, _lib.stream
ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")


def bench(M, N, K, tA, tB, inplace, engine, reps=20):
    g = torch.Generator("cuda").manual_seed(1)
    A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g)
    Cm = A if inplace else torch.zeros(M, N, device="cuda")
    o = _lib.Options(); lib.hb_options_init(C.byref(o)); o.gemm_engine = engine
    def run():
        return lib.hb_gemm_ws(P(A), A.shape[1], 0, tA, 0, P(B), B.shape[1], 0, tB, 0, P(Cm), Cm.shape[1], 0, 0, M, N, K, 1, 1.0,
                              0.0 if inplace else 1.0, None, 0, 0, 0, -50.0, 50.0, P(ws), ws.numel(), ST(), C.byref(o))
    rc = run(); torch.cuda.synchronize()
    if rc != 0:
        return None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


cases = [("panel solve (in place)", 60000, 128, 128, 0, 1, True), ("panel solve (in place)", 30000, 128, 128, 0, 1, True),
         ("panel solve (in place)", 8000, 128, 128, 0, 1, True),
         ("K=128 update", 60000, 128, 128, 0, 1, False), ("K=128 update", 30000, 128, 128, 0, 0, False),
         ("N=128, K=512", 60000, 128, 512, 0, 1, False), ("N=128, K=1024", 30000, 128, 1024, 0, 1, False),
         ("tall 128x128", 128, 128, 60000, 1, 0, False), ("tall 128x128", 128, 128, 20000, 1, 0, False),
         ("tall 256x256", 256, 256, 50000, 1, 0, False), ("256^3", 256, 256, 256, 0, 0, False), ("512^3", 512, 512, 512, 0, 0, False)]
for name, M, N, K, tA, tB, ip in cases:
    t1 = bench(M, N, K, tA, tB, ip, 1); t2 = bench(M, N, K, tA, tB, ip, 2); t0 = bench(M, N, K, tA, tB, ip, 0)
    f = lambda t: "   n/a" if t is None else f"{t:6.1f}"
    print(f"{name:24s} M={M:6d} N={N:4d} K={K:6d}: SIMT {f(t1)} us | tcgen05 {f(t2)} us | auto {f(t0)} us", flush=True)
