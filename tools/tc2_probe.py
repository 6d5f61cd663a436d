"""Generation-2 tcgen05 engine: correctness on all forms + timing against generation 1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from henbun_b200 import _lib
from tools.tc_forms_probe import run, lib, P, ST

def timeit(M, N, K, tA, tB, a_tri=0, c_tri=0, opt=0, reps=5, beta=0.0):
    g = torch.Generator("cuda").manual_seed(1)
    A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g)
    Cd = torch.zeros(M, N, device="cuda")
    wsb = lib.hb_gemm_tc_workspace_bytes(M, N, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    lib.hb_set_tc_option(opt); lib.hb_set_gemm_engine(2)
    def go():
        return lib.hb_gemm_ws(P(A), A.shape[1], 0, tA, a_tri, P(B), B.shape[1], 0, tB, 0, P(Cd), N, 0, c_tri, M, N, K, 1,
                              1.0, beta, None, 0, 0, 0, -50.0, 50.0, P(ws), wsb, ST())
    rc = go(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    lib.hb_set_tc_option(0); lib.hb_set_gemm_engine(0)
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * M * N * K * (0.5 if (c_tri or a_tri) else 1.0)
    print(f"time M={M} N={N} K={K} tA={tA} tB={tB} a_tri={a_tri} c_tri={c_tri} gen={'1' if opt & 2 else ('2-single' if opt & 4 else '2-pair')}{'-3xtf32' if opt & 8 else '-bf16x'}{'-nonpers' if opt & 16 else ''} rc={rc}: {ms:.3f} ms {fl / ms / 1e9:.1f} TF/s", flush=True)

if __name__ == "__main__":
    worst = 0.0
    for (tA, tB) in ((0, 1), (0, 0), (1, 0), (1, 1)):
        worst = max(worst, run(128, 128, 32, tA, tB))
        worst = max(worst, run(256, 384, 200, tA, tB))
        worst = max(worst, run(300, 500, 100, tA, tB, alpha=-1.0, beta=1.0))
        worst = max(worst, run(1000, 100, 333, tA, tB))
    for a_tri in (1, 2, 3, 4):
        worst = max(worst, run(512, 384, 512, 0, 0, a_tri=a_tri, alpha=-2.0, beta=1.0))
        worst = max(worst, run(512, 384, 512, 1, 0, a_tri=a_tri, alpha=-2.0, beta=1.0))
    for b_tri in (1, 2, 3, 4):
        worst = max(worst, run(384, 512, 512, 0, 1, b_tri=b_tri))
        worst = max(worst, run(384, 512, 512, 0, 0, b_tri=b_tri))
    worst = max(worst, run(640, 640, 384, 1, 0, c_tri=1, alpha=-1.0, beta=1.0))
    worst = max(worst, run(640, 640, 384, 0, 1, c_tri=1, alpha=-1.0, beta=1.0))
    worst = max(worst, run(1000, 1000, 1000, 1, 0, a_tri=3))
    worst = max(worst, run(2048, 2048, 2048, 0, 1))
    print("worst tc2 err (1-CTA shapes)", worst, flush=True)
    worst = 0.0
    for (tA, tB) in ((0, 1), (0, 0), (1, 0), (1, 1)):      # CTA-pair kernel: >= 64 tiles of 256 x 256
        worst = max(worst, run(2100, 2180, 520, tA, tB, alpha=-1.0, beta=1.0))
    for a_tri in (1, 2, 3, 4):
        worst = max(worst, run(2048, 2048, 2048, 0, 0, a_tri=a_tri, alpha=-2.0, beta=1.0))
        worst = max(worst, run(2048, 2048, 2048, 1, 0, a_tri=a_tri))
    for b_tri in (1, 2, 3, 4):
        worst = max(worst, run(2048, 2304, 2304, 0, 1, b_tri=b_tri))
        worst = max(worst, run(2048, 2304, 2304, 0, 0, b_tri=b_tri))
    worst = max(worst, run(3000, 3000, 700, 1, 0, c_tri=1, alpha=-1.0, beta=1.0))
    worst = max(worst, run(3000, 3000, 700, 0, 1, c_tri=1, alpha=-1.0, beta=1.0))
    print("worst tc2 err (pair shapes)", worst, flush=True)
    for opt in (0,):
        timeit(65280, 256, 256, 0, 1, opt=opt, reps=20, beta=1.0)
        timeit(32768, 512, 512, 0, 1, opt=opt, reps=20, beta=1.0)
        timeit(32768, 1024, 1024, 0, 0, opt=opt, reps=10, beta=1.0)
        timeit(16384, 2048, 2048, 1, 0, opt=opt, reps=10, beta=1.0)
        timeit(4096, 4096, 4096, 0, 1, opt=opt)
        timeit(8192, 8192, 8192, 0, 1, c_tri=1, opt=opt)
        timeit(8192, 8192, 8192, 1, 0, opt=opt)
        timeit(8192, 8192, 8192, 0, 0, a_tri=1, opt=opt)
        timeit(16384, 16384, 4096, 0, 1, opt=opt, reps=3)
        timeit(16384, 128, 128, 0, 1, opt=opt, reps=20)
        timeit(2048, 2048, 512, 0, 1, opt=opt, reps=20)
        timeit(1024, 1024, 1024, 0, 0, opt=opt, reps=20)
