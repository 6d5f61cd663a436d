"""Pre-split fp16 hi/lo engine (csrc/gemm_h2.cu) through hb_gemm_presplit: parity vs fp64 on every operand layout,
block mask and scale mode, then timing against the in-kernel-split pair kernel (hb_gemm_ws) on the same shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from henbun_b200 import _lib
lib = _lib.load()
P, ST = _lib.ptr, _lib.stream


def run(M, N, K, tA, tB, c_tri=0, bmode=0, blockscale=0, alpha=1.0, beta=0.0, wide=False, seed=0):
    g = torch.Generator("cuda").manual_seed(seed)
    A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g)
    if wide:   # per-column-block magnitudes spread over 2^40, plus tiny entries
        ca = A.shape[1]
        A *= torch.exp2(torch.randint(-20, 20, (ca // 128,), device="cuda", generator=g).float()).repeat_interleave(128)[None, :]
    C0 = torch.randn(M, N, device="cuda", generator=g)
    C = C0.clone()
    wsb = lib.hb_gemm_presplit_workspace_bytes(M, N, K, tA, tB)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    rc = lib.hb_gemm_presplit(P(A), A.shape[1], tA, P(B), B.shape[1], tB, P(C), N, c_tri, M, N, K, alpha, beta, bmode, blockscale,
                              0, P(ws), wsb, ST())
    torch.cuda.synchronize()
    assert rc == 0, rc
    opA = (A.T if tA else A).double(); opB = (B.T if tB else B).double()
    if bmode:
        rb = torch.arange(M, device="cuda") // 128; kb = torch.arange(K, device="cuda") // 128
        keep = (kb[None, :] <= rb[:, None]) if bmode == 1 else (kb[None, :] > rb[:, None])
        opA = opA * keep
    ref = alpha * opA @ opB + beta * C0.double()
    got = C.double()
    if c_tri:
        tri = torch.tril(torch.ones(M, N, device="cuda", dtype=torch.bool))
        # tiles entirely above the diagonal are skipped (C0 kept), inside a computed tile everything above is left alone too
        err = torch.linalg.norm((got - ref)[tri]) / torch.linalg.norm(ref[tri])
    else:
        err = torch.linalg.norm(got - ref) / torch.linalg.norm(ref)
    return err.item()


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ok = True
for (M, N, K) in ((512, 512, 256), (2048, 2304, 1024), (4096 + 256, 2048, 640)):
    for tA in (0, 1):
        for tB in (0, 1):
            e = run(M, N, K, tA, tB)
            print(f"M={M} N={N} K={K} tA={tA} tB={tB}: rel err {e:.2e}", flush=True)
            ok &= e < 1e-6
e = run(2048, 2048, 2048, 0, 1, c_tri=1, alpha=-1.0, beta=1.0); print(f"c_tri, alpha=-1, beta=1: {e:.2e}"); ok &= e < 1e-6
for tA in (0, 1):
    for tB in (0, 1):
        e = run(2048, 2048, 2048, tA, tB, blockscale=1, wide=True); print(f"blockscale wide tA={tA} tB={tB}: {e:.2e}"); ok &= e < 1e-6
e = run(2048, 1024, 2048, 0, 0, bmode=1, blockscale=1, alpha=-2.0, beta=1.0); print(f"bmode 1 (K-major A): {e:.2e}"); ok &= e < 1e-6
e = run(2048, 1024, 2048, 1, 0, bmode=2, blockscale=1, alpha=-2.0, beta=1.0); print(f"bmode 2 (MN-major A): {e:.2e}"); ok &= e < 1e-6
e = run(2304, 1024, 2304, 0, 0, bmode=1); print(f"bmode 1, 18 blocks: {e:.2e}"); ok &= e < 1e-6
e = run(2304, 1024, 2304, 1, 0, bmode=2); print(f"bmode 2, 18 blocks: {e:.2e}"); ok &= e < 1e-6
for (M, N, K) in ((4096, 64, 2048), (2048 + 128, 32, 1024), (8192, 64, 640)):       # 256 x 64 pair tiles (B K-major)
    for tA in (0, 1):
        e = run(M, N, K, tA, 1); print(f"narrow M={M} N={N} K={K} tA={tA} tB=1: {e:.2e}", flush=True); ok &= e < 1e-6
print("PARITY", "OK" if ok else "FAILED", flush=True)

# timing
if len(sys.argv) > 1 and sys.argv[1] == "parity":
    sys.exit(0 if ok else 1)
for (M, N, K) in ((8192, 8192, 8192), (16384, 16384, 4096), (32768, 8192, 2048), (60000, 1024, 1024), (60000, 2048, 512)):
    for tA, tB in ((0, 1), (0, 0), (1, 0)):
        g = torch.Generator("cuda").manual_seed(1)
        A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g)
        B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g)
        C = torch.zeros(M, N, device="cuda")
        wsb = lib.hb_gemm_presplit_workspace_bytes(M, N, K, tA, tB)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        lib.hb_gemm_presplit(P(A), A.shape[1], tA, P(B), B.shape[1], tB, P(C), N, 0, M, N, K, 1.0, 1.0, 0, 0, 0, P(ws), wsb, ST())
        t_new = timeit(lambda: lib.hb_gemm_presplit(P(A), A.shape[1], tA, P(B), B.shape[1], tB, P(C), N, 0, M, N, K, -1.0, 1.0, 0, 0, 1,
                                                    P(ws), wsb, ST()))
        t_all = timeit(lambda: lib.hb_gemm_presplit(P(A), A.shape[1], tA, P(B), B.shape[1], tB, P(C), N, 0, M, N, K, -1.0, 1.0, 0, 0, 0,
                                                    P(ws), wsb, ST()))
        t_old = timeit(lambda: lib.hb_gemm_ws(P(A), A.shape[1], 0, tA, 0, P(B), B.shape[1], 0, tB, 0, P(C), N, 0, 0, M, N, K, 1, -1.0, 1.0,
                                              None, 0, 0, 0, -50.0, 50.0, None, 0, ST()))
        fl = 2.0 * M * N * K
        print(f"M={M} N={N} K={K} tA={tA} tB={tB}: presplit {t_new:.3f} ms = {fl / t_new / 1e9:.0f} TF/s (with split passes "
              f"{t_all:.3f} ms); in-kernel split {t_old:.3f} ms = {fl / t_old / 1e9:.0f} TF/s", flush=True)
        del A, B, C, ws
