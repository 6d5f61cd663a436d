"""Raster-group sweep of the pre-split pair kernel at the top-level shapes of the N=65536 factorisation."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from henbun_b200 import _lib
lib = _lib.load(); P, ST = _lib.ptr, _lib.stream
shapes = [(32768, 32768, 32768, 0, 1, 1), (32768, 32768, 32768, 0, 0, 0), (32768, 32768, 32768, 1, 0, 0), (16384, 16384, 16384, 0, 1, 0),
          (49152, 16384, 16384, 0, 0, 0)]
for (M, N, K, tA, tB, ctri) in shapes:
    g = torch.Generator("cuda").manual_seed(1)
    A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g)
    Cm = torch.zeros(M, N, device="cuda")
    wsb = lib.hb_gemm_presplit_workspace_bytes(M, N, K, tA, tB); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    lib.hb_gemm_presplit(P(A), A.shape[1], tA, P(B), B.shape[1], tB, P(Cm), N, ctri, M, N, K, 1.0, 1.0, 0, 0, 0, P(ws), wsb, ST())
    out = []
    for grp in (2, 4, 8, 16, 32):
        _lib.OPTIONS.tc_option = grp << 8
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.hb_gemm_presplit(P(A), A.shape[1], tA, P(B), B.shape[1], tB, P(Cm), N, ctri, M, N, K, -1.0, 1.0, 0, 0, 1, P(ws), wsb, ST())
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            lib.hb_gemm_presplit(P(A), A.shape[1], tA, P(B), B.shape[1], tB, P(Cm), N, ctri, M, N, K, -1.0, 1.0, 0, 0, 1, P(ws), wsb, ST())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        fl = 2.0 * M * N * K * (0.5 if ctri else 1.0)
        out.append(f"g{grp}: {ms:.1f} ms {fl / ms / 1e9:.0f} TF/s")
    _lib.OPTIONS.tc_option = 0
    print(f"M={M} N={N} K={K} tA={tA} tB={tB} c_tri={ctri}: " + " | ".join(out), flush=True)
    del A, B, Cm, ws
