"""Bring-up probe for the tcgen05 3xTF32 GEMM: accuracy vs fp64 on several shapes + timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from henbun_b200 import _lib
lib = _lib.load()
P, ST = _lib.ptr, _lib.stream

def run(M, N, K, alpha=1.0, beta=0.0, c_tri=0, opt=0, ldpad=0, same=False, reps=0):
    rng = np.random.RandomState(M + 3 * N + 7 * K)
    A = rng.randn(M, K + ldpad).astype(np.float32); B = A if same else rng.randn(N, K + ldpad).astype(np.float32)
    C0 = rng.randn(M, N).astype(np.float32)
    Ad = torch.from_numpy(A).cuda(); Bd = Ad if same else torch.from_numpy(B).cuda(); Cd = torch.from_numpy(C0).cuda()
    wsb = lib.hb_gemm_tc_workspace_bytes(M, N, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    lib.hb_set_tc_option(opt)
    rc = lib.hb_gemm_tn_tc(P(Ad), K + ldpad, P(Bd), K + ldpad, P(Cd), N, c_tri, M, N, K, alpha, beta, P(ws), wsb, ST())
    torch.cuda.synchronize()
    ref = alpha * A[:, :K].astype(np.float64) @ B[:, :K].astype(np.float64).T + beta * C0
    out = Cd.cpu().numpy().astype(np.float64)
    if c_tri:
        iu = np.triu_indices(M, 1, N)
        untouched = np.array_equal(out[iu], C0.astype(np.float64)[iu])
        out = np.tril(out); ref = np.tril(ref)
    else:
        untouched = True
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    # fp32 SIMT reference error for comparison
    Cs = torch.from_numpy(C0).cuda()
    lib.hb_set_gemm_engine(1)
    lib.hb_gemm(P(Ad), K + ldpad, 0, 0, 0, P(Bd), K + ldpad, 0, 1, 0, P(Cs), N, 0, c_tri, M, N, K, 1, alpha, beta, None, 0, 0, 0, -50.0, 50.0, ST())
    lib.hb_set_gemm_engine(0)
    torch.cuda.synchronize()
    outs = Cs.cpu().numpy().astype(np.float64)
    if c_tri: outs = np.tril(outs)
    errs = np.linalg.norm(outs - ref) / np.linalg.norm(ref)
    msg = f"M={M} N={N} K={K} a={alpha} b={beta} tri={c_tri} opt={opt} same={same} rc={rc} rel_err_tc={err:.3e} rel_err_simt={errs:.3e} upper_untouched={untouched}"
    if reps:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lib.hb_gemm_tn_tc(P(Ad), K + ldpad, P(Bd), K + ldpad, P(Cd), N, c_tri, M, N, K, alpha, 0.0, P(ws), wsb, ST())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * M * N * K * (0.5 if c_tri else 1.0)
        msg += f"  {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (incl. split)"
    print(msg, flush=True)

if __name__ == "__main__":
    run(128, 128, 32)
    run(128, 128, 64)
    run(128, 256, 128)
    run(256, 256, 256, opt=1)        # raw fp32 as the hi operand: does the MMA truncate?
    run(256, 256, 256, opt=0)
    run(300, 500, 100, alpha=-1.0, beta=1.0)
    run(384, 384, 96, c_tri=1, alpha=-1.0, beta=1.0, same=True)
    run(1000, 1000, 1000, ldpad=24)
    run(4096, 4096, 4096, reps=5)
    run(8192, 8192, 8192, reps=3, c_tri=1, same=True, alpha=-1.0)
    run(16384, 16384, 2048, reps=3)
    run(8192, 8192, 128, reps=5)
