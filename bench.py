#!/usr/bin/env python
"""Headline benchmark: ELBO + gradient (+ TF-1 Adam update) evaluations per second of the
variational GP of BASELINE.json config 3 (N=65536, D=8, S=64 MC samples, RBF kernel, blocked
Cholesky), on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path

A step = one pass of the hot path on one batch of synthetic inputs: K(X,X)+jI -> Cholesky ->
reparameterised sampler + one-sample KL -> F = sqrt(k_var) L (s z) -> Gaussian log-likelihood ->
full backward (incl. reverse-mode Cholesky and the lengthscale gradient) -> Adam.
evals = S * N per step.  Multi-GPU: weak scaling in S (each rank draws its own S samples and
replicates the factorisation), one NCCL all-reduce of the packed gradient per step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "elbo_grad_evals_per_sec"
UNIT = "evals/s (MC samples x data points / s)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--dim", type=int, default=8)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--cpu-n", type=int, default=16384, help="size of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engine", type=int, default=0, help="GEMM engine: 0 auto, 1 SIMT fp32, 2 tcgen05 3xTF32")
    return ap.parse_args()


def extrapolation_factor(n_full, n_sample, S):
    """CPU time at n_full / CPU time at n_sample from the step's flop counts: n^3 (Cholesky n^3/3 + its reverse mode
    2n^3/3) scales cubically, the 6 n^2 S of the three sample projections quadratically."""
    r = n_full / n_sample
    q = 6.0 * S / (6.0 * S + n_sample)          # quadratic share at the sample size
    return (1.0 - q) * r ** 3 + q * r ** 2


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1 to every rank: override it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def cpu_baseline(n_full, D, S, n_sample, steps=2, warmup=1):
    """Oracle (torch-CPU fp32) ELBO+grad+Adam step on a bounded sample, extrapolated by flop count."""
    import torch
    from oracle import cpu_baseline as cb
    n_sample = min(n_sample, n_full)
    if n_sample >= 8192:
        steps = 1
    t, _, threads = cb.time_gpr_steps(n_sample, D, S, steps=steps, warmup=warmup, threads=host_threads())
    scale = extrapolation_factor(n_full, n_sample, S)
    t_full = t * scale
    return {
        "value": S * n_full / t_full, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": (f"torch-CPU fp32 oracle, full ELBO+grad+Adam step at N={n_sample}, D={D}, S={S}: "
                   f"{t:.3f} s/step (median of {steps}); extrapolated x{scale:.1f} to N={n_full} by flop count "
                   f"(n^3 Cholesky + reverse mode cubically, 6 n^2 S sample projections quadratically)"),
        "sample_evals_per_sec": S * n_sample / t, "sample_s_per_step": t,
    }


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.idx = gpu_index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(pw)), "samples": len(sm)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    times = []
    from oracle import cpu_baseline as cb
    n_s = min(a.cpu_n, a.n)
    # same job as our arm at --gpus N: weak scaling in S, i.e. S * N samples per step on the one host
    S_tot = a.samples * max(1, a.gpus)
    t, _, threads = cb.time_gpr_steps(n_s, a.dim, S_tot, steps=max(1, min(a.steps, 5)), warmup=max(1, min(a.warmup, 1)),
                                      threads=host_threads())
    scale = extrapolation_factor(a.n, n_s, S_tot)
    val = S_tot * a.n / (t * scale)
    sample = (f"torch-CPU fp32 restatement of the reference graph (TensorFlow is not installable here), "
              f"ELBO+grad+Adam at N={n_s}, S={S_tot}: {t:.3f} s/step, extrapolated x{scale:.1f} to N={a.n} by flop count "
              f"(n^3 cubically, 6 n^2 S quadratically)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t * scale, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"variational GP regression N={a.n} D={a.dim} S={a.samples}/GPU x {max(1, a.gpus)} RBF mean-field q, "
                               f"ELBO+grad+Adam (BASELINE config 3)",
                   "timing": "host wall clock, CPU only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    from henbun_b200 import _lib
    from oracle import cpu_baseline as cb       # only for the synthetic problem generator + CPU leg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    lib.hb_set_gemm_engine(a.engine)

    n, D, S = a.n, a.dim, a.samples
    X, Y, p = cb.make_gp_problem(n, D, S, seed=0)
    order = ("q_mu", "q_sqrt", "scale", "lengthscales", "k_var", "var")
    params_h = np.concatenate([np.asarray(p[k], np.float32).ravel() for k in order])
    cfg = _lib.GpConfig(n, D, S, 1, 0, 1e-5, 1000 + rank, 0)
    npar = lib.hb_gp_param_count(C.byref(cfg))
    assert npar == params_h.size
    dev = torch.device("cuda", local)
    params = torch.from_numpy(params_h).to(dev)
    grads = torch.zeros(npar, device=dev)
    adam_m = torch.zeros(npar, device=dev); adam_v = torch.zeros(npar, device=dev)
    step_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
    out4 = torch.zeros(4, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    wsb = lib.hb_gp_elbo_workspace_bytes(C.byref(cfg))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    Xd = torch.from_numpy(X).to(dev); Yd = torch.from_numpy(Y).to(dev)
    Xh = torch.from_numpy(X).pin_memory(); Yh = torch.from_numpy(Y).pin_memory()
    out_h = torch.zeros(4).pin_memory()
    st = _lib.stream

    def step(it, host_io=False):
        if host_io:      # the reference feeds Data through feed_dict on every session.run (param.py:701-705)
            Xd.copy_(Xh, non_blocking=True); Yd.copy_(Yh, non_blocking=True)
        cfg.offset = C.c_ulonglong(it * ((S * n + 3) // 4 * 4))
        _lib.check(lib.hb_gp_elbo_step(C.byref(cfg), _lib.ptr(Xd), _lib.ptr(Yd), _lib.ptr(params), None, _lib.ptr(grads),
                                       _lib.ptr(out4), _lib.ptr(ws), wsb, _lib.ptr(err), st()), "hb_gp_elbo_step")
        if world > 1:
            dist.all_reduce(grads)          # one NCCL all-reduce of the packed gradient (sum); mean below
        _lib.check(lib.hb_increment_i32(_lib.ptr(step_ctr), st()), "hb_increment_i32")
        _lib.check(lib.hb_adam_tf1(_lib.ptr(params), _lib.ptr(grads), _lib.ptr(adam_m), _lib.ptr(adam_v), npar,
                                   -1.0 / world, 1e-3, 0.9, 0.999, 1e-8, _lib.ptr(step_ctr), 0, st()), "hb_adam_tf1")
        if host_io:
            out_h.copy_(out4, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, host_io, profile=False):
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = lib.hb_launch_count()
        if profile:
            lib.hb_profile_begin(200000)
        ev0.record()
        for i in range(nsteps):
            step(timed.it, host_io); timed.it += 1
        ev1.record()
        barrier()
        prof = None
        if profile:
            buf = (C.c_double * 8)()
            lib.hb_profile_end_ex(buf)
            prof = list(buf)
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
        return ms, lib.hb_launch_count() - l0, prof
    timed.it = 0

    for _ in range(a.warmup):
        step(timed.it); timed.it += 1
    barrier()
    if err.item() != 0:
        raise RuntimeError(f"Cholesky failed: non-positive pivot at row {err.item() - 1}")

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, launches, _ = timed(a.steps, host_io=False)
    clk = clocks.stop() if rank == 0 else None
    ms_e2e, _, _ = timed(a.steps, host_io=True)
    # separate, untimed passes for the evidence: per-launch GEMM events (roofline) and per-phase events
    ms_prof, _, prof = timed(1, host_io=False, profile=True)
    phases = None
    if True:                                  # every rank steps (the step holds a collective); rank 0 reports
        lib.hb_phase_begin()
        step(timed.it); timed.it += 1
        pbuf = (C.c_double * 16)()
        npz = lib.hb_phase_end(pbuf, 16)
        names = ["scalars+gram_fwd", "potrf", "sampler+F+loglik+W", "sampler_bwd+Lbar", "potrf_bwd", "gram_bwd+scalar_grads"]
        phases = {names[i] if i < len(names) else f"phase{i}": round(pbuf[i], 3) for i in range(max(npz, 0))}
    barrier()
    elbo = out4[0].item()
    if not math.isfinite(elbo) or err.item() != 0:
        raise RuntimeError(f"non-finite ELBO {elbo} / err flag {err.item()}")

    if rank == 0:
        evals = float(S) * n * world
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained")
        peak_src = "measured bf16 dense sustained (MEASURED_PEAKS.json)"
        if peak is None:
            peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
        n_gemm, gemm_ms, gemm_flop, n_pair, pair_ms, pair_flop, h2_ms, h2_flop = prof[:8]
        traffic = None
        try:    # DRAM bytes of one captured launch of the dominant kernel (ncu --set full, summary committed in profiles/)
            tr = {}
            for ln in open(os.path.join(ROOT, "profiles", "r1_ncu_prof_tc2_pair_bfx_r1.csv")):
                c = ln.strip().split(",")
                if len(c) >= 3 and c[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tr[c[0]] = float(c[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(c[2], 1.0)
            if len(tr) == 2:
                traffic = {"bytes_per_launch": sum(tr.values()), "launch": "gemm_tc2_pair_kernel, M=N=K=8192 (tools/gemm_one.py)",
                           "algorithmic_bytes": 3 * 8192 * 8192 * 4, "source": "profiles/r1_ncu_prof_tc2_pair_bfx_r1.csv"}
        except Exception:
            traffic = None
        achieved = gemm_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        pair_achieved = pair_flop / (pair_ms * 1e-3) / 1e12 if pair_ms > 0 else 0.0
        h2_achieved = h2_flop / (h2_ms * 1e-3) / 1e12 if h2_ms > 0 else 0.0
        eng = lib.hb_get_gemm_engine()
        line = {
            "metric": METRIC, "value": evals * a.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"variational GP regression N={n} D={D} S={S}/GPU RBF mean-field q, ELBO+grad+Adam (BASELINE config 3)",
                "parallelism": f"replicated factorisation, sample-sharded (weak scaling in S), {world} rank(s)",
                "l2": "inputs larger than L2 (K/L and Lbar/Kbar are N^2 fp32 = %.1f GB each)" % (4.0 * n * n / 1e9),
                "eps": "device Philox-4x32-10, regenerated in the backward",
                "gemm_engine": {0: "auto", 1: "simt-fp32", 2: "tcgen05-3xtf32"}[eng],
                "elbo_last": elbo,
            },
            "e2e": {"value": evals * a.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(Xh.numel() * 4 + Yh.numel() * 4), "d2h_bytes_per_step": 16,
                    "ms_per_step": ms_e2e / a.steps,
                    "api": "hb_gp_elbo_step + hb_adam_tf1 through the C ABI, X/Y fed from pinned host memory every step"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {
                "bound": "tensor", "kernel": "hb::gemm = gemm_tc2_pair_kernel / gemm_tc2_kernel (tcgen05 3xTF32) + short-K SIMT kernel: all level-3 work of potrf / potrf_bwd / sample projections",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": traffic, "peak_source": peak_src,
                "gemm_launches": int(n_gemm), "gemm_ms_per_step": gemm_ms,
                "gemm_share_of_step": gemm_ms / ms_prof if ms_prof > 0 else None,
                "useful_gemm_flop_per_step": gemm_flop,
                "pair_kernel": {"kernel": "gemm_tc2_pair_kernel (tcgen05 cta_group::2, 256x256 tiles)", "launches": int(n_pair),
                                "ms_per_step": pair_ms, "useful_flop_per_step": pair_flop, "achieved": pair_achieved,
                                "frac": pair_achieved / peak if peak else None,
                                "share_of_step": pair_ms / ms_prof if ms_prof > 0 else None},
                "presplit_kernel": {"kernel": "gemm_h2_pair_kernel (tcgen05 cta_group::2, fp16 hi/lo shadows, 3 MMAs per k-step)",
                                    "ms_per_step": h2_ms, "useful_flop_per_step": h2_flop, "achieved": h2_achieved,
                                    "frac": h2_achieved / peak if peak else None,
                                    "share_of_step": h2_ms / ms_prof if ms_prof > 0 else None},
                "note": ("fp32 parity (1e-5) needs a split product: hi*hi in TF32 (two K=8 MMAs per 16-wide k-block) + lo*hi and hi*lo "
                         "in bf16 (one K=16 MMA each) = 4 tensor-pipe slots where a bf16 GEMM needs 1, i.e. frac <= 0.25 against the "
                         "bf16 peak for this formulation"),
                "how": "CUDA-event pair around every GEMM launch of one extra (untimed) step; achieved = useful FLOP / summed launch time",
            },
            "phases_ms": phases,
        }
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(n, D, S, a.cpu_n)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
