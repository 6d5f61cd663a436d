"""BASELINE configs 1, 2, 4 and 5 at full size through the Henbun-shaped Python API (model.optimize): ms per Adam step.
Not bench lines (bench.py measures config 3) -- evidence that the other configs run at their named sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import henbun_b200 as hb
import henbun_b200.tf as tf
which = sys.argv[1] if len(sys.argv) > 1 else "5"


def timed(opt, steps, **kw):
    opt.optimize(maxiter=3, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    opt.optimize(maxiter=steps, **kw)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


if which in ("1", "2"):
    # config 1: notebooks/GaussianProcess.ipynb:109-148 at N=100, full-covariance q, S=10
    # config 2: notebooks/Expert_GPR.ipynb:101-149 at N=2000, 3 experts (q_s, q_l full-covariance, q_r mean-field)
    rng = np.random.RandomState(0)
    if which == "1":
        n, S, jitter = 100, 10, 1e-5
        X = np.linspace(0, 6, n).reshape(-1, 1); Y = np.sin(X) + 0.3 * rng.randn(n, 1)

        class GPR(hb.model.Model):
            def setUp(self):
                self.X = hb.param.Data(X); self.Y = hb.param.Data(Y)
                self.q = hb.variationals.Gaussian(shape=X.shape, q_shape='fullrank')
                self.kern = hb.gp.kernels.UnitRBF()
                self.k_var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)
                self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)

            @hb.model.AutoOptimize()
            def ELBO(self):
                y_fit = tf.matmul(self.kern.Cholesky(self.X), self.q) * tf.sqrt(self.k_var)
                return tf.reduce_sum(hb.densities.gaussian(self.Y, y_fit, self.var)) - self.KL()
        m = GPR()
    else:
        n, S, jitter = 2000, 10, 3e-4
        X = np.linspace(0, 6, n).reshape(-1, 1); Y = np.sin(0.1 * X ** 3) + 0.1 * rng.randn(n, 1)

        class ExpertGPR(hb.model.Model):
            def setUp(self):
                self.X = hb.param.Data(X); self.Y = hb.param.Data(Y)
                self.q_s = hb.variationals.Gaussian(shape=X.shape, q_shape='fullrank')
                self.q_l = hb.variationals.Gaussian(shape=X.shape, q_shape='fullrank')
                self.q_r = hb.variationals.Gaussian(shape=X.shape, q_shape='diagonal')
                self.kern_s = hb.gp.kernels.UnitRBF(np.ones(1) * 0.2)
                self.kern_l = hb.gp.kernels.UnitRBF(np.ones(1) * 1)
                self.kern_r = hb.gp.kernels.UnitRBF(np.ones(1) * 1)
                self.k_var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)
                self.k_var_r = hb.param.Variable(shape=[1], transform=hb.transforms.positive)
                self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)

            @hb.model.AutoOptimize()
            def ELBO(self):
                f_s = tf.matmul(self.kern_s.Cholesky(self.X), self.q_s)
                f_l = tf.matmul(self.kern_l.Cholesky(self.X), self.q_l)
                f_r = tf.matmul(self.kern_r.Cholesky(self.X), self.q_r) * tf.sqrt(self.k_var_r)
                fraction = tf.sigmoid(f_r)
                f = (fraction * f_s + (1 - fraction) * f_l) * self.k_var
                return tf.reduce_sum(hb.densities.gaussian(self.Y, f, self.var)) - self.KL()
        m = ExpertGPR()
        for nm in ("q_s", "q_l"):
            getattr(m, nm).q_sqrt = 0.3 * np.eye(n) + (0.2 / np.sqrt(n)) * np.tril(rng.randn(n, n))
    cfg = hb.settings.get_settings(); cfg.numerics.jitter_level = jitter
    with hb.settings.temp_settings(cfg):
        m.ELBO().compile(n_samples=S, verbose=False)
    ms = timed(m.ELBO(), 20)
    print(f"config {which} (N={n}, S={S}, jitter {jitter:g}): {ms:.3f} ms/step = {S * n / ms * 1e3:.3e} evals/s; "
          f"ELBO {float(m.ELBO().run()):.4f}", flush=True)
elif which == "1f":
    # config 1 through the fused C entry point (hb_gp_elbo_step + hb_adam_tf1, full-covariance q): eager launches and
    # replay of the same launches captured in a CUDA graph.  SURVEY.md 8d: this config is launch-latency-bound; the
    # floor is one kernel (~5-10 us).  (A captured graph replays the SAME Philox offset -- timing only.)
    import ctypes as C
    from henbun_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(0)
    n, D, S, jitter = 100, 1, 10, 1e-3
    X = np.linspace(0, 6, n).reshape(-1, 1).astype(np.float32); Y = (np.sin(X[:, 0]) + 0.3 * rng.randn(n)).astype(np.float32)
    cfg = _lib.GpConfig(n, D, S, 1, 1, jitter, 0, 0)
    npar = lib.hb_gp_param_count(C.byref(cfg))
    q_sqrt = (0.3 * np.eye(n) + 0.02 * np.tril(rng.randn(n, n))).astype(np.float32)
    params = torch.tensor(np.concatenate([0.1 * rng.randn(n), q_sqrt.ravel(), [0.54], [0.54], [0.54], [-0.5]]).astype(np.float32), device="cuda")
    assert params.numel() == npar
    grads = torch.zeros(npar, device="cuda"); am = torch.zeros(npar, device="cuda"); av = torch.zeros(npar, device="cuda")
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda"); out4 = torch.zeros(4, device="cuda"); err = torch.zeros(1, dtype=torch.int32, device="cuda")
    wsb = lib.hb_gp_elbo_workspace_bytes(C.byref(cfg)); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    Xd, Yd = torch.tensor(X, device="cuda"), torch.tensor(Y, device="cuda")

    def step(it):
        cfg.offset = C.c_ulonglong(it * ((S * n + 3) // 4 * 4))
        _lib.check(lib.hb_gp_elbo_step(C.byref(cfg), _lib.ptr(Xd), _lib.ptr(Yd), _lib.ptr(params), None, _lib.ptr(grads), _lib.ptr(out4),
                                       _lib.ptr(ws), wsb, _lib.ptr(err), _lib.stream()), "hb_gp_elbo_step")
        _lib.check(lib.hb_increment_i32(_lib.ptr(ctr), _lib.stream()), "inc")
        _lib.check(lib.hb_adam_tf1(_lib.ptr(params), _lib.ptr(grads), _lib.ptr(am), _lib.ptr(av), npar, -1.0, 1e-3, 0.9, 0.999, 1e-8,
                                   _lib.ptr(ctr), 0, _lib.stream()), "adam")
    for i in range(20):
        step(i)
    l0 = lib.hb_launch_count(); step(20); per_step = lib.hb_launch_count() - l0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); ev0.record()
    for i in range(200):
        step(21 + i)
    ev1.record(); torch.cuda.synchronize()
    eager = ev0.elapsed_time(ev1) / 200 * 1e3
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(300)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            step(301)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize(); ev0.record()
    for _ in range(500):
        g.replay()
    ev1.record(); torch.cuda.synchronize()
    graph = ev0.elapsed_time(ev1) / 500 * 1e3
    print(f"config 1 fused C entry (N={n}, full-covariance q, S={S}, {per_step} kernels/step): eager {eager:.1f} us/step = "
          f"{S * n / eager * 1e6:.3e} evals/s; CUDA-graph replay {graph:.1f} us/step = {S * n / graph * 1e6:.3e} evals/s; "
          f"ELBO {float(out4[0]):.3f}, err flag {int(err.item())}", flush=True)
    m = None
elif which == "5f":
    # config 5 through the fused C entry points (hb_linop_elbo_local / _update): per-phase CUDA-event times and the
    # HBM roofline of the step (algorithmic bytes: SURVEY.md 8d / csrc/linop.cu header)
    # Under torchrun (N ranks) the operator is row-sharded: rank r streams M/N rows, one NCCL all-reduce of [S*n + 4]
    # floats per step (strong scaling in M: the operator is fixed).
    from henbun_b200.fused import LinearOperatorStep
    from henbun_b200 import parallel
    import torch.distributed as dist
    world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    M, n, S = 65536, 16384, int(sys.argv[2]) if len(sys.argv) > 2 else 64       # S sweep: SURVEY.md 8d ({1, 8, 64}, headline 64)
    first, rows = parallel.shard_rows(M, world, rank)
    gen = torch.Generator(device="cuda").manual_seed(0)
    w_true = torch.randn(256, device="cuda", generator=gen)
    gen_r = torch.Generator(device="cuda").manual_seed(100 + rank)
    A = torch.randn(rows, n, device="cuda", generator=gen_r) / np.sqrt(n)
    y = A[:, :256] @ w_true + 0.1 * torch.randn(rows, device="cuda", generator=gen_r)
    st = LinearOperatorStep(A, y, S, m_total=M, lr=1e-3)
    st.q_sqrt.copy_(0.1 * torch.eye(n, device="cuda") + 1e-3 * torch.tril(torch.randn(n, n, device="cuda", generator=gen)))
    for i in range(3):
        st.step(None, i)
    steps = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * steps + 1)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev[0].record()
    for i in range(steps):
        st.local(None, 3 + i); ev[3 * i + 1].record()
        parallel.allreduce_sum_(st.zbar_stats); ev[3 * i + 2].record()
        st.update(); ev[3 * i + 3].record()
    torch.cuda.synchronize()
    t_local = np.median([ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(steps)])
    t_coll = np.median([ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(steps)])
    t_upd = np.median([ev[3 * i + 2].elapsed_time(ev[3 * i + 3]) for i in range(steps)])
    tt = torch.tensor([ev[0].elapsed_time(ev[-1]) / steps], device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt)
    tri = n * (n + 1) / 2
    bytes_local = 2 * 4.0 * rows * n + 4 * tri + 4.0 * S * (3 * rows + 4 * n)
    bytes_upd = 6 * 4 * tri + 4.0 * S * n * 4
    peak = 6550.1e9
    if rank == 0:
        print(f"config 5 fused (n={n}, A {M}x{n} over {world} GPU(s), S={S}): {ms:.3f} ms/step = {S * M / ms * 1e3:.3e} evals/s; ELBO {float(st.out4[0]):.1f}")
        print(f"  local  (sampler + F=ZA^T + loglik + Zbar=RA, {rows} rows): {t_local:.3f} ms, {bytes_local / 1e9:.2f} GB algorithmic -> "
              f"{bytes_local / t_local / 1e6:.0f} GB/s = {bytes_local / t_local / 1e-3 / peak:.2f} of HBM peak")
        print(f"  all-reduce of {st.zbar_stats.numel() * 4 / 1e6:.1f} MB: {t_coll:.3f} ms")
        print(f"  update (mu-bar, var-bar, fused Lbar+Adam on tril): {t_upd:.3f} ms, {bytes_upd / 1e9:.2f} GB algorithmic -> "
              f"{bytes_upd / t_upd / 1e6:.0f} GB/s = {bytes_upd / t_upd / 1e-3 / peak:.2f} of HBM peak")
        print(f"  step/GPU: {(bytes_local + bytes_upd) / 1e9:.2f} GB algorithmic -> {(bytes_local + bytes_upd) / ms / 1e6:.0f} GB/s = "
              f"{(bytes_local + bytes_upd) / ms / 1e-3 / peak:.2f} of HBM peak (6550 GB/s measured copy)", flush=True)
    if world > 1:
        # every rank must hold bit-identical parameters after the replicated update
        chk = torch.stack([st.params.double().sum(), st.params.double().abs().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("  replicas identical after", 3 + steps, "steps:", bool(torch.equal(lo, hi)), flush=True)
        dist.destroy_process_group()
    m = None
elif which == "4":
    class Amortised(hb.model.Model):
        def setUp(self, X=None, latent=64):
            self.X = hb.param.MinibatchData(X)
            self.enc = hb.nn.NeuralNet([784, 512, 512, 2 * latent], stddev=0.05)
            self.dec = hb.nn.NeuralNet([latent, 512, 512, 784], stddev=0.05)
            self.q_local = hb.variationals.Normal([latent], collections=hb.param.graph_key.LOCAL)
            self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)

        @hb.model.AutoOptimize()
        def ELBO(self):
            self.q_local = self.enc(self.X)
            x_rec = self.dec(self.q_local)
            return tf.reduce_sum(hb.densities.gaussian(self.X, x_rec, self.var)) - self.KL(hb.param.graph_key.LOCAL)

    n_data = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    X = np.random.RandomState(0).randn(n_data, 784).astype(np.float32)
    m = Amortised(X=X)
    m.ELBO().compile(n_samples=32, verbose=False)
    ms = timed(m.ELBO(), 10, minibatch_size=4096)
    print(f"config 4 (784-512-512-2x64 encoder + mirrored decoder, {n_data} points, minibatch 4096, S=32): {ms:.2f} ms/step "
          f"= {32 * 4096 / ms * 1e3:.3e} evals/s", flush=True)
else:
    class LinearOperator(hb.model.Model):
        def setUp(self, A=None, y=None):
            self.A = hb.param.Data(A)
            self.y = hb.param.Data(y)
            self.q = hb.variationals.Normal([A.shape[1]], q_shape='fullrank', stddev=0.1)
            self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)

        @hb.model.AutoOptimize()
        def ELBO(self):
            f = tf.matmul(self.q, self.A, transpose_b=True)
            return tf.reduce_sum(hb.densities.gaussian(self.y, f, self.var)) - self.KL()

    M, n, S = 65536, 16384, 64
    rng = np.random.RandomState(0)
    A = (rng.randn(M, n) / np.sqrt(n)).astype(np.float32)
    y = (A[:, :256] @ rng.randn(256) + 0.1 * rng.randn(M)).astype(np.float32)
    m = LinearOperator(A=A, y=y)
    m.ELBO().compile(n_samples=S, verbose=False)
    ms = timed(m.ELBO(), 5)
    print(f"config 5 (n=16384 full-covariance q, A 65536x16384, S=64): {ms:.2f} ms/step = {S * M / ms * 1e3:.3e} evals/s", flush=True)

if os.environ.get("HB_TORCH_PROF") == "1" and which == "5f":
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(3):
            st.step(None, 20 + i)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=90))
elif os.environ.get("HB_TORCH_PROF") == "1":
    from torch.profiler import profile, ProfilerActivity
    opt = m.ELBO()
    kw = dict(minibatch_size=4096) if which == "4" else {}
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        opt.optimize(maxiter=3, **kw)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
