"""BASELINE configs 4 and 5 at full size through the Henbun-shaped Python API (model.optimize): ms per Adam step.
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


if which == "4":
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

if os.environ.get("HB_TORCH_PROF") == "1":
    from torch.profiler import profile, ProfilerActivity
    opt = m.ELBO()
    kw = dict(minibatch_size=4096) if which == "4" else {}
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        opt.optimize(maxiter=3, **kw)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
