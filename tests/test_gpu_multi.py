"""World-size-2 NCCL run of the public API (needs two GPUs; skipped on a one-GPU box): the sample-sharded GP step must
reproduce the single-GPU step with the same total number of samples -- ranks read disjoint windows of one Philox stream,
so the union of their draws IS the single-GPU draw and the all-reduced gradient is the same estimate."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import henbun_b200 as hb, henbun_b200.tf as tf
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
rng = np.random.RandomState(0)
n, D, S = 192, 4, 8
X = rng.randn(n, D); Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.randn(n, 1)
class GPR(hb.model.Model):
    def setUp(self):
        self.X = hb.param.Data(X); self.Y = hb.param.Data(Y)
        self.q = hb.variationals.Gaussian(shape=[n, 1], q_shape='diagonal')
        self.kern = hb.gp.kernels.UnitRBF()
        self.k_var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)
        self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)
    @hb.model.AutoOptimize()
    def ELBO(self):
        y_fit = tf.matmul(self.kern.Cholesky(self.X), self.q) * tf.sqrt(self.k_var)
        return tf.reduce_sum(hb.densities.gaussian(self.Y, y_fit, self.var)) - self.KL()
out = {}
for fused in (True, False):
    np.random.seed(5)
    m = GPR()
    m.ELBO().compile(optimizer=tf.train.AdamOptimizer(0.01), n_samples=S, seed=3, shard='samples', verbose=False, fused=fused)
    for _ in range(4):
        m.ELBO().optimize(maxiter=1)
    out[str(fused)] = [float(x) for x in np.concatenate([m.q.q_mu._free_numpy().ravel()[:16], m.kern.lengthscales._free_numpy().ravel(),
                                                          m.var._free_numpy().ravel()])]
if rank == 0:
    print("RESULT " + json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
'''


def _run(world):
    code = WORKER % {"root": ROOT}
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    if world == 1:
        cmd = [sys.executable, "-c", code]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", "29731", "--no-python", sys.executable, "-c", code]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_sample_sharded_api_step_on_two_gpus_equals_one_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    one, two = _run(1), _run(2)
    for k in ("True", "False"):
        assert np.allclose(one[k], two[k], rtol=2e-4, atol=2e-6), k
    assert np.allclose(one["True"], one["False"], rtol=2e-4, atol=2e-6)
