"""hb_amortised_elbo_step (csrc/amortised.cu), the whole-step entry point of BASELINE config 4, against the oracle's
amortised_elbo in fp64: ELBO and every gradient, at small ragged sizes (all activations, 1-3 layers) and at the named
size (784-512-512-2x64 + mirrored decoder, minibatch 4096, S = 32); and its binding behind the Henbun API."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import henbun_oracle as O

pytestmark = pytest.mark.gpu
ACT = {"sigmoid": 1, "relu": 2, "tanh": 3}


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def run_case(enc_nodes, dec_nodes, enc_acts, dec_acts, B, S, seed, device_oracle="cpu", scale=0.3):
    from henbun_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(seed)
    lat = dec_nodes[0]
    X = rng.randn(B, enc_nodes[0]).astype(np.float32)
    U = rng.randn(S, B, lat).astype(np.float32)
    p, flat = {}, []
    for name, nodes in (("enc", enc_nodes), ("dec", dec_nodes)):
        for i in range(len(nodes) - 1):
            w = (scale * rng.randn(nodes[i], nodes[i + 1]) / np.sqrt(nodes[i]) * 3).astype(np.float32)
            b = (0.1 * rng.randn(1, nodes[i + 1])).astype(np.float32)
            p[f"{name}.w{i}"], p[f"{name}.b{i}"] = w.astype(np.float64), b.astype(np.float64)
            flat += [w.ravel(), b.ravel()]
    var = np.array([0.3], np.float32); p["var"] = var.astype(np.float64); flat.append(var)
    cfg = _lib.AmortisedConfig()
    cfg.B, cfg.S, cfg.latent = B, S, lat
    cfg.n_enc, cfg.n_dec = len(enc_nodes) - 1, len(dec_nodes) - 1
    for i, w in enumerate(enc_nodes): cfg.enc_nodes[i] = w
    for i, w in enumerate(dec_nodes): cfg.dec_nodes[i] = w
    for i, a in enumerate(enc_acts): cfg.enc_act[i] = ACT[a]
    for i, a in enumerate(dec_acts): cfg.dec_act[i] = ACT[a]
    npar = lib.hb_amortised_param_count(C.byref(cfg))
    params = torch.from_numpy(np.concatenate(flat)).cuda()
    assert npar == params.numel()
    grads = torch.zeros(npar, device="cuda"); out4 = torch.zeros(4, device="cuda")
    wsb = lib.hb_amortised_workspace_bytes(C.byref(cfg)); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    Xd, Ud = torch.from_numpy(X).cuda(), torch.from_numpy(U).cuda()
    rc = lib.hb_amortised_elbo_step(C.byref(cfg), _lib.ptr(Xd), _lib.ptr(params), _lib.ptr(Ud), _lib.ptr(grads), _lib.ptr(out4),
                                    _lib.ptr(ws), wsb, _lib.stream())
    torch.cuda.synchronize()
    assert rc == 0
    ref, gref = O.value_and_grads(lambda pp, X_, U_: O.amortised_elbo(pp, X_, U_, list(enc_acts), list(dec_acts)), p,
                                  X.astype(np.float64), U.astype(np.float64), device=device_oracle)
    got = grads.double().cpu().numpy()
    errs, off = {}, 0
    for name, nodes in (("enc", enc_nodes), ("dec", dec_nodes)):
        for i in range(len(nodes) - 1):
            for k, sz in ((f"{name}.w{i}", nodes[i] * nodes[i + 1]), (f"{name}.b{i}", nodes[i + 1])):
                errs[k] = rel_err(got[off:off + sz], gref[k].ravel()); off += sz
    errs["var"] = rel_err(got[off:off + 1], gref["var"])
    return out4.double().cpu().numpy(), ref, errs


@pytest.mark.parametrize("enc,dec,ea,da,B,S", [
    ([12, 16, 8], [4, 16, 12], ["sigmoid"], ["sigmoid"], 32, 5),
    ([7, 6], [3, 7], [], [], 9, 1),                                        # single linear layers
    ([20, 33, 17, 10], [5, 9, 31, 20], ["tanh", "relu"], ["relu", "tanh"], 50, 3),
    ([64, 128, 128, 32], [16, 128, 128, 64], ["sigmoid", "sigmoid"], ["sigmoid", "sigmoid"], 512, 8),
])
def test_amortised_step_matches_oracle(enc, dec, ea, da, B, S):
    out4, ref, errs = run_case(enc, dec, ea, da, B, S, seed=B + S)
    assert abs(out4[0] - ref) <= 1e-5 * abs(ref)
    bar = 5e-5 if "relu" in ea + da else 2e-5        # relu: a pre-activation within rounding of 0 may flip its branch
    for k, e in errs.items():
        assert e < bar, (k, e)


def test_amortised_step_named_size_matches_fp64_oracle():
    """784-512-512-2x64 encoder + mirrored decoder, minibatch 4096, S = 32: the big MatBias products run on the tcgen05
    engine, the dW reductions over 131072 rows by split-K."""
    out4, ref, errs = run_case([784, 512, 512, 128], [64, 512, 512, 784], ["sigmoid", "sigmoid"], ["sigmoid", "sigmoid"], 4096, 32,
                               seed=0, device_oracle="cuda", scale=0.05 / 3 * np.sqrt(512))
    assert abs(out4[0] - ref) <= 1e-5 * abs(ref)
    for k, e in errs.items():
        assert e < 1e-5, (k, e)


def test_amortised_binding_matches_the_eager_tape():
    import henbun_b200 as hb
    import henbun_b200.tf as tf

    class Amortised(hb.model.Model):
        def setUp(self, X=None, latent=4, hidden=16):
            self.X = hb.param.MinibatchData(X)
            self.enc = hb.nn.NeuralNet([X.shape[1], hidden, 2 * latent], stddev=0.3, neuron_types=tf.tanh)
            self.dec = hb.nn.NeuralNet([latent, hidden, X.shape[1]], stddev=0.3)
            self.q_local = hb.variationals.Normal([latent], collections=hb.param.graph_key.LOCAL)
            self.var = hb.param.Variable(shape=[1], transform=hb.transforms.positive)

        @hb.model.AutoOptimize()
        def ELBO(self):
            self.q_local = self.enc(self.X)
            x_rec = self.dec(self.q_local)
            return tf.reduce_sum(hb.densities.gaussian(self.X, x_rec, self.var)) - self.KL(hb.param.graph_key.LOCAL)

    rng = np.random.RandomState(0)
    Xall = rng.randn(500, 12).astype(np.float32)
    res = []
    for fused in (True, False):
        np.random.seed(11)
        m = Amortised(X=Xall)
        m.ELBO().compile(optimizer=tf.train.AdamOptimizer(0.01), n_samples=5, seed=2, verbose=False, fused=fused)
        assert (m.ELBO().fused_entry == 'AmortisedElboBinding') == fused
        objs = [float(m.ELBO().optimize(maxiter=1, minibatch_size=64)) for _ in range(4)]
        res.append((objs, [v._free_numpy().astype(np.float64).ravel() for v in m.get_variables() if v.is_parameter]))
    assert np.allclose(res[0][0], res[1][0], rtol=2e-5)
    for a, b in zip(res[0][1], res[1][1]):
        assert np.allclose(a, b, rtol=1e-4, atol=2e-5)
