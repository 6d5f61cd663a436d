"""world_size-2 gloo test (CPU) of the multi-rank plumbing used by bench.py / Optimizer: sample-axis
sharding, disjoint Philox windows, and the single all-reduce of the packed gradient."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from henbun_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = parallel.shard_samples(7, world, rank)
    g = torch.full((10,), float(rank + 1)) * count           # "gradient summed over my samples"
    parallel.allreduce_sum_(g)
    w, r = parallel.world()
    out[rank] = (first, count, g.numpy().copy(), w, r)
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    world = 2
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (f0, c0, g0, w0, r0), (f1, c1, g1, w1, r1) = out[0], out[1]
    assert (f0, c0, f1, c1) == (0, 4, 4, 3) and (w0, r0, w1, r1) == (2, 0, 2, 1)
    assert np.allclose(g0, 1 * 4 + 2 * 3) and np.allclose(g0, g1)      # every rank holds the same summed gradient


def test_philox_windows_are_disjoint_and_cover_the_stream():
    S, per = 64, 1000
    for world in (1, 2, 4, 8):
        spans = []
        for r in range(world):
            first, cnt = parallel.shard_samples(S, world, r)
            off = parallel.rank_philox_offset(3, first, per, S)
            assert off % 4 == 0
            spans.append((off, off + cnt * per))
        spans.sort()
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 <= b0
        assert spans[0][0] == 3 * S * per and spans[-1][1] == 4 * S * per
