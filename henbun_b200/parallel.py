"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  The model-level collective of a step is ONE all-reduce of the packed gradient: the
MC-sample axis (or the minibatch, or the operator's rows) is sharded.  The dense-GP step
additionally shares one column-block-cyclic Cholesky + reverse mode between the ranks; those panel
broadcasts run inside the library on its own NCCL communicator (``block_cyclic_env`` below, csrc/comm.cu).
The reference has no distributed concept at all."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_samples(n_samples_total: int, world_size: int, rank: int):
    """Contiguous split of the global sample axis; returns (first_sample, count) for this rank."""
    base, rem = divmod(int(n_samples_total), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def shard_rows(n_rows_total: int, world_size: int, rank: int):
    """Contiguous split of the rows of a dense operator / data set (config 5: A is row-sharded);
    returns (first_row, count).  Same rule as shard_samples."""
    return shard_samples(n_rows_total, world_size, rank)


def rank_philox_offset(step: int, first_sample: int, per_sample: int, total_samples: int) -> int:
    """Philox stream position of this rank's first draw at `step`: ranks read disjoint windows of ONE
    global stream, so the union over ranks equals the single-GPU draw of the same step."""
    stride = (int(total_samples) * int(per_sample) + 3) // 4 * 4
    return int(step) * stride + (int(first_sample) * int(per_sample)) // 4 * 4


def allreduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks of the packed gradient buffer (the step's single collective)."""
    w, _ = world()
    if w > 1:
        dist.all_reduce(flat)
    return flat


# ---- column-block-cyclic factorisations (csrc/linalg.cu: potrf_flat / chol_rev_flat) --------------------------------
_block_comm = {}


def default_block_cyclic(world_size: int):
    """(block, batch, turn, block_bwd) measured best at N=65536 (tools/flat_probe.py, profiles/README.md): 2048-column blocks on 2
    ranks (each rank is busy ~95 % of the time; narrower blocks only add exchanges).  From 4 ranks on the block chain is the
    critical path: blocks are dealt two at a time (the broadcast of a block overlaps the owner's next one), 1024 columns wide
    in the forward pass and 2048 in the reverse mode (potrf + reverse at 4 GPUs: 2048x1 151 + 253 ms, 1024x2 134 + 268,
    2048x2 140 + 241)."""
    return (1024, 1, 2, 2048) if world_size >= 4 else (2048, 1, 1, 0)


def block_cyclic_env(block: int = None, batch: int = None, shard_samples: bool = False, turn: int = None, block_bwd: int = None):
    """``_lib.Dist`` describing this rank's place in a column-block-cyclic Cholesky / reverse mode over all ranks of the
    default process group.  The library runs its own NCCL communicator (panel broadcasts on a dedicated stream): rank 0
    draws the unique id, torch.distributed carries the 128 bytes to the other ranks, every rank joins with its current
    CUDA device.  One communicator per process, created on first use."""
    import ctypes as C
    from . import _lib
    w, r = world()
    dflt = default_block_cyclic(w)
    if block_bwd is None:                         # an explicit block width applies to both passes unless told otherwise
        block_bwd = dflt[3] if block is None else 0
    block = dflt[0] if block is None else block
    batch = dflt[1] if batch is None else batch
    turn = dflt[2] if turn is None else turn
    if w == 1:
        return _lib.Dist(None, 0, 1, int(block), 0, int(batch), int(block_bwd), int(turn))
    comm = _block_comm.get("comm")
    if comm is None:
        lib = _lib.load()
        ident = (C.c_ubyte * 128)()
        if r == 0:
            _lib.check(lib.hb_comm_unique_id(C.cast(ident, C.c_void_p)), "hb_comm_unique_id")
        t = torch.tensor(list(ident), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().tolist())
        buf = (C.c_ubyte * 128).from_buffer_copy(raw)
        out = C.c_void_p()
        # NCCL prints its version banner to stdout when a communicator is created: keep stdout clean for callers that print
        # machine-readable lines there (bench.py) by pointing fd 1 at stderr for the duration of the call
        import os
        import sys
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            rc = lib.hb_comm_create(C.cast(buf, C.c_void_p), r, w, C.byref(out))
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        _lib.check(rc, "hb_comm_create")
        comm = out.value
        _block_comm["comm"] = comm
    return _lib.Dist(comm, r, w, int(block), 1 if shard_samples else 0, int(batch), int(block_bwd), int(turn))
