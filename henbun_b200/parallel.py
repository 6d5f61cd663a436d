"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  The ELBO hot path needs exactly one collective per step -- an all-reduce of the packed
gradient -- because the MC-sample axis is sharded and every other quantity is replicated
(SURVEY.md 8e).  The reference has no distributed concept at all."""
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
