"""Batch-sharded data parallelism of the RAU step: host-side helpers.

The reference is single-GPU (F:128-146); the only contract is "N ranks on a batch split N ways == one rank on the whole
batch".  Rank r owns rows [lo, hi) of feats / tokens / lengths / labels, every rank holds a full replica of the
parameters and optimizer state, local losses and dscore are scaled by 1/B_global (rau_batch.B_global), and the three flat
gradients are summed over ranks between backward and noise/clip/optimizer.  On the GPU the sum is ncclAllReduce issued
by librau itself (rau_comm.cu); `allreduce_flat` below is the same reduction over torch.distributed for host-side tests
(gloo) and for callers that own the process group.
"""
from __future__ import annotations


def shard_rows(B_global: int, rank: int, world: int):
    """[lo, hi) rows of the global batch owned by `rank`: contiguous, sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(B_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(feats, tokens, lengths, labels, rank: int, world: int):
    """Slices of one global batch for `rank` (tokens are [T, B]: the batch is dim 1)."""
    lo, hi = shard_rows(feats.shape[0], rank, world)
    return feats[lo:hi], tokens[:, lo:hi], lengths[lo:hi], labels[lo:hi]


def allreduce_flat(grads, group=None):
    """Sum each flat gradient over the ranks of a torch.distributed group, in place; returns the list."""
    import torch.distributed as dist
    for g in grads:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    return grads
