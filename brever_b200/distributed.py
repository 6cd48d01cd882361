"""Multi-GPU plumbing for the front-end: shard utterances by rank, reduce metrics.

The hot path has no data-path collective — every kernel is per utterance — so the
only communication is the final metric reduction the reference performs once per
epoch (``brever/training.py:369-373``: ``dist.reduce`` then ``/ world_size``).
Here it is one ``all_reduce(SUM)`` of ``[sum, count]`` so that uneven shards
still give the exact global mean.  Backend: NCCL on GPUs (gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world_size):
    """Contiguous [start, stop) slice of `n_items` utterances owned by `rank`.

    Same partitioning idea as the reference's DistributedBatchSamplerWrapper
    (``brever/batching.py:279-290``) at utterance granularity: sizes differ by
    at most one and every item belongs to exactly one rank.
    """
    if not 0 <= rank < world_size:
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard(tensor, rank=None, world_size=None):
    """The slice of dim 0 owned by this rank."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    start, stop = shard_bounds(tensor.shape[0], rank, world_size)
    return tensor[start:stop]


def global_mean(values):
    """Mean of a per-utterance metric over ALL ranks (one 2-element all-reduce)."""
    packed = torch.stack([values.double().sum(),
                          torch.tensor(float(values.numel()), dtype=torch.float64,
                                       device=values.device)])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    return (packed[0] / packed[1].clamp_min(1)).to(values.dtype)
