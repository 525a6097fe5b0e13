"""Data-parallel helpers for the two paths that shard by batch: inference needs no collective; the
reconstruction step needs exactly one all-reduce(SUM) of the flat factor-gradient buffer per iteration
(the reference reaches the same collective through DistributedDataParallel, hooks/finetune.py:155-160)."""
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """[lo, hi) of the rank's contiguous share of n samples (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_flat_(flat, group=None):
    """In-place all-reduce(SUM).  Callers pre-scale by 1/B_global, so the sum over ranks IS the full-batch gradient
    even when shards are ragged (the loss is a mean over samples of per-sample norms, hooks/finetune.py:320)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
