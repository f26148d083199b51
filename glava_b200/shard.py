"""Multi-GPU plumbing: streams are independent, so the batch is sharded stream-wise across ranks
with NO data-path collective.  torch.distributed is used only to line ranks up
(barrier) and to reduce the timing (max over ranks)."""


def shard_streams(total_streams, world, rank):
    """contiguous block partition: (first_stream, count) of `rank`; counts differ by at most 1"""
    base, extra = divmod(int(total_streams), int(world))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (identity when torch.distributed is not initialised)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_objects(obj):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
