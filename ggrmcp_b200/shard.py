"""Multi-GPU plan of the tools/call path: items are independent, so a job shards by
batch index - rank r of W owns the contiguous block [r * per_rank, (r + 1) * per_rank) - and there is
no collective on the data path.  torch.distributed is used for the launch barrier and for the
max-over-ranks reduction of the timings only (NCCL on the GPUs, gloo in the CPU tests)."""


def shard_range(per_rank, rank, world):
    """(first_item, n_items) of `rank` in a weak-scaling job with `per_rank` items per rank."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return rank * per_rank, per_rank


def split_batch(n_total, world):
    """Contiguous blocks of a fixed batch (strong scaling): [(first, count)] per rank, sizes differ by <= 1."""
    base, extra = divmod(n_total, world)
    out, first = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


def max_over_ranks(dist, value, device="cpu"):
    """max of a python float over all ranks (identity when dist is None / world is 1)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
